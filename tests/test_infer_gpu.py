"""InferencePipeline (double-buffered host -> device -> host loop) returns what the module returns."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_direct_calls():
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[1, 1], dims=[32, 64], num_classes=5).cuda().eval()
    g = torch.Generator().manual_seed(1)
    batches = [torch.randn(3, 3, 32, 32, generator=g).pin_memory() for _ in range(5)]
    batches.append(torch.randn(2, 3, 32, 32, generator=g).pin_memory())          # ragged last batch
    pipe = mm.InferencePipeline(net, autocast_dtype=None)
    got = list(pipe.stream(batches))
    assert len(got) == len(batches)
    with torch.no_grad():
        for x, y in zip(batches, got):
            want = net(x.cuda()).float().cpu()
            assert y.shape == want.shape and not y.is_cuda
            assert torch.equal(y, want)
    one = pipe(batches[0])
    assert torch.equal(one, got[0])
    assert list(pipe.stream([])) == []


def test_pipeline_rejects_cpu_module():
    import medmamba_b200 as mm
    with pytest.raises(RuntimeError):
        mm.InferencePipeline(mm.VSSM(depths=[1], dims=[32], num_classes=2))

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


@pytest.mark.parametrize("amp", [None, torch.bfloat16])
def test_graphed_forward_replays_the_eager_forward(amp):
    """GraphedForward: the forward captured as one CUDA graph (the two-stream branch overlap included) returns, for new
    contents of its static input, exactly what the eager forward returns."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[1, 2], dims=[32, 64], num_classes=5).cuda().eval()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    gf = mm.GraphedForward(net, x, amp)
    for seed in (1, 2, 3):
        x.copy_(torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(seed)))
        got = gf.replay().clone()
        with torch.no_grad(), torch.autocast("cuda", dtype=amp or torch.bfloat16, enabled=amp is not None):
            want = net(x)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    # the pipeline picks the graph for small batches on its own and the eager path when told to
    batches = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(10 + i)).pin_memory() for i in range(4)]
    a = list(mm.InferencePipeline(net, autocast_dtype=amp, cuda_graph="auto").stream(batches))
    b = list(mm.InferencePipeline(net, autocast_dtype=amp, cuda_graph=False).stream(batches))
    assert all(torch.equal(p, q) for p, q in zip(a, b))


def test_pipeline_rejects_cpu_module():
    import medmamba_b200 as mm
    with pytest.raises(RuntimeError):
        mm.InferencePipeline(mm.VSSM(depths=[1], dims=[32], num_classes=2))


def test_gradcam_style_hooks_on_the_reference_target_layer():
    """Grad-CAM consumer contract (test.py:101-108, grad_cam/utils.py:5-49): eval-mode model, forward and full
    backward hooks on net.layers[-1].blocks[-1].conv33conv33conv11[-2], batch-1 backward from the top logit.
    The hooked layer must be called (no fused bypass), activations / gradients must match the module path."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[1, 1], dims=[32, 64], num_classes=3).cuda().eval()
    target = net.layers[-1].blocks[-1].conv33conv33conv11[-2]
    acts, grads = [], []
    h1 = target.register_forward_hook(lambda m, i, o: acts.append(o.detach()))
    h2 = target.register_full_backward_hook(lambda m, gi, go: grads.append(go[0].detach()))
    x = torch.randn(1, 3, 64, 64, device="cuda")
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # both paths in true fp32
    try:
        with torch.no_grad():
            net(x)
        assert len(acts) == 1, "a hooked layer must be called even under no_grad (fast path must step aside)"
        acts.clear()
        logits = net(x)
        net.zero_grad()
        logits[0, logits.argmax()].backward()
        assert len(acts) == 1 and len(grads) == 1 and acts[0].shape == grads[0].shape
        fused_act, fused_grad = acts[0].clone(), grads[0].clone()
        for m in net.modules():                      # reference op order everywhere
            if isinstance(m, (mm.SS2D, mm.PatchEmbed2D, mm.PatchMerging2D)):
                m.fused = False
        acts.clear(); grads.clear()
        logits_ref = net(x)
        net.zero_grad()
        logits_ref[0, logits_ref.argmax()].backward()
    finally:
        h1.remove(); h2.remove()
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert torch.allclose(logits, logits_ref, rtol=1e-3, atol=1e-4)
    assert torch.allclose(fused_act, acts[0], rtol=1e-3, atol=1e-4)
    assert torch.allclose(fused_grad, grads[0], rtol=1e-2, atol=1e-5)
    with torch.no_grad():                            # hooks removed: the fast path is back
        y = net(x)
    assert torch.allclose(y, logits.detach(), rtol=1e-3, atol=1e-4)


def test_branch_overlap_is_bitwise_neutral(monkeypatch):
    """The CNN branch on a side stream (SS_Conv_SSM.forward) changes scheduling only: same logits bit for bit."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[2, 2], dims=[32, 64], num_classes=4).cuda().eval()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    outs = {}
    for flag in ("1", "0", "1"):
        monkeypatch.setenv("MMB_BRANCH_OVERLAP", flag)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(3):
                y = net(x)
        torch.cuda.synchronize()
        outs.setdefault(flag, []).append(y.float().cpu())
    assert torch.equal(outs["1"][0], outs["0"][0]) and torch.equal(outs["1"][0], outs["1"][1])

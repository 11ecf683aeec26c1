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


def test_graphed_pipeline_follows_weight_updates():
    """A captured graph computes with tensors derived from the weights at capture time (folded BN, -exp(A_logs), packed
    x_proj).  After load_state_dict / an in-place update the pipeline must capture again: logits equal the eager ones."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[1, 1], dims=[32, 64], num_classes=5).cuda().eval()
    pipe = mm.InferencePipeline(net, autocast_dtype=None, cuda_graph=True)
    batches = [torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(i)).pin_memory() for i in range(3)]

    def eager():
        with torch.no_grad():
            return [net(b.cuda()).float().cpu() for b in batches]

    first = list(pipe.stream(batches))
    assert all(torch.equal(p, q) for p, q in zip(first, eager()))
    sd = {k: (v * 1.25 if v.is_floating_point() else v.clone()) for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    second = list(pipe.stream(batches))
    assert all(torch.equal(p, q) for p, q in zip(second, eager()))
    assert not any(torch.equal(p, q) for p, q in zip(first, second))
    with torch.no_grad():
        net.head.weight.mul_(-1.0)
    third = list(pipe.stream(batches))
    assert all(torch.equal(p, q) for p, q in zip(third, eager()))
    gf = pipe._graphs[0][1]
    assert not gf.stale()
    with torch.no_grad():
        net.layers[0].blocks[0].self_attention.A_logs.add_(0.1)
    assert gf.stale()


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


@pytest.mark.parametrize("amp", [None, torch.bfloat16])
def test_derived_parameter_cache_follows_the_parameters(amp, monkeypatch):
    """ops._derived_params caches what the inference path derives from an SS2D module's parameters (packed x_proj weight,
    A = -exp(A_logs), sum_k Ds_k).  Every way a parameter can change -- an in-place update (optimizer step), load_state_dict,
    rebinding `.data` -- must give the logits of an uncached forward, bit for bit."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    net = mm.VSSM(depths=[1, 1], dims=[32, 64], num_classes=5).cuda().eval()
    x = torch.randn(3, 3, 64, 64, device="cuda")

    def run(cache):
        monkeypatch.setenv("MMB_PARAM_CACHE", "1" if cache else "0")
        with torch.no_grad(), torch.autocast("cuda", dtype=amp or torch.bfloat16, enabled=amp is not None):
            return net(x).float().clone()

    ss = [m for m in net.modules() if isinstance(m, mm.SS2D)]

    def derived_are_current():
        # the logits of this tiny net need not move under bf16 rounding, so the cached tensors themselves are checked too
        monkeypatch.setenv("MMB_PARAM_CACHE", "1")
        for m in ss:
            with torch.no_grad():
                w, Wdt, b, A, Dc, dsum = mm.ops._derived_params(m.x_proj_weight, m.dt_projs_weight, m.dt_projs_bias, m.A_logs,
                                                               m.Ds, m.d_state, m.dt_rank, amp is not None)
                wp = mm.ops.pack_x_proj(m.x_proj_weight.float(), m.d_state, m.dt_rank)
                assert torch.equal(w, wp.to(torch.bfloat16) if amp is not None else wp)
                assert torch.equal(Wdt, m.dt_projs_weight.float()) and torch.equal(b, m.dt_projs_bias.float())
                assert torch.equal(A, -torch.exp(m.A_logs.float())) and torch.equal(Dc, m.Ds.float())
                assert torch.equal(dsum, m.Ds.float().view(4, -1).sum(0))

    first = run(True)
    assert torch.equal(first, run(True)) and torch.equal(first, run(False))
    derived_are_current()
    with torch.no_grad():
        ss[0].x_proj_weight.mul_(1.5)                          # in-place: _version changes
        ss[1].A_logs.add_(0.25)
        ss[0].Ds.mul_(-2.0)
    second = run(True)
    assert not torch.equal(first, second)
    assert torch.equal(second, run(False))
    derived_are_current()
    with torch.no_grad():
        ss[1].dt_projs_bias.data = ss[1].dt_projs_bias.data + 0.5      # rebinding: data_ptr changes
    third = run(True)
    assert torch.equal(third, run(False))
    derived_are_current()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    for k in sd:
        if k.endswith("A_logs") or k.endswith("x_proj_weight"):
            sd[k] = sd[k] * 0.9
    net.load_state_dict(sd)
    fourth = run(True)
    assert torch.equal(fourth, run(False))
    derived_are_current()
    # training mode / gradients wanted: nothing cached, gradients reach the parameters
    net.train()
    out = net(x)
    out.sum().backward()
    assert ss[0].x_proj_weight.grad is not None and ss[1].A_logs.grad is not None

"""Training path: gradients of a whole SS2D block / tiny VSSM through the CUDA kernels against torch
autograd through the functional oracle (small shapes: autograd through selective_scan_ref is O(L^2))."""
import pytest
import torch

from oracle import medmamba_ref
from tests.util import assert_close
from oracle.selective_scan_ref import selective_scan_ref

pytestmark = pytest.mark.gpu


def _ref_scan64(*a, **k):
    return selective_scan_ref(*a, **k)


def test_ss2d_block_gradients_vs_oracle_autograd():
    import medmamba_b200 as mm
    torch.manual_seed(0)
    m = mm.SS2D(d_model=16).train()
    with torch.no_grad():
        m.A_logs.add_(0.2 * torch.randn_like(m.A_logs))
        m.x_proj_weight.mul_(3.0)
    x = torch.randn(2, 6, 5, 16)
    # oracle: autograd through the functional restatement, fp32
    sd = {k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    xr = x.clone().requires_grad_()
    yr = medmamba_ref.ss2d_forward(sd, "", xr)
    gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1))
    yr.backward(gy)
    # product, on the GPU
    m = m.cuda()
    xg = x.cuda().requires_grad_()
    yg = m(xg)
    yg.backward(gy.cuda())
    scale = lambda t: max(t.abs().max().item(), 1e-12)
    assert (yg.detach().cpu() - yr.detach()).abs().max().item() / scale(yr) < 1e-4
    assert (xg.grad.cpu() - xr.grad).abs().max().item() / scale(xr.grad) < 1e-3
    for name, p in m.named_parameters():
        ref = sd[name].grad
        err = (p.grad.cpu() - ref).abs().max().item() / scale(ref)
        assert err < 2e-3, f"{name}: {err:.2e}"


def test_tiny_vssm_training_step_matches_oracle():
    import medmamba_b200 as mm
    torch.manual_seed(1)
    net = mm.VSSM(depths=[1, 1, 1, 1], dims=[16, 32, 64, 128], num_classes=3, drop_path_rate=0.0).eval()  # eval: BN running stats
    x = torch.randn(2, 3, 32, 32)
    target = torch.tensor([0, 2])
    sd = {k: (v.detach().clone().requires_grad_() if (v.is_floating_point() and "running_" not in k) else v.detach().clone())
          for k, v in net.state_dict().items()}
    loss_r = torch.nn.functional.cross_entropy(medmamba_ref.vssm_forward(sd, x, depths=(1, 1, 1, 1)), target)
    loss_r.backward()
    net = net.cuda()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        loss_g = torch.nn.functional.cross_entropy(net(x.cuda()), target.cuda())
        loss_g.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(loss_g.item() - loss_r.item()) < 1e-5
    worst = 0.0
    for name, p in net.named_parameters():
        ref = sd[name].grad
        if ref is None:
            continue
        err = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-8)
        worst = max(worst, err)
        assert err < 5e-3, f"{name}: {err:.2e}"
    print("worst relative gradient error", worst)


@pytest.mark.parametrize("d_model,H,W,B", [(48, 20, 20, 2), (96, 9, 13, 3), (192, 14, 14, 2), (384, 7, 7, 2), (8, 3, 5, 1),
                                            (48, 56, 56, 1)])
def test_fused_backward_matches_interface_backward(d_model, H, W, B):
    """The fused path's hand-written backward kernels against the reference-order path whose only custom
    backward is mmb_scan_bwd (itself pinned to the fp64 oracle in tests/test_scan_bwd_gpu.py)."""
    import medmamba_b200 as mm
    torch.manual_seed(d_model + H)
    m = mm.SS2D(d_model=d_model).cuda().train()
    with torch.no_grad():
        m.A_logs.add_(0.2 * torch.randn_like(m.A_logs))
        m.x_proj_weight.mul_(3.0)
        m.Ds.add_(0.3 * torch.randn_like(m.Ds))
    x = torch.randn(B, H, W, d_model, device="cuda")
    gy = torch.randn(B, H, W, d_model, device="cuda")
    grads = {}
    for fused in (True, False):
        m.fused = fused
        m.zero_grad()
        xin = x.clone().requires_grad_()
        y = m(xin)
        y.backward(gy)
        grads[fused] = dict({n: p.grad.clone() for n, p in m.named_parameters()}, x=xin.grad.clone(), y=y.detach())
    for name in grads[True]:
        a, b = grads[True][name], grads[False][name]
        err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
        assert err < 2e-3, f"{name}: fused vs interface backward differ by {err:.2e}"


def _core_problem(B, H, W, D, R, seed, N=16):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    rp = {1: 4, 2: 4, 3: 4, 6: 8, 12: 12, 24: 24}[R]
    xc = 0.3 * rn(B, H, W, D)
    proj = 0.2 * rn(B, H, W, 4, 32 + rp)
    proj[..., 32 + R:] = 0
    Wdt = rn(4, D, R) * R ** -0.5
    bias = rn(4, D) - 3.0
    A = -torch.exp(0.3 * rn(4 * D, N)) * torch.arange(1, N + 1)
    Ds = rn(4 * D)
    dy = rn(B, H, W, D)
    return xc, proj, Wdt, bias, A.contiguous(), Ds, dy


@pytest.mark.parametrize("xc_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,D,R", [(2, 56, 56, 96, 3), (2, 28, 28, 192, 6), (4, 14, 14, 384, 12), (4, 7, 7, 768, 24),
                                        (1, 9, 5, 40, 3), (3, 33, 6, 24, 2)])
def test_core_bwd_vs_analytic_fp64_oracle(B, H, W, D, R, xc_dtype):
    """mmb_ss2d_core_bwd DIRECTLY against the analytic fp64 backward of the scan (oracle/scan_ref.c, SURVEY.md
    Appendix B) composed with the cross-scan index maps of Appendix A -- at the four MedMamba-T stage shapes, not
    through the repo's other backward kernel.  Every output within 5e-4 of its own max-norm."""
    from oracle import cscan
    from medmamba_b200 import ops
    from medmamba_b200.fused_autograd import core_bwd
    N = 16
    xc, proj, Wdt, bias, A, Ds, dy = _core_problem(B, H, W, D, R, seed=H * W + D + R)
    if xc_dtype != torch.float32:
        xc = xc.to(xc_dtype).float()              # the oracle sees the values the kernel reads
    L = H * W
    src = torch.from_numpy(medmamba_ref.cross_scan_index(H, W))                       # (4, L): position read at step l
    xt, pt, dyt = xc.view(B, L, D), proj.view(B, L, 4, -1), dy.view(B, L, D)
    u = torch.stack([xt[:, src[k]].transpose(1, 2) for k in range(4)], 1).reshape(B, 4 * D, L)
    dout = torch.stack([dyt[:, src[k]].transpose(1, 2) for k in range(4)], 1).reshape(B, 4 * D, L)
    dtr = [pt[:, src[k], k, 32:32 + R] for k in range(4)]                            # (B, L, R) per direction
    draw = torch.stack([torch.einsum("blr,dr->bdl", dtr[k], Wdt[k]) for k in range(4)], 1).reshape(B, 4 * D, L)
    Bm = torch.stack([pt[:, src[k], k, 0:N].transpose(1, 2) for k in range(4)], 1)   # (B, 4, N, L)
    Cm = torch.stack([pt[:, src[k], k, 16:16 + N].transpose(1, 2) for k in range(4)], 1)
    want = cscan.scan_bwd(u.contiguous(), draw.contiguous(), A, Bm.contiguous(), Cm.contiguous(), Ds, None,
                          bias.reshape(-1), True, dout.contiguous())
    w_dudir = torch.zeros(B, L, 4, D, dtype=torch.float64)
    w_dproj = torch.zeros(B, L, 4, proj.shape[-1], dtype=torch.float64)
    w_dW = torch.zeros(4, D, R, dtype=torch.float64)
    for k in range(4):
        du_k = want["du"][:, k * D:(k + 1) * D]                                        # (B, D, L) in time order
        dd_k = want["ddelta"][:, k * D:(k + 1) * D]
        w_dudir[:, src[k], k] = du_k.transpose(1, 2)
        w_dproj[:, src[k], k, 0:N] = want["dB"][:, k].transpose(1, 2)
        w_dproj[:, src[k], k, 16:16 + N] = want["dC"][:, k].transpose(1, 2)
        w_dproj[:, src[k], k, 32:32 + R] = torch.einsum("bdl,dr->blr", dd_k, Wdt[k].double())
        w_dW[k] = torch.einsum("bdl,blr->dr", dd_k, dtr[k].double())
    c = lambda t: t.cuda().contiguous()
    xg = c(xc).to(xc_dtype)
    _, hsave = ops.ss2d_core(xg, c(proj), c(Wdt), c(bias), c(A), c(Ds), N, R, save_states=True)
    dudir, dproj, dA, dW, dD, db = core_bwd(xg, c(proj), c(dy), c(Wdt), c(bias), c(A), c(Ds), hsave, N, R)
    torch.cuda.synchronize()
    checks = [("dudir", dudir.view(B, L, 4, D), w_dudir), ("dproj", dproj.view(B, L, 4, -1), w_dproj),
              ("dA", dA, want["dA"]), ("dWdt", dW, w_dW), ("dDs", dD, want["dD"]),
              ("d dt_bias", db.reshape(-1), want["ddelta_bias"])]
    for name, got, ref in checks:
        assert got.shape == ref.shape, name
        err = (got.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err < 5e-4, f"{name}: rel-to-max error {err:.2e}"


@pytest.mark.parametrize("split", ["1", "2"])
def test_fused_backward_both_lane_splits(split, monkeypatch):
    """The backward kernel's one- and two-lanes-per-channel instantiations agree with the interface backward."""
    monkeypatch.setenv("MMB_BWD_S", split)
    test_fused_backward_matches_interface_backward(96, 12, 10, 2)
    test_fused_backward_matches_interface_backward(40, 7, 7, 1)


def test_fused_backward_deterministic_and_bf16():
    import medmamba_b200 as mm
    torch.manual_seed(0)
    m = mm.SS2D(d_model=96).cuda().train()
    x = torch.randn(2, 28, 28, 96, device="cuda")
    gy = torch.randn(2, 28, 28, 96, device="cuda")
    runs = []
    for _ in range(2):
        m.zero_grad()
        xin = x.clone().requires_grad_()
        m(xin).backward(gy)
        runs.append([p.grad.clone() for p in m.parameters()] + [xin.grad.clone()])
    assert all(torch.equal(a, b) for a, b in zip(*runs)), "fused backward is not bit-reproducible"
    # bf16 autocast training step: gradients close to the fp32 ones
    m.zero_grad()
    xin = x.clone().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(xin)
    y.float().backward(gy)
    for a, p in zip(runs[0], list(m.parameters())):
        err = (p.grad.float() - a).abs().max().item() / max(a.abs().max().item(), 1e-12)
        assert err < 5e-2, err


def test_training_side_stream_is_bit_identical(monkeypatch):
    """The CNN branch of every block runs on a side stream in training (forward, and through autograd backward).  Three
    steps with the side stream give the logits and every gradient of the single-stream run bit for bit: the fork / join
    and the caching allocator's cross-stream reuse leave no race."""
    import medmamba_b200 as mm
    torch.manual_seed(0)
    x = torch.randn(8, 3, 96, 96, device="cuda")
    y = torch.randint(0, 5, (8,), device="cuda")

    def run(overlap):
        monkeypatch.setenv("MMB_TRAIN_BRANCH_OVERLAP", overlap)
        torch.manual_seed(1)
        net = mm.VSSM(depths=[2, 2, 2, 2], dims=[32, 64, 128, 256], num_classes=5).cuda().train()
        out = []
        for _ in range(3):
            net.zero_grad(set_to_none=True)
            torch.manual_seed(2)                    # DropPath draws
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = net(x)
            torch.nn.functional.cross_entropy(logits.float(), y).backward()
            torch.cuda.synchronize()
            out.append([logits.detach().clone()] + [p.grad.clone() for p in net.parameters()])
        return out

    one, two = run("0"), run("1")
    for a, b in zip(one, two):
        assert len(a) == len(b) and all(torch.equal(u, v) for u, v in zip(a, b))
    assert all(torch.equal(u, v) for u, v in zip(two[0], two[2])), "side-stream run differs from step to step"


def test_use_checkpoint_matches_plain_backward():
    """VSSM(use_checkpoint=True) (MedMamba.py:359-422: torch.utils.checkpoint around every block) recomputes each block's
    forward -- fused kernels, state checkpoints, side-stream fork and all -- inside backward: same logits and gradients."""
    import medmamba_b200 as mm
    cfg = dict(depths=[1, 2, 1, 1], dims=[32, 64, 128, 256], num_classes=5, drop_path_rate=0.0)
    x = torch.randn(4, 3, 64, 64, device="cuda")
    y = torch.randint(0, 5, (4,), device="cuda")
    res = []
    for ckpt in (False, True):
        torch.manual_seed(3)
        net = mm.VSSM(use_checkpoint=ckpt, **cfg).cuda().train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(x)
        torch.nn.functional.cross_entropy(logits.float(), y).backward()
        res.append([logits.detach()] + [p.grad.clone() for p in net.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,H,W,C,strided,bf16_out", [(2, 5, 7, 48, True, True), (1, 3, 3, 96, False, False),
                                                      (2, 14, 14, 384, True, True), (3, 4, 4, 512, False, False),
                                                      (2, 6, 6, 16, True, False)])
def test_layernorm_fn_gradients(B, H, W, C, strided, bf16_out):
    """LayerNormFn (mmb_layernorm_fwd / mmb_layernorm_bwd) against torch's LayerNorm autograd in fp64; the
    strided case is ln_1's input: the right half of the residual stream."""
    from medmamba_b200.fused_autograd import LayerNormFn
    g = torch.Generator().manual_seed(C + H)
    full = torch.randn(B, H, W, 2 * C if strided else C, generator=g)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    dy = torch.randn(B, H, W, C, generator=g)
    if bf16_out:
        dy = dy.bfloat16().float()
    xr = (full[..., C:] if strided else full).double().clone().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-6).backward(dy.double())
    fc = full.cuda().requires_grad_(True)
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    x = fc[..., C:] if strided else fc
    y = LayerNormFn.apply(x, wc, bc, 1e-6, torch.bfloat16 if bf16_out else torch.float32)
    y.backward(dy.cuda().to(y.dtype))
    gx = fc.grad[..., C:] if strided else fc.grad
    assert_close(gx, xr.grad, 1e-4, 1e-5, "LayerNormFn dx")
    assert_close(wc.grad, wr.grad, 1e-4, 1e-4, "LayerNormFn dgamma")
    assert_close(bc.grad, br.grad, 1e-4, 1e-4, "LayerNormFn dbeta")
    if strided:
        assert float(fc.grad[..., :C].abs().max()) == 0.0

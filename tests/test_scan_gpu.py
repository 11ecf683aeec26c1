"""Parity of the CUDA selective scan (through the C ABI) against the oracle.

fp32 tolerance is the one BASELINE.json's north_star states: rtol 1e-4 / atol 1e-5, checked on
`out` AND on `out - u*D` (at random init the SSM term is ~0.3 % of out).
bf16 I/O: 1e-2 relative.
"""
import pytest
import torch

from oracle import cscan
from oracle.selective_scan_ref import selective_scan_ref
from tests.util import STAGE_SHAPES, assert_close, make_scan_inputs, rel_err

pytestmark = pytest.mark.gpu


def _gpu(d):
    return {k: (v.cuda() if v is not None else None) for k, v in d.items()}


def _run(inp, softplus=True, last=False, dtype=torch.float32):
    from medmamba_b200 import selective_scan_fn
    g = _gpu(inp)
    cast = lambda t: t.to(dtype) if t is not None else None
    return selective_scan_fn(cast(g["u"]), cast(g["delta"]), g["A"], g["B"], g["C"], g["D"], cast(g["z"]),
                             g["delta_bias"], softplus, last)


@pytest.mark.parametrize("family", ["model", "stress"])
@pytest.mark.parametrize("KD,L", STAGE_SHAPES)
@pytest.mark.parametrize("batch", [2, 64])
def test_scan_fwd_stage_shapes_fp32(family, KD, L, batch):
    if batch == 64 and family == "model" and L < 3136:
        pytest.skip("covered by the stress family at this size")
    inp = make_scan_inputs(family, batch, KD, L, seed=KD + L)
    out, last = _run(inp, last=True)
    torch.cuda.synchronize()
    want, want_last = cscan.scan_fwd(**inp, delta_softplus=True, precision="f64")
    assert out.dtype == torch.float32 and out.shape == (batch, KD, L)
    # atol is stated for unit-scale outputs; the stress family reaches |out| ~ 1e2, so scale it
    sc = max(1.0, want.abs().max().item())
    assert_close(out.double().cpu() / sc, want / sc, 1e-4, 1e-5, "out")
    uD = (inp["u"] * inp["D"][None, :, None]).double()
    assert_close((out.double().cpu() - uD) / sc, (want - uD) / sc, 1e-4, 1e-5, "out - u*D")
    assert_close(last.double().cpu() / sc, want_last / sc, 1e-4, 1e-5, "last_state")


@pytest.mark.parametrize("with_z", [False, True])
@pytest.mark.parametrize("layout", ["NL", "LN"])
@pytest.mark.parametrize("batch,KD,L,G,N", [
    (1, 4, 1, 4, 16), (3, 8, 5, 4, 16), (2, 24, 37, 4, 16), (2, 40, 130, 1, 16), (5, 36, 67, 2, 8),
    (1, 96, 257, 4, 16), (2, 12, 49, 4, 3), (9, 132, 200, 4, 16),
])
def test_scan_fwd_ragged_shapes(batch, KD, L, G, N, with_z, layout):
    inp = make_scan_inputs("stress", batch, KD, L, N=N, G=G, seed=L, with_z=with_z, layout=layout)
    out, last = _run(inp, last=True)
    want, want_last = selective_scan_ref(**inp, delta_softplus=True, return_last_state=True,
                                         compute_dtype=torch.float64)
    sc = max(1.0, want.abs().max().item())
    assert_close(out.double().cpu() / sc, want / sc, 1e-4, 1e-5, "out")
    assert_close(last.double().cpu() / sc, want_last / sc, 1e-4, 1e-5, "last_state")


def test_scan_fwd_options():
    inp = make_scan_inputs("stress", 2, 16, 33, seed=7)
    # no softplus, no bias, no D, 3-d B/C (single group)
    from medmamba_b200 import selective_scan_fn
    g = _gpu(inp)
    dl = torch.nn.functional.softplus(g["delta"])
    out = selective_scan_fn(g["u"], dl, g["A"], g["B"][:, 0], g["C"][:, 0])
    want = selective_scan_ref(inp["u"], dl.cpu(), inp["A"], inp["B"][:, 0], inp["C"][:, 0],
                              compute_dtype=torch.float64)
    assert_close(out, want, 1e-4, 1e-5, "plain")
    # strided u / delta rows (views into a larger buffer)
    big = torch.randn(2, 16, 80, device="cuda")
    u = big[:, :, 3:36]
    out = selective_scan_fn(u, g["delta"], g["A"], g["B"], g["C"], g["D"], None, g["delta_bias"], True)
    want = selective_scan_ref(u.cpu(), inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None,
                              inp["delta_bias"], True, compute_dtype=torch.float64)
    assert_close(out, want, 1e-4, 1e-5, "strided")
    # softplus threshold: large raw delta passes through unchanged
    dbig = torch.full_like(g["delta"], 25.0)
    out = selective_scan_fn(g["u"], dbig, g["A"], g["B"], g["C"], g["D"], None, None, True)
    want = selective_scan_ref(inp["u"], dbig.cpu(), inp["A"], inp["B"], inp["C"], inp["D"], None, None, True,
                              compute_dtype=torch.float64)
    assert_close(out, want, 1e-4, 1e-5, "threshold")


def test_scan_fwd_empty_and_errors():
    from medmamba_b200 import selective_scan_fn
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = selective_scan_fn(z(0, 8, 5), z(0, 8, 5), -torch.ones(8, 16, device="cuda"), z(0, 4, 16, 5), z(0, 4, 16, 5))
    assert out.shape == (0, 8, 5)
    with pytest.raises(ValueError):
        selective_scan_fn(z(1, 8, 5), z(1, 8, 4), -torch.ones(8, 16, device="cuda"), z(1, 4, 16, 5), z(1, 4, 16, 5))
    with pytest.raises(ValueError):          # dstate 32 runs (two groups of 16), but B / C must carry 32 states too
        selective_scan_fn(z(1, 8, 5), z(1, 8, 5), -torch.ones(8, 32, device="cuda"), z(1, 4, 16, 5), z(1, 4, 32, 5))
    with pytest.raises(ValueError):          # the reference operator's own limit (dstate <= 256)
        selective_scan_fn(z(1, 8, 5), z(1, 8, 5), -torch.ones(8, 257, device="cuda"), z(1, 4, 257, 5), z(1, 4, 257, 5))
    assert selective_scan_fn(z(1, 8, 5), z(1, 8, 5), -torch.ones(8, 32, device="cuda"), z(1, 4, 32, 5), z(1, 4, 32, 5)).shape == (1, 8, 5)
    with pytest.raises(RuntimeError):
        selective_scan_fn(torch.zeros(1, 8, 5), torch.zeros(1, 8, 5), -torch.ones(8, 16), torch.zeros(1, 4, 16, 5),
                          torch.zeros(1, 4, 16, 5))


@pytest.mark.parametrize("batch", [4, 64])
@pytest.mark.parametrize("KD,L", STAGE_SHAPES)
def test_scan_fwd_bf16_io(KD, L, batch):
    """BASELINE config 2 in bf16 I/O (state fp32), at the config's own batch of 64 and at a small one."""
    inp = make_scan_inputs("stress", batch, KD, L, seed=L)
    q = lambda t: t.bfloat16().float()
    inp_q = dict(inp, u=q(inp["u"]), delta=q(inp["delta"]))
    out = _run(inp_q, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16
    want, _ = cscan.scan_fwd(**inp_q, delta_softplus=True, precision="f64")
    assert rel_err(out.float().cpu(), want) < 1e-2


@pytest.mark.parametrize("bc_same", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("L", [64, 200, 50, 37, 8, 1])
def test_scan_fwd_16bit_io_paths(L, dtype, tol, bc_same):
    """Every staging path of the 16-bit forward: 16-byte pieces (L % 8 == 0), 4-byte pieces (even L), the synchronous
    kernel (odd L); B / C in fp32 or in the I/O dtype; with the silu(z) gate and the last state."""
    from medmamba_b200 import selective_scan_fn
    inp = make_scan_inputs("stress", 3, 40, L, seed=L, with_z=True)
    q = lambda t: t.to(dtype).float()
    inp_q = dict(inp, u=q(inp["u"]), delta=q(inp["delta"]), z=q(inp["z"]))
    if bc_same:
        inp_q.update(B=q(inp["B"]), C=q(inp["C"]))
    g = _gpu(inp_q)
    bc = (lambda t: t.to(dtype)) if bc_same else (lambda t: t)
    out, last = selective_scan_fn(g["u"].to(dtype), g["delta"].to(dtype), g["A"], bc(g["B"]), bc(g["C"]), g["D"],
                                  g["z"].to(dtype), g["delta_bias"], True, True)
    assert out.dtype == dtype
    want, want_last = cscan.scan_fwd(**inp_q, delta_softplus=True, precision="f64")
    assert rel_err(out.float().cpu(), want) < tol
    assert rel_err(last.float().cpu(), want_last) < 1e-4


def test_scan_fwd_linearity_full_size():
    """Size-independent property at the BASELINE size: out is linear in u for fixed delta, B, C."""
    inp = _gpu(make_scan_inputs("stress", 64, 384, 3136, seed=1))
    from medmamba_b200 import selective_scan_fn
    f = lambda u: selective_scan_fn(u, inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None,
                                    inp["delta_bias"], True)
    u2 = torch.randn_like(inp["u"])
    lhs = f(inp["u"] + 2.0 * u2)
    rhs = f(inp["u"]) + 2.0 * f(u2)
    assert rel_err(lhs, rhs) < 1e-5


@pytest.mark.parametrize("with_z", [False, True])
@pytest.mark.parametrize("batch,KD,L,G,N", [(2, 24, 37, 4, 22), (1, 8, 130, 1, 17), (3, 12, 64, 2, 48), (1, 4, 9, 4, 256)])
def test_scan_wide_state_forward(batch, KD, L, G, N, with_z):
    """dstate > 16 (temp.py:27-36 allows up to 256): the interface runs groups of 16 states as separate launches and adds
    their outputs; out, out - u*D and last_state against the fp64 oracle at the fp32 bar."""
    inp = make_scan_inputs("stress", batch, KD, L, N=N, G=G, seed=N + L, with_z=with_z)
    out, last = _run(inp, last=True)
    want, want_last = selective_scan_ref(**inp, delta_softplus=True, return_last_state=True,
                                         compute_dtype=torch.float64)
    assert out.dtype == torch.float32 and last.shape == (batch, KD, N)
    sc = max(1.0, want.abs().max().item())
    assert_close(out.double().cpu() / sc, want / sc, 1e-4, 1e-5, "out")
    assert_close(last.double().cpu() / sc, want_last.double() / sc, 1e-4, 1e-5, "last_state")
    if not with_z:
        uD = (inp["u"] * inp["D"][None, :, None]).double()
        assert_close((out.double().cpu() - uD) / sc, (want - uD) / sc, 1e-4, 1e-5, "out - u*D")
    out16 = _run(inp, dtype=torch.bfloat16)
    assert out16.dtype == torch.bfloat16 and rel_err(out16.double().cpu(), want) < 1e-2


def test_scan_wide_state_backward():
    """Gradients of the 22-state scan (two launches, composed by autograd) against the analytic fp64 backward."""
    from medmamba_b200 import selective_scan_fn
    from oracle.selective_scan_ref import selective_scan_bwd_ref
    batch, KD, L, N = 2, 16, 50, 22
    inp = make_scan_inputs("stress", batch, KD, L, N=N, G=4, seed=5, with_z=True)
    g = {k: (v.cuda().requires_grad_(True) if v is not None else None) for k, v in inp.items()}
    out = selective_scan_fn(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["z"], g["delta_bias"], True)
    torch.manual_seed(1)
    dout = torch.randn(batch, KD, L)
    out.backward(dout.cuda())
    want = selective_scan_bwd_ref(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                  inp["delta_bias"], True, dout)
    for name in ["u", "delta", "A", "B", "C", "D", "z", "delta_bias"]:
        w = want["d" + name]
        got = g[name].grad.double().cpu()
        sc = max(1.0, w.abs().max().item())
        assert_close(got / sc, w.double().reshape(got.shape) / sc, 2e-3, 2e-4, "d" + name)

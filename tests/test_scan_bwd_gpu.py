"""Gradients of selective_scan_fn (CUDA, mmb_scan_bwd through the C ABI) against the analytic float64
oracle (oracle/scan_ref.c; itself checked against autograd in tests/test_oracle.py).
Bar: every gradient within 2e-4 of its own max-norm (fp32 accumulation over L and the batch)."""
import pytest
import torch

from oracle import cscan
from tests.util import STAGE_SHAPES, make_scan_inputs

pytestmark = pytest.mark.gpu
NAMES = ["u", "delta", "A", "B", "C", "D", "z", "delta_bias"]
GRADS = ["du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias"]


def _check(inp, softplus=True, tol=2e-4, seed=0):
    from medmamba_b200 import selective_scan_fn
    args = [inp[n].cuda().requires_grad_() if inp[n] is not None else None for n in NAMES]
    out = selective_scan_fn(*args[:6], args[6], args[7], softplus)
    dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed))
    out.backward(dout.cuda())
    torch.cuda.synchronize()
    want = cscan.scan_bwd(*[inp[n].contiguous() if inp[n] is not None else None for n in NAMES], softplus, dout)
    for a, key in zip(args, GRADS):
        if a is None:
            continue
        g, w = a.grad.double().cpu(), want[key]
        assert g.shape == w.shape, key
        err = (g - w).abs().max().item() / max(w.abs().max().item(), 1e-30)
        assert err < tol, f"{key}: rel-to-max error {err:.2e}"


@pytest.mark.parametrize("with_z", [False, True])
@pytest.mark.parametrize("layout", ["NL", "LN"])
@pytest.mark.parametrize("batch,KD,L,G,N", [
    (1, 4, 1, 4, 16), (2, 8, 5, 4, 16), (2, 24, 37, 4, 16), (2, 40, 130, 1, 16), (3, 36, 67, 2, 8), (1, 96, 257, 4, 16),
    (2, 12, 16, 4, 3), (5, 132, 48, 4, 16),
])
def test_scan_bwd_ragged(batch, KD, L, G, N, with_z, layout):
    _check(make_scan_inputs("stress", batch, KD, L, N=N, G=G, seed=L + KD, with_z=with_z, layout=layout))


@pytest.mark.parametrize("family", ["model", "stress"])
@pytest.mark.parametrize("KD,L", STAGE_SHAPES)
def test_scan_bwd_stage_shapes(family, KD, L):
    _check(make_scan_inputs(family, 2, KD, L, seed=KD), tol=5e-4)


def test_scan_bwd_options():
    inp = make_scan_inputs("stress", 2, 16, 33, seed=7)
    inp["delta"] = torch.nn.functional.softplus(inp["delta"])
    plain = dict(inp, D=None, delta_bias=None)
    _check(plain, softplus=False)                                         # no D, no bias, no softplus
    inp3 = dict(plain, B=inp["B"][:, 0].contiguous(), C=inp["C"][:, 0].contiguous(), D=inp["D"])
    _check(inp3, softplus=False)                                          # 3-d (ungrouped) B / C


def test_scan_bwd_deterministic():
    from medmamba_b200 import selective_scan_fn
    inp = make_scan_inputs("stress", 4, 96, 200, seed=3)
    outs = []
    for _ in range(2):
        args = [inp[n].cuda().requires_grad_() if inp[n] is not None else None for n in NAMES]
        out = selective_scan_fn(*args[:6], args[6], args[7], True)
        out.backward(torch.ones_like(out))
        outs.append([a.grad.clone() for a in args if a is not None])
    for a, b in zip(*outs):
        assert torch.equal(a, b)

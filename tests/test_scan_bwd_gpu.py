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


IO = {"u", "delta", "z"}        # tensors that travel in the I/O dtype (temp.py:27-36); parameters and B/C stay fp32


def _check(inp, softplus=True, tol=2e-4, seed=0, dtype=torch.float32):
    from medmamba_b200 import selective_scan_fn
    if dtype != torch.float32:      # the oracle sees exactly the values the kernel reads
        inp = {k: (v.to(dtype).float() if (k in IO and v is not None) else v) for k, v in inp.items()}
    args = [(inp[n].to(dtype) if n in IO else inp[n]).cuda().requires_grad_() if inp[n] is not None else None
            for n in NAMES]
    out = selective_scan_fn(*args[:6], args[6], args[7], softplus)
    assert out.dtype == dtype
    dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed)).to(dtype)
    out.backward(dout.cuda())
    dout = dout.float()
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
@pytest.mark.parametrize("batch", [2, 64])
def test_scan_bwd_stage_shapes(family, KD, L, batch):
    """BASELINE config 2 backward, fp32, at the config's own batch of 64 (and a quick batch-2 case)."""
    if batch == 64 and family == "model" and L < 3136:
        pytest.skip("covered by the stress family at this size")
    _check(make_scan_inputs(family, batch, KD, L, seed=KD), tol=5e-4)


@pytest.mark.parametrize("KD,L", STAGE_SHAPES)
@pytest.mark.parametrize("batch", [2, 64])
def test_scan_bwd_stage_shapes_bf16_io(KD, L, batch):
    """BASELINE config 2 backward with bf16 u / delta / dout (state, parameters and reductions fp32): every gradient
    within 1e-2 of its own max-norm (north_star's bf16 bar) of the fp64 oracle evaluated on the same bf16 inputs."""
    _check(make_scan_inputs("stress", batch, KD, L, seed=KD + 1), tol=1e-2, dtype=torch.bfloat16)


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

"""CPU suite: the oracle against the golden fixtures (outputs of the unmodified reference) and,
when /root/reference is present (build container), against the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import cscan, medmamba_ref, refload
from oracle.selective_scan_ref import selective_scan_bwd_ref, selective_scan_ref
from tests.util import assert_close, make_scan_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _cases(d):
    names = sorted({k.split(".")[0] for k in d})
    return {n: {k[len(n) + 1:]: v for k, v in d.items() if k.startswith(n + ".")} for n in names}


SCAN = _cases(_load("scan_small.npz"))


@pytest.mark.parametrize("name", sorted(SCAN))
def test_scan_restatement_matches_reference_outputs(name):
    c = SCAN[name]
    out, last = selective_scan_ref(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), c.get("z"),
                                   c.get("delta_bias"), bool(c["softplus"]), True)
    assert_close(out, c["out"], 1e-5, 1e-5, "python restatement")
    assert_close(last, c["last_state"], 1e-5, 1e-5, "python restatement last_state")
    out, last = cscan.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), c.get("z"),
                               c.get("delta_bias"), bool(c["softplus"]))
    assert_close(out, c["out"], 1e-5, 1e-5, "C restatement")
    assert_close(last, c["last_state"], 1e-5, 1e-5, "C restatement last_state")
    out64, _ = cscan.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), c.get("z"),
                              c.get("delta_bias"), bool(c["softplus"]), precision="f64")
    assert_close(out64, c["out"], 1e-4, 1e-4, "fp64 truth vs fp32 reference")


def test_index_maps_bit_exact():
    d = _cases(_load("index_maps.npz"))
    for name, c in d.items():
        H, W = map(int, name.split("x"))
        x = c["x"]
        B, D = x.shape[:2]
        xs = medmamba_ref.cross_scan(x)
        assert torch.equal(xs.reshape(B, 4 * D, H * W), c["xs"])
        src = torch.from_numpy(medmamba_ref.cross_scan_index(H, W))
        flat = x.reshape(B, D, H * W)
        for k in range(4):   # closed form of SURVEY.md Appendix A
            assert torch.equal(flat[:, :, src[k]], xs[:, k])
        ys = medmamba_ref.cross_merge(xs, H, W)
        for i, y in enumerate(ys):
            assert torch.equal(y.contiguous(), c[f"y{i + 1}"])
            assert torch.equal(y.contiguous(), flat)       # merge(scan(x)) puts every token back
        # the product's own host helpers (used on the reference-order path)
        from medmamba_b200 import ops
        assert torch.equal(ops.cross_scan(x), xs)
        for a, b_ in zip(ops.cross_merge(xs, H, W), ys):
            assert torch.equal(a, b_)


def test_ss2d_oracle_matches_reference_module():
    for name, c in _cases(_load("ss2d_small.npz")).items():
        sd = {k[3:]: v for k, v in c.items() if k.startswith("sd.")}
        y = medmamba_ref.ss2d_forward(sd, "", c["x"])
        assert_close(y, c["y"], 1e-5, 1e-6, f"ss2d {name}")
        ys = medmamba_ref.ss2d_core(c["conv_out"], sd["x_proj_weight"], sd["dt_projs_weight"],
                                    sd["dt_projs_bias"], sd["A_logs"], sd["Ds"])
        for i, yy in enumerate(ys):
            assert_close(yy, c[f"y{i + 1}"], 1e-5, 1e-6, f"ss2d {name} y{i + 1}")


def test_vssm_oracle_matches_reference_tiny():
    c = _load("vssm_tiny.npz")
    sd = {k[3:]: v for k, v in c.items() if k.startswith("sd.")}
    logits = medmamba_ref.vssm_forward(sd, c["x"], depths=tuple(int(v) for v in c["depths"]))
    assert_close(logits, c["logits"], 1e-4, 1e-5, "tiny VSSM logits")


def test_model_mirror_state_dict_and_seeded_init():
    """Keys / shapes of the product modules equal the golden reference state_dict, and seeded
    construction reproduces the reference's random-init weights (config 1 depends on it)."""
    import medmamba_b200 as mm
    c = _load("vssm_tiny.npz")
    sd = {k[3:]: v for k, v in c.items() if k.startswith("sd.")}
    torch.manual_seed(3)
    net = mm.VSSM(depths=[int(v) for v in c["depths"]], dims=[int(v) for v in c["dims"]], num_classes=5)
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:
        assert mine[k].shape == sd[k].shape, k
        if "running_" not in k:      # BN statistics were randomised after construction for the fixture
            assert torch.equal(mine[k], sd[k]), k
    g = _load("vssm_t_config1.npz")
    torch.manual_seed(int(g["weight_seed"]))
    t = mm.medmamba_t(num_classes=6)
    assert sum(p.numel() for p in t.parameters()) == 14457222
    assert torch.equal(t.head.weight, g["head_weight"])
    assert torch.equal(t.layers[0].blocks[0].self_attention.x_proj_weight, g["first_x_proj"])
    net.load_state_dict(sd)          # reference checkpoints load into the mirror


def test_vssm_oracle_config1_logits():
    import medmamba_b200 as mm
    g = _load("vssm_t_config1.npz")
    torch.manual_seed(int(g["weight_seed"]))
    sd = mm.medmamba_t(num_classes=6).state_dict()
    torch.manual_seed(int(g["input_seed"]))
    x = torch.randn(8, 3, 224, 224)
    with torch.no_grad():
        logits = medmamba_ref.vssm_forward(sd, x[:2], scan_fn=lambda *a, **k: cscan_scan(*a, **k))
    assert_close(logits, g["logits"][:2], 1e-4, 1e-5, "config-1 logits (first 2 images)")
    assert torch.equal(logits.argmax(1), g["logits"][:2].argmax(1))


def cscan_scan(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
    out, last = cscan.scan_fwd(u, delta, A, B.contiguous(), C.contiguous(), D, z, delta_bias, delta_softplus)
    return (out, last) if return_last_state else out


def test_backward_oracle_matches_autograd():
    inp = make_scan_inputs("stress", 2, 12, 23, seed=5, with_z=True)
    names = ["u", "delta", "A", "B", "C", "D", "z", "delta_bias"]
    args = [inp[n].double().requires_grad_() for n in names]
    out = selective_scan_ref(*args, delta_softplus=True, compute_dtype=torch.float64)
    dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    grads = torch.autograd.grad(out, args, dout)
    py = selective_scan_bwd_ref(*[inp[n] for n in names], True, dout)
    cc = cscan.scan_bwd(*[inp[n] for n in names], True, dout)
    for key, g in zip(["du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias"], grads):
        assert_close(py[key], g, 1e-9, 1e-9, "python analytic " + key)
        # the C entry takes fp32 dout
        assert_close(cc[key], g, 1e-5, 1e-5, "C analytic " + key)


@needs_ref
def test_restatement_matches_reference_text():
    ref = refload.reference_scan_ref()
    for seed, (wz, layout) in enumerate([(False, "NL"), (True, "LN")]):
        inp = make_scan_inputs("stress", 2, 24, 41, seed=seed, with_z=wz, layout=layout)
        o1, l1 = ref(**inp, delta_softplus=True, return_last_state=True)
        o2, l2 = selective_scan_ref(**inp, delta_softplus=True, return_last_state=True)
        assert_close(o2, o1, 1e-5, 1e-5, "restatement vs reference text")
        assert_close(l2, l1, 1e-5, 1e-5, "last_state")


@needs_ref
def test_seeded_init_equals_unmodified_reference():
    import medmamba_b200 as mm
    mod, _ = refload.load_reference()
    cfg = dict(depths=[1, 1, 2, 1], dims=[16, 32, 64, 128], num_classes=4)
    torch.manual_seed(11); a = mod.VSSM(**cfg).state_dict()
    torch.manual_seed(11); b = mm.VSSM(**cfg).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_flops_model_matches_survey_numbers():
    from medmamba_b200.flops import flops_selective_scan_ref
    assert flops_selective_scan_ref(B=64, L=3136, D=384, N=16) == pytest.approx(4.39e9, rel=0.01)


def test_reference_checkpoint_format_round_trip(tmp_path):
    """train.py:310-319 writes {'epoch', 'model_state_dict', 'optimizer_state_dict', 'best_acc', 'num_classes',
    'class_indices', 'scheduler_state_dict'}; train.py:208-249 resumes from it.  The mirror's parameter order and
    names are the reference's, so both the model and an AdamW / MultiStepLR state (indexed by parameter
    position) written by one load into the other."""
    import medmamba_b200 as mm
    torch.manual_seed(1)
    cfg = dict(depths=[1, 1], dims=[16, 32], num_classes=3)
    src = mm.VSSM(**cfg)
    opt = torch.optim.AdamW(src.parameters(), lr=1e-4, weight_decay=1e-4)            # train.py:190-192
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2, 4], gamma=0.1)
    path = tmp_path / "ckpt.pth"
    torch.save({"epoch": 3, "model_state_dict": src.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "best_acc": 0.5, "num_classes": 3, "class_indices": {"0": "a", "1": "b", "2": "c"},
                "scheduler_state_dict": sched.state_dict()}, path)
    ck = torch.load(path, map_location="cpu")
    torch.manual_seed(2)
    dst = mm.VSSM(**cfg)
    dst.load_state_dict(ck["model_state_dict"])                                       # train.py:214
    opt2 = torch.optim.AdamW(dst.parameters(), lr=1e-4, weight_decay=1e-4)
    opt2.load_state_dict(ck["optimizer_state_dict"])                                  # train.py:217
    sched2 = torch.optim.lr_scheduler.MultiStepLR(opt2, milestones=[2, 4], gamma=0.1)
    sched2.load_state_dict(ck["scheduler_state_dict"])
    assert ck["epoch"] + 1 == 4 and ck["best_acc"] == 0.5
    for (ka, a), (kb, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    assert [n for n, _ in src.named_parameters()] == [n for n, _ in dst.named_parameters()]

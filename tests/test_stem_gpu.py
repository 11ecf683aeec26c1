"""Parity of the fused stem kernels (patch embedding, patch merging) with the reference's op sequence
(MedMamba.py:54-76, :93-117) evaluated in fp64 by torch on the CPU."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Hin,Win,E", [(2, 224, 224, 96), (1, 32, 40, 32), (3, 8, 4, 128), (1, 512, 512, 96), (2, 12, 60, 64)])
@pytest.mark.parametrize("bias", [True, False])
def test_patch_embed_ln(B, Hin, Win, E, bias):
    from medmamba_b200 import ops
    g = torch.Generator().manual_seed(Hin + E)
    x = torch.randn(B, 3, Hin, Win, generator=g)
    w = torch.randn(E, 3, 4, 4, generator=g) * 0.2
    cb = torch.randn(E, generator=g) if bias else None
    gm, bt = torch.randn(E, generator=g), torch.randn(E, generator=g)
    want = F.conv2d(x.double(), w.double(), None if cb is None else cb.double(), stride=4).permute(0, 2, 3, 1)
    want = F.layer_norm(want, (E,), gm.double(), bt.double(), 1e-5)
    got = ops.patch_embed_ln(x.cuda(), w.cuda(), None if cb is None else cb.cuda(), gm.cuda(), bt.cuda(), 1e-5)
    assert got.shape == (B, Hin // 4, Win // 4, E) and got.dtype == torch.float32
    assert_close(got, want, 1e-4, 1e-5, "patch_embed_ln fp32")
    got16 = ops.patch_embed_ln(x.cuda().bfloat16(), w.cuda(), None if cb is None else cb.cuda(), gm.cuda(), bt.cuda(), 1e-5)
    want16 = F.conv2d(x.bfloat16().double(), w.double(), None if cb is None else cb.double(), stride=4).permute(0, 2, 3, 1)
    want16 = F.layer_norm(want16, (E,), gm.double(), bt.double(), 1e-5)
    assert_close(got16, want16, 1e-4, 1e-5, "patch_embed_ln bf16 input")


@pytest.mark.parametrize("B,Hin,Win,E", [(2, 224, 224, 96), (1, 512, 512, 96), (3, 40, 24, 32), (2, 36, 72, 64), (1, 4, 4, 128),
                                         (2, 20, 512, 96), (5, 12, 8, 96), (1, 28, 516, 96)])
@pytest.mark.parametrize("bias", [True, False])
def test_patch_embed_ln_tensor_core_path(B, Hin, Win, E, bias):
    """math_mode 1 (what bf16 autocast selects): operands rounded to bf16, fp32 accumulation on HMMA, bias + LayerNorm in
    fp32 -- against the fp64 convolution of the SAME bf16-rounded operands (the only difference left is the fp32
    accumulation order), strips that straddle token rows, ragged last strips, and the > 128 tokens-per-row fall-back."""
    from medmamba_b200 import ops
    g = torch.Generator().manual_seed(Hin + Win + E)
    x = torch.randn(B, 3, Hin, Win, generator=g)
    w = torch.randn(E, 3, 4, 4, generator=g) * 0.2
    cb = torch.randn(E, generator=g) if bias else None
    gm, bt = torch.randn(E, generator=g), torch.randn(E, generator=g)
    q = lambda t: t.bfloat16().double()
    fallback = Win // 4 > 128
    want = F.conv2d(x.double() if fallback else q(x), w.double() if fallback else q(w), None if cb is None else cb.double(),
                    stride=4).permute(0, 2, 3, 1)
    want = F.layer_norm(want, (E,), gm.double(), bt.double(), 1e-5)
    got = ops.patch_embed_ln(x.cuda(), w.cuda(), None if cb is None else cb.cuda(), gm.cuda(), bt.cuda(), 1e-5, bf16_math=True)
    assert got.shape == (B, Hin // 4, Win // 4, E) and got.dtype == torch.float32
    assert_close(got, want, 1e-4, 1e-5, "patch_embed_ln tensor-core path")
    # and it is what the module picks under bf16 autocast, within bf16 accuracy of the exact convolution
    exact = F.layer_norm(F.conv2d(x.double(), w.double(), None if cb is None else cb.double(), stride=4).permute(0, 2, 3, 1),
                         (E,), gm.double(), bt.double(), 1e-5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        auto = ops.patch_embed_ln(x.cuda(), w.cuda(), None if cb is None else cb.cuda(), gm.cuda(), bt.cuda(), 1e-5)
    assert torch.equal(auto, got)
    assert (auto.double().cpu() - exact).abs().max().item() < 3e-2 * max(1.0, exact.abs().max().item())


@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 96), (1, 7, 9, 8), (2, 28, 28, 192), (1, 14, 14, 384), (3, 5, 6, 20),
                                     (1, 4, 4, 512)])
def test_patch_merge_ln(B, H, W, C):
    from medmamba_b200 import ops
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(B, H, W, C, generator=g)
    gm, bt = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    h2, w2 = H // 2, W // 2
    quads = [x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]]      # MedMamba.py:100-103
    cat = torch.cat([q[:, :h2, :w2] for q in quads], -1).double()
    want = F.layer_norm(cat, (4 * C,), gm.double(), bt.double(), 1e-5)
    got = ops.patch_merge_ln(x.cuda(), gm.cuda(), bt.cuda(), 1e-5)
    assert got.shape == (B, h2, w2, 4 * C)
    assert_close(got, want, 1e-5, 1e-5, "patch_merge_ln fp32")
    got16 = ops.patch_merge_ln(x.cuda(), gm.cuda(), bt.cuda(), 1e-5, out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    assert_close(got16.float(), want, 1e-2, 1e-2, "patch_merge_ln bf16 out")


def test_model_stem_paths_match_module_paths():
    """PatchEmbed2D / PatchMerging2D: fused kernels against the torch module sequence of the same weights."""
    import medmamba_b200 as mm
    torch.manual_seed(3)
    pe = mm.PatchEmbed2D(4, 3, 96, torch.nn.LayerNorm).cuda().eval()
    pm = mm.PatchMerging2D(96).cuda().eval()
    x = torch.randn(2, 3, 64, 96, device="cuda")
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        a = pe(x)
        pe.fused = False
        b = pe(x)
        assert_close(a, b.double(), 1e-4, 1e-5, "PatchEmbed2D fused vs modules")
        c = pm(a)
        pm.fused = False
        d = pm(a)
        assert_close(c, d.double(), 1e-4, 1e-5, "PatchMerging2D fused vs modules")

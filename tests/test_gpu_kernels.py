"""Kernel-level parity on the B200 (pytest -m gpu): every libspecyolo kernel, called through the C-ABI
(ctypes wrappers in specyolo.ops), against the CPU oracle / a plain fp32 restatement of the same op on
identical inputs.  Tolerances are stated per test: bf16 operands + bf16 outputs bound conv outputs to
~2^-8 relative; fp32 kernels (decode, STFT) are compared at 1e-3 px / 1e-4; NMS is bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _rel_err(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _fmap(x_nchw_f32):
    """fp32 NCHW (cpu) -> NHWC bf16 feature map on the GPU via plain torch (test plumbing)."""
    return x_nchw_f32.to(DEV).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
def test_layout_roundtrip(lib):
    from specyolo import ops

    g = torch.Generator().manual_seed(0)
    x = torch.rand((2, 37, 9, 13), generator=g)
    y = ops.to_nhwc_bf16(x.to(DEV))
    assert y.shape == x.shape and y.stride(1) == 1
    back = ops.to_nchw_f32(y).cpu()
    assert torch.equal(back, _bf(x))
    u8 = (x * 255).to(torch.uint8)
    y8 = ops.to_nchw_f32(ops.to_nhwc_bf16(u8.to(DEV))).cpu()
    assert torch.allclose(y8, _bf(u8.float() / 255), atol=4e-3)


CONV_CASES = [
    # cin, cout, k, s, d, g, B, H, W, act, residual
    (64, 64, 1, 1, 1, 1, 2, 40, 40, True, False),
    (96, 128, 1, 1, 1, 1, 2, 20, 20, True, False),      # kc = 32
    (16, 32, 1, 1, 1, 1, 1, 24, 24, True, False),       # kc = 16
    (512, 512, 1, 1, 1, 1, 2, 20, 20, True, False),     # two N tiles
    (1024, 512, 1, 1, 1, 1, 1, 20, 20, True, False),    # long K
    (128, 2, 1, 1, 1, 1, 2, 20, 20, False, False),      # nc=2 head conv (n_pad 16)
    (128, 80, 1, 1, 1, 1, 1, 20, 20, False, False),
    (32, 16, 3, 1, 1, 1, 2, 32, 32, True, False),
    (16, 32, 3, 1, 1, 1, 2, 32, 32, True, True),        # bottleneck cv2 + shortcut
    (64, 64, 3, 1, 1, 1, 3, 20, 20, True, True),        # tile spans images
    (128, 128, 3, 1, 1, 1, 2, 40, 40, True, False),
    (32, 64, 3, 2, 1, 1, 2, 64, 64, True, False),       # stride 2
    (128, 128, 3, 2, 1, 1, 2, 80, 80, True, False),
    (256, 512, 3, 2, 1, 1, 2, 40, 40, True, False),
    (128, 128, 7, 2, 2, 8, 2, 32, 32, True, False),     # DDWConv.conv1 (layer 11)
    (256, 128, 3, 2, 2, 8, 2, 40, 40, True, False),     # DDWConv.conv1 (layer 13)
    (64, 64, 3, 1, 1, 1, 1, 7, 9, True, False),         # ragged spatial size
    # halo-tile kernel: resident weights, one halo box per tile, taps as shifted UMMA descriptors
    (128, 64, 3, 1, 1, 1, 2, 24, 40, True, False),      # two 64-channel weight chunks per tap
    (64, 32, 3, 1, 1, 1, 16, 64, 64, True, True),       # 512 tiles: every persistent CTA walks several
    (32, 32, 3, 1, 2, 1, 2, 20, 20, True, False),       # dilation 2, stride 1
    (32, 48, 5, 1, 1, 1, 1, 19, 21, True, False),       # 5x5, Cout not a multiple of 16... of 32
    (64, 64, 7, 2, 2, 4, 2, 36, 28, True, False),       # grouped, sub-lattice stride 2
    (128, 128, 3, 1, 1, 1, 2, 20, 20, True, True),      # weights too large to stay resident: per-tap kernel
    # thin tiles (<= 32 columns): paired-task epilogue over super-tiles of 4
    (16, 32, 3, 1, 1, 1, 3, 19, 21, True, True),        # ragged edges, tile count not a multiple of the super-tile
    (32, 16, 3, 1, 1, 1, 5, 17, 40, True, False),       # 16 columns: the two warps of a quadrant alternate sub-tiles
    (32, 32, 3, 1, 1, 1, 7, 40, 24, False, True),       # no activation + shortcut
]


@pytest.mark.parametrize("cin,cout,k,s,d,g,B,H,W,act,res", CONV_CASES)
def test_conv_igemm(lib, cin, cout, k, s, d, g, B, H, W, act, res):
    from specyolo import ops

    gen = torch.Generator().manual_seed(cin * 131 + cout * 7 + k)
    x = torch.randn((B, cin, H, W), generator=gen)
    w = torch.randn((cout, cin // g, k, k), generator=gen) * math.sqrt(2.0 / (cin // g * k * k))
    gamma = torch.rand(cout, generator=gen) + 0.5
    beta = torch.randn(cout, generator=gen) * 0.1
    mean = torch.randn(cout, generator=gen) * 0.1
    var = torch.rand(cout, generator=gen) + 0.5
    eps = 1e-3
    ke = d * (k - 1) + 1
    p = ke // 2
    pc = ops.fold_pack(w.to(DEV), None, [t.to(DEV) for t in (gamma, beta, mean, var)], eps, s, p, d, g, act)
    xf = _fmap(x)
    Ho, Wo = pc.out_hw(H, W)
    r = torch.randn((B, cout, Ho, Wo), generator=gen) if res else None
    rf = _fmap(r) if res else None
    y = ops.conv2d(xf, pc, residual=rf)
    torch.cuda.synchronize()
    # reference: fp32 conv on the bf16-rounded operands (what the tensor core multiplies)
    scale = gamma / torch.sqrt(var + eps)
    wf = _bf(w * scale.view(-1, 1, 1, 1))
    bf = beta - mean * scale
    ref = F.conv2d(_bf(x), wf, bf, s, p, d, g)
    if act:
        ref = F.silu(ref)
    if res:
        ref = ref + _bf(r)
    got = y.float().cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    tol = 1.5e-2 * ref.abs() + 1.5e-2
    assert bool((err <= tol).all()), f"max err {err.max().item()} rel {_rel_err(got, ref)}"
    assert _rel_err(got, ref) < 6e-3


def test_conv_concat_slices(lib):
    """Input read from / output written into channel windows of wider NHWC buffers."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(5)
    B, H, W = 2, 20, 20
    big_in = torch.randn((B, 192, H, W), generator=gen)
    w = torch.randn((64, 64, 3, 3), generator=gen) * 0.05
    pc = ops.fold_pack(w.to(DEV), torch.randn(64, generator=gen).to(DEV) * 0.1, None, 0.0, 1, 1, 1, 1, True)
    fin = _fmap(big_in)
    out_big = ops.new_act(B, 256, H, W, DEV)
    out_big.zero_()
    ops.conv2d(fin[:, 64:128], pc, out=out_big[:, 128:192])
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(_bf(big_in[:, 64:128]), _bf(w), pc.bias[:64].cpu(), 1, 1))
    got = out_big.float().cpu()
    assert torch.count_nonzero(got[:, :128]) == 0 and torch.count_nonzero(got[:, 192:]) == 0
    assert _rel_err(got[:, 128:192], ref) < 6e-3


def test_conv_fp32_out(lib):
    from specyolo import ops

    gen = torch.Generator().manual_seed(6)
    B, H, W, no_stride = 2, 10, 10, 68
    x = torch.randn((B, 64, H, W), generator=gen)
    w = torch.randn((64, 64, 1, 1), generator=gen) * 0.1
    b = torch.randn(64, generator=gen)
    pc = ops.fold_pack(w.to(DEV), b.to(DEV), None, 0.0, 1, 0, 1, 1, False)
    buf = torch.zeros((B, H * W, no_stride), device=DEV)
    view = buf.view(B, H, W, no_stride).permute(0, 3, 1, 2)
    ops.conv2d(_fmap(x), pc, out=view[:, :64], out_fp32=True)
    torch.cuda.synchronize()
    ref = F.conv2d(_bf(x), _bf(w), b)
    got = view[:, :64].cpu()
    assert torch.allclose(got, ref, atol=2e-3, rtol=2e-3)
    assert torch.count_nonzero(buf[..., 64:]) == 0


def test_stem_conv(lib):
    from specyolo import ops

    gen = torch.Generator().manual_seed(7)
    x = torch.rand((2, 3, 64, 96), generator=gen)
    w = torch.randn((32, 3, 3, 3), generator=gen) * 0.2
    bn = [torch.rand(32, generator=gen) + 0.5, torch.randn(32, generator=gen) * 0.1,
          torch.randn(32, generator=gen) * 0.1, torch.rand(32, generator=gen) + 0.5]
    pc = ops.fold_pack(w.to(DEV), None, [t.to(DEV) for t in bn], 1e-3, 2, 1, 1, 1, True)
    scale = bn[0] / torch.sqrt(bn[3] + 1e-3)
    ref = F.silu(F.conv2d(x, w * scale.view(-1, 1, 1, 1), bn[1] - bn[2] * scale, 2, 1))
    y = ops.stem_conv(x.to(DEV), pc).float().cpu()
    assert _rel_err(y, ref) < 5e-3
    u8 = (x * 255).round().to(torch.uint8)
    y8 = ops.stem_conv(u8.to(DEV), pc).float().cpu()
    ref8 = F.silu(F.conv2d(u8.float() / 255, w * scale.view(-1, 1, 1, 1), bn[1] - bn[2] * scale, 2, 1))
    assert _rel_err(y8, ref8) < 5e-3


def test_stem_blocked_pair(lib):
    """Stem with 2x2-blocked output + 3x3/s2 conv reading it as a 2x2 conv == the two plain convs."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(21)
    x = torch.rand((2, 3, 64, 96), generator=gen)
    w0 = torch.randn((32, 3, 3, 3), generator=gen) * 0.3
    w1 = torch.randn((64, 32, 3, 3), generator=gen) * 0.08
    b0 = torch.randn(32, generator=gen) * 0.1
    b1 = torch.randn(64, generator=gen) * 0.1
    pc0 = ops.fold_pack(w0.to(DEV), b0.to(DEV), None, 0.0, 2, 1, 1, 1, True)
    xb = ops.stem_conv(x.to(DEV), pc0, blocked_out=True)
    assert xb.shape == (2, 128, 16, 24)
    y0 = F.silu(F.conv2d(_bf(x), _bf(w0), b0, 2, 1))                      # [2, 32, 32, 48]
    got0 = ops.to_nchw_f32(xb).cpu().view(2, 2, 2, 32, 16, 24).permute(0, 3, 4, 1, 5, 2).reshape(2, 32, 32, 48)
    assert _rel_err(got0, y0) < 6e-3
    pc1 = ops.pack_from_blocked(w1.to(DEV), b1.to(DEV), None, 0.0, True)
    y = ops.conv2d(xb, pc1).float().cpu()
    ref = F.silu(F.conv2d(_bf(got0), _bf(w1), b1, 2, 1))
    assert y.shape == ref.shape == (2, 64, 16, 24)
    assert _rel_err(y, ref) < 6e-3


def test_conv_thin_concat_slices(lib):
    """Thin-tile path with input, output and residual all being channel windows of wider buffers."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(15)
    B, H, W = 3, 26, 22
    big_in = torch.randn((B, 96, H, W), generator=gen)
    big_res = torch.randn((B, 80, H, W), generator=gen)
    w = torch.randn((32, 32, 3, 3), generator=gen) * 0.08
    b = torch.randn(32, generator=gen) * 0.1
    pc = ops.fold_pack(w.to(DEV), b.to(DEV), None, 0.0, 1, 1, 1, 1, True)
    fin, fres = _fmap(big_in), _fmap(big_res)
    out_big = ops.new_act(B, 64, H, W, DEV)
    out_big.zero_()
    ops.conv2d(fin[:, 32:64], pc, out=out_big[:, 16:48], residual=fres[:, 48:80])
    ref = F.silu(F.conv2d(_bf(big_in[:, 32:64]), _bf(w), b, 1, 1)) + _bf(big_res[:, 48:80])
    got = out_big.float().cpu()
    assert float(got[:, :16].abs().max()) == 0.0 and float(got[:, 48:].abs().max()) == 0.0
    assert _rel_err(got[:, 16:48], ref) < 6e-3


@pytest.mark.parametrize("shape", [(2, 64, 96, 32, 64), (3, 128, 160, 32, 64), (1, 36, 48, 32, 64), (2, 64, 96, 16, 64),
                                   (1, 96, 64, 16, 96)])
def test_stem_pair_fused(lib, shape):
    """Layers 0 + 1 in one kernel (uint8 input) == the two plain convs (torch fp32 reference on bf16-rounded operands)
    and == the layer-by-layer route of this library; sizes with partial tiles in both directions."""
    from specyolo import ops

    B, H, W, c0, c1 = shape
    gen = torch.Generator().manual_seed(33)
    u8 = (torch.rand((B, 3, H, W), generator=gen) * 255).round().to(torch.uint8)
    w0 = torch.randn((c0, 3, 3, 3), generator=gen) * 0.3
    w1 = torch.randn((c1, c0, 3, 3), generator=gen) * (0.08 * math.sqrt(32 / c0))
    b0 = torch.randn(c0, generator=gen) * 0.1
    b1 = torch.randn(c1, generator=gen) * 0.1
    pc0 = ops.fold_pack(w0.to(DEV), b0.to(DEV), None, 0.0, 2, 1, 1, 1, True)
    pc1 = ops.pack_from_blocked(w1.to(DEV), b1.to(DEV), None, 0.0, True)
    assert ops.stem_pair_ok(u8.to(DEV), pc0, pc1)
    y = ops.stem_pair(u8.to(DEV), pc0, pc1)
    assert y.shape == (B, c1, H // 4, W // 4)
    y0 = _bf(F.silu(F.conv2d(u8.float(), _bf(w0 / 255.0), b0, 2, 1)))
    ref = F.silu(F.conv2d(y0, _bf(w1), b1, 2, 1))
    assert _rel_err(y.float().cpu(), ref) < 6e-3
    two = ops.conv2d(ops.stem_conv(u8.to(DEV), pc0, blocked_out=True), pc1)
    assert _rel_err(y.float().cpu(), two.float().cpu()) < 6e-3
    # the output may be a channel slice of a wider buffer
    buf = ops.new_act(B, c1 + 32, H // 4, W // 4, DEV).zero_()
    ops.stem_pair(u8.to(DEV), pc0, pc1, out=buf[:, 16:16 + c1])
    assert torch.equal(buf[:, 16:16 + c1].float().cpu(), y.float().cpu())
    assert float(buf[:, :16].float().abs().max()) == 0.0 and float(buf[:, 16 + c1:].float().abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 32, 16, 56, 56, True), (3, 64, 32, 40, 40, True), (1, 64, 32, 29, 47, True),
                                   (2, 32, 32, 20, 20, True), (1, 32, 16, 160, 160, True), (2, 64, 32, 33, 15, False),
                                   (5, 64, 16, 14, 14, True), (1, 32, 16, 13, 31, False)])
def test_bottleneck_fused(lib, shape):
    """Bottleneck (3x3 -> 3x3 + shortcut, block.py:713-726) in one kernel == the two plain convs: torch fp32 reference on
    bf16-rounded operands with the intermediate rounded to bf16 (what the two-launch route stores), and == the
    two-launch route of this library.  Sizes with partial 14 x 14 tiles in both directions, tiles that are entirely
    border, and the output written into a channel slice of a wider buffer (as C3k2 does)."""
    from specyolo import ops

    B, C, Cm, H, W, add = shape
    gen = torch.Generator().manual_seed(57 + C + Cm + H)
    x = torch.randn((B, C, H, W), generator=gen)
    w1 = torch.randn((Cm, C, 3, 3), generator=gen) * math.sqrt(2.0 / (9 * C))
    w2 = torch.randn((C, Cm, 3, 3), generator=gen) * math.sqrt(2.0 / (9 * Cm))
    b1 = torch.randn(Cm, generator=gen) * 0.2
    b2 = torch.randn(C, generator=gen) * 0.2
    pc1 = ops.fold_pack(w1.to(DEV), b1.to(DEV), None, 0.0, 1, 1, 1, 1, True)
    pc2 = ops.fold_pack(w2.to(DEV), b2.to(DEV), None, 0.0, 1, 1, 1, 1, True)
    xf = _fmap(x)
    assert ops.bottleneck_ok(xf, pc1, pc2)
    y = ops.bottleneck(xf, pc1, pc2, add)
    assert y.shape == (B, C, H, W)
    mid = _bf(F.silu(F.conv2d(_bf(x), _bf(w1), b1, 1, 1)))
    ref = F.silu(F.conv2d(mid, _bf(w2), b2, 1, 1)) + (_bf(x) if add else 0)
    got = y.float().cpu()
    assert _rel_err(got, ref) < 6e-3, _rel_err(got, ref)
    assert float((got - ref).abs().max()) < 0.06 * float(ref.abs().max())
    two = ops.conv2d(ops.conv2d(xf, pc1), pc2, residual=xf if add else None)
    assert _rel_err(got, two.float().cpu()) < 6e-3
    # x read from, and y written into, channel slices of wider buffers
    src = ops.new_act(B, C + 32, H, W, DEV).zero_()
    src[:, 16:16 + C].copy_(xf)
    buf = ops.new_act(B, C + 48, H, W, DEV).zero_()
    ops.bottleneck(src[:, 16:16 + C], pc1, pc2, add, out=buf[:, 32:32 + C])
    assert torch.equal(buf[:, 32:32 + C].float().cpu(), got)
    assert float(buf[:, :32].float().abs().max()) == 0.0 and float(buf[:, 32 + C:].float().abs().max()) == 0.0


def test_depthwise(lib):
    from specyolo import ops

    gen = torch.Generator().manual_seed(8)
    x = torch.randn((2, 128, 20, 20), generator=gen)
    w = torch.randn((128, 1, 3, 3), generator=gen) * 0.3
    bn = [torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.1,
          torch.randn(128, generator=gen) * 0.1, torch.rand(128, generator=gen) + 0.5]
    for act in (True, False):
        pc = ops.fold_pack(w.to(DEV), None, [t.to(DEV) for t in bn], 1e-3, 1, 1, 1, 128, act)
        scale = bn[0] / torch.sqrt(bn[3] + 1e-3)
        ref = F.conv2d(_bf(x), _bf(w * scale.view(-1, 1, 1, 1)), bn[1] - bn[2] * scale, 1, 1, 1, 128)
        ref = F.silu(ref) if act else ref
        y = ops.conv2d(_fmap(x), pc).float().cpu()
        assert _rel_err(y, ref) < 5e-3


def test_depthwise_ragged(lib):
    """C not a multiple of the warp's 128-channel span, W not a multiple of the run length."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(18)
    x = torch.randn((3, 80, 11, 13), generator=gen)
    w = torch.randn((80, 1, 3, 3), generator=gen) * 0.3
    b = torch.randn(80, generator=gen) * 0.1
    pc = ops.fold_pack(w.to(DEV), b.to(DEV), None, 0.0, 1, 1, 1, 80, True)
    ref = F.silu(F.conv2d(_bf(x), _bf(w), b, 1, 1, 1, 80))
    y = ops.conv2d(_fmap(x), pc).float().cpu()
    assert _rel_err(y, ref) < 5e-3


@pytest.mark.parametrize("C,cout,B,H,W", [(128, 128, 2, 40, 40), (64, 128, 3, 20, 20), (128, 80, 1, 21, 13)])
def test_dwconv_pwconv_fused(lib, C, cout, B, H, W):
    """Fused DWConv 3x3 + SiLU -> Conv 1x1 + SiLU (Detect.cv3 pair) vs the two fp32 convs (bf16 intermediate)."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(31 + C + cout)
    x = torch.randn((B, C, H, W), generator=gen)
    wd = torch.randn((C, 1, 3, 3), generator=gen) * 0.3
    bd = torch.randn(C, generator=gen) * 0.1
    wp = torch.randn((cout, C, 1, 1), generator=gen) * math.sqrt(2.0 / C)
    bp = torch.randn(cout, generator=gen) * 0.1
    pw = ops.fold_pack(wp.to(DEV), bp.to(DEV), None, 0.0, 1, 0, 1, 1, True)
    assert ops.dwconv_pwconv_ok(C, pw)
    dw_w = wd.view(C, 9).t().contiguous().to(DEV)
    y = ops.dwconv_pwconv(_fmap(x), dw_w, bd.to(DEV), pw).float().cpu()
    mid = _bf(F.silu(F.conv2d(_bf(x), wd, bd, 1, 1, 1, C)))
    ref = F.silu(F.conv2d(mid, _bf(wp), bp))
    assert y.shape == ref.shape
    assert _rel_err(y, ref) < 8e-3, _rel_err(y, ref)
    # fused class head: the closing nn.Conv2d(cout, nc, 1) of a Detect.cv3 branch (head.py:56) in the epilogue, fp32
    # logits written into channels [64, 64 + nc) of the [B, H*W, 68] decode buffer; nothing else of that buffer is touched
    for nc in (1, 2, 4):
        wh = torch.randn((nc, cout), generator=gen) * math.sqrt(1.0 / cout)
        bh = torch.randn(nc, generator=gen) * 0.5
        buf = torch.full((B, H * W, 68), 7.0, device=DEV)
        view = buf.view(B, H, W, 68).permute(0, 3, 1, 2)[:, 64:64 + nc]
        r = ops.dwconv_pwconv(_fmap(x), dw_w, bd.to(DEV), pw, head=(wh.to(DEV), bh.to(DEV), view))
        assert r is None
        ref_h = F.conv2d(ref, wh.view(nc, cout, 1, 1), bh)                    # on the fp32 activations
        got = buf.view(B, H, W, 68)[..., 64:64 + nc].permute(0, 3, 1, 2).cpu()
        assert float((got - ref_h).abs().max()) < 2e-2 * max(1.0, float(ref_h.abs().max())), float((got - ref_h).abs().max())
        rest = buf.clone()
        rest.view(B, H, W, 68)[..., 64:64 + nc] = 7.0
        assert bool((rest == 7.0).all())


@pytest.mark.parametrize("c,B,H,W", [(32, 2, 20, 20), (256, 3, 20, 20), (128, 2, 40, 40), (64, 2, 13, 27), (256, 1, 5, 3),
                                     (128, 1, 33, 17)])
def test_sppf_pool(lib, c, B, H, W):
    """three chained MaxPool2d(5, 1, 2) (block.py:194-198): both CTA shapes (64 / 16 channels), ragged and tiny maps,
    maps smaller than the window, the 40 x 40 map of a 1280^2 input."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(9)
    y0 = torch.randn((B, c, H, W), generator=gen)
    buf = ops.new_act(B, 4 * c, H, W, DEV)
    buf.zero_()
    buf[:, :c].copy_(y0.to(DEV))
    ops.sppf_pool(buf, c)
    y = [_bf(y0)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    assert torch.equal(buf.float().cpu(), torch.cat(y, 1))     # max of bf16 values is exact


@pytest.mark.parametrize("k,H,W,up,c", [(3, 16, 16, (1, 0, 0), 128), (2, 10, 10, (0, 0), 128), (3, 40, 40, (0, 0, 0), 128),
                                        (2, 12, 20, (1, 0), 64), (3, 8, 8, (0, 0, 0), 32), (2, 20, 20, (0, 1), 256),
                                        (3, 80, 80, (1, 0, 0), 128)])
def test_fusion_eschannel(lib, k, H, W, up, c):
    from oracle.yolo_ref import Ref
    from specyolo import ops

    gen = torch.Generator().manual_seed(10 + k)
    B = 2           # every channel width the kernels are instantiated for (c / 32 lanes per pixel), ragged and full-size maps
    xs = [torch.randn((B, c, H >> u, W >> u), generator=gen) for u in up]
    sd = {"f.gsc%d.alpha" % k: torch.rand((1, k * c, 1, 1), generator=gen) * 0.5 + 0.75,
          "f.gsc%d.gamma" % k: torch.randn((1, k * c, 1, 1), generator=gen) * 0.3,
          "f.gsc%d.beta" % k: torch.randn((1, k * c, 1, 1), generator=gen) * 0.3,
          "f.sab.cv1.weight": torch.randn((1, 2, 3, 3), generator=gen) * 0.5}
    xs_full = [F.interpolate(_bf(x), scale_factor=2, mode="nearest") if u else _bf(x) for x, u in zip(xs, up)]
    ref = Ref(sd).fusion(xs_full, "f")
    got = ops.fusion_eschannel([_fmap(x) for x in xs], list(up), sd["f.gsc%d.alpha" % k].to(DEV),
                               sd["f.gsc%d.gamma" % k].to(DEV), sd["f.gsc%d.beta" % k].to(DEV), 1e-5,
                               sd["f.sab.cv1.weight"].to(DEV)).float().cpu()
    assert _rel_err(got, ref) < 5e-3


@pytest.mark.parametrize("B,c,H,W,window", [(2, 128, 80, 80, False), (3, 256, 40, 40, False), (2, 512, 20, 20, False), (1, 64, 7, 9, False),
                                            (2, 192, 5, 33, True), (1, 8, 3, 3, False), (5, 128, 1, 1, True)])
def test_sobel_spatial_attention(lib, B, c, H, W, window):
    """SobelSpatialAttention (conv.py:1184-1198) vs the oracle's restatement on bf16-rounded inputs: every lane-group
    width (c / 8 = 1 ... 64 vectors per pixel, incl. one that is not a power of two), ragged / single-pixel maps,
    in place on a channel window of a wider buffer and out of place."""
    from specyolo import ops
    from specyolo.nn.modules import SobelSpatialAttention

    gen = torch.Generator().manual_seed(40 + c)
    x = torch.randn((B, c, H, W), generator=gen)
    m = SobelSpatialAttention(7)
    with torch.no_grad():
        for cv in m.sobel.convs:
            cv.weight.add_(torch.randn(cv.weight.shape, generator=gen) * 0.3)
        m.cv1.weight.copy_(torch.randn(m.cv1.weight.shape, generator=gen) * 0.7)
    xb = _bf(x)
    mm = torch.cat([xb.mean(1, keepdim=True), xb.max(1, keepdim=True)[0]], 1)
    e = sum(F.conv2d(mm, cv.weight.detach(), None, 1, 1, 1, 2) for cv in m.sobel.convs)
    ref = xb * torch.sigmoid(F.conv2d(e, m.cv1.weight.detach()))
    if window:
        buf = ops.new_act(B, c + 16, H, W, DEV)
        buf.zero_()
        xin = buf[:, 8:8 + c]
        xin.copy_(x.to(DEV))
        got = m(xin)                                           # in place on the window
        assert got.data_ptr() == xin.data_ptr()
        assert float(buf[:, :8].float().abs().max()) == 0.0 and float(buf[:, 8 + c:].float().abs().max()) == 0.0
    else:
        xin = _fmap(x)
        keep = xin.clone()
        out = ops.new_act(B, c, H, W, DEV)
        got = ops.sobel_spatial_attention(xin, m.stencil(), out)
        assert torch.equal(xin, keep)                          # out of place leaves x alone
    got = got.float().cpu()
    assert (got - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())      # one bf16 rounding of the product
    assert _rel_err(got, ref) < 3e-3


@pytest.mark.parametrize("B,c,H,W,window", [(2, 64, 80, 80, False), (3, 128, 40, 40, False), (2, 256, 20, 20, True), (1, 64, 7, 9, False),
                                            (2, 200, 33, 70, False), (1, 8, 3, 3, True), (4, 512, 17, 5, False), (3, 64, 1, 1, False)])
def test_msc_spatial_attention(lib, B, c, H, W, window):
    """MSCSpatialAttention (conv.py:1200-1243) vs the oracle's restatement on bf16-rounded inputs: the 80 x 80 x 64 shape of
    the *_OMN config, every lane-group width (one, two vectors per lane; a channel count that is not a power of two), maps
    smaller than the 31 x 31 kernel, ragged tiles, in place on a channel window of a wider buffer and out of place."""
    from oracle.yolo_ref import Ref
    from specyolo import ops
    from specyolo.nn.modules import MSCSpatialAttention

    gen = torch.Generator().manual_seed(70 + c)
    x = torch.randn((B, c, H, W), generator=gen) * 0.8 + 0.1
    m = MSCSpatialAttention(c)
    with torch.no_grad():
        m.cv1[0].weight.copy_(torch.randn(m.cv1[0].weight.shape, generator=gen) * 0.05)
        m.cv2[0].weight.copy_(torch.randn(m.cv2[0].weight.shape, generator=gen) * 0.4)
        m.fc.weight.copy_(torch.randn(m.fc.weight.shape, generator=gen) * (2.0 / c) ** 0.5)
        m.fc.bias.copy_(torch.randn(m.fc.bias.shape, generator=gen) * 0.2 + 0.3)
    ref = Ref({"m." + k: v for k, v in m.state_dict().items()}).msc(_bf(x), "m")
    m.to(DEV)
    if window:
        buf = ops.new_act(B, c + 16, H, W, DEV)
        buf.zero_()
        xin = buf[:, 8:8 + c]
        xin.copy_(x.to(DEV))
        got = m(xin)                                           # in place on the window
        assert got.data_ptr() == xin.data_ptr()
        assert float(buf[:, :8].float().abs().max()) == 0.0 and float(buf[:, 8 + c:].float().abs().max()) == 0.0
    else:
        xin = _fmap(x)
        keep = xin.clone()
        out = ops.new_act(B, c, H, W, DEV)
        got = m(xin, out=out)
        assert torch.equal(xin, keep)                          # out of place leaves x alone
        again = m(xin, out=ops.new_act(B, c, H, W, DEV))
        assert torch.equal(got, again)                         # fixed-order reductions: bit-identical run to run
    got = got.float().cpu()
    assert (got - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())      # one bf16 rounding of the result
    assert _rel_err(got, ref) < 3e-3


@pytest.mark.parametrize("B,c,H,W,window", [(2, 32, 160, 160, False), (3, 32, 24, 32, True), (2, 16, 40, 56, False), (1, 32, 20, 12, False),
                                            (2, 32, 1, 1, False), (1, 16, 7, 15, True), (2, 32, 96, 128, False), (1, 16, 64, 192, False),
                                            (1, 16, 32, 224, False), (1, 32, 32, 256, True), (1, 32, 104, 88, False), (1, 16, 152, 136, False)])
def test_bottlenect_fgm(lib, B, c, H, W, window):
    """BottleNect + FGM (block.py:782-861) vs the oracle's restatement (torch.fft) on bf16-rounded inputs: the 160 x 160 x 32
    planes of the *_GC config at 640^2, every radix of the mixed-radix transform (4, 2, 3, 5, 7), an odd number of stages, every
    width of the register path (sides of 32 ... 256), non-square and single-pixel planes, both channel counts, in place on a channel window and out of place."""
    from oracle.yolo_ref import Ref
    from specyolo import ops
    from specyolo.nn.modules import BottleNect

    gen = torch.Generator().manual_seed(90 + c + H)
    x = torch.randn((B, c, H, W), generator=gen) * 0.8 + 0.2
    m = BottleNect(c)
    with torch.no_grad():
        for name, prm in m.named_parameters():
            if name.endswith("alpha"):
                prm.copy_(torch.rand(prm.shape, generator=gen) * 0.5 + 0.75)
            elif name.endswith("beta"):
                prm.copy_(torch.randn(prm.shape, generator=gen) * 0.5)
            elif name.endswith("bias"):
                prm.copy_(torch.randn(prm.shape, generator=gen) * 0.3)
            else:
                prm.copy_(torch.randn(prm.shape, generator=gen) * (1.5 / prm[0].numel()) ** 0.5)
    ref = Ref({"m." + k: v for k, v in m.state_dict().items()}).bottlenect(_bf(x), "m")
    m.to(DEV)
    if window:
        buf = ops.new_act(B, c + 16, H, W, DEV)
        buf.zero_()
        xin = buf[:, 8:8 + c]
        xin.copy_(x.to(DEV))
        got = m(xin)                                           # in place on the window
        assert got.data_ptr() == xin.data_ptr()
        assert float(buf[:, :8].float().abs().max()) == 0.0 and float(buf[:, 8 + c:].float().abs().max()) == 0.0
    else:
        xin = _fmap(x)
        keep = xin.clone()
        out = ops.new_act(B, c, H, W, DEV)
        got = m(xin, out=out)
        assert torch.equal(xin, keep)                          # out of place leaves x alone
        again = m(xin, out=ops.new_act(B, c, H, W, DEV))
        assert torch.equal(got, again)                         # fixed-order reductions: bit-identical run to run
    got = got.float().cpu()
    assert (got - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())      # one bf16 rounding of the result
    assert _rel_err(got, ref) < 3e-3


@pytest.mark.parametrize("H,W,heads", [(20, 20, 4), (8, 12, 2), (16, 16, 2), (5, 7, 1), (16, 32, 2), (24, 24, 2), (40, 40, 1)])
def test_psa_attention(lib, H, W, heads):
    """N = 400 (640^2 input), ragged / tiny / exactly 256 and 512 tokens on the resident-S kernel (N <= 512), 576 and
    1600 tokens on the two-sweep kernel."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(11)
    B, kd, hd = 2, 32, 64
    C = heads * hd
    qkv = torch.randn((B, heads * (2 * kd + hd), H, W), generator=gen)
    pe_w = torch.randn((C, 9), generator=gen) * 0.2
    pe_b = torch.randn((C,), generator=gen) * 0.1
    scale = kd ** -0.5
    got = ops.psa_attention(_fmap(qkv), heads, kd, hd, scale, pe_w.to(DEV), pe_b.to(DEV)).float().cpu()
    # block.py:1922-1933
    N = H * W
    q, k, v = _bf(qkv).view(B, heads, 2 * kd + hd, N).split([kd, kd, hd], dim=2)
    attn = ((q.transpose(-2, -1) @ k) * scale).softmax(dim=-1)
    ref = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + F.conv2d(v.reshape(B, C, H, W), pe_w.view(C, 1, 3, 3), pe_b, 1, 1, 1, C)
    assert _rel_err(got, ref) < 6e-3


def _rand_head_logits(gen, B, hw, nc, no_stride):
    bufs = []
    for (h, w) in hw:
        t = torch.zeros((B, h * w, no_stride))
        t[..., :64] = torch.randn((B, h * w, 64), generator=gen) * 2.0
        t[..., 64:64 + nc] = torch.randn((B, h * w, nc), generator=gen) * 1.5 - 2.0
        bufs.append(t)
    return bufs


@pytest.mark.parametrize("nc,hw", [(2, [(80, 80), (40, 40), (20, 20)]), (80, [(16, 12), (8, 6), (4, 3)])])
def test_detect_decode(lib, nc, hw):
    from oracle import yolo_ref
    from specyolo import ops

    gen = torch.Generator().manual_seed(12)
    B = 2
    no = 64 + nc
    no_stride = (no + 3) // 4 * 4
    bufs = _rand_head_logits(gen, B, hw, nc, no_stride)
    strides = [8.0, 16.0, 32.0]
    raw = [b[..., :no].view(B, h, w, no).permute(0, 3, 1, 2).contiguous() for b, (h, w) in zip(bufs, hw)]
    ref = yolo_ref.detect_decode(raw, strides, nc)
    conf = 0.25
    y, cand, seg = ops.detect_decode([b.to(DEV) for b in bufs], hw, strides, nc, want_dense=True, conf_thres=conf)
    y = y.cpu()
    # fp32 kernel vs fp32 oracle on identical logits: 1e-3 px on boxes, 1e-6 on scores
    assert torch.allclose(y[:, :4], ref[:, :4], atol=1e-3, rtol=0)
    assert torch.allclose(y[:, 4:], ref[:, 4:], atol=1e-6, rtol=0)
    # fused candidates == thresholding the kernel's own dense output, in anchor order
    A = y.shape[2]
    cand, seg = cand.cpu(), seg.cpu()
    for b in range(B):
        sc, cl = y[b, 4:].max(0)
        idx = torch.nonzero(sc > conf).flatten()
        rows = torch.cat([cand[b, s, : seg[b, s]] for s in range(seg.shape[1])])
        assert rows.shape[0] == idx.numel()
        xy, wh = y[b, :2, idx].T, y[b, 2:4, idx].T / 2
        exp = torch.cat((xy - wh, xy + wh, sc[idx, None], cl[idx, None].float()), 1)
        assert torch.equal(rows, exp)


@pytest.mark.parametrize("nc", [2, 16, 17, 21, 80])
def test_detect_decode_scores_first(lib, nc):
    """The candidate path without the dense output (what `predict` runs): class maxima first — gathered cooperatively by
    eight lanes per anchor for nc >= 16 —, boxes only for survivors.  Candidates must equal, bit for bit and in anchor order,
    those of the dense path; row padding beyond nc (garbage here) must never be read as a class."""
    from specyolo import ops

    gen = torch.Generator().manual_seed(50 + nc)
    B, hw, strides, conf = 3, [(20, 28), (10, 14), (5, 7)], [8.0, 16.0, 32.0], 0.25
    no_stride = (64 + nc + 3) // 4 * 4 + 4                      # at least one whole padding vector per row
    bufs = _rand_head_logits(gen, B, hw, nc, no_stride)
    for t in bufs:
        t[..., 64 + nc:] = 50.0
    dev = [t.to(DEV) for t in bufs]
    _, cand_d, seg_d = ops.detect_decode(dev, hw, strides, nc, want_dense=True, conf_thres=conf)
    _, cand_s, seg_s = ops.detect_decode(dev, hw, strides, nc, want_dense=False, conf_thres=conf)
    assert torch.equal(seg_d, seg_s) and int(seg_s.sum()) > 20
    seg = seg_s.cpu()
    for b in range(B):
        for q in range(seg.shape[1]):
            n = int(seg[b, q])
            assert torch.equal(cand_d[b, q, :n], cand_s[b, q, :n])


def _nms_inputs(gen, B, nc, A, dup=True):
    xy = torch.rand((B, 2, A), generator=gen) * 600 + 20
    wh = torch.rand((B, 2, A), generator=gen) * 192 + 8
    scores = torch.distributions.Beta(0.5, 8.0).sample((B, nc, A))
    pred = torch.cat((xy, wh, scores), 1).float()
    if dup:   # exact score ties and duplicated boxes
        pred[:, :, 1::7] = pred[:, :, 0:-1:7][:, :, : pred[:, :, 1::7].shape[2]]
    return pred


@pytest.mark.parametrize("nc,A,conf,iou,agn,ml", [
    (2, 8400, 0.25, 0.7, False, False),
    (2, 8400, 0.05, 0.45, False, False),
    (80, 2100, 0.1, 0.7, False, False),
    (80, 2100, 0.1, 0.5, True, False),
    (3, 3000, 0.02, 0.6, False, True),       # multi_label (validation path)
    (2, 8400, 0.001, 0.7, False, False),     # thousands of candidates -> global-memory sort path
])
def test_nms_bit_exact(lib, nc, A, conf, iou, agn, ml):
    from oracle import nms_ref
    from specyolo.utils.ops import non_max_suppression

    gen = torch.Generator().manual_seed(13 + nc + A)
    torch.manual_seed(13 + nc + A)
    B = 3
    pred = _nms_inputs(gen, B, nc, A)
    ref, ref_idx = nms_ref.non_max_suppression(pred.numpy(), conf, iou, agnostic=agn, multi_label=ml, return_indices=True)
    got, got_idx = non_max_suppression(pred.to(DEV), conf, iou, agnostic=agn, multi_label=ml, return_idxs=True)
    for b in range(B):
        g = got[b].cpu().numpy()
        assert g.shape == ref[b].shape, (g.shape, ref[b].shape)
        assert np.array_equal(got_idx[b].cpu().numpy(), ref_idx[b])        # keep indices: exact
        assert np.array_equal(g, ref[b])                                   # boxes, conf, cls: bit-exact


def test_nms_vs_torchvision(lib):
    """Same inputs through the reference's own third-party kernel (torchvision CPU nms)."""
    torchvision = pytest.importorskip("torchvision")
    from specyolo.utils.ops import non_max_suppression

    gen = torch.Generator().manual_seed(99)
    torch.manual_seed(99)
    pred = _nms_inputs(gen, 2, 1, 4000)
    conf, iou = 0.05, 0.5
    got, idx = non_max_suppression(pred.to(DEV), conf, iou, return_idxs=True)
    for b in range(2):
        x = pred[b].T
        m = x[:, 4] > conf
        x = x[m]
        boxes = torch.cat((x[:, :2] - x[:, 2:4] / 2, x[:, :2] + x[:, 2:4] / 2), 1)
        keep = torchvision.ops.nms(boxes, x[:, 4], iou)[:300]
        assert torch.equal(idx[b].cpu(), keep)


def test_nms_edge_cases(lib):
    from specyolo.utils.ops import non_max_suppression

    pred = torch.zeros((2, 6, 100))
    pred[:, 2:4] = 10
    out = non_max_suppression(pred.to(DEV), 0.25, 0.7)
    assert all(o.shape == (0, 6) for o in out)                    # no candidate at all
    # boxes exactly at the IoU threshold: IoU = 1/3 with thr = 1/3 -> NOT > thr in fp32/double compare
    p = torch.zeros((1, 5, 2))
    p[0, :, 0] = torch.tensor([10.0, 10.0, 20.0, 20.0, 0.9])
    p[0, :, 1] = torch.tensor([20.0, 10.0, 20.0, 20.0, 0.8])      # overlap 10x20 of two 20x20 boxes: IoU = 1/3
    from oracle import nms_ref
    for thr in (1.0 / 3.0, 0.33, 0.34):
        ref = nms_ref.non_max_suppression(p.numpy(), 0.25, thr)
        got = non_max_suppression(p.to(DEV), 0.25, thr)
        assert got[0].shape[0] == ref[0].shape[0]
    with pytest.raises(AssertionError):
        non_max_suppression(p.to(DEV), 1.5, 0.5)


def test_stft_letterbox(lib):
    from oracle import stft_ref
    from specyolo import ops
    from specyolo.nn.init import synth_iq

    iq = synth_iq(2, 1 << 16, seed=3)           # 2^16 samples -> 1024 x 253 spectrogram (up-sampled in time)
    ref = stft_ref.iq_to_letterbox(iq.numpy(), out_hw=(640, 640))
    got = ops.iq_to_letterbox(iq.to(DEV), out_dtype=torch.float32).cpu().numpy()
    assert got.shape == ref.shape
    # fp32 FFT + log vs float64 spec: 1e-4 on the [0,1] image, a few px at steep edges excepted
    err = np.abs(got - ref)
    assert err.max() < 5e-3 and np.mean(err > 1e-4) < 1e-3, (err.max(), np.mean(err > 1e-4))
    got_bf = ops.iq_to_letterbox(iq.to(DEV), out_dtype=torch.bfloat16).float().cpu().numpy()
    assert np.abs(got_bf - ref).max() < 8e-3


@pytest.mark.parametrize("L,hop,out_hw", [
    (50_000, 200, (250, 333)),      # odd output width (scalar stores, scalar padding), hop not a multiple of 32
    (1024, 256, (1024, 8)),         # a single frame -> one content column, all 1024 bins as rows (no vertical resampling)
    (3000, 96, (640, 640)),         # 21 frames stretched over a 640-row band (tallest staging tile)
    (200_000, 256, (320, 1288)),    # wide output, fp32 vector padding with a partial last column tile
])
def test_stft_letterbox_ragged(lib, L, hop, out_hw):
    """Geometries off the north-star path: every store / padding variant of the kernel against the float64 oracle."""
    from oracle import stft_ref
    from specyolo import ops
    from specyolo.nn.init import synth_iq

    iq = synth_iq(3, L, seed=L % 97)
    ref = stft_ref.iq_to_letterbox(iq.numpy(), hop=hop, out_hw=out_hw)
    got = ops.iq_to_letterbox(iq.to(DEV), hop=hop, out_hw=out_hw, out_dtype=torch.float32).cpu().numpy()
    err = np.abs(got - ref)
    assert err.max() < 5e-3 and np.mean(err > 1e-4) < 1e-3, (err.max(), np.mean(err > 1e-4))
    got_bf = ops.iq_to_letterbox(iq.to(DEV), hop=hop, out_hw=out_hw, out_dtype=torch.bfloat16).float().cpu().numpy()
    assert np.abs(got_bf - ref).max() < 8e-3


def test_stft_letterbox_full_burst(lib):
    """North-star geometry: 2^20 samples, hop 256 -> 1024 x 4093 -> 160 x 640 band."""
    from oracle import stft_ref
    from specyolo import ops
    from specyolo.nn.init import synth_iq

    iq = synth_iq(1, 1 << 20, seed=4)
    ref = stft_ref.iq_to_letterbox(iq.numpy(), out_hw=(640, 640))
    got = ops.iq_to_letterbox(iq.to(DEV), out_dtype=torch.float32).cpu().numpy()
    err = np.abs(got - ref)
    assert err.max() < 5e-3 and np.mean(err > 1e-4) < 1e-3, (err.max(), np.mean(err > 1e-4))
    assert np.all(got[0, :, :240] == np.float32(114.0 / 255.0)) and np.all(got[0, :, 400:] == np.float32(114.0 / 255.0))

"""Tensor-level wrappers over the C-ABI: unpack torch tensors into pointers / strides and call
libspecyolo on the current CUDA stream.  PyTorch is plumbing only (allocation, streams).

Activation convention: a feature map is a bf16 tensor of logical shape [B, C, H, W] whose memory is
NHWC ("channels last", stride(1) == 1).  A channel slice of a wider buffer is a valid feature map,
which is how concat buffers are written in place.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (BneckArgs, BottleNectArgs, ConvArgs, DecodeArgs, DwpwArgs, FusionArgs, MscGateArgs, NmsArgs, SpatialGateArgs, StemPairArgs,
                   StftArgs, check)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def new_act(B: int, Cc: int, H: int, W: int, device, dtype=torch.bfloat16) -> torch.Tensor:
    """Uninitialised [B,C,H,W] feature map with NHWC memory."""
    return torch.empty((B, H, W, Cc), device=device, dtype=dtype).permute(0, 3, 1, 2)


def nhwc_meta(t: torch.Tensor):
    """(B, C, H, W, pixel stride) of an NHWC-memory feature map; raises if the layout is anything else."""
    if t.dim() != 4 or not t.is_cuda:
        raise ValueError(f"expected a 4-D CUDA feature map, got {tuple(t.shape)} on {t.device}")
    B, Cc, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    pix = sw if W > 1 else (sh if H > 1 else (sb if B > 1 else Cc))
    ok = (sc == 1 or Cc == 1) and (W == 1 or sw == pix) and (H == 1 or sh == W * pix) and (B == 1 or sb == H * W * pix)
    if not ok or pix < Cc:
        raise ValueError(f"feature map is not NHWC-contiguous: shape {tuple(t.shape)} strides {t.stride()}")
    return B, Cc, H, W, pix


# ----------------------------------------------------------------------------------------------
# concurrency: independent chains of small kernels on side streams (fork / join)
# ----------------------------------------------------------------------------------------------
_side_streams: dict = {}
CONCURRENT = True     # bench.py's per-kernel timing pass switches this off so kernels are timed alone


def run_concurrently(fns):
    """Run independent callables on separate CUDA streams, forked from and joined back into the current stream.

    The small layers of the neck and the six Detect branches are 200-3200 tiles each — not enough to fill 148 SMs
    alone — and they do not depend on each other, so they are issued as parallel branches (under CUDA-graph
    capture this becomes a fork/join in the graph).  Returns the list of results; tensors inside them are marked as
    used by the joining stream so the caching allocator cannot recycle them early.
    """
    if len(fns) <= 1 or not CONCURRENT:
        return [f() for f in fns]
    cur = torch.cuda.current_stream()
    dev = cur.device
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < len(fns) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    fork = cur.record_event()
    results, joins = [None] * len(fns), []
    for i, f in enumerate(fns):
        if i == 0:
            continue
        st = pool[i - 1]
        st.wait_event(fork)
        with torch.cuda.stream(st):
            results[i] = f()
            joins.append(st.record_event())
    results[0] = fns[0]()           # first chain stays on the current stream
    for ev in joins:
        cur.wait_event(ev)

    def mark(o):
        if isinstance(o, torch.Tensor):
            o.record_stream(cur)
        elif isinstance(o, (list, tuple)):
            for t in o:
                mark(t)
        elif hasattr(o, "src") and isinstance(getattr(o, "src"), torch.Tensor):
            o.src.record_stream(cur)
    for r in results[1:]:
        mark(r)
    return results


# ----------------------------------------------------------------------------------------------
# layout
# ----------------------------------------------------------------------------------------------
_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.uint8: _lib.DT_U8}


def to_nhwc_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NCHW-contiguous fp32 / bf16 / uint8 (scaled by 1/255) -> NHWC bf16 feature map."""
    _lib.init_device()
    if x.dtype not in _DT:
        raise TypeError(f"unsupported input dtype {x.dtype}")
    x = x.contiguous()
    B, Cc, H, W = x.shape
    if out is None:
        out = new_act(B, Cc, H, W, x.device)
    _, _, _, _, pix = nhwc_meta(out)
    check(_lib.load().specyolo_nchw_to_nhwc_bf16(x.data_ptr(), _DT[x.dtype], 1.0, B, Cc, H, W, out.data_ptr(), pix,
                                                 _lib.stream_ptr()))
    return out


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Upsample(None, 2, 'nearest') of an NHWC bf16 feature map, optionally into a channel slice of a concat buffer
    (specyolo_upsample2x)."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("upsample2x expects bf16")
    if out is None:
        out = new_act(B, Cc, 2 * H, 2 * W, x.device)
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, Cc, 2 * H, 2 * W) or out.dtype != torch.bfloat16:
        raise ValueError("upsample2x: out shape / dtype mismatch")
    check(_lib.load().specyolo_upsample2x(x.data_ptr(), xpix, B, H, W, Cc, out.data_ptr(), ypix, _lib.stream_ptr()))
    return out


def to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    """NHWC bf16 feature map -> NCHW-contiguous fp32 tensor."""
    B, Cc, H, W, pix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("to_nchw_f32 expects bf16")
    y = torch.empty((B, Cc, H, W), device=x.device, dtype=torch.float32)
    check(_lib.load().specyolo_nhwc_bf16_to_nchw_f32(x.data_ptr(), pix, B, Cc, H, W, y.data_ptr(), _lib.stream_ptr()))
    return y


# ----------------------------------------------------------------------------------------------
# conv
# ----------------------------------------------------------------------------------------------
@dataclass
class PackedConv:
    """BN-folded, repacked conv weights (device) + geometry."""
    w: torch.Tensor          # bf16 [groups*n_pad, kh*kw*cin_g]
    bias: torch.Tensor       # fp32 [groups*n_pad]
    cin: int
    cout: int
    k: int
    s: int
    p: int
    d: int
    g: int
    n_pad: int
    act: int
    g_orig: int = 1          # groups of the source conv (g is the packed, possibly merged, group count)
    cin_true: int = 0        # input channels of the source conv (cin may be zero-padded to a multiple of 16)
    w_folded_f32: Optional[torch.Tensor] = None   # stem only: fp32 OIHW folded weights
    alg_k: int = 0           # algorithmic reduction length when it differs from cin/g*k*k (space-to-depth stem: 27)
    trim: int = 0            # far-edge output rows / columns that are not computed (space-to-depth stem: 1)
    stem_w0: Optional[torch.Tensor] = None   # stem only: fp16 [cout, 32] im2col weights / 255 for ops.stem_pair
    stem_b0: Optional[torch.Tensor] = None   # stem only: folded bias - 1024 * sum(stem_w0) (see fold_pack)
    s2d: Optional[dict] = None   # stem only: {"u8": PackedConv, "f": PackedConv} 2x2 convs over the blocked image

    def out_hw(self, H: int, W: int):
        ke = self.d * (self.k - 1) + 1
        return (H + 2 * self.p - ke) // self.s + 1 - self.trim, (W + 2 * self.p - ke) // self.s + 1 - self.trim


def fold_pack(weight: torch.Tensor, conv_bias: Optional[torch.Tensor], bn: Optional[Sequence[torch.Tensor]],
              eps: float, stride: int, pad: int, dil: int, groups: int, act: bool) -> PackedConv:
    """fuse_conv_and_bn + repack on the device (specyolo_fold_pack_conv).  `bn` = (gamma, beta, mean, var)."""
    _lib.init_device()
    lib = _lib.load()
    w = weight.detach().to(torch.float32).contiguous()
    if not w.is_cuda:
        raise ValueError("fold_pack: weights must live on the CUDA device")
    cout, cin_g, kh, kw = w.shape
    if kh != kw:
        raise ValueError("square kernels only")
    cin_true = cin_g * groups
    if groups == 1 and cin_g > 3 and cin_g % 16:
        # the UMMA K step is 16 channels: thin convs (8 channels in yolo11n's first C3k2) get zero weight columns;
        # conv2d() zero-extends the activation to match (a plumbing copy on a layer that is 0.2 % of the FLOPs)
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 16 - cin_g % 16))
        cin_g = w.shape[1]
    merge = lib.specyolo_conv_merge(cin_g * groups, cout, groups, kh, stride, pad, dil)   # grouped convs on the per-tap kernel: fuse groups up to 64-ch K chunks
    pgroups = groups // merge
    n_pad = lib.specyolo_conv_npad(cout, pgroups)
    if n_pad <= 0:
        raise ValueError("bad conv shape")
    wp = torch.empty((pgroups * n_pad, kh * kw * cin_g * merge), device=w.device, dtype=torch.bfloat16)
    bias = torch.empty((pgroups * n_pad,), device=w.device, dtype=torch.float32)
    bnt = [None] * 4 if bn is None else [t.detach().to(torch.float32).contiguous() for t in bn]
    cb = None if conv_bias is None else conv_bias.detach().to(torch.float32).contiguous()
    check(lib.specyolo_fold_pack_conv(w.data_ptr(), _p(cb), _p(bnt[0]), _p(bnt[1]), _p(bnt[2]), _p(bnt[3]),
                                      float(eps), cout, cin_g, kh, kw, groups, merge, n_pad, wp.data_ptr(),
                                      bias.data_ptr(), _lib.stream_ptr()))
    pc = PackedConv(wp, bias, cin_g * groups, cout, kh, stride, pad, dil, pgroups, n_pad,
                    _lib.ACT_SILU if act else _lib.ACT_NONE, g_orig=groups)
    pc.cin_true = cin_true
    if cin_g * groups == 3:  # stem: CUDA-core kernel consumes folded fp32 OIHW weights
        if bn is not None:
            scale = bnt[0] / torch.sqrt(bnt[3] + eps)
            pc.w_folded_f32 = (w * scale.view(-1, 1, 1, 1)).contiguous()
        else:
            pc.w_folded_f32 = w
        if kh == 3 and stride == 2 and pad == 1 and dil == 1 and groups == 1 and cin_g == 3:
            # tensor-core route: the same conv as a 2x2 / stride-1 conv over the 2x2-blocked (space-to-depth) image
            w2 = torch.zeros((cout, 16, 2, 2), device=w.device, dtype=torch.float32)
            for ty in range(2):
                for tx in range(2):
                    for dy in range(2):
                        for dx in range(2):
                            ky, kx = 2 * ty + dy - 1, 2 * tx + dx - 1
                            if 0 <= ky < 3 and 0 <= kx < 3:
                                c0 = (dy * 2 + dx) * 3
                                w2[:, c0:c0 + 3, ty, tx] = w[:, :, ky, kx]
            # fused stem kernel (ops.stem_pair): fp16 weights / 255, column c*9 + ky*3 + kx; the kernel's A operand is
            # 1024 + pixel (fp16 bits 0x6400 | byte), so 1024 * sum(w) comes off the bias
            w0p = torch.zeros((cout, 32), device=w.device, dtype=torch.float16)
            w0p[:, :27] = (pc.w_folded_f32 * (1.0 / 255.0)).reshape(cout, 27).to(torch.float16)
            pc.stem_w0 = w0p
            pc.stem_b0 = (bias[:cout].double() - 1024.0 * w0p.double().sum(dim=1)).float().contiguous()
            pc.s2d = {}
            for key, scale in (("u8", 1.0 / 255.0), ("f", 1.0)):
                q = fold_pack(w2 * scale, conv_bias, bn, eps, 1, 1, 1, 1, act)
                q.trim = 1
                q.alg_k = 27
                pc.s2d[key] = q
    return pc


def conv2d(x: torch.Tensor, pc: PackedConv, out: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out_fp32: bool = False, blocked_out: bool = False) -> torch.Tensor:
    """y = act(conv(x) + b) [+ residual] on NHWC feature maps (specyolo_conv2d_bias_act).

    blocked_out: the result is written 2x2-blocked (space-to-depth) as a [B, 4*cout, Ho/2, Wo/2] feature map
    (channel ((oh%2)*2 + ow%2)*cout + c) for a following stride-2 conv packed with `pack_from_blocked`."""
    B, Cin, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("conv2d expects a bf16 feature map")
    if Cin != pc.cin:
        if Cin != pc.cin_true:
            raise ValueError(f"conv2d: input has {Cin} channels, weights expect {pc.cin_true}")
        xp = torch.zeros((B, H, W, pc.cin), device=x.device, dtype=x.dtype).permute(0, 3, 1, 2)
        xp[:, :Cin].copy_(x)
        x, Cin, xpix = xp, pc.cin, pc.cin
    Ho, Wo = pc.out_hw(H, W)
    oshape = (B, 4 * pc.cout, Ho // 2, Wo // 2) if blocked_out else (B, pc.cout, Ho, Wo)
    if blocked_out and (Ho % 2 or Wo % 2 or out_fp32 or residual is not None):
        raise ValueError("conv2d: blocked output needs even Ho/Wo, bf16 output and no residual")
    if out is None:
        out = new_act(*oshape, x.device, torch.float32 if out_fp32 else torch.bfloat16)
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != oshape:
        raise ValueError(f"conv2d: out shape {tuple(out.shape)} != {oshape}")
    if out.dtype != (torch.float32 if out_fp32 else torch.bfloat16):
        raise TypeError("conv2d: out dtype mismatch")
    a = ConvArgs()
    a.x, a.B, a.H, a.W, a.Cin, a.x_pixstride, a.x_upshift = x.data_ptr(), B, H, W, Cin, xpix, 0
    a.w_packed, a.bias, a.Cout, a.n_pad = pc.w.data_ptr(), pc.bias.data_ptr(), pc.cout, pc.n_pad
    a.kh = a.kw = pc.k
    a.stride, a.pad, a.dil, a.groups, a.act = pc.s, pc.p, pc.d, pc.g, pc.act
    a.y, a.Ho, a.Wo, a.y_pixstride, a.y_fp32 = out.data_ptr(), Ho, Wo, ypix, int(out_fp32)
    a.y_s2d = int(blocked_out)
    if residual is not None:
        rB, rC, rH, rW, rpix = nhwc_meta(residual)
        if (rB, rC, rH, rW) != (B, pc.cout, Ho, Wo) or residual.dtype != torch.bfloat16:
            raise ValueError("conv2d: residual shape/dtype mismatch")
        a.residual, a.r_pixstride = residual.data_ptr(), rpix
    else:
        a.residual, a.r_pixstride = None, 0
    check(_lib.load().specyolo_conv2d_bias_act(C.byref(a), _lib.stream_ptr()))
    return out


def dwconv_pwconv_ok(C_in: int, pw: PackedConv) -> bool:
    """Shapes the fused depthwise + pointwise kernel takes (specyolo_dwconv_pwconv)."""
    return C_in in (64, 128) and pw.k == 1 and pw.g == 1 and pw.cin == C_in and 64 <= pw.n_pad <= 256 and \
        pw.act == _lib.ACT_SILU


DWPW_HEAD_MAX_NC = 4


def dwconv_pwconv(x: torch.Tensor, dw_w: torch.Tensor, dw_b: torch.Tensor, pw: PackedConv,
                  out: Optional[torch.Tensor] = None, head: Optional[tuple] = None) -> Optional[torch.Tensor]:
    """SiLU(conv1x1(SiLU(dwconv3x3(x) + b_dw)) + b_pw) in one kernel; dw_w fp32 [9, C] (tap-major), dw_b fp32 [C].

    head = (w [nc, Cout] fp32, b [nc] fp32, y fp32 view [B, nc, H, W] with NHWC memory): the closing Conv2d(Cout, nc, 1)
    of a Detect.cv3 branch (nc <= DWPW_HEAD_MAX_NC) evaluated in the epilogue; the Cout-channel tensor is then never
    written and the function returns None."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16 or dw_w.shape != (9, Cc) or dw_w.dtype != torch.float32 or not dw_w.is_contiguous():
        raise ValueError("dwconv_pwconv: x must be bf16 NHWC, dw_w contiguous fp32 [9, C]")
    a = DwpwArgs()
    a.x, a.B, a.H, a.W, a.C, a.x_pixstride = x.data_ptr(), B, H, W, Cc, xpix
    a.dw_w, a.dw_b, a.pw_packed, a.pw_bias = dw_w.data_ptr(), dw_b.data_ptr(), pw.w.data_ptr(), pw.bias.data_ptr()
    a.Cout, a.n_pad = pw.cout, pw.n_pad
    if head is not None:
        hw, hb, hy = head
        nc = hw.shape[0]
        hB, hC, hH, hW_, hpix = nhwc_meta(hy)
        if hw.shape != (nc, pw.cout) or hw.dtype != torch.float32 or not hw.is_contiguous() or hb.shape != (nc,) or \
                hb.dtype != torch.float32 or hy.dtype != torch.float32 or (hB, hC, hH, hW_) != (B, nc, H, W) or \
                nc > DWPW_HEAD_MAX_NC:
            raise ValueError("dwconv_pwconv: bad fused-head tensors")
        a.y, a.y_pixstride = None, pw.cout
        a.head_w, a.head_b, a.head_y, a.head_nc, a.head_pixstride = hw.data_ptr(), hb.data_ptr(), hy.data_ptr(), nc, hpix
        check(_lib.load().specyolo_dwconv_pwconv(C.byref(a), _lib.stream_ptr()))
        return None
    if out is None:
        out = new_act(B, pw.cout, H, W, x.device)
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, pw.cout, H, W) or out.dtype != torch.bfloat16:
        raise ValueError("dwconv_pwconv: out shape / dtype mismatch")
    a.y, a.y_pixstride = out.data_ptr(), ypix
    a.head_w = a.head_b = a.head_y = None
    a.head_nc = a.head_pixstride = 0
    check(_lib.load().specyolo_dwconv_pwconv(C.byref(a), _lib.stream_ptr()))
    return out


def bottleneck_ok(x: torch.Tensor, pc1: PackedConv, pc2: PackedConv, prefer: bool = False) -> bool:
    """Shapes the fused Bottleneck kernel takes (specyolo_bottleneck_ok): two dense 3x3 / s1 / p1 SiLU convs, C -> Cmid -> C.
    prefer=True (the model's dispatch): additionally only where the fused kernel measured faster than two launches —
    not for the thinnest pair (32 -> 16 -> 32: 151 us fused against 145 us as two halo-kernel launches at 160^2, batch
    64; every MMA of either form costs the same ~40 cycles of shared-memory operand reads whatever N is, and the fused
    form issues 1.7x as many because of the halo recompute)."""
    for pc in (pc1, pc2):
        if pc.k != 3 or pc.s != 1 or pc.p != 1 or pc.d != 1 or pc.g != 1 or pc.g_orig != 1 or pc.act != _lib.ACT_SILU:
            return False
    if x.dtype != torch.bfloat16 or x.shape[1] != pc1.cin or pc1.cout != pc2.cin or pc1.cin != pc1.cin_true or \
            pc2.cin != pc2.cin_true:
        return False
    if prefer and pc1.cin == 32 and pc1.cout == 16:
        return False
    return bool(_lib.load().specyolo_bottleneck_ok(pc1.cin, pc1.cout, pc2.cout, pc1.n_pad, pc2.n_pad))


def bottleneck(x: torch.Tensor, pc1: PackedConv, pc2: PackedConv, add: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = [x +] conv3x3(conv3x3(x)) (Conv + BN + SiLU each) in one kernel (specyolo_bottleneck): the intermediate stays
    in shared memory, the shortcut is added in the second epilogue."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if not bottleneck_ok(x, pc1, pc2) or (add and pc2.cout != Cc):
        raise ValueError("bottleneck: unsupported shapes / weights")
    if out is None:
        out = new_act(B, pc2.cout, H, W, x.device)
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, pc2.cout, H, W) or out.dtype != torch.bfloat16:
        raise ValueError("bottleneck: out shape / dtype mismatch")
    a = BneckArgs()
    a.x, a.B, a.H, a.W, a.C, a.x_pixstride = x.data_ptr(), B, H, W, Cc, xpix
    a.w1_packed, a.b1, a.Cmid, a.n_pad1 = pc1.w.data_ptr(), pc1.bias.data_ptr(), pc1.cout, pc1.n_pad
    a.w2_packed, a.b2, a.Cout, a.n_pad2 = pc2.w.data_ptr(), pc2.bias.data_ptr(), pc2.cout, pc2.n_pad
    a.add, a.y, a.y_pixstride = int(bool(add)), out.data_ptr(), ypix
    check(_lib.load().specyolo_bottleneck(C.byref(a), _lib.stream_ptr()))
    return out


def stem_space_to_depth(x: torch.Tensor) -> torch.Tensor:
    """NCHW network input (fp32 | bf16 | uint8) -> 2x2-blocked NHWC bf16 [B, 16, H/2, W/2] (specyolo_stem_space_to_depth)."""
    B, _, H, W = x.shape
    blocked = new_act(B, 16, H // 2, W // 2, x.device)
    check(_lib.load().specyolo_stem_space_to_depth(x.data_ptr(), _DT[x.dtype], B, H, W, blocked.data_ptr(), 16,
                                                   _lib.stream_ptr()))
    return blocked


def pack_from_blocked(weight: torch.Tensor, conv_bias: Optional[torch.Tensor], bn, eps: float, act: bool) -> PackedConv:
    """Weights of a 3x3 / stride-2 / pad-1 conv [cout, c, 3, 3] repacked as the equivalent 2x2 / stride-1 conv over the
    2x2-blocked input (4c channels, taps at block offsets -1, 0): w2[:, (dy*2+dx)*c + ci, ty, tx] = w[:, ci, 2ty+dy-1,
    2tx+dx-1].  K grows from 9c to 16c (zeros), which is free while the layer is HBM-bound, and the input is then read
    once through one 128-byte-row TMA box per tile instead of nine strided 64-byte-row boxes."""
    w = weight.detach().to(torch.float32)
    cout, c, kh, kw = w.shape
    if kh != 3 or kw != 3:
        raise ValueError("pack_from_blocked: 3x3 kernels only")
    w2 = torch.zeros((cout, 4 * c, 2, 2), device=w.device, dtype=torch.float32)
    for ty in range(2):
        for tx in range(2):
            for dy in range(2):
                for dx in range(2):
                    ky, kx = 2 * ty + dy - 1, 2 * tx + dx - 1
                    if 0 <= ky < 3 and 0 <= kx < 3:
                        c0 = (dy * 2 + dx) * c
                        w2[:, c0:c0 + c, ty, tx] = w[:, :, ky, kx]
    q = fold_pack(w2, conv_bias, bn, eps, 1, 1, 1, 1, act)
    q.trim = 1
    q.alg_k = 9 * c
    return q


def stem_pair_ok(x: torch.Tensor, pc0: PackedConv, pc1b: PackedConv) -> bool:
    """Shapes the fused two-layer stem kernel takes (specyolo_stem_pair_ok): uint8 NCHW input, SiLU on both layers."""
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[1] != 3 or pc0.stem_w0 is None:
        return False
    if pc0.act != _lib.ACT_SILU or pc1b.act != _lib.ACT_SILU or pc1b.k != 2 or pc1b.g != 1 or pc1b.cin != 4 * pc0.cout:
        return False
    return bool(_lib.load().specyolo_stem_pair_ok(x.shape[2], x.shape[3], pc0.cout, pc1b.cout, pc1b.n_pad))


def stem_pair(x: torch.Tensor, pc0: PackedConv, pc1b: PackedConv, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Layers 0 and 1 of the trunk in one kernel (specyolo_stem_pair): uint8 NCHW image -> bf16 NHWC [B, c1, H/4, W/4].
    pc0 = fold_pack of the 3->c0 stem, pc1b = pack_from_blocked of the c0->c1 3x3/s2 conv."""
    _lib.init_device()
    if not stem_pair_ok(x, pc0, pc1b):
        raise ValueError("stem_pair: unsupported input / weights")
    x = x.contiguous()
    B, _, H, W = x.shape
    if out is None:
        out = new_act(B, pc1b.cout, H // 4, W // 4, x.device)
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, pc1b.cout, H // 4, W // 4) or out.dtype != torch.bfloat16:
        raise ValueError("stem_pair: out shape / dtype mismatch")
    a = StemPairArgs()
    a.x, a.B, a.H, a.W = x.data_ptr(), B, H, W
    a.w0, a.b0, a.c0 = pc0.stem_w0.data_ptr(), pc0.stem_b0.data_ptr(), pc0.cout
    a.w1_packed, a.b1, a.Cout, a.n_pad = pc1b.w.data_ptr(), pc1b.bias.data_ptr(), pc1b.cout, pc1b.n_pad
    a.y, a.y_pixstride = out.data_ptr(), ypix
    check(_lib.load().specyolo_stem_pair(C.byref(a), _lib.stream_ptr()))
    return out


def stem_conv(x: torch.Tensor, pc: PackedConv, out: Optional[torch.Tensor] = None, blocked_out: bool = False) -> torch.Tensor:
    """3-channel 3x3/s2 stem reading the NCHW network input (fp32 | bf16 | uint8/255)."""
    _lib.init_device()
    if x.dtype not in _DT:
        raise TypeError(f"unsupported input dtype {x.dtype}")
    if pc.w_folded_f32 is None or pc.k != 3 or pc.s != 2 or pc.p != 1 or pc.act != _lib.ACT_SILU:
        raise ValueError("stem_conv: weights are not a 3->C 3x3/s2 SiLU stem")
    x = x.contiguous()
    B, Cin, H, W = x.shape
    if pc.s2d is not None and H % 2 == 0 and W % 2 == 0:
        # tensor-core route: space-to-depth (one streaming pass) + a K = 64 implicit GEMM
        return conv2d(stem_space_to_depth(x), pc.s2d["u8" if x.dtype == torch.uint8 else "f"], out=out,
                      blocked_out=blocked_out)
    if blocked_out:
        raise ValueError("stem_conv: blocked output needs the space-to-depth route (even H, W)")
    Ho, Wo = pc.out_hw(H, W)
    if out is None:
        out = new_act(B, pc.cout, Ho, Wo, x.device)
    _, _, _, _, ypix = nhwc_meta(out)
    check(_lib.load().specyolo_stem_conv3x3s2(x.data_ptr(), _DT[x.dtype], B, H, W, pc.w_folded_f32.data_ptr(),
                                              pc.bias.data_ptr(), pc.cout, out.data_ptr(), ypix, _lib.stream_ptr()))
    return out


# ----------------------------------------------------------------------------------------------
# neck glue
# ----------------------------------------------------------------------------------------------
def sppf_pool(buf: torch.Tensor, c: int) -> None:
    """Fill channels [c,4c) of the SPPF concat buffer with the three chained 5x5 max-pools of [0,c)."""
    B, Cc, H, W, pix = nhwc_meta(buf)
    if Cc != 4 * c:
        raise ValueError("sppf_pool: buffer must have 4*c channels")
    check(_lib.load().specyolo_sppf_pool(buf.data_ptr(), B, H, W, c, pix, _lib.stream_ptr()))


def fusion_eschannel(xs: Sequence[torch.Tensor], upshift: Sequence[int], alpha: torch.Tensor, gamma: torch.Tensor,
                     beta: torch.Tensor, eps: float, sab_w: torch.Tensor,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fusion('ESChannel'); inputs flagged in `upshift` are read through a x2 nearest upsample."""
    k = len(xs)
    metas = [nhwc_meta(x) for x in xs]
    B, c = metas[0][0], metas[0][1]
    H, W = metas[0][2] << upshift[0], metas[0][3] << upshift[0]
    for m, u in zip(metas, upshift):
        if (m[0], m[1], m[2] << u, m[3] << u) != (B, c, H, W):
            raise ValueError("fusion: input shapes differ")
    if out is None:
        out = new_act(B, c, H, W, xs[0].device)
    lib = _lib.load()
    ws = torch.empty(lib.specyolo_fusion_ws_bytes(k, B, H, W, c), device=xs[0].device, dtype=torch.uint8)
    a = FusionArgs()
    a.k = k
    for i in range(k):
        a.x[i], a.pixstride[i], a.upshift[i] = xs[i].data_ptr(), metas[i][4], int(upshift[i])
    a.B, a.H, a.W, a.c = B, H, W, c
    a.alpha, a.gamma, a.beta, a.gct_eps = alpha.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps)
    a.sab_w = sab_w.data_ptr()
    a.y, a.y_pixstride = out.data_ptr(), nhwc_meta(out)[4]
    a.ws = ws.data_ptr()
    check(lib.specyolo_fusion_eschannel(C.byref(a), _lib.stream_ptr()))
    return out


def sobel_spatial_attention(x: torch.Tensor, w18: Sequence[float], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SobelSpatialAttention (conv.py:1184-1198): x * sigmoid(stencil(mean_c x, max_c x)); `w18` is the folded
    2 x 3 x 3 stencil (see include/specyolo.h).  out=None gates x in place."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("sobel_spatial_attention expects bf16")
    if out is None:
        out = x
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, Cc, H, W) or out.dtype != torch.bfloat16:
        raise ValueError("sobel_spatial_attention: out shape / dtype mismatch")
    mm = torch.empty((B, 2, H, W), device=x.device, dtype=torch.float32)
    a = SpatialGateArgs()
    a.x, a.x_pixstride, a.y, a.y_pixstride = x.data_ptr(), xpix, out.data_ptr(), ypix
    a.B, a.H, a.W, a.C = B, H, W, Cc
    for i, v in enumerate(w18):
        a.w[i] = float(v)
    a.mm = mm.data_ptr()
    check(_lib.load().specyolo_sobel_spatial_attention(C.byref(a), _lib.stream_ptr()))
    return out


_BOTTLENECT_FIELDS = ("in_w", "in_b", "fac_w", "fac_b", "sca_w", "sca_b", "dw1_w", "dw1_b", "dw2_w", "dw2_b", "alpha", "beta")


def bottlenect(x: torch.Tensor, weights: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """BottleNect + FGM (block.py:782-861): relu(|ifft2(x1 * fft2(x2))| * alpha + x_sca * beta), see include/specyolo.h.
    `weights`: fp32 device tensors in the order in_conv.0 (w, b), fac_conv (w, b), conv (w, b), fgm.dwconv1 (w, b),
    fgm.dwconv2 (w, b), fgm.alpha, fgm.beta.  out=None runs in place."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("bottlenect expects bf16")
    if out is None:
        out = x
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, Cc, H, W) or out.dtype != torch.bfloat16:
        raise ValueError("bottlenect: out shape / dtype mismatch")
    if len(weights) != len(_BOTTLENECT_FIELDS):
        raise ValueError("bottlenect: 12 weight tensors expected")
    for t, name in zip(weights, _BOTTLENECT_FIELDS):
        n = Cc * Cc if name.endswith("_w") else Cc
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n or t.device != x.device:
            raise ValueError(f"bottlenect: {name} must be a contiguous fp32 tensor of {n} elements on x's device")
    lib = _lib.load()
    ws = torch.empty(lib.specyolo_bottlenect_ws_bytes(B, H, W, Cc), device=x.device, dtype=torch.uint8)
    a = BottleNectArgs()
    a.x, a.x_pixstride, a.y, a.y_pixstride = x.data_ptr(), xpix, out.data_ptr(), ypix
    a.B, a.H, a.W, a.C = B, H, W, Cc
    for t, name in zip(weights, _BOTTLENECT_FIELDS):
        setattr(a, name, t.data_ptr())
    a.ws = ws.data_ptr()
    check(lib.specyolo_bottlenect(C.byref(a), _lib.stream_ptr()))
    return out


def msc_spatial_attention(x: torch.Tensor, w_big: torch.Tensor, w_small: torch.Tensor, fc_w: torch.Tensor, fc_b: torch.Tensor,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MSCSpatialAttention (conv.py:1200-1243): y = x * s * g + x with s = relu(cv1(mm)) + relu(cv2(mm)) over the channel
    mean / max planes mm and g = relu(fc(mean_hw(x * s))).  Weights are fp32 device tensors: w_big [1,2,31,31],
    w_small [1,2,3,3], fc_w [C,C,1,1], fc_b [C].  out=None runs in place."""
    B, Cc, H, W, xpix = nhwc_meta(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("msc_spatial_attention expects bf16")
    if out is None:
        out = x
    oB, oC, oH, oW, ypix = nhwc_meta(out)
    if (oB, oC, oH, oW) != (B, Cc, H, W) or out.dtype != torch.bfloat16:
        raise ValueError("msc_spatial_attention: out shape / dtype mismatch")
    for t, n in ((w_big, 2 * w_big.shape[-1] ** 2), (w_small, 18), (fc_w, Cc * Cc), (fc_b, Cc)):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n or t.device != x.device:
            raise ValueError("msc_spatial_attention: weights must be contiguous fp32 tensors on x's device")
    lib = _lib.load()
    ws = torch.empty(lib.specyolo_msc_ws_bytes(B, H, W, Cc), device=x.device, dtype=torch.uint8)
    a = MscGateArgs()
    a.x, a.x_pixstride, a.y, a.y_pixstride = x.data_ptr(), xpix, out.data_ptr(), ypix
    a.B, a.H, a.W, a.C = B, H, W, Cc
    a.w_big, a.k_big, a.w_small = w_big.data_ptr(), int(w_big.shape[-1]), w_small.data_ptr()
    a.fc_w, a.fc_b, a.ws = fc_w.data_ptr(), fc_b.data_ptr(), ws.data_ptr()
    check(lib.specyolo_msc_spatial_attention(C.byref(a), _lib.stream_ptr()))
    return out


def psa_attention(qkv: torch.Tensor, heads: int, key_dim: int, head_dim: int, scale: float, pe_w: torch.Tensor,
                  pe_b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, Cc, H, W, pix = nhwc_meta(qkv)
    if Cc != heads * (2 * key_dim + head_dim):
        raise ValueError("psa_attention: qkv channel count mismatch")
    if out is None:
        out = new_act(B, heads * head_dim, H, W, qkv.device)
    check(_lib.load().specyolo_psa_attention(qkv.data_ptr(), pix, B, H, W, heads, key_dim, head_dim, float(scale),
                                             pe_w.data_ptr(), pe_b.data_ptr(), out.data_ptr(), nhwc_meta(out)[4],
                                             _lib.stream_ptr()))
    return out


# ----------------------------------------------------------------------------------------------
# head
# ----------------------------------------------------------------------------------------------
def num_segments(A: int) -> int:
    return (A + _lib.DECODE_SEG - 1) // _lib.DECODE_SEG


def detect_decode(logits: Sequence[torch.Tensor], hw: Sequence[tuple], strides: Sequence[float], nc: int,
                  want_dense: bool = True, conf_thres: Optional[float] = None):
    """Fused DFL + dist2bbox + sigmoid (+ score threshold).

    logits[l]: fp32 [B, h*w, no_stride] (64 DFL bins then nc class logits per anchor).
    Returns (y_dense [B,4+nc,A] or None, cand [B,nseg,256,6] or None, seg_count [B,nseg] or None).
    """
    _lib.init_device()
    B, _, no_stride = logits[0].shape
    A = sum(h * w for h, w in hw)
    dev = logits[0].device
    a = DecodeArgs()
    a.nl = len(logits)
    for i, (t, (h, w), s) in enumerate(zip(logits, hw, strides)):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != (B, h * w, no_stride):
            raise ValueError("detect_decode: logits must be contiguous fp32 [B, h*w, no_stride]")
        a.logits[i], a.h[i], a.w[i], a.stride[i] = t.data_ptr(), h, w, float(s)
    a.no_stride, a.B, a.nc, a.reg_max = no_stride, B, nc, 16
    y = torch.empty((B, 4 + nc, A), device=dev, dtype=torch.float32) if want_dense else None
    cand = seg = None
    if conf_thres is not None:
        nseg = num_segments(A)
        cand = torch.empty((B, nseg, _lib.DECODE_SEG, 6), device=dev, dtype=torch.float32)
        seg = torch.empty((B, nseg), device=dev, dtype=torch.int32)
    a.y, a.conf_thres, a.cand, a.seg_count = _p(y), float(conf_thres or 0.0), _p(cand), _p(seg)
    check(_lib.load().specyolo_detect_decode(C.byref(a), _lib.stream_ptr()))
    return y, cand, seg


def nms(prediction: Optional[torch.Tensor] = None, cand: Optional[torch.Tensor] = None,
        seg_count: Optional[torch.Tensor] = None, *, B: int, nc: int, A: int, conf_thres: float, iou_thres: float,
        agnostic: bool = False, multi_label: bool = False, max_det: int = 300, max_nms: int = 30000,
        max_wh: float = 7680.0, classes: Optional[torch.Tensor] = None, clip_hw: Optional[tuple] = None):
    """Batched NMS; returns (out [B,max_det,6], count [B], keep_idx [B,max_det], n_cand [B]) on device.
    clip_hw = (H, W): kept boxes are clamped to the image as they are written (clip_boxes, ops.py:335-354)."""
    _lib.init_device()
    lib = _lib.load()
    dev = (prediction if prediction is not None else cand).device
    out = torch.zeros((B, max_det, 6), device=dev, dtype=torch.float32)
    cnt = torch.empty((B,), device=dev, dtype=torch.int32)
    keep = torch.zeros((B, max_det), device=dev, dtype=torch.int32)
    ncand = torch.empty((B,), device=dev, dtype=torch.int32)
    ml = bool(multi_label and nc > 1)
    ws = torch.empty(lib.specyolo_nms_ws_bytes(B, nc, A, int(ml)), device=dev, dtype=torch.uint8)
    a = NmsArgs()
    a.B, a.nc, a.A = B, nc, A
    a.prediction, a.cand, a.seg_count = _p(prediction), _p(cand), _p(seg_count)
    a.conf_thres, a.iou_thres = float(conf_thres), float(iou_thres)
    a.agnostic, a.multi_label, a.max_det, a.max_nms, a.max_wh = int(agnostic), int(ml), max_det, max_nms, float(max_wh)
    a.classes, a.n_classes = _p(classes), 0 if classes is None else classes.numel()
    a.out, a.out_count, a.keep_idx, a.n_cand, a.ws = out.data_ptr(), cnt.data_ptr(), keep.data_ptr(), ncand.data_ptr(), ws.data_ptr()
    a.clip_h, a.clip_w = (0.0, 0.0) if clip_hw is None else (float(clip_hw[0]), float(clip_hw[1]))
    check(lib.specyolo_nms(C.byref(a), _lib.stream_ptr()))
    return out, cnt, keep, ncand


def scale_boxes_(out: torch.Tensor, cnt: torch.Tensor, img1_shape, img0_shape, ratio_pad=None) -> None:
    """In-place scale_boxes + clip_boxes on the NMS output (ops.py:92-127 geometry; `ratio_pad` = ((rh, rw), (left, top))
    as the validation dataloader records it: gain = rh, pads as given)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad_w = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
        pad_h = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    else:
        gain = ratio_pad[0][0]
        pad_w, pad_h = ratio_pad[1]
    B, max_det, _ = out.shape
    check(_lib.load().specyolo_scale_boxes(out.data_ptr(), cnt.data_ptr(), B, max_det, float(gain), float(pad_w),
                                           float(pad_h), float(img0_shape[1]), float(img0_shape[0]),
                                           _lib.stream_ptr()))


def letterbox_u8(imgs: torch.Tensor, geometry, swap_rb: bool = True, chw: bool = True, pad_value: int = 114,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,H,W,3] uint8 (CUDA) -> letterboxed [B,3,oh,ow] (chw) or [B,oh,ow,3]; geometry = (new_w, new_h, left, top, oh, ow)
    from specyolo.data.LetterBox.geometry (specyolo_letterbox_u8, bit-exact with cv2 INTER_LINEAR)."""
    _lib.init_device()
    if imgs.dtype != torch.uint8 or imgs.dim() != 4 or imgs.shape[3] != 3 or not imgs.is_cuda or not imgs.is_contiguous():
        raise ValueError("letterbox_u8 expects a contiguous CUDA uint8 tensor [B,H,W,3]")
    B, H, W, _ = imgs.shape
    new_w, new_h, left, top, oh, ow = (int(v) for v in geometry)
    shape = (B, 3, oh, ow) if chw else (B, oh, ow, 3)
    if out is None:
        out = torch.empty(shape, device=imgs.device, dtype=torch.uint8)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError(f"letterbox_u8: out must be a contiguous uint8 tensor of shape {shape}")
    check(_lib.load().specyolo_letterbox_u8(imgs.data_ptr(), B, H, W, out.data_ptr(), oh, ow, new_w, new_h, left, top,
                                            int(pad_value), int(swap_rb), int(chw), _lib.stream_ptr()))
    return out


def match_predictions(out: torch.Tensor, cnt: torch.Tensor, labels: torch.Tensor, label_off: torch.Tensor,
                      max_labels_per_image: int, iouv: Sequence[float]) -> torch.Tensor:
    """Detections [B,max_det,6] + counts [B] vs ground-truth rows [n,5] (cls, x1,y1,x2,y2) grouped by image
    (label_off [B+1], int32) -> correct [B, max_det, niou] bool (specyolo_match_predictions)."""
    _lib.init_device()
    B, max_det, _ = out.shape
    niou = len(iouv)
    if out.dtype != torch.float32 or not out.is_contiguous() or cnt.dtype != torch.int32:
        raise ValueError("match_predictions: out must be contiguous fp32 [B,max_det,6], cnt int32")
    labels = labels.to(device=out.device, dtype=torch.float32).contiguous()
    label_off = label_off.to(device=out.device, dtype=torch.int32).contiguous()
    correct = torch.empty((B, max_det, niou), device=out.device, dtype=torch.uint8)
    arr = (C.c_float * niou)(*[float(v) for v in iouv])
    check(_lib.load().specyolo_match_predictions(out.data_ptr(), cnt.data_ptr(), B, max_det,
                                                 labels.data_ptr() if labels.numel() else None, label_off.data_ptr(),
                                                 int(max_labels_per_image), arr, niou, correct.data_ptr(), _lib.stream_ptr()))
    return correct.bool()


# ----------------------------------------------------------------------------------------------
# front end
# ----------------------------------------------------------------------------------------------
def iq_to_letterbox(iq: torch.Tensor, nfft: int = 1024, hop: int = 256, db_min: float = -100.0, db_max: float = 0.0,
                    out_hw=(640, 640), pad_value: float = 114.0 / 255.0, out_dtype=torch.bfloat16,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """complex64 bursts [B, L] -> letterboxed spectrogram images [B, 3, H, W] (NCHW) in [0,1]."""
    _lib.init_device()
    if iq.dtype == torch.complex64:
        iq = torch.view_as_real(iq)
    if iq.dtype != torch.float32 or iq.dim() != 3 or iq.shape[-1] != 2 or not iq.is_contiguous():
        raise ValueError("iq must be complex64 [B, L] (or float32 [B, L, 2]), contiguous")
    B, L, _ = iq.shape
    if out is None:
        out = torch.empty((B, 3, out_hw[0], out_hw[1]), device=iq.device, dtype=out_dtype)
    a = StftArgs()
    a.iq, a.B, a.L, a.nfft, a.hop = iq.data_ptr(), B, L, nfft, hop
    a.db_min, a.db_max, a.out_h, a.out_w, a.pad_value = db_min, db_max, out_hw[0], out_hw[1], pad_value
    a.out, a.out_fp32 = out.data_ptr(), int(out.dtype == torch.float32)
    check(_lib.load().specyolo_iq_to_letterbox(C.byref(a), _lib.stream_ptr()))
    return out

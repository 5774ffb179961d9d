"""B200 versions of the reference's operator blocks — same class names, constructor signatures,
sub-module names and state_dict keys as ultralytics.nn.modules (so reference checkpoints / state
dicts load unchanged), but every forward runs libspecyolo kernels on NHWC bf16 feature maps.

torch.nn.Conv2d / BatchNorm2d instances are used purely as parameter containers (their forward is
never called): BatchNorm is folded into the conv and the weights are repacked for the tcgen05
implicit-GEMM kernel on first use (`Conv.packed()`), the device-side equivalent of
BaseModel.fuse() (ultralytics/nn/tasks.py:223-251).

Concats never materialise as copies inside a block: each block allocates its concat buffer once per
call and hands channel-slice views to the producing convs (`out=`).
"""
from __future__ import annotations

import math
import weakref
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops

BN_EPS = 1e-3  # initialize_weights (ultralytics/utils/torch_utils.py:410-420)


def autopad(k, p=None, d=1):
    """Same-shape padding (ultralytics/nn/modules/conv.py:56-62)."""
    if isinstance(k, (tuple, list)):
        k = k[0]
    if d > 1:
        k = d * (k - 1) + 1
    return k // 2 if p is None else p


def _as_fmap(x: torch.Tensor) -> torch.Tensor:
    """Accept what reference callers pass (NCHW fp32/bf16/uint8) and bring it to NHWC bf16."""
    if x.dtype == torch.bfloat16 and x.dim() == 4 and (x.stride(1) == 1 or x.shape[1] == 1):
        return x
    if x.dtype == torch.float16:        # the reference's half=True path hands fp16 tensors to the first block
        x = x.float()
    return ops.to_nhwc_bf16(x)


class Conv(nn.Module):
    """Conv2d + BatchNorm2d + SiLU (ultralytics/nn/modules/conv.py:65-83), run as one fused kernel."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        if isinstance(k, (tuple, list)):   # C2f passes k=((3, 3), (3, 3)); only square kernels exist on the path
            assert k[0] == k[1]
            k = k[0]
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=BN_EPS, momentum=0.03)
        if not (act is True or act is False or isinstance(act, (nn.SiLU, nn.Identity))):
            raise NotImplementedError("specyolo Conv supports SiLU or no activation")
        self.act = nn.SiLU() if (act is True or isinstance(act, nn.SiLU)) else nn.Identity()
        self._packed: Optional[ops.PackedConv] = None

    # -- weight prep ---------------------------------------------------------------------------
    def _bn_terms(self):
        """(conv bias or None, (gamma, beta, mean, var) or None, eps): `bn` is gone once the reference's BaseModel.fuse()
        (nn/tasks.py:223-251) has folded it into `conv` — the instance may be a reference Conv behind the shim."""
        bn = getattr(self, "bn", None)
        if bn is None:
            return self.conv.bias, None, 0.0
        return self.conv.bias, (bn.weight, bn.bias, bn.running_mean, bn.running_var), bn.eps

    def packed(self) -> ops.PackedConv:
        if getattr(self, "_packed", None) is None:
            c = self.conv
            cb, bn, eps = self._bn_terms()
            self._packed = ops.fold_pack(c.weight, cb, bn, eps, c.stride[0], c.padding[0], c.dilation[0], c.groups,
                                         isinstance(self.act, nn.SiLU))
        return self._packed

    def packed_from_blocked(self) -> ops.PackedConv:
        """This 3x3 / stride-2 conv as a 2x2 conv over a 2x2-blocked input (ops.pack_from_blocked)."""
        if getattr(self, "_packed_blocked", None) is None:
            cb, bn, eps = self._bn_terms()
            self._packed_blocked = ops.pack_from_blocked(self.conv.weight, cb, bn, eps, isinstance(self.act, nn.SiLU))
        return self._packed_blocked

    def folded_depthwise(self):
        """(w [9, C] fp32 tap-major, b [C] fp32): BN-folded weights of a depthwise 3x3 conv for ops.dwconv_pwconv
        (fuse_conv_and_bn, ultralytics/utils/torch_utils.py:238-265, on the device; one-off)."""
        if getattr(self, "_folded_dw", None) is None:
            c = self.conv
            cb, bn, eps = self._bn_terms()
            w = c.weight.detach().float()
            b = torch.zeros(c.out_channels, device=w.device) if cb is None else cb.detach().float()
            if bn is not None:
                scale = bn[0].detach().float() / torch.sqrt(bn[3].detach().float() + eps)
                w = w * scale.view(-1, 1, 1, 1)
                b = bn[1].detach().float() + (b - bn[2].detach().float()) * scale
            self._folded_dw = (w.view(c.out_channels, 9).t().contiguous(), b.contiguous())
        return self._folded_dw

    def is_depthwise3x3(self) -> bool:
        c = self.conv
        return c.groups == c.in_channels == c.out_channels and c.kernel_size == (3, 3) and c.stride == (1, 1) and \
            c.padding == (1, 1) and c.dilation == (1, 1) and isinstance(self.act, nn.SiLU)

    def invalidate(self):
        self._packed = None
        self._packed_blocked = None
        self._folded_dw = None

    def _apply(self, fn, *a, **k):
        self._packed = None  # parameters moved / cast: repack lazily
        self._packed_blocked = None
        self._folded_dw = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        self._packed_blocked = None
        self._folded_dw = None
        return super()._load_from_state_dict(*a, **k)

    # -- forward -------------------------------------------------------------------------------
    def is_stem(self) -> bool:
        c = self.conv
        return c.in_channels == 3 and c.groups == 1 and c.kernel_size == (3, 3) and c.stride == (2, 2) and \
            c.padding == (1, 1) and c.dilation == (1, 1)

    def takes_blocked(self) -> bool:
        """3x3 / s2 / p1 dense conv thin enough to be HBM-bound with K = 16 c: can read a 2x2-blocked input."""
        c = self.conv
        return c.groups == 1 and c.kernel_size == (3, 3) and c.stride == (2, 2) and c.padding == (1, 1) and \
            c.dilation == (1, 1) and c.in_channels % 4 == 0 and c.in_channels <= 32 and c.out_channels <= 256

    def forward(self, x, out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None):
        if self.conv.in_channels == 3 and self.conv.groups == 1 and x.dim() == 4 and x.stride(1) != 1:
            return ops.stem_conv(x, self.packed(), out)  # NCHW network input straight into the stem
        return ops.conv2d(_as_fmap(x), self.packed(), out, residual)

    forward_fuse = forward


class DWConv(Conv):
    """Depth-wise conv (conv.py:687-692)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class DDWConv(nn.Module):
    """Grouped (g=8) dilated strided conv followed by a 1x1 conv (conv.py:694-710)."""

    def __init__(self, c1, c2, k=3, s=2, d=1, act=True):
        super().__init__()
        self.conv1 = Conv(c1, c2, k, s, g=8, d=d, act=act)
        self.kz = k
        self.conv2 = Conv(c2, c2, k=1, s=1)

    def forward(self, x, out=None):
        return self.conv2(self.conv1(x), out=out)


def tensor_version(t: torch.Tensor) -> int:
    """In-place modification counter of a tensor, for cache keys; tensors created under torch.inference_mode (the
    reference fuses its models inside `smart_inference_mode`) do not track one — and cannot be modified in place either."""
    try:
        return t._version
    except RuntimeError:
        return -1


class SobelConv(nn.Module):
    """Parameter container of SobelConv (conv.py:1153-1182): three depthwise 3x3 convs (groups = out_channels, no bias)
    initialised to the Sobel-x, Sobel-x + Sobel-y and Sobel-y kernels; their outputs are summed."""

    def __init__(self, in_channels=1, out_channels=16):
        super().__init__()
        sx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])
        sy = torch.tensor([[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]])
        self.convs = nn.ModuleList()
        for kern in (sx, sx + sy, sy):
            conv = nn.Conv2d(in_channels, out_channels, 3, padding=1, groups=out_channels, bias=False)
            conv.weight = nn.Parameter(kern.view(1, 1, 3, 3).repeat(out_channels, 1, 1, 1))
            self.convs.append(conv)


class SobelSpatialAttention(nn.Module):
    """x * sigmoid(cv1(sobel(cat(mean_c x, max_c x)))) (conv.py:1184-1198): one statistics pass + one gate pass
    (csrc/spatial_gate.cu); the seven small linear weights are folded into one 2 x 3 x 3 stencil on the host."""

    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in {3, 7}
        self.sobel = SobelConv(in_channels=2, out_channels=2)
        self.cv1 = nn.Conv2d(2, 1, 1, padding=0, bias=False)
        self._w18 = None

    def stencil(self):
        ws = [c.weight for c in self.sobel.convs] + [self.cv1.weight]
        key = tuple((w.data_ptr(), tensor_version(w)) for w in ws)
        if getattr(self, "_w18", None) is None or self._w18[0] != key:
            k = sum(c.weight.detach().float() for c in self.sobel.convs).reshape(2, 9)      # [c][ky*3+kx]
            w = (self.cv1.weight.detach().float().reshape(2, 1) * k).reshape(18)
            self._w18 = (key, [float(v) for v in w.cpu()])
        return self._w18[1]

    def forward(self, x, out=None):
        return ops.sobel_spatial_attention(_as_fmap(x), self.stencil(), out)


class ConvHCA(nn.Module):
    """Conv(k, s) followed by SobelSpatialAttention (conv.py:829-844; *_convHCA configs, backbone layers 3 / 5 / 7)."""

    def __init__(self, c1, c2, k=3, s=2, d=1, act=True):
        super().__init__()
        self.kz, self.stride, self.dilation = k, s, d
        self.conv2 = Conv(c1, c2, k=k, s=s)
        self.hca = SobelSpatialAttention(7)

    def forward(self, x, out=None):
        return self.hca(self.conv2(x, out=out))        # the gate runs in place on the conv's output window


class Bottleneck(nn.Module):
    """Two convs with an optional shortcut added in the second conv's epilogue (block.py:713-726)."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        x = _as_fmap(x)
        p1, p2 = self.cv1.packed(), self.cv2.packed()
        if ops.bottleneck_ok(x, p1, p2, prefer=True):       # thin 3x3 pairs: one kernel, the intermediate never leaves the SM
            return ops.bottleneck(x, p1, p2, self.add, out)
        return self.cv2(self.cv1(x), out=out, residual=x if self.add else None)


class C3(nn.Module):
    """CSP bottleneck with 3 convs (block.py:490-504); concat buffer written in place."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=((1, 1), (3, 3)), e=1.0) for _ in range(n)))
        self.c_ = c_

    def forward(self, x, out=None):
        x = _as_fmap(x)
        B, _, H, W = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(B, 2 * c_, H, W, x.device)
        t = self.cv1(x)
        last = len(self.m) - 1
        for j, m in enumerate(self.m):
            t = m(t, out=cat[:, : c_] if j == last else None)
        if last < 0:
            cat[:, : c_].copy_(t)
        self.cv2(x, out=cat[:, c_:])
        return self.cv3(cat, out=out)


class MSCSpatialAttention(nn.Module):
    """x * s * g + x with s = relu(conv31(mm)) + relu(conv3(mm)) over the channel mean / max planes and
    g = relu(fc(mean_hw(x * s))) (conv.py:1200-1243; x8 == x9 there): csrc/spatial_gate.cu, four launches."""

    def __init__(self, c1, kernel_size=7):
        super().__init__()
        self.cv1 = nn.Sequential(nn.Conv2d(2, 1, (31, 31), padding=(15, 15), bias=False), nn.ReLU())
        self.cv2 = nn.Sequential(nn.Conv2d(2, 1, (3, 3), padding=(1, 1), bias=False), nn.ReLU())
        self.fc = nn.Conv2d(c1, c1, 1, 1, 0, bias=True)
        self._f32 = None

    def weights_f32(self):
        ws = (self.cv1[0].weight, self.cv2[0].weight, self.fc.weight, self.fc.bias)
        key = tuple((w.data_ptr(), tensor_version(w)) for w in ws)
        if getattr(self, "_f32", None) is None or self._f32[0] != key:
            self._f32 = (key, tuple(w.detach().float().contiguous() for w in ws))
        return self._f32[1]

    def forward(self, x, out=None):
        return ops.msc_spatial_attention(_as_fmap(x), *self.weights_f32(), out=out)


class C3x(C3):
    """C3 whose inner block is ONE MSCSpatialAttention (block.py:522-529; *_OMN configs, head layer 21)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.c_ = int(c2 * e)
        self.m = MSCSpatialAttention(self.c_)

    def forward(self, x, out=None):
        x = _as_fmap(x)
        B, _, H, W = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(B, 2 * c_, H, W, x.device)
        self.m(self.cv1(x), out=cat[:, : c_])
        self.cv2(x, out=cat[:, c_:])
        return self.cv3(cat, out=out)


class C3k(C3):
    """C3 with k x k bottlenecks (block.py:1672-1680)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))


class C2f(nn.Module):
    """CSP bottleneck with 2 convs and n inner blocks (block.py:444-464)."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))
        self.n = n

    def forward(self, x, out=None):
        x = _as_fmap(x)
        B, _, H, W = x.shape
        c = self.c
        cat = ops.new_act(B, (2 + len(self.m)) * c, H, W, x.device)
        self.cv1(x, out=cat[:, : 2 * c])           # the two chunks of cv1 are the first 2c channels
        for j, m in enumerate(self.m):
            m(cat[:, (1 + j) * c: (2 + j) * c], out=cat[:, (2 + j) * c: (3 + j) * c])
        return self.cv2(cat, out=out)


class C3k2(C2f):
    """C2f whose inner blocks are C3k or Bottleneck (block.py:1659-1671)."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g) for _ in range(n))


class FGM(nn.Module):
    """Parameter container of FGM (block.py:838-861): out = |ifft2(dwconv1(x) * fft2(dwconv2(x)))| * alpha + x * beta.
    `conv` is declared by the reference and never used; it is kept for state_dict compatibility."""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim * 2, 3, 1, 1)
        self.dwconv1 = nn.Conv2d(dim, dim, 1, 1, groups=1)
        self.dwconv2 = nn.Conv2d(dim, dim, 1, 1, groups=1)
        self.alpha = nn.Parameter(torch.zeros(dim, 1, 1))
        self.beta = nn.Parameter(torch.ones(dim, 1, 1))


class BottleNect(nn.Module):
    """The block.py BottleNect (block.py:782-836, the one C3k2GC instantiates): 1x1 conv + GELU, the two pooled channel
    gates around an FFT identity, FGM, ReLU — csrc/bottlenect.cu.  `out_conv` and `dw_11` are declared by the reference and
    never used; kept for state_dict compatibility."""

    def __init__(self, dim):
        super().__init__()
        self.in_conv = nn.Sequential(nn.Conv2d(dim, dim, kernel_size=1, padding=0, stride=1), nn.GELU())
        self.out_conv = nn.Conv2d(dim, dim, kernel_size=1, padding=0, stride=1)
        self.dw_11 = nn.Conv2d(dim, dim, kernel_size=3, padding=1, stride=1, groups=dim)
        self.act = nn.ReLU()
        self.conv = nn.Conv2d(dim, dim, kernel_size=1, padding=0, stride=1, groups=1, bias=True)
        self.fac_conv = nn.Conv2d(dim, dim, kernel_size=1, padding=0, stride=1, groups=1, bias=True)
        self.fgm = FGM(dim)
        self._f32 = None

    def weights_f32(self):
        ws = (self.in_conv[0].weight, self.in_conv[0].bias, self.fac_conv.weight, self.fac_conv.bias, self.conv.weight,
              self.conv.bias, self.fgm.dwconv1.weight, self.fgm.dwconv1.bias, self.fgm.dwconv2.weight, self.fgm.dwconv2.bias,
              self.fgm.alpha, self.fgm.beta)
        key = tuple((w.data_ptr(), tensor_version(w)) for w in ws)
        if getattr(self, "_f32", None) is None or self._f32[0] != key:
            self._f32 = (key, tuple(w.detach().float().contiguous() for w in ws))
        return self._f32[1]

    def forward(self, x, out=None):
        return ops.bottlenect(_as_fmap(x), self.weights_f32(), out=out)


class C3k2GC(C2f):
    """C2f whose inner blocks are BottleNect (block.py:1706-1714; *_GC configs, backbone layer 2).  The c3k=True branch
    (C3kGC, scales m / l / x) is not built."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut, g, e)
        if c3k:
            raise NotImplementedError("C3k2GC(c3k=True) (C3kGC, scales m/l/x) is outside the hot path this package implements")
        self.m = nn.ModuleList(BottleNect(self.c) for _ in range(n))


class SPPF(nn.Module):
    """cv1 -> 3 chained 5x5 max-pools -> cat -> cv2 (block.py:179-198)."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        if k != 5:
            raise NotImplementedError("SPPF kernel is specialised for k=5")
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.c_ = c_

    def forward(self, x, out=None):
        x = _as_fmap(x)
        B, _, H, W = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(B, 4 * c_, H, W, x.device)
        self.cv1(x, out=cat[:, : c_])
        ops.sppf_pool(cat, c_)
        return self.cv2(cat, out=out)


class Attention(nn.Module):
    """Multi-head self-attention over the H*W tokens + depthwise positional conv (block.py:1878-1933)."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)
        self._pe_f32 = None

    def _pe_weights(self):
        cur = getattr(self, "_pe_f32", None)
        if cur is None or cur[0].device != self.pe.conv.weight.device:
            w9, b = self.pe.folded_depthwise()               # [9, C] tap-major, BN folded (or already fused)
            self._pe_f32 = (w9.t().contiguous(), b)
            self.pe._folded_dw = None
        return self._pe_f32

    def _apply(self, fn, *a, **k):
        self._pe_f32 = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._pe_f32 = None
        return super()._load_from_state_dict(*a, **k)

    def forward(self, x, out=None, residual=None):
        qkv = self.qkv(x)
        pe_w, pe_b = self._pe_weights()
        y = ops.psa_attention(qkv, self.num_heads, self.key_dim, self.head_dim, self.scale, pe_w, pe_b)
        return self.proj(y, out=out, residual=residual)


class PSABlock(nn.Module):
    """x + attn(x); x + ffn(x) (block.py:1973-2007); both residual adds live in conv epilogues."""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut

    def forward(self, x, out=None):
        x = self.attn(x, residual=x if self.add else None)
        return self.ffn[1](self.ffn[0](x), out=out, residual=x if self.add else None)


class C2PSA(nn.Module):
    """cv1 -> split(a, b) -> b = PSABlocks(b) -> cv2(cat(a, b)) (block.py:2100-2139)."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x, out=None):
        x = _as_fmap(x)
        B, _, H, W = x.shape
        cat = ops.new_act(B, 2 * self.c, H, W, x.device)
        self.cv1(x, out=cat)
        b = cat[:, self.c:]
        last = len(self.m) - 1
        for j, m in enumerate(self.m):
            b = m(b, out=cat[:, self.c:] if j == last else None)   # last block overwrites b in place
        return self.cv2(cat, out=out)


class Concat(nn.Module):
    """torch.cat along channels (conv.py:1810-1820) — used by the stock YOLO11 necks."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, xs: Sequence[torch.Tensor]):
        """Fallback form (inputs copied): DetectionModel._run_trunk normally plans the buffer ahead and lets the producers
        write their channel slices in place (tasks.py `_concat_plan`), so nothing is copied."""
        B, _, H, W = xs[0].shape
        dev = (xs[0].src if isinstance(xs[0], UpsampledView) else xs[0]).device
        out = ops.new_act(B, sum(x.shape[1] for x in xs), H, W, dev)
        o = 0
        for x in xs:
            c = x.shape[1]
            if isinstance(x, UpsampledView):
                x.materialise(out=out[:, o: o + c])
            else:
                out[:, o: o + c].copy_(_as_fmap(x))   # plumbing copy (torch)
            o += c
        return out


class Upsample2x(nn.Module):
    """nn.Upsample(None, 2, 'nearest').  In the Spectrogram cfg the only consumer is Fusion, which reads
    through the upsample by index math, so this module just tags its input (no kernel, no bytes)."""

    def __init__(self, size=None, scale_factor=2, mode="nearest"):
        super().__init__()
        if size is not None or scale_factor != 2 or mode != "nearest":
            raise NotImplementedError("only x2 nearest upsampling is on the hot path")

    def forward(self, x):
        return UpsampledView(_as_fmap(x))


class UpsampledView:
    """Lazy x2 nearest-neighbour upsample of an NHWC feature map."""

    def __init__(self, src: torch.Tensor):
        self.src = src

    @property
    def shape(self):
        B, C, H, W = self.src.shape
        return torch.Size((B, C, 2 * H, 2 * W))

    def materialise(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The upsampled map itself, optionally written into a channel slice of the consumer's concat buffer."""
        if self.src.shape[1] % 8 == 0:
            return ops.upsample2x(self.src, out)
        s = self.src                                   # odd channel counts: torch plumbing
        B, C, H, W = s.shape
        res = ops.new_act(B, C, 2 * H, 2 * W, s.device)
        v = res.permute(0, 2, 3, 1).view(B, H, 2, W, 2, C)
        v.copy_(s.permute(0, 2, 3, 1)[:, :, None, :, None, :].expand(B, H, 2, W, 2, C))
        if out is not None:
            out.copy_(res)
            return out
        return res


class GCT(nn.Module):
    """Parameter container for the gated channel transformation (conv.py:2284-2301)."""

    def __init__(self, num_channels, epsilon=1e-5, mode="l2", after_relu=False):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, num_channels, 1, 1))
        self.gamma = nn.Parameter(torch.zeros(1, num_channels, 1, 1))
        self.beta = nn.Parameter(torch.zeros(1, num_channels, 1, 1))
        self.epsilon = epsilon


class WeightedSpatialAttention(nn.Module):
    """Parameter container for the 2->1 spatial-attention conv (conv.py:1839-1852)."""

    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in {3, 7}
        self.cv1 = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)


class Fusion(nn.Module):
    """Fusion('ESChannel') (conv.py:1854-1937, 2113-2127): GCT gate over the concatenated inputs plus a
    shared spatial-attention gate per input, summed — two HBM-bound kernels, no concat, no upsample copy."""

    def __init__(self, inc_list, fusion="ESChannel", c1=128):
        super().__init__()
        if fusion != "ESChannel":
            raise NotImplementedError("only Fusion('ESChannel') is reachable from the target configs")
        self.fusion = fusion
        self.sab = WeightedSpatialAttention(3)
        self.gsc2 = GCT(c1 * 2)
        self.gsc3 = GCT(c1 * 3)

    def forward(self, xs, out=None):
        srcs, ups = [], []
        for x in xs:
            if isinstance(x, UpsampledView):
                srcs.append(x.src)
                ups.append(1)
            else:
                srcs.append(_as_fmap(x))
                ups.append(0)
        g = self.gsc2 if len(xs) == 2 else self.gsc3
        return ops.fusion_eschannel(srcs, ups, g.alpha.detach().float().contiguous(), g.gamma.detach().float().contiguous(),
                                    g.beta.detach().float().contiguous(), g.epsilon,
                                    self.sab.cv1.weight.detach().float().contiguous(), out)


class DFL(nn.Module):
    """Parameter container kept for state_dict parity (block.py:65-83); the expectation over the 16 bins is
    computed inside the fused decode kernel."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1


class _HeadConv(nn.Conv2d):
    """Plain nn.Conv2d(c, n, 1) with bias ending each Detect branch; fp32 output into the decode buffer."""

    def packed(self):
        return head_packed(self)


# Weight packs of the plain nn.Conv2d tails live OUTSIDE the modules (weakly keyed by the module, validated against the
# identity of its parameters): the tails of a reference model bound in by the shim are stock torch modules, and nothing of
# this package may end up in their __dict__ — the reference's trainer pickles them into its checkpoints.
_TAIL_CACHE: "weakref.WeakKeyDictionary[nn.Module, dict]" = weakref.WeakKeyDictionary()


def _tail_entry(conv: nn.Conv2d) -> dict:
    w, b = conv.weight, conv.bias
    key = (w.data_ptr(), tensor_version(w), w.dtype, w.device, None if b is None else (b.data_ptr(), tensor_version(b)))
    e = _TAIL_CACHE.get(conv)
    if e is None or e["key"] != key:
        e = _TAIL_CACHE[conv] = {"key": key}
    return e


def head_packed(conv: nn.Conv2d) -> ops.PackedConv:
    """Packed weights of the plain nn.Conv2d(c, n, 1) (with bias) that ends a Detect branch; cached per module."""
    e = _tail_entry(conv)
    if "packed" not in e:
        e["packed"] = ops.fold_pack(conv.weight, conv.bias, None, 0.0, 1, 0, 1, 1, act=False)
    return e["packed"]


def head_f32(conv: nn.Conv2d):
    """(w [n, c] fp32, b [n] fp32) of a plain 1x1 nn.Conv2d for the fused class head of ops.dwconv_pwconv; cached."""
    e = _tail_entry(conv)
    if "f32" not in e:
        e["f32"] = (conv.weight.detach().float().reshape(conv.out_channels, conv.in_channels).contiguous(),
                    conv.bias.detach().float().contiguous())
    return e["f32"]


class Detect(nn.Module):
    """YOLO Detect head (ultralytics/nn/modules/head.py:21-172).

    forward(list of 3 maps) -> (y [B, 4+nc, A] fp32, [raw maps [B, no, h, w] fp32]) in eval mode, like
    the reference.  The last 1x1 conv of every branch writes fp32 logits straight into the
    [B, h*w, no_stride] buffer that the fused decode kernel reads.
    """

    dynamic = False
    export = False
    format = None
    end2end = False
    max_det = 300
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    legacy = False

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.legacy = bool(type(self).legacy)   # frozen per instance: a later model with another flag must not change this one
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), _HeadConv(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = (
            nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), _HeadConv(c3, self.nc, 1)) for x in ch)
            if self.legacy
            else nn.ModuleList(
                nn.Sequential(
                    nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                    nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                    _HeadConv(c3, self.nc, 1),
                )
                for x in ch
            )
        )
        self.dfl = DFL(self.reg_max)
        self.no_stride = (self.no + 3) // 4 * 4

    def bias_init(self):
        """head.py:133-144."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)

    def head_logits(self, xs: Sequence[torch.Tensor]):
        """Runs the cv2 / cv3 branches; returns per-level fp32 [B, h*w, no_stride] buffers and (h, w).
        The 2 x nl branches are independent chains of small convs: they run as parallel streams."""
        bufs, hw, views, fns = [], [], [], []
        xs = [_as_fmap(x) for x in xs]
        no_stride = (self.no + 3) // 4 * 4
        for x in xs:
            B, _, H, W = x.shape
            buf = torch.empty((B, H * W, no_stride), device=x.device, dtype=torch.float32)
            views.append(buf.view(B, H, W, no_stride).permute(0, 3, 1, 2))   # [B, no_stride, H, W], NHWC memory
            bufs.append(buf)
            hw.append((H, W))

        def box_branch(i):
            t = self.cv2[i][1](self.cv2[i][0](xs[i]))
            ops.conv2d(t, head_packed(self.cv2[i][2]), out=views[i][:, : 4 * self.reg_max], out_fp32=True)

        def dw_pw(pair, x):
            """Sequential(DWConv 3x3, Conv 1x1) (head.py:51-58): one fused kernel when the shapes allow it."""
            dw, pw = pair[0], pair[1]
            if dw.is_depthwise3x3() and ops.dwconv_pwconv_ok(x.shape[1], pw.packed()):
                w, b = dw.folded_depthwise()
                return ops.dwconv_pwconv(x, w, b, pw.packed())
            return pw(dw(x))

        def cls_branch(i):
            cls_view = views[i][:, 4 * self.reg_max: self.no]
            if self.legacy:
                t = self.cv3[i][1](self.cv3[i][0](xs[i]))
            else:
                t = dw_pw(self.cv3[i][0], xs[i])
                dw, pw, last = self.cv3[i][1][0], self.cv3[i][1][1], self.cv3[i][2]
                if self.nc <= ops.DWPW_HEAD_MAX_NC and dw.is_depthwise3x3() and ops.dwconv_pwconv_ok(t.shape[1], pw.packed()):
                    # few classes (the 5G / LTE detector has 2): the closing Conv2d(c3, nc, 1) rides in the epilogue of the
                    # second DWConv + Conv pair; its c3-channel output is never written
                    w, b = dw.folded_depthwise()
                    ops.dwconv_pwconv(t, w, b, pw.packed(), head=head_f32(last) + (cls_view,))
                    return
                t = dw_pw(self.cv3[i][1], t)
            ops.conv2d(t, head_packed(self.cv3[i][2]), out=cls_view, out_fp32=True)

        for i in range(len(xs)):
            fns.append(lambda i=i: box_branch(i))
            fns.append(lambda i=i: cls_branch(i))
        ops.run_concurrently(fns)
        return bufs, hw

    def forward(self, xs: List[torch.Tensor]):
        bufs, hw = self.head_logits(xs)
        y, _, _ = ops.detect_decode(bufs, hw, [float(s) for s in self.stride], self.nc, want_dense=True)
        raw = [b.view(b.shape[0], h, w, b.shape[2])[..., : self.no].permute(0, 3, 1, 2) for b, (h, w) in zip(bufs, hw)]
        return y if self.export else (y, raw)

"""TEST INFRASTRUCTURE — CPU restatement (PyTorch fp32, functional) of the reference's detector forward.

This is the oracle for the floating-point part of the hot path: every function cites the reference
file:line it restates (paths relative to the upstream repo, a fork of Ultralytics 8.3.70).  It is
pure function of (model yaml dict, state_dict, input) — it owns no weights and imports nothing from
the product package.  Pinned against the real reference by the fixtures in tests/golden/ that oracle/gen_golden.py produced from
the real reference (checked by tests/test_oracle_golden.py), and — where the vendored copy oracle/_ref exists —
directly by tests/test_reference_live.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path never does.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # initialize_weights sets BatchNorm2d.eps = 1e-3 (ultralytics/utils/torch_utils.py:410-420)


def make_divisible(x, divisor):
    """ultralytics/utils/ops.py:130-143."""
    return math.ceil(x / divisor) * divisor


def autopad(k, p=None, d=1):
    """ultralytics/nn/modules/conv.py:56-62."""
    if d > 1:
        k = d * (k - 1) + 1
    return k // 2 if p is None else p


# ------------------------------------------------------------------------------------------------
# graph description: a restatement of parse_model (ultralytics/nn/tasks.py:963-1168) restricted to the
# modules of the three target configs.  Returns a list of dicts {i, f, type, args..., prefix}.
# ------------------------------------------------------------------------------------------------
def guess_scale(path: str) -> str:
    """ultralytics/nn/tasks.py:1187-1202: the scale letter sits after 'yolo11' / 'yolov8' in the file name."""
    m = re.search(r"yolo[v]?\d+([nslmx])", str(path).rsplit("/", 1)[-1])
    return m.group(1) if m else ""


def parse_graph(d: dict, scale: str, nc: int | None = None, ch: int = 3) -> List[dict]:
    nc = nc if nc is not None else d["nc"]
    depth, width, max_ch = d["scales"][scale] if scale else (1.0, 1.0, float("inf"))
    chans: List[int] = []
    layers: List[dict] = []
    cin_first = ch
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = [nc if a == "nc" else a for a in args]
        n = max(round(n * depth), 1) if n > 1 else n  # tasks.py:1085
        c1 = (cin_first if i == 0 else chans[f]) if isinstance(f, int) else None
        L = dict(i=i, f=f, type=m, prefix=f"model.{i}")
        if m in ("Conv", "ConvHCA", "C3k2", "C3k2GC", "C3x", "SPPF", "C2PSA", "DDWConv"):
            c2 = make_divisible(min(args[0], max_ch) * width, 8)  # tasks.py:1088-1089
            rest = list(args[1:])
            if m == "Conv":
                k = rest[0] if len(rest) > 0 else 1
                s = rest[1] if len(rest) > 1 else 1
                L.update(c1=c1, c2=c2, k=k, s=s)
            elif m == "ConvHCA":  # conv.py:830: k=3, s=2 defaults
                L.update(c1=c1, c2=c2, k=rest[0] if len(rest) > 0 else 3, s=rest[1] if len(rest) > 1 else 2)
            elif m == "C3k2":
                c3k = rest[0] if len(rest) > 0 else False
                e = rest[1] if len(rest) > 1 else 0.5
                if scale in "mlx":
                    c3k = True  # tasks.py:1098-1101
                L.update(c1=c1, c2=c2, n=n, c3k=c3k, e=e)
            elif m == "C3k2GC":  # block.py:1706-1714; tasks.py:1110-1112 forces c3k for m/l/x (C3kGC: not restated)
                c3k = (rest[0] if len(rest) > 0 else False) or scale in "mlx"
                if c3k:
                    raise NotImplementedError("oracle does not restate C3kGC")
                L.update(c1=c1, c2=c2, n=n, e=rest[1] if len(rest) > 1 else 0.5)
            elif m == "C3x":  # block.py:522-529: C3(c1, c2, n, shortcut, g, e=0.5) with m = MSCSpatialAttention(c_)
                L.update(c1=c1, c2=c2, n=n)
            elif m == "SPPF":
                L.update(c1=c1, c2=c2, k=rest[0] if rest else 5)
            elif m == "C2PSA":
                L.update(c1=c1, c2=c2, n=n, e=0.5)
            elif m == "DDWConv":
                k = rest[0] if len(rest) > 0 else 3
                s = rest[1] if len(rest) > 1 else 2
                dd = rest[2] if len(rest) > 2 else 1
                L.update(c1=c1, c2=c2, k=k, s=s, d=dd)
        elif m == "nn.Upsample":
            c2 = c1
            L.update(scale=args[1], mode=args[2])
        elif m == "Concat":
            c2 = sum(chans[x] for x in f)
        elif m == "Fusion":
            c2 = chans[f[0]]  # tasks.py:1132-1135: every Fusion becomes 'ESChannel', c1 stays 128
            L.update(k=len(f), c=128)
        elif m == "Detect":
            c2 = None
            L.update(nc=args[0], ch=[chans[x] for x in f])
        else:
            raise NotImplementedError(f"oracle does not restate module {m}")
        chans.append(c2)
        layers.append(L)
    return layers


# ------------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------------
class Ref:
    """Functional forward over a state_dict."""

    def __init__(self, sd: Dict[str, torch.Tensor], fuse: bool = False, device=None, dtype=torch.float32,
                 bf16_storage: bool = False):
        """bf16_storage: emulate a bf16 pipeline with fp32 accumulation on the CPU — BN-folded weights rounded to bf16,
        every Conv output rounded to bf16 (what any bf16 implementation stores between layers), all arithmetic fp32.
        Used by the parity tests to measure the error FLOOR of bf16 storage against the fp32 reference, i.e. how far a
        numerically ideal bf16 implementation sits from fp32 on the same weights."""
        self.sd = {k: v.detach().to(device=device, dtype=dtype) for k, v in sd.items() if v.is_floating_point()}
        self.fuse = fuse or bf16_storage
        self.bf16_storage = bf16_storage
        self._folded: Dict[str, tuple] = {}

    def folded(self, p):
        """fuse_conv_and_bn (ultralytics/utils/torch_utils.py:238-265): what the predict path runs after
        AutoBackend's model.fuse() (nn/autobackend.py:152) — W' = diag(gamma / sqrt(var + eps)) W, b' = beta - mean * scale."""
        t = self._folded.get(p)
        if t is None:
            sd = self.sd
            scale = sd[p + ".bn.weight"].float() / torch.sqrt(sd[p + ".bn.running_var"].float() + BN_EPS)
            w = (sd[p + ".conv.weight"].float() * scale.view(-1, 1, 1, 1)).to(sd[p + ".conv.weight"].dtype)
            b = (sd[p + ".bn.bias"].float() - sd[p + ".bn.running_mean"].float() * scale).to(w.dtype)
            if self.bf16_storage:
                w = w.to(torch.bfloat16).to(w.dtype)
            t = self._folded[p] = (w, b)
        return t

    # Conv = Conv2d(bias=False) + BatchNorm2d(eval) + SiLU  (conv.py:65-79); forward_fuse (conv.py:81-83) when fused
    def conv(self, x, p, k=1, s=1, g=1, d=1, act=True):
        sd = self.sd
        if self.fuse:
            w, b = self.folded(p)
            if self.bf16_storage:
                x = x.to(torch.bfloat16).to(x.dtype)
            y = F.conv2d(x, w, b, s, autopad(k, None, d), d, g)
            y = F.silu(y) if act else y
            return y.to(torch.bfloat16).to(y.dtype) if self.bf16_storage else y
        w = sd[p + ".conv.weight"]
        y = F.conv2d(x, w, None, s, autopad(k, None, d), d, g)
        y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                         sd[p + ".bn.bias"], False, 0.0, BN_EPS)
        return F.silu(y) if act else y

    # Bottleneck (block.py:713-726)
    def bottleneck(self, x, p, k=(3, 3), shortcut=True):
        y = self.conv(self.conv(x, p + ".cv1", k[0]), p + ".cv2", k[1])
        c1 = x.shape[1]
        return x + y if (shortcut and c1 == y.shape[1]) else y

    # C3k = C3 with n Bottleneck(k=(3,3), e=1.0) (block.py:490-504, 1672-1680)
    def c3k(self, x, p, n=2):
        a = self.conv(x, p + ".cv1")
        for j in range(n):
            a = self.bottleneck(a, f"{p}.m.{j}", (3, 3))
        return self.conv(torch.cat((a, self.conv(x, p + ".cv2")), 1), p + ".cv3")

    # C3k2 / C2f (block.py:444-464, 1659-1671)
    def c3k2(self, x, p, n, c3k):
        y = list(self.conv(x, p + ".cv1").chunk(2, 1))
        for j in range(n):
            y.append(self.c3k(y[-1], f"{p}.m.{j}") if c3k else self.bottleneck(y[-1], f"{p}.m.{j}", (3, 3)))
        return self.conv(torch.cat(y, 1), p + ".cv2")

    # SPPF (block.py:179-198)
    def sppf(self, x, p, k=5):
        y = [self.conv(x, p + ".cv1")]
        for _ in range(3):
            y.append(F.max_pool2d(y[-1], k, 1, k // 2))
        return self.conv(torch.cat(y, 1), p + ".cv2")

    # Attention (block.py:1896-1933)
    def attention(self, x, p, num_heads):
        B, C, H, W = x.shape
        N = H * W
        head_dim = C // num_heads
        key_dim = int(head_dim * 0.5)
        scale = key_dim ** -0.5
        qkv = self.conv(x, p + ".qkv", act=False)
        q, k, v = qkv.reshape(B, num_heads, key_dim * 2 + head_dim, N).split([key_dim, key_dim, head_dim], dim=2)
        attn = (q.transpose(-2, -1) @ k) * scale
        attn = attn.softmax(dim=-1)
        y = (v @ attn.transpose(-2, -1)).reshape(B, C, H, W) + self.conv(v.reshape(B, C, H, W), p + ".pe", 3, 1, g=C, act=False)
        return self.conv(y, p + ".proj", act=False)

    # PSABlock (block.py:1995-2007), C2PSA (block.py:2125-2139)
    def c2psa(self, x, p, n):
        y = self.conv(x, p + ".cv1")
        c = y.shape[1] // 2
        a, b = y.split((c, c), 1)
        for j in range(n):
            q = f"{p}.m.{j}"
            b = b + self.attention(b, q + ".attn", c // 64)
            b = b + self.conv(self.conv(b, q + ".ffn.0"), q + ".ffn.1", act=False)
        return self.conv(torch.cat((a, b), 1), p + ".cv2")

    # DDWConv (conv.py:694-710): Conv(k, s, g=8, d) then Conv 1x1
    def ddwconv(self, x, p, k, s, d):
        return self.conv(self.conv(x, p + ".conv1", k, s, g=8, d=d), p + ".conv2")

    # ConvHCA (conv.py:829-844) = Conv then SobelSpatialAttention (conv.py:1184-1198) over SobelConv (conv.py:1153-1182)
    def convhca(self, x, p, k, s):
        sd = self.sd
        x1 = self.conv(x, p + ".conv2", k, s)
        m = torch.cat([x1.mean(1, keepdim=True), x1.max(1, keepdim=True)[0]], 1)
        e = sum(F.conv2d(m, sd[f"{p}.hca.sobel.convs.{j}.weight"], None, 1, 1, 1, 2) for j in range(3))
        return x1 * torch.sigmoid(F.conv2d(e, sd[p + ".hca.cv1.weight"]))

    # BottleNect (block.py:782-836) with FGM (block.py:838-861); torch.fft on the CPU where the reference calls cuFFT
    def bottlenect(self, x, p):
        sd = self.sd
        out = F.gelu(F.conv2d(x, sd[p + ".in_conv.0.weight"], sd[p + ".in_conv.0.bias"]))
        x_att = F.conv2d(out.mean((2, 3), keepdim=True), sd[p + ".fac_conv.weight"], sd[p + ".fac_conv.bias"])
        x_fca = torch.abs(torch.fft.ifft2(x_att * torch.fft.fft2(out, norm="backward"), dim=(-2, -1), norm="backward"))
        x_att = F.conv2d(x_fca.mean((2, 3), keepdim=True), sd[p + ".conv.weight"], sd[p + ".conv.bias"])
        x_sca = x_att * x_fca
        x1 = F.conv2d(x_sca, sd[p + ".fgm.dwconv1.weight"], sd[p + ".fgm.dwconv1.bias"])
        x2 = F.conv2d(x_sca, sd[p + ".fgm.dwconv2.weight"], sd[p + ".fgm.dwconv2.bias"])
        o = torch.abs(torch.fft.ifft2(x1 * torch.fft.fft2(x2, norm="backward"), dim=(-2, -1), norm="backward"))
        return F.relu(o * sd[p + ".fgm.alpha"] + x_sca * sd[p + ".fgm.beta"])

    # C3k2GC (block.py:1706-1714) = C2f.forward (block.py:459-464) over BottleNect blocks
    def c3k2gc(self, x, p, n):
        y = list(self.conv(x, p + ".cv1").chunk(2, 1))
        for j in range(n):
            y.append(self.bottlenect(y[-1], f"{p}.m.{j}"))
        return self.conv(torch.cat(y, 1), p + ".cv2")

    # MSCSpatialAttention (conv.py:1200-1243); x8 and x9 are the same tensor there
    def msc(self, x, p):
        sd = self.sd
        x1 = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
        x2 = F.relu(F.conv2d(x1, sd[p + ".cv1.0.weight"], None, 1, 15))
        x3 = F.relu(F.conv2d(x1, sd[p + ".cv2.0.weight"], None, 1, 1))
        x4, x5 = x * x2, x * x3
        x8 = F.relu(F.conv2d((x4 + x5).mean((2, 3), keepdim=True), sd[p + ".fc.weight"], sd[p + ".fc.bias"]))
        return x4 * x8 + x5 * x8 + x

    # C3x (block.py:522-529) = C3.forward (block.py:501-504) with m = MSCSpatialAttention
    def c3x(self, x, p):
        return self.conv(torch.cat((self.msc(self.conv(x, p + ".cv1"), p + ".m"), self.conv(x, p + ".cv2")), 1), p + ".cv3")

    # GCT (conv.py:2284-2301)
    def gct(self, x, p, eps=1e-5):
        sd = self.sd
        emb = (x.pow(2).sum((2, 3), keepdim=True) + eps).pow(0.5) * sd[p + ".alpha"]
        norm = sd[p + ".gamma"] / (emb.pow(2).mean(dim=1, keepdim=True) + eps).pow(0.5)
        return x * (1.0 + torch.tanh(emb * norm + sd[p + ".beta"]))

    # WeightedSpatialAttention (conv.py:1839-1852)
    def sab(self, x, p):
        m = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
        return x * torch.sigmoid(F.conv2d(m, self.sd[p + ".cv1.weight"], None, 1, 1))

    # Fusion 'ESChannel' (conv.py:2113-2127)
    def fusion(self, xs: Sequence[torch.Tensor], p):
        a_b = self.gct(torch.cat(list(xs), 1), p + (".gsc2" if len(xs) == 2 else ".gsc3"))
        chunks = torch.chunk(a_b, len(xs), dim=1)
        return sum(ch + self.sab(xs[i], p + ".sab") for i, ch in enumerate(chunks))

    # Detect.forward (head.py:64-74) — raw per-level maps [B, 64+nc, h, w]
    def detect_raw(self, xs: Sequence[torch.Tensor], p, nc, legacy=False):
        sd = self.sd
        out = []
        for i, x in enumerate(xs):
            b = self.conv(self.conv(x, f"{p}.cv2.{i}.0", 3), f"{p}.cv2.{i}.1", 3)
            b = F.conv2d(b, sd[f"{p}.cv2.{i}.2.weight"], sd[f"{p}.cv2.{i}.2.bias"])
            if legacy:
                c = self.conv(self.conv(x, f"{p}.cv3.{i}.0", 3), f"{p}.cv3.{i}.1", 3)
            else:
                c = self.conv(x, f"{p}.cv3.{i}.0.0", 3, g=x.shape[1])
                c = self.conv(c, f"{p}.cv3.{i}.0.1")
                c = self.conv(c, f"{p}.cv3.{i}.1.0", 3, g=c.shape[1])
                c = self.conv(c, f"{p}.cv3.{i}.1.1")
            c = F.conv2d(c, sd[f"{p}.cv3.{i}.2.weight"], sd[f"{p}.cv3.{i}.2.bias"])
            out.append(torch.cat((b, c), 1))
        return out


# make_anchors (utils/tal.py:334-346)
def make_anchors(hw: Sequence[tuple], strides: Sequence[float], offset=0.5, device=None):
    pts, st = [], []
    for (h, w), s in zip(hw, strides):
        sx = torch.arange(w, dtype=torch.float32, device=device) + offset
        sy = torch.arange(h, dtype=torch.float32, device=device) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32, device=device))
    return torch.cat(pts), torch.cat(st)


# Detect._inference (head.py:100-131) + DFL (block.py:80-83) + dist2bbox (tal.py:349-358)
def detect_decode(raw: Sequence[torch.Tensor], strides: Sequence[float], nc: int, reg_max: int = 16):
    B = raw[0].shape[0]
    no = nc + 4 * reg_max
    x_cat = torch.cat([xi.reshape(B, no, -1) for xi in raw], 2).float()
    anchors, st = make_anchors([tuple(xi.shape[2:]) for xi in raw], strides, device=x_cat.device)
    anchors, st = anchors.transpose(0, 1), st.transpose(0, 1)
    box, cls = x_cat.split((reg_max * 4, nc), 1)
    b, _, a = box.shape
    proj = torch.arange(reg_max, dtype=torch.float32, device=x_cat.device).view(1, reg_max, 1, 1)
    dist = (box.view(b, 4, reg_max, a).transpose(2, 1).softmax(1) * proj).sum(1)  # DFL conv with weights 0..15
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


# ------------------------------------------------------------------------------------------------
# whole model: BaseModel._predict_once (tasks.py:161-188)
# ------------------------------------------------------------------------------------------------
def forward(graph: List[dict], sd, x: torch.Tensor, strides=(8.0, 16.0, 32.0),
            return_layers: bool = False, fuse: bool = False):
    """Returns (y [B,4+nc,A], raw list) like Detect in eval mode; optionally every layer output.
    `sd` is a state_dict or a prepared `Ref` (device / dtype / folded weights kept across calls); `fuse=True` runs
    Conv.forward_fuse on BN-folded weights as the reference's predict path does (autobackend.py:152)."""
    R = sd if isinstance(sd, Ref) else Ref(sd, fuse=fuse)
    x = x.to(next(iter(R.sd.values())).dtype)
    ys: List[torch.Tensor] = []
    legacy = not any(L["type"] == "C3k2" for L in graph)  # tasks.py:1097 sets legacy False for YOLO11
    for L in graph:
        f = L["f"]
        if isinstance(f, int):
            inp = x if f == -1 else ys[f]
        else:
            inp = [x if j == -1 else ys[j] for j in f]
        t, p = L["type"], L["prefix"]
        if t == "Conv":
            x = R.conv(inp, p, L["k"], L["s"])
        elif t == "ConvHCA":
            x = R.convhca(inp, p, L["k"], L["s"])
        elif t == "C3k2":
            x = R.c3k2(inp, p, L["n"], L["c3k"])
        elif t == "C3k2GC":
            x = R.c3k2gc(inp, p, L["n"])
        elif t == "C3x":
            x = R.c3x(inp, p)
        elif t == "SPPF":
            x = R.sppf(inp, p, L["k"])
        elif t == "C2PSA":
            x = R.c2psa(inp, p, L["n"])
        elif t == "DDWConv":
            x = R.ddwconv(inp, p, L["k"], L["s"], L["d"])
        elif t == "nn.Upsample":
            x = F.interpolate(inp, scale_factor=L["scale"], mode=L["mode"])
        elif t == "Concat":
            x = torch.cat(inp, 1)
        elif t == "Fusion":
            x = R.fusion(inp, p)
        elif t == "Detect":
            raw = R.detect_raw(inp, p, L["nc"], legacy)
            y = detect_decode(raw, strides, L["nc"])
            x = (y, raw)
        ys.append(x)
    if return_layers:
        return x, ys
    return x

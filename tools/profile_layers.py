"""Per-launch timing of one eager forward (CUDA events; a device-side sleep lets the host run ahead)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo import ops
from specyolo.nn.init import synth_images, synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sys.argv[2] if len(sys.argv) > 2 else "yolo11s_fusion_sand3_new.yaml"
nc = 2 if "fusion" in cfg else 80
imgsz = int(sys.argv[3]) if len(sys.argv) > 3 else 640
yolo = specyolo.YOLO(cfg, nc=nc)
yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda"); yolo.fuse()
x = synth_images(B, imgsz, seed=0, dtype=torch.uint8).cuda()
rec = []
orig = {}
ops.CONCURRENT = False
def wrap(name):
    f = getattr(ops, name); orig[name] = f
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record()
        desc, fl, by = name, 0.0, 0.0
        if name == "conv2d":
            xx, pc = a[0], a[1]
            Bq, Cin, H, W = xx.shape; Ho, Wo = pc.out_hw(H, W)
            fl = 2.0 * Bq * Ho * Wo * pc.cout * (pc.cin // pc.g_orig) * pc.k * pc.k
            by = 2.0 * (xx.numel() + r.numel() * (2 if r.dtype == torch.float32 else 1)) + 2.0 * pc.w.numel()
            if len(a) > 3 and a[3] is not None or k.get("residual") is not None: by += 2.0 * r.numel()
            desc = f"conv {Cin:4d}->{pc.cout:4d} k{pc.k} s{pc.s} d{pc.d} g{pc.g_orig:3d} {H:3d}x{W:3d} M={Bq*Ho*Wo:8d} K={(pc.cin//pc.g_orig)*pc.k*pc.k:5d}"
        if name == "dwconv_pwconv":
            xx, pw = a[0], a[3]
            Bq, Cc, H, W = xx.shape
            fl = 2.0 * Bq * H * W * Cc * (9 + pw.cout); by = 2.0 * (xx.numel() + r.numel())
            desc = f"dw3x3+pw {Cc:4d}->{pw.cout:4d} fused          {H:3d}x{W:3d} M={Bq*H*W:8d} K={Cc+9:5d}"
        if name == "stem_pair":
            xx, pc0, pc1 = a[0], a[1], a[2]
            Bq, _, H, W = xx.shape
            fl = 2.0 * Bq * ((H // 2) * (W // 2) * pc0.cout * 27 + (H // 4) * (W // 4) * pc1.cout * 9 * pc0.cout)
            by = xx.numel() + 2.0 * r.numel()
            desc = f"fused stem 3->{pc0.cout}->{pc1.cout} (u8 in)            {H:3d}x{W:3d}"
        rec.append((desc, e0, e1, fl, by)); return r
    setattr(ops, name, g)
for n in ("conv2d", "dwconv_pwconv", "stem_pair", "stem_space_to_depth", "sppf_pool", "fusion_eschannel", "psa_attention", "detect_decode", "nms"):
    wrap(n)
for _ in range(2):
    yolo.model.detect_fused(x)
rec.clear()
torch.cuda.synchronize(); torch.cuda._sleep(int(6e8))
yolo.model.detect_fused(x)
torch.cuda.synchronize()
tot = sum(e0.elapsed_time(e1) for _, e0, e1, _, _ in rec)
print(f"B={B} cfg={cfg} total {tot:.3f} ms over {len(rec)} calls")
rows = [(e0.elapsed_time(e1), d, fl, by) for d, e0, e1, fl, by in rec]
for i, (t, d, fl, by) in enumerate(rows):
    print(f"{i:3d} {t*1e3:8.1f} us  {d:70s} {fl/t/1e9 if fl else 0:7.1f} TF/s {by/t/1e6 if by else 0:7.0f} GB/s")

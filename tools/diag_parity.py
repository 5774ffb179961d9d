"""Diagnostic: anchor-level error of the CUDA path vs the fp32 oracle on the bench's input (uint8, 640^2)."""
import sys
from pathlib import Path
import numpy as np, torch, yaml
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from oracle import yolo_ref
from specyolo import ops
from specyolo.nn.init import synth_images, synth_state_dict
from specyolo.nn.modules import UpsampledView

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
sd = synth_state_dict(yolo.model, seed=0); yolo.load_state_dict(sd); yolo.to("cuda")
x = synth_images(64, 640, seed=0, dtype=torch.uint8)[:n]
g = yolo_ref.parse_graph(yaml.safe_load((ROOT / "spectrogram-yolov11_b200/specyolo/cfg/yolo11_fusion_sand3_new.yaml").read_text()), "s", 2)
with torch.no_grad():
    (y_ref, raw_ref), layers_ref = yolo_ref.forward(g, sd, x.float() / 255, return_layers=True)
for name, xin in (("uint8 (fused stem)", x.cuda()), ("float (layered stem)", (x.float() / 255).cuda())):
    y, raw = yolo.model(xin)
    y = y.cpu()
    sc_ref = y_ref[:, 4:].amax(1)
    m = sc_ref > 0.25
    eb = (y[:, :4] - y_ref[:, :4]).abs().amax(1)
    es = (y[:, 4:] - y_ref[:, 4:]).abs().amax(1)
    print(f"{name}: candidates {int(m.sum())}; box err on candidates: mean {eb[m].mean():.3f} p50 {eb[m].median():.3f} "
          f"p90 {eb[m].quantile(0.9):.3f} p99 {eb[m].quantile(0.99):.3f} max {eb[m].max():.3f}; score err mean {es[m].mean():.4f} "
          f"p99 {es[m].quantile(0.99):.4f} max {es[m].max():.4f}; all anchors box max {eb.max():.3f} score max {es.max():.4f}")
    lv = [6400, 1600, 400]; o = 0
    for L, s in zip(lv, (8, 16, 32)):
        mm = m[:, o:o + L]
        if mm.any():
            print(f"   stride {s}: cand {int(mm.sum())} box mean {eb[:, o:o+L][mm].mean():.3f} p99 {eb[:, o:o+L][mm].quantile(0.99):.3f} max {eb[:, o:o+L][mm].max():.3f}")
        o += L
    for i, (r, rr) in enumerate(zip(raw, raw_ref)):
        print(f"   raw{i} rel-L2 {((r.cpu() - rr).norm() / rr.norm()).item():.4f}")
# per-layer rel-L2, uint8 path replayed layer by layer (non-fused stem) for localisation
outs, t = [], (x.float() / 255).cuda()
for mod in yolo.model.model[:-1]:
    if mod.f != -1:
        t = outs[mod.f] if isinstance(mod.f, int) else [t if j == -1 else outs[j] for j in mod.f]
    t = mod(t); outs.append(t)
for i, (o, r) in enumerate(zip(outs, layers_ref)):
    if isinstance(o, UpsampledView): o = o.materialise()
    got = ops.to_nchw_f32(o).cpu()
    print(f"layer {i:2d} {yolo.model.model[i].type:10s} rel-L2 {((got - r).norm() / r.norm()).item():.4f}  max|ref| {r.abs().max():.2f}")

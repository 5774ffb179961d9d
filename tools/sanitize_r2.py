"""Round-2 kernels under compute-sanitizer memcheck: fused Bottleneck (both channel configs, partial / border tiles), resident-S
attention (ragged N), NMS bitmask path, separable SPPF pool (both CTA shapes), DWConv+1x1 with the fused class head, and one
small end-to-end predict (eager, no graph) of the Spectrogram cfg."""
import sys, math
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo import ops
from specyolo.nn.init import synth_images, synth_state_dict
g = torch.Generator().manual_seed(0)
def fmap(x): return x.cuda().to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
for (B, C, Cm, H, W) in ((2, 32, 16, 29, 17), (1, 64, 32, 15, 33), (2, 32, 32, 20, 20)):
    w1 = torch.randn((Cm, C, 3, 3), generator=g) * 0.1; w2 = torch.randn((C, Cm, 3, 3), generator=g) * 0.1
    p1 = ops.fold_pack(w1.cuda(), torch.zeros(Cm).cuda(), None, 0.0, 1, 1, 1, 1, True)
    p2 = ops.fold_pack(w2.cuda(), torch.zeros(C).cuda(), None, 0.0, 1, 1, 1, 1, True)
    buf = ops.new_act(B, 3 * C, H, W, "cuda").normal_()
    ops.bottleneck(buf[:, C:2 * C], p1, p2, True, out=buf[:, 2 * C:])
for (H, W, heads) in ((20, 20, 2), (5, 7, 1), (16, 32, 1)):
    qkv = torch.randn((1, heads * 128, H, W), generator=g)
    ops.psa_attention(fmap(qkv), heads, 32, 64, 32 ** -0.5, (torch.randn((heads * 64, 9), generator=g) * 0.2).cuda(),
                      torch.zeros(heads * 64).cuda())
for (c, H, W) in ((64, 20, 20), (32, 13, 27), (128, 40, 40)):
    buf = ops.new_act(1, 4 * c, H, W, "cuda").normal_(); ops.sppf_pool(buf, c)
for n in (50, 700, 1500):
    pred = torch.zeros((2, 6, n)); pred[:, :2] = torch.rand((2, 2, n), generator=g) * 300 + 20
    pred[:, 2:4] = torch.rand((2, 2, n), generator=g) * 80 + 8; pred[:, 4:] = torch.rand((2, 2, n), generator=g)
    ops.nms(prediction=pred.cuda().contiguous(), B=2, nc=2, A=n, conf_thres=0.25, iou_thres=0.6)
x = torch.randn((2, 128, 21, 13), generator=g)
pw = ops.fold_pack((torch.randn((128, 128, 1, 1), generator=g) * 0.1).cuda(), torch.zeros(128).cuda(), None, 0.0, 1, 0, 1, 1, True)
buf = torch.zeros((2, 21 * 13, 68), device="cuda")
view = buf.view(2, 21, 13, 68).permute(0, 3, 1, 2)[:, 64:66]
ops.dwconv_pwconv(fmap(x), (torch.randn((9, 128), generator=g) * 0.2).cuda(), torch.zeros(128).cuda(), pw,
                  head=((torch.randn((2, 128), generator=g) * 0.1).cuda(), torch.zeros(2).cuda(), view))
yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda")
r = yolo.predict(synth_images(2, 160, seed=1, dtype=torch.uint8).cuda(), conf=0.05, use_graph=False)
torch.cuda.synchronize()
print("ok", [len(a) for a in r])

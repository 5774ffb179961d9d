#!/bin/bash
# Sibling configs through the product API: graph / eager equality, cost per 64-image step next to the base config
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py -k "sibling" -x > gpurun_out/t_variants.log 2>&1; echo "variants exit $?"; tail -8 gpurun_out/t_variants.log
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_shim.py -k "variant" -x > gpurun_out/t_shim_variants.log 2>&1; echo "shim variants exit $?"; tail -8 gpurun_out/t_shim_variants.log
timeout 600 python - <<'PY' | tee gpurun_out/variants_step.txt
import sys, torch
sys.path.insert(0, "spectrogram-yolov11_b200")
import specyolo
from specyolo.nn.init import synth_images, synth_state_dict
for cfg in ("yolo11s_fusion_sand3_new.yaml", "yolo11s_fusion_sand3_new_convHCA.yaml", "yolo11s_fusion_sand3_new_OMN.yaml", "yolo11s_fusion_sand3_new_GC.yaml"):
    y = specyolo.YOLO(cfg, nc=2); y.load_state_dict(synth_state_dict(y.model, seed=0)); y.to("cuda")
    x = synth_images(64, 640, seed=0, dtype=torch.uint8).cuda()
    for _ in range(3): r = y.predict(x, conf=0.25, iou=0.7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): r = y.predict(x, conf=0.25, iou=0.7)
    e1.record(); torch.cuda.synchronize()
    print(cfg, f"{e0.elapsed_time(e1)/20:.3f} ms per 64 images (predict, sync API)", sum(len(q.boxes) for q in r), "detections")
PY

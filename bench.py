#!/usr/bin/env python
"""Headline benchmark: spectrogram images/s of the Spectrogram-YOLOv11 forward + NMS at 640^2, bf16, batch 64
per GPU (BASELINE.json configs[1]), N GPUs of one node, images sharded by GPU with no collective on the
data path (weak scaling).

    python bench.py --gpus N --steps K --warmup W                # product arm (CUDA, libspecyolo)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` = images/s with the uint8 input batch resident in HBM;
`e2e` = same metric through YOLO.predict() from pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` = the tcgen05 implicit-GEMM conv kernel (all its launches of one step) against the measured
bf16 peak; `stages` adds the same arithmetic for the HBM-bound kernels; `cpu_baseline` = the oracle port
(CPU restatement of the reference) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "spectrogram-yolov11_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

CFG = "yolo11s_fusion_sand3_new.yaml"
CFG_FILE = "yolo11_fusion_sand3_new.yaml"
NC = 2
IMGSZ = 640
CONF, IOU, MAX_DET = 0.25, 0.7, 300
METRIC = "spectrogram images/sec (YOLO11 fwd+NMS, 640^2, bf16)"


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.is_file():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        # samples taken inside the timed region (nvidia-smi reports ~50 ms late: allow that much slack at the end)
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.06)]
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's predict path
# ------------------------------------------------------------------------------------------------
def cpu_port_images_per_s(batch: int, repeats: int, warmup: int, seed: int = 0):
    import numpy as np
    import torch
    import yaml

    import specyolo
    from oracle import nms_ref, yolo_ref
    from specyolo.nn.init import synth_images, synth_state_dict

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = specyolo.DetectionModel(CFG, nc=NC)           # parameter shapes only (CPU tensors, never run)
    sd = synth_state_dict(m, seed=seed)
    graph = yolo_ref.parse_graph(yaml.safe_load((PKG / "specyolo" / "cfg" / CFG_FILE).read_text()), "s", NC)
    x = synth_images(batch, IMGSZ, seed=seed, dtype=torch.uint8)

    def step():
        with torch.no_grad():
            im = x.float() / 255                      # predictor.py:133-135
            y, _ = yolo_ref.forward(graph, sd, im)
        return nms_ref.non_max_suppression(y.numpy(), CONF, IOU, max_det=MAX_DET)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), threads, times


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    ips, threads, times = cpu_port_images_per_s(args.cpu_batch, args.steps, args.warmup)
    sample = f"{args.cpu_batch} synthetic 640^2 uint8 images per step, {args.steps} timed steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "spectrogram-yolov11-s nc=2 640^2 predict (fwd + NMS)", "cpu_batch": args.cpu_batch,
                   "path": "oracle port of the reference CPU predict path (PyTorch CPU fp32 + numpy NMS)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def stage_roofline(model, x_dev, peaks):
    """One eager (un-graphed) step with CUDA events around every libspecyolo call; a device-side sleep is
    queued first so the host runs ahead and the events see kernel time only."""
    import torch

    from specyolo import ops

    rec = []
    orig = {}
    ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)      # flop per byte

    def wrap(name, flops_fn=None, bytes_fn=None):
        f = getattr(ops, name)
        orig[name] = f

        def g(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*a, **k)
            e1.record()
            nm = name
            fl = flops_fn(*a, **k) if flops_fn else 0.0
            by = bytes_fn(r, *a, **k) if bytes_fn else 0.0
            if name in ("conv2d", "dwconv_pwconv", "stem_pair"):     # class by arithmetic intensity against the ridge of the measured peaks
                nm = "conv2d_tensor_bound" if fl / max(by, 1.0) >= ridge else "conv2d_hbm_bound"
            rec.append((nm, e0, e1, fl, by))
            return r
        setattr(ops, name, g)

    def conv_flops(x, pc, *a, **k):
        B, _, H, W = x.shape
        Ho, Wo = pc.out_hw(H, W)
        kk = pc.alg_k or (pc.cin // pc.g_orig) * pc.k * pc.k                       # algorithmic (source groups, no padding)
        return 2.0 * B * Ho * Wo * pc.cout * kk

    def conv_bytes(r, x, pc, *a, **k):
        B, Cin, H, W = x.shape
        return 2.0 * (B * Cin * H * W + r.numel()) + 2.0 * pc.w.numel()

    def dwpw_flops(x, dw_w, dw_b, pw, *a, **k):
        B, Cc, H, W = x.shape
        return 2.0 * B * H * W * Cc * (9 + pw.cout)           # depthwise 3x3 + pointwise 1x1

    def stem_pair_flops(x, pc0, pc1b, *a, **k):      # layer 0 (K = 27) at H/2 x W/2 + layer 1 (K = 9 c0) at H/4 x W/4
        B, _, H, W = x.shape
        return 2.0 * B * ((H // 2) * (W // 2) * pc0.cout * 27 + (H // 4) * (W // 4) * pc1b.cout * 9 * pc0.cout)

    wrap("conv2d", conv_flops, conv_bytes)
    wrap("stem_pair", stem_pair_flops, lambda r, x, pc0, pc1b, *a, **k: x.numel() * x.element_size() + 2.0 * r.numel() + 2.0 * pc1b.w.numel())
    wrap("dwconv_pwconv", dwpw_flops, lambda r, x, *a, **k: 2.0 * (x.numel() + r.numel()))
    wrap("stem_space_to_depth", None, lambda r, x, *a, **k: x.numel() * x.element_size() + 2.0 * r.numel())
    wrap("sppf_pool", None, lambda r, buf, c: 2.0 * buf.numel())
    wrap("fusion_eschannel", None, lambda r, xs, *a, **k: 2.0 * (2 * sum(t.numel() for t in xs) + r.numel()))
    wrap("psa_attention", None, lambda r, qkv, *a, **k: 2.0 * (qkv.numel() + r.numel()))
    wrap("detect_decode", None, lambda r, logits, *a, **k: 4.0 * sum(t.numel() for t in logits))
    wrap("nms", None, None)
    # modules bind `ops.<fn>` at call time through the module attribute, so patching ops is enough
    ops.CONCURRENT = False          # time every kernel alone (the graph runs independent branches concurrently)
    reps = 5
    try:
        for _ in range(2):          # warm the eager path (caching allocator: a cudaMalloc would serialise the host)
            model.detect_fused(x_dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET)
        rec.clear()
        for _ in range(reps):
            torch.cuda.synchronize()
            torch.cuda._sleep(int(4e8))
            model.detect_fused(x_dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET)
        torch.cuda.synchronize()
    finally:
        ops.CONCURRENT = True
        for n, f in orig.items():
            setattr(ops, n, f)
    agg = {}
    per = len(rec) // reps                           # every pass issues the same launch sequence
    for i in range(per):                             # per launch: MEDIAN over the passes (an eager pass now and then
        name, _, _, fl, by = rec[i]                  # catches one launch behind a host hiccup; the mean let a single
        ts = sorted(rec[i + r * per][1].elapsed_time(rec[i + r * per][2]) for r in range(reps))   # outlier move the group)
        a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
        a[0] += ts[reps // 2] * 1e-3
        a[1] += fl
        a[2] += by
        a[3] += 1
    total = sum(a[0] for a in agg.values())
    stages = {}
    for name, (t, fl, by, n) in agg.items():
        s = {"launches": n, "ms": 1e3 * t, "share": t / total if total else None}
        if fl:
            s.update(tflops=fl / t / 1e12, frac_tensor=fl / t / 1e12 / peaks["tf_sust"])
        if by:
            s.update(gbs=by / t / 1e9, frac_hbm=by / t / 1e9 / peaks["hbm"])
        stages[name] = s
    c = [sum(agg[n][j] for n in agg if n.startswith("conv2d")) for j in range(4)]
    traffic = None
    tf = ROOT / "profiles" / "conv_dram_traffic.json"      # written from an ncu capture by tools/ncu_traffic.py
    if tf.is_file():
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_step")
    # 61 of the 86 conv launches (54 % of the conv time) sit below the ridge point of the measured peaks, i.e. are
    # HBM-bound at bf16: the headline roofline of the group is therefore the bandwidth one; the tensor-pipe view of the
    # same launches is reported beside it, and `stages` splits the launches at the ridge.
    gbs = c[2] / c[0] / 1e9
    tfs = c[1] / c[0] / 1e12
    roof = {"bound": "hbm", "kernel": "conv_igemm_kernel + conv_halo_kernel + dwpw_kernel + stem_pair_kernel: every tcgen05 implicit-GEMM conv launch of one step "
                                       "(most layers are below the ridge point, i.e. HBM-bound; stages.conv2d_tensor_bound / conv2d_hbm_bound split them)",
            "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
            "traffic": traffic, "algorithmic_bytes_per_step": c[2], "peak_source": peaks["src"] + " (HBM copy bandwidth; bf16 sustained for the tensor view)",
            "flops_per_step": c[1], "ms_per_step": 1e3 * c[0], "launches_per_step": c[3],
            "tflops_same_launches": tfs, "frac_tensor_same_launches": tfs / peaks["tf_sust"], "tensor_peak": peaks["tf_sust"]}
    return roof, stages


def run_product_arm(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    B = args.batch

    yolo = specyolo.YOLO(CFG, nc=NC)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to(dev)
    yolo.fuse()
    x_host = synth_images(B, IMGSZ, seed=rank, dtype=torch.uint8).pin_memory()   # what predict() uploads (uint8)
    x_dev = x_host.to(dev)
    pred_args = dict(conf=CONF, iou=IOU, max_det=MAX_DET)

    # ---- resident-input throughput: graph replay, inputs already in HBM ----
    predictor = specyolo.DetectionPredictor(yolo.model, pred_args)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()             # started early: nvidia-smi needs ~0.2 s before its first sample
    # several batches in flight (DetectionPredictor.infer_pipelined, --inflight): every step is a full forward + decode + NMS of B
    # images; consecutive steps overlap on the GPU the way consecutive batches of predict(stream=True) do
    predictor.infer_pipelined(x_dev, max(args.warmup, 3), args.inflight)
    torch.cuda.synchronize()
    launches_per_step = predictor.last_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = predictor.infer_pipelined(x_dev, args.steps, args.inflight)
    out, cnt = last[(args.steps - 1) % args.inflight]
    e1.record()
    barrier()
    sampler.mark_end()
    t_dev = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([t_dev], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    n_det = int(cnt.sum().item())

    # ---- end to end through the public API: pinned host uint8 -> predict() -> Results on the host ----
    # predict(source, stream=True): every step uploads its batch from pinned host memory (copy stream, overlapped
    # with the previous step's compute) and lands its [B, max_det, 6] result in host memory before it is yielded
    for res in yolo.predict([x_host] * 3, stream=True, **pred_args):
        pass
    barrier()
    t0 = time.perf_counter()
    n_e2e = 0
    for res in yolo.predict([x_host] * args.steps, stream=True, **pred_args):
        n_e2e += len(res)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert n_e2e == B * args.steps
    barrier()
    te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e_max = float(te.item())
    h2d = x_host.numel() * x_host.element_size()
    d2h = B * MAX_DET * 6 * 4 + B * 4

    if rank == 0:
        roof, stages = stage_roofline(yolo.model, x_dev, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ips, threads, times = cpu_port_images_per_s(args.cpu_batch, 2, 1)
            cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_batch} of the same synthetic 640^2 images per step, 2 timed steps, oracle port "
                             f"(PyTorch CPU fp32 forward + numpy NMS)"}
        line = {
            "metric": METRIC, "value": world * B * args.steps / t_max, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "spectrogram-yolov11-s (yolo11s_fusion_sand3_new, 6.82M params, nc=2) predict: "
                                   "fwd + fused decode + NMS, 640^2, batch 64 per GPU (BASELINE configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "imgsz": IMGSZ, "conf": CONF, "iou": IOU,
                       "input": "uint8 NCHW, resident in HBM for `value`, pinned host for `e2e`",
                       "pipelining": f"{args.inflight} batches in flight (one CUDA-graph instance + stream each), every step = full fwd+decode+NMS",
                       "l2": "no flush: one step streams ~9 GB of activations, >> 126 MB L2",
                       "sharding": "images split across GPUs, no collective on the data path",
                       "detections_last_step": n_det},
            "e2e": {"value": world * B * args.steps / t_e2e_max, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "specyolo.YOLO.predict(iterable of pinned uint8 batches, stream=True)"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": roof, "stages": stages, "clocks": clocks, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="specyolo", choices=["specyolo", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight in the resident-input loop")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_product_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

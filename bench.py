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
# CPU arm: the reference's own predict path on the host cores (oracle/_ref = the unmodified reference installed by
# oracle/Makefile), or — where that directory did not travel — the oracle port with BN folded as the reference's
# predict path does (nn/autobackend.py:152)
# ------------------------------------------------------------------------------------------------
def _reference_yolo(sd):
    """The REAL reference `YOLO` object carrying the synthetic weights, or None when oracle/_ref is absent."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        return None
    ultralytics = ref_loader.import_reference()
    from ultralytics.nn.tasks import DetectionModel as RefModel

    cfg = Path(ultralytics.__file__).parent / "cfg" / "models" / "11" / CFG
    m = RefModel(str(cfg), nc=NC, verbose=False)
    m.load_state_dict(sd, strict=True)
    y = ultralytics.YOLO(str(cfg), task="detect")
    y.model = m.eval()
    return y


def cpu_reference_images_per_s(batch: int, repeats: int, warmup: int, seed: int = 0):
    """(images/s, threads, per-step times, kind): `YOLO(cfg).predict(x, device='cpu')` of the real reference when
    oracle/_ref is present (kind "reference"), else the oracle port (kind "port")."""
    import torch
    import yaml

    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    threads = os.cpu_count() or 1
    m = specyolo.DetectionModel(CFG, nc=NC)           # parameter shapes only (CPU tensors, never run)
    sd = synth_state_dict(m, seed=seed)
    x = synth_images(batch, IMGSZ, seed=seed, dtype=torch.uint8)
    ref = _reference_yolo(sd)
    if ref is not None:
        kind = "reference"
        xf = x.float() / 255                          # what LoadTensor hands to preprocess for a tensor source

        def step():
            return ref.predict(xf, device="cpu", conf=CONF, iou=IOU, max_det=MAX_DET, verbose=False)

        step()                                        # builds the predictor (select_device caps torch threads at min(8, n-1))
        torch.set_num_threads(threads)                # ... lift that cap: the arm gets every host thread
    else:
        kind = "port"
        from oracle import nms_ref, yolo_ref

        torch.set_num_threads(threads)
        graph = yolo_ref.parse_graph(yaml.safe_load((PKG / "specyolo" / "cfg" / CFG_FILE).read_text()), "s", NC)
        R = yolo_ref.Ref(sd, fuse=True)               # BN folded once, as AutoBackend does before predict

        def step():
            with torch.no_grad():
                y, _ = yolo_ref.forward(graph, R, x.float() / 255)
            return nms_ref.non_max_suppression(y.numpy(), CONF, IOU, max_det=MAX_DET)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), threads, times, kind


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    ips, threads, times, kind = cpu_reference_images_per_s(args.cpu_batch, args.steps, args.warmup)
    sample = f"{args.cpu_batch} synthetic 640^2 uint8 images per step, {args.steps} timed steps"
    path = ("the unmodified reference (oracle/_ref): ultralytics.YOLO(cfg).predict(x, device='cpu'), fp32, BN fused, "
            "torchvision NMS, all host threads" if kind == "reference" else
            "oracle port of the reference CPU predict path (PyTorch CPU fp32, BN folded, numpy NMS)")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "spectrogram-yolov11-s nc=2 640^2 predict (fwd + NMS)", "cpu_batch": args.cpu_batch,
                   "path": path},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def graph_roofline(fn, peaks, kernel_desc: str):
    """`roofline` + `stages` of one captured step from IN-GRAPH kernel durations (specyolo/utils/kprof.py: the step's
    graph on one stream without PDL, CUPTI activity records over 5 replays, per-launch median).  `serial_ms` is the step
    time of that same execution, so the group's kernel time is <= it by construction."""
    from specyolo.utils import kprof

    prof = kprof.profile_graph(fn)
    c, stages, total = kprof.summarise(prof, peaks)
    traffic = None
    tf = ROOT / "profiles" / "conv_dram_traffic.json"      # written from an ncu capture by tools/ncu_traffic.py
    if tf.is_file():
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_step")
    gbs = c[2] / c[0] / 1e9
    tfs = c[1] / c[0] / 1e12
    # most conv launches (and most of their time) sit below the ridge point of the measured peaks, i.e. are HBM-bound at
    # bf16: the headline roofline of the group is the bandwidth one, the tensor-pipe view of the same launches is
    # reported beside it, and `stages` splits the launches at the ridge.
    roof = {"bound": "hbm", "kernel": kernel_desc,
            "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
            "traffic": traffic, "algorithmic_bytes_per_step": c[2],
            "peak_source": peaks["src"] + " (HBM copy bandwidth; bf16 sustained for the tensor view)",
            "flops_per_step": c[1], "group_ms_per_step": 1e3 * c[0], "launches_per_step": c[3],
            "tflops_same_launches": tfs, "frac_tensor_same_launches": tfs / peaks["tf_sust"], "tensor_peak": peaks["tf_sust"],
            "timing": {"source": prof["source"], "note": prof["note"],
                       "serial_graph_ms_per_step": prof["serial_ms"], "all_kernels_ms_per_step": 1e3 * total,
                       "what": "kernel durations inside a single-stream, PDL-off capture of the step; the headline "
                               "ms_per_step is the same kernels with PDL, parallel branches and several batches in flight"}}
    return roof, stages, prof


def time_pipelined(predictor, x_dev, steps, inflight, barrier):
    """steps full passes with `inflight` graph instances; returns (device seconds, last (out, cnt))."""
    import torch

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = predictor.infer_pipelined(x_dev, steps, inflight)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) * 1e-3, last[(steps - 1) % inflight]


def run_product_arm(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import specyolo
    from specyolo.nn.init import (EMISSION_DB_RANGE, IQ_CLS_BIAS, YOLO11S_1280_CLS_BIAS, synth_images,
                                  synth_iq_emissions, synth_state_dict)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from specyolo.dist import bind_to_gpu_numa

    numa = bind_to_gpu_numa(local_rank)     # before any pinned allocation: first touch puts the staging buffers on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    yolo = specyolo.YOLO(CFG, nc=NC)
    sd = synth_state_dict(yolo.model, seed=0)
    yolo.load_state_dict(sd)
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
    sampler.mark_begin()
    t_dev, (out, cnt) = time_pipelined(predictor, x_dev, args.steps, args.inflight, barrier)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    t_max = max_over_ranks(t_dev)
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
    t_e2e_max = max_over_ranks(t_e2e)
    h2d = x_host.numel() * x_host.element_size()
    d2h = B * MAX_DET * 6 * 4 + B * 4

    roof = stages = None
    if rank == 0:
        roof, stages, _ = graph_roofline(
            lambda: yolo.model.detect_fused(x_dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET), peaks,
            "conv_igemm_kernel + conv_halo_kernel + dwpw_kernel + stem_pair_kernel: every tcgen05 implicit-GEMM conv launch "
            "of one step (most layers are below the ridge point, i.e. HBM-bound; stages.conv2d_tensor_bound / "
            "conv2d_hbm_bound split them)")
    # free the C2 graphs before the larger configurations are captured
    yolo.predictor = None
    del predictor
    torch.cuda.empty_cache()

    # ---- c3: raw IQ -> boxes (BASELINE configs[2]: 256 bursts of 2^20 samples over 8 GPUs = 32 per GPU), the STFT
    # kernel inside the captured step; IQ resident in HBM for `value` (8.4 MB per burst: a host feed would be PCIe-bound)
    c3 = None
    if not args.no_extra:
        nb = args.iq_bursts
        iq = torch.view_as_real(synth_iq_emissions(nb, 1 << 20, seed=100 + rank)).contiguous().to(dev)
        front = specyolo.engine.IQFrontEnd(db_min=EMISSION_DB_RANGE[0], db_max=EMISSION_DB_RANGE[1], out_hw=(IMGSZ, IMGSZ))
        # same architecture and seed, class-logit bias raised for letterboxed spectrograms (specyolo/nn/init.py)
        yolo.load_state_dict(synth_state_dict(yolo.model, seed=0, cls_bias=IQ_CLS_BIAS))
        yolo.fuse()
        p3 = specyolo.DetectionPredictor(yolo.model, pred_args, front=front)
        p3.infer_pipelined(iq, 3, args.inflight)
        torch.cuda.synchronize()
        l3 = p3.last_launches
        steps3 = max(10, args.steps // 2)
        t3, (_, cnt3) = time_pipelined(p3, iq, steps3, args.inflight, barrier)
        t3 = max_over_ranks(t3)
        c3 = {"workload": f"raw IQ -> STFT 1024/256 -> letterbox 640^2 -> spectrogram-yolov11-s -> NMS, {nb} bursts of 2^20 "
                          f"complex64 samples per GPU (BASELINE configs[2]), STFT kernel inside the captured graph, IQ resident in HBM",
              "value": world * nb * steps3 / t3, "unit": "bursts/s", "ms_per_step": 1e3 * t3 / steps3, "steps": steps3,
              "bursts_per_gpu": nb, "launches_per_step": l3, "detections_last_step": int(cnt3.sum().item()),
              "parity": "STFT stage unpinned (the reference has no IQ code); detector + NMS pinned"}
        # the same workload end to end through the public API: pinned host IQ -> predict_iq(stream=True) -> host results
        # (268 MB of samples per 32-burst step: bound by the host-to-device copy, not by the kernels).  No collective inside
        # the try block: a rank that fails must not leave the others waiting.
        steps3e = max(6, steps3 // 4)
        t3e, err3, h2d3 = -1.0, "", nb * (1 << 20) * 8
        try:
            iq_host = iq.cpu().pin_memory()
            kw3 = dict(db_min=EMISSION_DB_RANGE[0], db_max=EMISSION_DB_RANGE[1], imgsz=IMGSZ, **pred_args)
            for _ in yolo.predict_iq([iq_host] * 3, stream=True, **kw3):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n3 = sum(len(r) for r in yolo.predict_iq([iq_host] * steps3e, stream=True, **kw3))
            torch.cuda.synchronize()
            t3e = time.perf_counter() - t0
            assert n3 == nb * steps3e
            yolo._iq_predictor = None
            del iq_host
        except Exception as ex:      # an extra key must never take the bench line down
            t3e, err3 = -1.0, f"{type(ex).__name__}: {ex}"[:200]
        ok3 = -max_over_ranks(-(1.0 if t3e > 0 else 0.0)) > 0.5          # every rank measured
        t3e = max_over_ranks(t3e)
        if ok3:
            c3["e2e"] = {"value": world * nb * steps3e / t3e, "unit": "bursts/s", "h2d_bytes_per_step": h2d3,
                         "d2h_bytes_per_step": nb * MAX_DET * 6 * 4 + nb * 4, "h2d_gbs_per_gpu": h2d3 * steps3e / t3e / 1e9,
                         "api": "specyolo.YOLO.predict_iq(iterable of pinned complex64 batches, stream=True)"}
        else:
            c3["e2e"] = {"error": err3 or "failed on another rank"}
        if rank == 0:
            from specyolo.utils import kprof

            prof3 = kprof.profile_graph(lambda: yolo.model.detect_fused(front(iq), conf_thres=CONF, iou_thres=IOU,
                                                                        max_det=MAX_DET))
            st = [c for c in prof3["calls"] if c["stage"] == "stft_letterbox"]
            if st:
                us = sum(c["us"] for c in st)
                by = sum(c["bytes"] for c in st)
                c3["roofline"] = {"bound": "hbm", "kernel": "stft_letterbox_kernel", "achieved": by / us / 1e3,
                                  "peak": peaks["hbm"], "unit": "GB/s", "frac": by / us / 1e3 / peaks["hbm"],
                                  "algorithmic_bytes_per_launch": by, "us_per_launch": us, "traffic": None,
                                  "timing": prof3["source"], "serial_graph_ms_per_step": prof3["serial_ms"]}
        del p3, iq
        torch.cuda.empty_cache()

    # ---- c4: stock yolo11s (nc=80) at 1280^2, batch 128 per GPU (BASELINE configs[3]) ----
    c4 = None
    if not args.no_extra:
        B4 = args.c4_batch
        y4 = specyolo.YOLO("yolo11s.yaml", nc=80)
        y4.load_state_dict(synth_state_dict(y4.model, seed=3, cls_bias=YOLO11S_1280_CLS_BIAS))
        y4.to(dev)
        y4.fuse()
        x4 = synth_images(B4, 1280, seed=200 + rank, dtype=torch.uint8).to(dev)
        p4 = specyolo.DetectionPredictor(y4.model, pred_args)
        p4.infer_pipelined(x4, 3, 2)
        torch.cuda.synchronize()
        l4 = p4.last_launches
        steps4 = max(6, args.steps // 10)
        t4, (_, cnt4) = time_pipelined(p4, x4, steps4, 2, barrier)
        t4 = max_over_ranks(t4)
        c4 = {"workload": f"yolo11s (nc=80) predict at 1280^2, batch {B4} per GPU, bf16 (BASELINE configs[3]), uint8 input resident in HBM, "
                          f"2 batches in flight",
              "value": world * B4 * steps4 / t4, "unit": "images/s", "ms_per_step": 1e3 * t4 / steps4, "steps": steps4,
              "batch_per_gpu": B4, "launches_per_step": l4, "detections_last_step": int(cnt4.sum().item()),
              "tflops": world * B4 * steps4 / t4 * 87.8e9 / 1e12,
              "frac_tensor": B4 * steps4 / t4 * 87.8e9 / 1e12 / peaks["tf_sust"]}
        del p4
        torch.cuda.empty_cache()
        if rank == 0:
            r4, s4, _ = graph_roofline(lambda: y4.model.detect_fused(x4, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET),
                                       peaks, "every tcgen05 implicit-GEMM conv launch of one yolo11s 1280^2 step")
            r4["traffic"] = None
            c4["roofline"] = r4
            c4["stages"] = {k: {kk: v[kk] for kk in ("launches", "ms", "frac_tensor", "frac_hbm") if kk in v} for k, v in s4.items()}
        del y4, x4
        torch.cuda.empty_cache()

    # ---- c5: DDP training step (BASELINE configs[4]) ----
    c5 = None
    if not args.no_extra:
        c5 = c5_train_step(sd, dev, rank, world, barrier, max_over_ranks, args.c5_batch)

    if rank == 0:
        cpu = lib_bar = None
        if world == 1 and not args.no_cpu_baseline:
            ips, threads, times, kind = cpu_reference_images_per_s(args.cpu_batch, 2, 1)
            cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                   "sample": f"{args.cpu_batch} of the same synthetic 640^2 images per step, 2 timed steps, " +
                             ("the unmodified reference's YOLO.predict(device='cpu') from oracle/_ref (fp32, BN fused, "
                              "torchvision NMS)" if kind == "reference" else
                              "oracle port (PyTorch CPU fp32 forward with BN folded + numpy NMS)")}
        crit = None
        if world == 1 and not args.no_extra:
            lib_bar = library_bar(sd, x_dev, dev)
            crit = train_criterion_bar(dev, B)
            if c5 is not None and "value" in c5:
                c5["cuda_graph"] = c5_graph_subprocess(args.c5_batch, local_rank)
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
                    "d2h_bytes_per_step": d2h, "h2d_gbs_per_gpu": h2d * args.steps / t_e2e_max / 1e9, "numa": numa,
                    "api": "specyolo.YOLO.predict(iterable of pinned uint8 batches, stream=True)"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": roof, "stages": stages, "clocks": clocks, "cpu_baseline": cpu,
            "c3": c3, "c4": c4, "c5": c5, "library_bar": lib_bar, "train_criterion": crit,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5_train_step(sd, dev, rank, world, barrier, max_over_ranks, batch):
    """BASELINE configs[4] / SURVEY 8 f2: one data-parallel TRAINING step of the Spectrogram detector on synthetic images and
    labels, one process per GPU: forward, detection criterion, backward with the NCCL gradient all-reduce of
    DistributedDataParallel (engine/trainer.py:272-273), gradient clipping + SGD step + EMA update (trainer.py:380-399,
    590-600), fp32 as the reference trainer runs when its AMP self-check cannot (offline) — under fp16 autocast GCT's sum of squares
    over an 80 x 80 map overflows and the loss is NaN for the stock criterion as well.  The network's forward / backward are the REFERENCE's own
    PyTorch modules (oracle/_ref; cuDNN / cuBLAS library kernels — this package has no backward pass); what this package
    contributes to the step is the criterion (`specyolo_det_loss`) and the EMA update (`specyolo_ema_update`), bound in through
    `ultralytics_shim.install()`.  Two arms on the same weights and batch: stock and shim.  Max over ranks."""
    import torch

    out = {"workload": f"spectrogram-yolov11-s nc={NC} DDP training step, batch {batch} per GPU, {IMGSZ}^2, fp32 (cuDNN TF32 convs), SGD + EMA, "
                       "synthetic images / 8 boxes per image (BASELINE configs[4]); forward / backward = the reference's PyTorch "
                       "modules (library kernels), criterion + EMA update = this package in the shim arm",
           "unit": "images/s", "batch_per_gpu": batch, "n_gpus": world,
           "allreduce": "torch DistributedDataParallel over NCCL" if world > 1 else "none (one GPU)"}
    try:
        import copy
        from types import SimpleNamespace

        import torch.distributed as dist

        from oracle import ref_loader                       # comparison / host leg: the REFERENCE's modules are the network
        from specyolo import ultralytics_shim as shim
        from specyolo.nn.init import synth_det_batch, synth_images

        # every rank must take the same path through the collectives below: agree first on whether the leg can run at all
        ready, why = 1.0, ""
        try:
            if not ref_loader.reference_available():
                raise RuntimeError("oracle/_ref (the reference package) did not travel to this box")
            ultralytics = ref_loader.import_reference()
            from ultralytics.nn.tasks import DetectionModel as RefModel
        except Exception as ex:
            ready, why = 0.0, f"{type(ex).__name__}: {ex}"[:200]
        if -max_over_ranks(-ready) < 1.0:                               # min over ranks
            out["unavailable"] = why or "the reference package is missing on another rank"
            return out

        torch.manual_seed(0)        # a freshly initialised model, as `YOLO(cfg).train()` starts from (the calibrated synthetic
        base = RefModel(str(Path(ultralytics.__file__).parent / "cfg" / "models" / "11" / CFG), nc=NC, verbose=False)
        # inference weights overflow fp16 in GCT's sum of squares once BatchNorm switches to batch statistics)
        labels = {k: v.to(dev) for k, v in synth_det_batch(batch, IMGSZ, NC, 8, seed=11 + rank).items()}
        labels["img"] = synth_images(batch, IMGSZ, seed=300 + rank).to(dev)
        steps, warm = 8, 3

        def arm():
            from ultralytics.utils.torch_utils import ModelEMA          # the shim rebinds this name while installed

            model = copy.deepcopy(base).to(dev).train()
            for name, prm in model.named_parameters():                  # trainer.py:236-252: everything trains but the DFL projection
                prm.requires_grad_(".dfl" not in name)
            model.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
            model.criterion = None
            ema = ModelEMA(model)
            net = (torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
                   if world > 1 else model)                        # trainer.py:272-273
            opt = torch.optim.SGD([q for q in model.parameters() if q.requires_grad], lr=1e-4, momentum=0.937, nesterov=True)

            def step():
                loss, _items = net(labels)                               # BaseModel.forward(dict) -> self.loss(batch) (tasks.py:100-116)
                (loss * world).backward()                                # trainer.py:385: loss *= world_size under DDP
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
                opt.step()
                opt.zero_grad()
                ema.update(model)
                return loss

            for _ in range(warm):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                last = step()
            barrier()
            t = max_over_ranks(time.perf_counter() - t0)
            lv = float(last.detach())
            del net, opt, ema, model
            torch.cuda.empty_cache()
            return t, lv

        t_stock, l_stock = arm()
        shim.install()
        try:
            t_shim, l_shim = arm()
        finally:
            shim.uninstall()
        out.update({"value": world * batch * steps / t_shim, "ms_per_step": 1e3 * t_shim / steps, "steps": steps,
                    "stock_value": world * batch * steps / t_stock, "stock_ms_per_step": 1e3 * t_stock / steps,
                    "loss_last_step": {"stock": l_stock, "shim": l_shim},
                    "timing": "host wall clock between barriers (+ device synchronize), max over ranks"})
        if world > 1:
            dist.barrier()
    except Exception as ex:      # an extra key must never take the bench line down
        out["error"] = f"{type(ex).__name__}: {ex}"[:300]
    return out


def c5_graph_worker(batch: int, local_rank: int):
    """`bench.py --c5-graph`: the c5 training step of ONE GPU as a single CUDA graph — forward of the reference's modules,
    this package's criterion on pre-packed static targets (`pack_batch_targets`: no host work, no host synchronisation),
    backward, gradient clipping and the SGD step captured once and replayed; the EMA update (its decay changes per step) runs
    eagerly after each replay.  The reference's own criterion cannot be captured (`if fg_mask.sum()`, `max(target_scores
    .sum(), 1)` and the per-image loop of `preprocess` synchronise with the host).  Run as a SEPARATE PROCESS by the parent
    bench so that a capture failure can never take the bench line down; prints one JSON object."""
    import copy
    from types import SimpleNamespace

    import torch

    out = {}
    try:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        from oracle import ref_loader                       # host leg: the REFERENCE's modules are the network
        from specyolo import ultralytics_shim as shim
        from specyolo.nn.init import synth_det_batch, synth_images
        from specyolo.utils.loss import pack_batch_targets

        ultralytics = ref_loader.import_reference()
        from ultralytics.nn.tasks import DetectionModel as RefModel

        torch.manual_seed(0)
        base = RefModel(str(Path(ultralytics.__file__).parent / "cfg" / "models" / "11" / CFG), nc=NC, verbose=False)
        labels = synth_det_batch(batch, IMGSZ, NC, 8, seed=11)
        gbatch = {k: v.to(dev) for k, v in labels.items()}
        gbatch["img"] = synth_images(batch, IMGSZ, seed=300).to(dev)
        gbatch["packed_targets"] = pack_batch_targets(labels, batch, (IMGSZ, IMGSZ), dev, max_boxes=8)
        shim.install()
        from ultralytics.utils.torch_utils import ModelEMA

        model = copy.deepcopy(base).to(dev).train()
        for name, prm in model.named_parameters():
            prm.requires_grad_(".dfl" not in name)
        model.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
        model.criterion = None
        ema = ModelEMA(model)
        params = [q for q in model.parameters() if q.requires_grad]
        opt = torch.optim.SGD(params, lr=1e-4, momentum=0.937, nesterov=True)

        def step_body():
            loss, _items = model(gbatch)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, max_norm=10.0)
            opt.step()
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                                            # warm-up: cuDNN plans, criterion caches, momentum buffers
                opt.zero_grad(set_to_none=True)
                step_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = step_body()
        steps = 20
        for _ in range(3):
            graph.replay()
            ema.update(model)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            graph.replay()
            ema.update(model)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        out = {"value": batch * steps / t, "ms_per_step": 1e3 * t / steps, "steps": steps, "loss_last_step": float(static_loss.detach()),
               "what": "forward + criterion + backward + clip + SGD step replayed as one CUDA graph, EMA update eager; one GPU"}
    except Exception as ex:
        out = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    print("C5GRAPH " + json.dumps(out), flush=True)


def c5_graph_subprocess(batch: int, local_rank: int):
    """Runs c5_graph_worker in its own process (bounded time); returns its JSON object or an error record."""
    import subprocess

    try:
        r = subprocess.run([sys.executable, str(Path(__file__).resolve()), "--c5-graph", "--c5-batch", str(batch)],
                           capture_output=True, text=True, timeout=240,
                           env={**os.environ, "LOCAL_RANK": str(local_rank), "WORLD_SIZE": "1", "RANK": "0"})
        for ln in r.stdout.splitlines():
            if ln.startswith("C5GRAPH "):
                return json.loads(ln[len("C5GRAPH "):])
        return {"error": f"worker exited {r.returncode}: {(r.stderr or '')[-200:]}"}
    except Exception as ex:
        return {"error": f"{type(ex).__name__}: {ex}"[:200]}


def train_criterion_bar(dev, BATCH):
    """SURVEY 8 f2, first slice: the detection criterion of a training step (v8DetectionLoss: assigner + BCE / CIoU / DFL,
    forward + backward into the head maps) at the bench's batch — specyolo_det_loss (six launches) against the reference's
    own criterion in PyTorch eager on the same GPU (oracle/_ref), same seeded head maps and labels."""
    import torch

    out = {"unit": "ms per forward+backward", "batch": BATCH, "imgsz": IMGSZ, "nc": NC, "boxes_per_image": 8}
    try:
        from types import SimpleNamespace

        from specyolo.nn.init import synth_det_batch
        from specyolo.utils.loss import v8DetectionLoss

        feats, batch = synth_det_batch(BATCH, IMGSZ, NC, 8, seed=5, head_maps=True)
        f = [x.to(dev).requires_grad_(True) for x in feats]
        fake = SimpleNamespace(model=[SimpleNamespace(nc=NC, reg_max=16, stride=torch.tensor([8.0, 16.0, 32.0]))],
                               args={"box": 7.5, "cls": 0.5, "dfl": 1.5}, parameters=lambda: iter([torch.zeros(1, device=dev)]))

        def timeit(c, reps=10):
            def step():
                for x in f:
                    x.grad = None
                total, items = c(f, batch)
                total.backward()
                return items
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                items = step()
            torch.cuda.synchronize()
            return 1e3 * (time.perf_counter() - t0) / reps, [float(v) for v in items]

        out["specyolo"], out["loss_items"] = timeit(v8DetectionLoss(fake))
        out["what"] = "host wall clock per call incl. torch cat / permute glue, label packing and autograd"
        from oracle import ref_loader

        if ref_loader.reference_available():
            ref_loader.import_reference()
            from ultralytics.utils.loss import v8DetectionLoss as RefLoss

            out["reference_eager"], out["reference_loss_items"] = timeit(RefLoss(SimpleNamespace(
                model=fake.model, args=SimpleNamespace(box=7.5, cls=0.5, dfl=1.5), parameters=fake.parameters)))
    except Exception as ex:      # an extra key must never take the bench line down
        out["error"] = f"{type(ex).__name__}: {ex}"[:200]
    return out


def library_bar(sd, x_dev, dev):
    """BASELINE.md 3.7 / SURVEY 2.1: the same network through the LIBRARY path on the same GPU — the reference's own
    `YOLO.predict(x, device=0, half=True)` (cuDNN / cuBLAS fp16 + torchvision CUDA nms) when oracle/_ref is present, and
    the oracle port's functional forward in bf16 channels-last + torchvision nms otherwise.  Images/s on the bench's
    batch; this is the bar every hand-written kernel has to beat, not the CPU."""
    import torch
    import torchvision
    import yaml

    out = {"unit": "images/s", "batch": int(x_dev.shape[0])}
    xf = (x_dev.float() / 255).contiguous()

    def timeit(step, reps=5):
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        return xf.shape[0] * reps / (time.perf_counter() - t0)

    try:
        ref = _reference_yolo(sd)
        if ref is not None:
            idx = dev.index or 0
            for half in (True, False):
                ref.predictor = None
                step = lambda: ref.predict(xf, device=idx, half=half, conf=CONF, iou=IOU, max_det=MAX_DET, verbose=False)  # noqa: E731
                out["reference_predict_fp16" if half else "reference_predict_fp32"] = timeit(step)
            out["what"] = ("the unmodified reference's YOLO.predict(tensor, device=0) on this GPU: PyTorch eager cuDNN/cuBLAS "
                           "(half=True: fp16) + its Python NMS loop over torchvision.ops.nms (CUDA)")
    except Exception as ex:      # the library bar must never take the bench line down
        out["reference_error"] = f"{type(ex).__name__}: {ex}"[:200]
    try:
        from oracle import yolo_ref

        graph = yolo_ref.parse_graph(yaml.safe_load((PKG / "specyolo" / "cfg" / CFG_FILE).read_text()), "s", NC)
        R = yolo_ref.Ref(sd, fuse=True, device=dev, dtype=torch.bfloat16)
        xb = xf.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

        def step_port():
            with torch.no_grad():
                y, _ = yolo_ref.forward(graph, R, xb)
                y = y.float()
                for b in range(y.shape[0]):          # the reference's per-image loop, on the device (ops.py:262-327)
                    p = y[b].t()
                    sc, j = p[:, 4:].max(1)
                    m = sc > CONF
                    p, sc, j = p[m], sc[m], j[m]
                    bx = torch.cat((p[:, :2] - p[:, 2:4] / 2, p[:, :2] + p[:, 2:4] / 2), 1)
                    torchvision.ops.nms(bx + j[:, None].float() * 7680, sc, IOU)[:MAX_DET]
        out["torch_bf16_channels_last"] = timeit(step_port)
        out["what_port"] = ("the oracle port's functional forward on this GPU in bf16 channels-last (cuDNN/cuBLAS, BN folded) + "
                            "threshold + torchvision.ops.nms (CUDA) per image")
    except Exception as ex:
        out["port_error"] = f"{type(ex).__name__}: {ex}"[:200]
    return out


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
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 / c5 / library_bar keys")
    ap.add_argument("--c5-batch", type=int, default=16, help="images per GPU of the c5 training step (the reference's default batch)")
    ap.add_argument("--c5-graph", action="store_true", help="(internal) run only the CUDA-graph arm of c5 and print its JSON object")
    ap.add_argument("--iq-bursts", type=int, default=32, help="IQ bursts per GPU per step of the c3 workload")
    ap.add_argument("--c4-batch", type=int, default=128, help="images per GPU per step of the c4 workload (yolo11s 1280^2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.c5_graph:
        c5_graph_worker(args.c5_batch, local_rank)
        return
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_product_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""In-graph per-kernel timing of one captured step (bench.py `roofline` / `stages`, tools/profile_layers.py).

What is measured: the step's own CUDA graph, captured on ONE stream with programmatic dependent launch off
(`SPECYOLO_NO_PDL=1`, `ops.CONCURRENT = False`), so that kernels run strictly one after the other and a kernel's
CUPTI duration is its execution alone — with PDL on, a kernel is resident (spinning in `griddepcontrol.wait`) while
its predecessor drains and the durations overlap.  Durations come from CUPTI activity records (torch.profiler /
kineto) over `reps` replays of that graph; the per-launch figure is the median over the replays.  The same graph is
also replayed WITHOUT the profiler between CUDA events: `serial_ms` is the step time of the very execution the kernel
durations belong to, so the sum of any group's kernel time is <= `serial_ms` by construction.  The headline `value`
of bench.py is a different execution of the same kernels (PDL on, independent branches on side streams, several
batches in flight) and is therefore shorter than `serial_ms`.

Work per launch: an `OpRecorder` wraps the `specyolo.ops` entry points for one eager pass of the same function and
notes, per call, the stage name, algorithmic FLOPs and bytes and how many kernels the call launched
(`specyolo_launch_count` delta); the graph replays the same launches in the same order, so record i <-> kernel(s) i.
"""
from __future__ import annotations

import json
import os
import statistics
import tempfile
from typing import Callable, List, Optional

import torch

from .. import _lib, ops


def _conv_k(pc) -> int:
    return pc.alg_k or (pc.cin // pc.g_orig) * pc.k * pc.k            # algorithmic reduction length (no padding)


def _conv_work(r, x, pc, out=None, residual=None, out_fp32=False, blocked_out=False):
    B, Cin, H, W = x.shape
    Ho, Wo = pc.out_hw(H, W)
    fl = 2.0 * B * Ho * Wo * pc.cout * _conv_k(pc)
    by = 2.0 * (B * Cin * H * W) + r.numel() * r.element_size() + 2.0 * pc.w.numel()
    if residual is not None:
        by += 2.0 * residual.numel()
    label = (f"conv {Cin:4d}->{pc.cout:4d} k{pc.k} s{pc.s} d{pc.d} g{pc.g_orig:3d} {H:4d}x{W:4d} "
             f"M={B * Ho * Wo:8d} K={_conv_k(pc):5d}" + (" +res" if residual is not None else ""))
    return "conv2d", fl, by, label


def _dwpw_work(r, x, dw_w, dw_b, pw, out=None, head=None):
    B, C, H, W = x.shape
    if head is not None:        # + the closing Conv2d(cout, nc, 1); only nc fp32 logits per pixel are written
        nc = head[0].shape[0]
        return ("conv2d", 2.0 * B * H * W * (C * (9 + pw.cout) + pw.cout * nc), 2.0 * x.numel() + 4.0 * B * H * W * nc + 2.0 * pw.w.numel(),
                f"dw3x3+pw+head {C:4d}->{pw.cout:4d}->{nc} fused {H:4d}x{W:4d} M={B * H * W:8d}")
    return ("conv2d", 2.0 * B * H * W * C * (9 + pw.cout), 2.0 * (x.numel() + r.numel()) + 2.0 * pw.w.numel(),
            f"dw3x3+pw {C:4d}->{pw.cout:4d} fused {H:4d}x{W:4d} M={B * H * W:8d}")


def _stem_pair_work(r, x, pc0, pc1b, out=None):
    B, _, H, W = x.shape
    fl = 2.0 * B * ((H // 2) * (W // 2) * pc0.cout * 27 + (H // 4) * (W // 4) * pc1b.cout * 9 * pc0.cout)
    return ("conv2d", fl, x.numel() * x.element_size() + 2.0 * r.numel() + 2.0 * pc1b.w.numel(),
            f"fused stem 3->{pc0.cout}->{pc1b.cout} (u8 in) {H:4d}x{W:4d}")


def _bneck_work(r, x, pc1, pc2, add, out=None):
    B, C, H, W = x.shape
    fl = 2.0 * B * H * W * 9 * (C * pc1.cout + pc1.cout * pc2.cout)
    return ("conv2d", fl, 2.0 * (x.numel() + r.numel()) + 2.0 * (pc1.w.numel() + pc2.w.numel()),
            f"bottleneck {C:4d}->{pc1.cout:3d}->{pc2.cout:4d} 3x3+3x3 fused {H:4d}x{W:4d} M={B * H * W:8d}" + (" +res" if add else ""))


_WORK = {
    "conv2d": _conv_work,
    "bottleneck": _bneck_work,
    "dwconv_pwconv": _dwpw_work,
    "stem_pair": _stem_pair_work,
    "stem_space_to_depth": lambda r, x: ("stem_space_to_depth", 0.0, x.numel() * x.element_size() + 2.0 * r.numel(),
                                         f"space-to-depth {tuple(x.shape)}"),
    "sppf_pool": lambda r, buf, c: ("sppf_pool", 0.0, 2.0 * buf.numel(), f"sppf_pool c={c} {tuple(buf.shape)}"),
    "fusion_eschannel": lambda r, xs, *a, **k: ("fusion_eschannel", 0.0, 2.0 * (2 * sum(t.numel() for t in xs) + r.numel()),
                                                f"fusion k={len(xs)} -> {tuple(r.shape)}"),
    "psa_attention": lambda r, qkv, heads, key_dim, head_dim, *a, **k: (
        "psa_attention", 2.0 * qkv.shape[0] * heads * (qkv.shape[2] * qkv.shape[3]) ** 2 * (key_dim + head_dim),
        2.0 * (qkv.numel() + r.numel()), f"psa_attention N={qkv.shape[2] * qkv.shape[3]} heads={heads}"),
    # with the fused threshold only the class logits of every anchor (and the 64 DFL bins of the few survivors) are read
    "detect_decode": lambda r, logits, hw, strides, nc, want_dense=True, conf_thres=None: (
        "detect_decode", 0.0,
        4.0 * sum(t.numel() for t in logits) if (want_dense or conf_thres is None) else
        4.0 * sum(t.shape[0] * t.shape[1] * max(nc, 8) for t in logits),       # (a 32-byte sector per row at least)
        "detect_decode" + ("" if want_dense else " (scores first)")),
    "nms": lambda r, *a, **k: ("nms", 0.0, 0.0, "nms"),
    "iq_to_letterbox": lambda r, iq, *a, **k: ("stft_letterbox", 0.0, 4.0 * iq.numel() + r.numel() * r.element_size(),
                                               f"iq_to_letterbox {tuple(iq.shape)}"),
    "upsample2x": lambda r, x, out=None: ("layout", 0.0, 2.0 * x.numel() + 2.0 * r.numel(), f"upsample x2 {tuple(x.shape)}"),
    "to_nhwc_bf16": lambda r, x, *a, **k: ("layout", 0.0, x.numel() * x.element_size() + 2.0 * r.numel(), "nchw->nhwc"),
}


class OpRecorder:
    """Context manager: every `specyolo.ops.<fn>` call made inside is appended to `.calls` as
    {stage, label, flops, bytes, kernels}.  `timed=True` additionally brackets each call with CUDA events (eager)."""

    def __init__(self, timed: bool = False):
        self.calls: List[dict] = []
        self.timed = timed
        self._orig = {}

    def __enter__(self):
        lib = _lib.load()
        for name, work in _WORK.items():
            f = getattr(ops, name)
            self._orig[name] = f

            def g(*a, _f=f, _work=work, _name=name, **k):
                n0 = lib.specyolo_launch_count()
                if self.timed:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                r = _f(*a, **k)
                if self.timed:
                    e1.record()
                nk = int(lib.specyolo_launch_count() - n0)      # wrapped ops never call each other
                stage, fl, by, label = _work(r, *a, **k)
                rec = {"stage": stage, "label": label, "flops": float(fl), "bytes": float(by), "kernels": nk, "op": _name}
                if self.timed:
                    rec["events"] = (e0, e1)
                self.calls.append(rec)
                return r
            setattr(ops, name, g)
        return self

    def __exit__(self, *exc):
        for name, f in self._orig.items():
            setattr(ops, name, f)
        self.calls = [c for c in self.calls if c["kernels"] > 0]
        return False


def profile_graph(fn: Callable[[], object], reps: int = 5, trace_path: Optional[str] = None) -> dict:
    """See the module docstring.  Returns {"calls": [... + "us" (median), "us_all"], "serial_ms", "kernels_per_step",
    "source": "cupti-in-graph" | "cuda-events-eager", "note"}."""
    lib = _lib.load()
    prev_env, prev_conc = os.environ.get("SPECYOLO_NO_PDL"), ops.CONCURRENT
    os.environ["SPECYOLO_NO_PDL"] = "1"
    ops.CONCURRENT = False
    try:
        with torch.no_grad():
            for _ in range(2):
                fn()
            with OpRecorder() as rec:
                n0 = lib.specyolo_launch_count()
                fn()
                n_step = int(lib.specyolo_launch_count() - n0)
            calls = rec.calls
            torch.cuda.synchronize()
            note = None
            if sum(c["kernels"] for c in calls) != n_step:
                note = f"recorder saw {sum(c['kernels'] for c in calls)} of {n_step} launches"
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                keep = fn()                      # noqa: F841  (outputs stay alive with the graph)
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            serial_ms = e0.elapsed_time(e1) / reps
            durs = None
            try:
                from torch.profiler import ProfilerActivity, profile

                # one profiler session per replay: a session then holds exactly one step, and a session in which CUPTI
                # dropped a record (it happens now and then) is recognised by its kernel count and repeated
                good, names, attempts = [], None, 0
                tmpdir = tempfile.mkdtemp(prefix="specyolo_kprof_")
                while len(good) < reps and attempts < reps + 4:
                    attempts += 1
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        graph.replay()
                        torch.cuda.synchronize()
                    path = trace_path or os.path.join(tmpdir, "trace.json")
                    prof.export_chrome_trace(path)
                    with open(path) as f:
                        ev = json.load(f)["traceEvents"]
                    ks = sorted((e for e in ev if e.get("cat") == "kernel" and "specyolo" in e.get("name", "")),
                                key=lambda e: e["ts"])
                    if len(ks) == n_step:
                        good.append([float(e["dur"]) for e in ks])
                        names = [e["name"] for e in ks]
                if note is None and len(good) >= 2:
                    durs = good
                    if attempts != len(good):
                        note = f"cupti dropped records in {attempts - len(good)} of {attempts} sessions (repeated)"
                elif note is None:
                    note = f"cupti sessions never held {n_step} kernels (last saw {len(ks)})"
            except Exception as ex:      # CUPTI not usable on this box: fall back to eager events, say so
                note = f"profiler unavailable ({type(ex).__name__}: {ex})"
            del graph
            if durs is not None:
                k = 0
                for c in calls:
                    per_rep = [sum(d[k:k + c["kernels"]]) for d in durs]
                    c["us"] = statistics.median(per_rep)
                    c["kernel_names"] = [n.split("(")[0][-80:] for n in names[k:k + c["kernels"]]]
                    k += c["kernels"]
                return {"calls": calls, "serial_ms": serial_ms, "kernels_per_step": n_step, "source": "cupti-in-graph",
                        "note": note, "replays_used": len(durs)}
            # fallback: eager CUDA events around every call, a device-side sleep queued first so the host runs ahead
            acc = None
            for _ in range(reps):
                torch.cuda.synchronize()
                torch.cuda._sleep(int(4e8))
                with OpRecorder(timed=True) as rec2:
                    fn()
                torch.cuda.synchronize()
                t = [c["events"][0].elapsed_time(c["events"][1]) * 1e3 for c in rec2.calls]
                acc = [[v] for v in t] if acc is None else [a + [v] for a, v in zip(acc, t)]
            for c, a in zip(calls, acc):
                c["us"] = statistics.median(a)
            return {"calls": calls, "serial_ms": serial_ms, "kernels_per_step": n_step, "source": "cuda-events-eager",
                    "note": note}
    finally:
        ops.CONCURRENT = prev_conc
        if prev_env is None:
            os.environ.pop("SPECYOLO_NO_PDL", None)
        else:
            os.environ["SPECYOLO_NO_PDL"] = prev_env


def summarise(prof: dict, peaks: dict) -> tuple:
    """(roofline of the conv group, stages) from profile_graph() output.  peaks: {"hbm": GB/s, "tf_sust": TFLOP/s}."""
    ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
    agg = {}
    for c in prof["calls"]:
        nm = c["stage"]
        if nm == "conv2d":
            nm = "conv2d_tensor_bound" if c["flops"] / max(c["bytes"], 1.0) >= ridge else "conv2d_hbm_bound"
        a = agg.setdefault(nm, [0.0, 0.0, 0.0, 0])
        a[0] += c["us"] * 1e-6
        a[1] += c["flops"]
        a[2] += c["bytes"]
        a[3] += c["kernels"]
    total = sum(a[0] for a in agg.values())
    stages = {}
    for nm, (t, fl, by, n) in agg.items():
        s = {"launches": n, "ms": 1e3 * t, "share": t / total if total else None}
        if fl:
            s.update(tflops=fl / t / 1e12, frac_tensor=fl / t / 1e12 / peaks["tf_sust"])
        if by:
            s.update(gbs=by / t / 1e9, frac_hbm=by / t / 1e9 / peaks["hbm"])
        stages[nm] = s
    c = [sum(v[j] for k, v in agg.items() if k.startswith("conv2d")) for j in range(4)]
    return c, stages, total

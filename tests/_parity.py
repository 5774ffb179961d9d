"""Detection-level comparison helpers for the parity tests (VERDICT r1 "next" item 1).

The reference's own half-precision precedent is `check_amp`: identical detection count and
`allclose(atol=0.5)` on [x1, y1, x2, y2, conf, cls] (ultralytics/utils/checks.py:691-699).  These helpers
measure exactly those quantities between two detection lists (oracle fp32 vs CUDA bf16)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def box_iou_np(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:4], b[None, :, 2:4])
    inter = np.clip(rb - lt, 0, None).prod(2)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-12)


def compare_detections(ref_list, got_list, score_floor: float = 0.30, match_iou: float = 0.5) -> dict:
    """ref_list / got_list: per image [n, 6] arrays (x1, y1, x2, y2, conf, cls).

    Every oracle detection with conf > score_floor is matched greedily (highest oracle score first) to the unmatched
    CUDA detection of the same class with the highest IoU >= match_iou.  Returns the matched rate over those oracle
    detections, the worst / mean |dbox| and |dscore| on the matches, and the per-image count differences over ALL
    detections (both directions are visible: a detection the CUDA path invents shows up in `count_diff`)."""
    n_ref = n_hit = 0
    dbox, dscore, count_diff = [], [], []
    unmatched = []
    for i, (r, g) in enumerate(zip(ref_list, got_list)):
        r, g = np.asarray(r, dtype=np.float64).reshape(-1, 6), np.asarray(g, dtype=np.float64).reshape(-1, 6)
        count_diff.append(len(g) - len(r))
        sel = np.nonzero(r[:, 4] > score_floor)[0]
        if len(sel) == 0:
            continue
        n_ref += len(sel)
        if len(g) == 0:
            unmatched.extend((i, float(r[k, 4])) for k in sel)
            continue
        iou = box_iou_np(r[sel, :4], g[:, :4])
        iou[r[sel, 5][:, None] != g[None, :, 5]] = -1.0
        used = np.zeros(len(g), dtype=bool)
        for k in np.argsort(-r[sel, 4], kind="stable"):
            row = np.where(used, -1.0, iou[k])
            j = int(np.argmax(row))
            if row[j] >= match_iou:
                used[j] = True
                n_hit += 1
                dbox.append(np.abs(r[sel[k], :4] - g[j, :4]).max())
                dscore.append(abs(r[sel[k], 4] - g[j, 4]))
            else:
                unmatched.append((i, float(r[sel[k], 4])))
    dbox, dscore = np.asarray(dbox), np.asarray(dscore)
    return {
        "images": len(ref_list),
        "ref_detections_total": int(sum(len(np.asarray(r).reshape(-1, 6)) for r in ref_list)),
        "got_detections_total": int(sum(len(np.asarray(g).reshape(-1, 6)) for g in got_list)),
        "ref_detections_above_floor": int(n_ref), "matched": int(n_hit),
        "matched_rate": float(n_hit / n_ref) if n_ref else 1.0,
        "max_dbox_px": float(dbox.max()) if len(dbox) else 0.0,
        "p99_dbox_px": float(np.quantile(dbox, 0.99)) if len(dbox) else 0.0,
        "mean_dbox_px": float(dbox.mean()) if len(dbox) else 0.0,
        "frac_dbox_le_0p5": float((dbox <= 0.5).mean()) if len(dbox) else 1.0,
        "frac_dbox_le_2": float((dbox <= 2.0).mean()) if len(dbox) else 1.0,
        "median_dbox_px": float(np.median(dbox)) if len(dbox) else 0.0,
        "max_dscore": float(dscore.max()) if len(dscore) else 0.0,
        "mean_dscore": float(dscore.mean()) if len(dscore) else 0.0,
        "images_with_count_diff": int(sum(1 for c in count_diff if c)),
        "max_abs_count_diff": int(max((abs(c) for c in count_diff), default=0)),
        "sum_count_diff": int(sum(count_diff)),
        "unmatched_ref_scores": [round(s, 4) for _, s in unmatched][:32],
        "score_floor": score_floor, "match_iou": match_iou,
    }


def record(name: str, stats: dict) -> None:
    """Measured parity numbers go to profiles/parity_<name>.json (committed: DESIGN section 5 quotes them) and, on the
    GPU box, to gpurun_out/ so they come back with the call."""
    for d in (ROOT / "gpurun_out", ROOT / "profiles"):
        try:
            d.mkdir(exist_ok=True)
            (d / f"parity_{name}.json").write_text(json.dumps(stats, indent=1) + "\n")
        except OSError:
            pass


def clip_boxes_np(d: np.ndarray, h: int, w: int) -> np.ndarray:
    """clip_boxes (ultralytics/utils/ops.py:335-354) on [n, 6] rows — what construct_result applies after NMS."""
    d = np.array(d, copy=True)
    d[:, [0, 2]] = d[:, [0, 2]].clip(0, w)
    d[:, [1, 3]] = d[:, [1, 3]].clip(0, h)
    return d

"""Build-container tool: measures, with the CPU oracle, the RMS of every conv's pre-BatchNorm output under
the synthetic weight recipe (specyolo.nn.init.synth_state_dict) and writes them to
spectrogram-yolov11_b200/specyolo/cfg/<cfg>.synth.json.  synth_state_dict scales the BatchNorm running
statistics with these numbers, which is what training would have done: activations stay O(1) through the
~100 layers instead of growing geometrically, so the synthetic detector produces a realistic few-percent of
confident anchors.  Pure test/bench data preparation — not part of the product path.
"""
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))

import specyolo  # noqa: E402
from oracle import yolo_ref  # noqa: E402
from specyolo.nn.init import synth_images, synth_state_dict  # noqa: E402


class CalibRef(yolo_ref.Ref):
    def __init__(self, sd):
        super().__init__(sd)
        self.rms = {}

    def conv(self, x, p, k=1, s=1, g=1, d=1, act=True):
        sd = self.sd
        y = F.conv2d(x, sd[p + ".conv.weight"], None, s, yolo_ref.autopad(k, None, d), d, g)
        r = float(y.pow(2).mean().sqrt())
        self.rms[p] = r
        sd[p + ".bn.running_var"] = sd[p + ".bn.running_var"] * (r * r)
        sd[p + ".bn.running_mean"] = sd[p + ".bn.running_mean"] * r
        y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                         sd[p + ".bn.bias"], False, 0.0, yolo_ref.BN_EPS)
        return F.silu(y) if act else y

    def detect_raw(self, xs, p, nc, legacy=False):
        # final plain convs: record the RMS of (W x) so the synth recipe can normalise the logits
        sd = self.sd
        orig = F.conv2d

        def rec(inp, w, b=None, *a, **k):
            for key in self._final:
                if w is sd[key]:
                    self.rms[key[: -len(".weight")]] = float(orig(inp, w, None, *a, **k).pow(2).mean().sqrt())
            return orig(inp, w, b, *a, **k)

        self._final = [k for k in sd if k.startswith(p) and k.endswith(".2.weight")]
        yolo_ref.F.conv2d = rec
        try:
            return super().detect_raw(xs, p, nc, legacy)
        finally:
            yolo_ref.F.conv2d = orig


def main():
    cases = [("yolo11s_fusion_sand3_new.yaml", "yolo11_fusion_sand3_new.yaml", "s", 2),
             ("yolo11n.yaml", "yolo11.yaml", "n", 80), ("yolo11s.yaml", "yolo11.yaml", "s", 80),
             ("yolo11s_fusion_sand3_new_convHCA.yaml", "yolo11_fusion_sand3_new_convHCA.yaml", "s", 2),
             ("yolo11s_fusion_sand3_new_OMN.yaml", "yolo11_fusion_sand3_new_OMN.yaml", "s", 2),
             ("yolo11s_fusion_sand3_new_GC.yaml", "yolo11_fusion_sand3_new_GC.yaml", "s", 2)]
    only = set(sys.argv[1:])
    for cfg, cfg_file, scale, nc in cases:
        if only and cfg not in only:
            continue
        m = specyolo.DetectionModel(cfg, nc=nc)
        d = yaml.safe_load((ROOT / "spectrogram-yolov11_b200" / "specyolo" / "cfg" / cfg_file).read_text())
        g = yolo_ref.parse_graph(d, scale, nc)
        sd = synth_state_dict(m, seed=0, calibrated=False)
        x = synth_images(2, 320, seed=0)
        R = CalibRef(sd)
        with torch.no_grad():
            yolo_ref.forward(g, R, x)          # a prepared Ref is accepted in place of the state_dict
        out = ROOT / "spectrogram-yolov11_b200" / "specyolo" / "cfg" / (Path(cfg).stem + ".synth.json")
        out.write_text(json.dumps({k: round(v, 5) for k, v in R.rms.items()}, indent=0))
        print(cfg, len(R.rms), "convs; max rms", max(R.rms.values()), "min", min(R.rms.values()))


if __name__ == "__main__":
    main()

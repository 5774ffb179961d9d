"""Pins the CPU oracle (oracle/yolo_ref.py) LAYER BY LAYER against the real reference running in this process: every
top-level layer of the reference model is hooked, the oracle's restatement of that layer is fed the very tensor(s) the
reference layer received, and the outputs must agree element by element.  (The golden fixtures pin whole-model outputs and
per-layer fingerprints; this test localises a disagreement to one block and needs no stored data.)  Needs the reference
package (oracle/_ref or /root/reference): skipped elsewhere."""
from pathlib import Path

import pytest
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
CFG = ROOT / "spectrogram-yolov11_b200" / "specyolo" / "cfg"

CASES = [
    ("yolo11s_fusion_sand3_new.yaml", "yolo11_fusion_sand3_new.yaml", "s", 2),
    ("yolo11s_fusion_sand3_new_convHCA.yaml", "yolo11_fusion_sand3_new_convHCA.yaml", "s", 2),
    ("yolo11s_fusion_sand3_new_OMN.yaml", "yolo11_fusion_sand3_new_OMN.yaml", "s", 2),
    ("yolo11s_fusion_sand3_new_GC.yaml", "yolo11_fusion_sand3_new_GC.yaml", "s", 2),
    ("yolo11n.yaml", "yolo11.yaml", "n", 80),
]


@pytest.mark.reference
@pytest.mark.parametrize("cfg,cfg_file,scale,nc", CASES)
def test_oracle_layers_match_live_reference(cfg, cfg_file, scale, nc):
    from oracle import yolo_ref
    from oracle.ref_loader import import_reference, reference_available

    if not reference_available():
        pytest.skip("needs the reference package (oracle/_ref or /root/reference)")
    ultralytics = import_reference()
    from ultralytics.nn.tasks import DetectionModel as RefModel

    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    sd = synth_state_dict(specyolo.DetectionModel(cfg, nc=nc), seed=7)
    ref = RefModel(str(Path(ultralytics.__file__).parent / "cfg" / "models" / "11" / cfg), nc=nc, verbose=False).eval()
    ref.load_state_dict(sd, strict=True)
    seen = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, idx=idx: seen.__setitem__(idx, (i[0], o))) for idx, m in enumerate(ref.model)]
    x = synth_images(1, 640, seed=3)[:, :, :160, :128].contiguous()
    with torch.no_grad():
        ref(x)
    for h in hooks:
        h.remove()

    graph = yolo_ref.parse_graph(yaml.safe_load((CFG / cfg_file).read_text()), scale, nc)
    assert len(graph) == len(ref.model)
    R = yolo_ref.Ref(sd)
    legacy = not any(L["type"] == "C3k2" for L in graph)
    checked = 0
    with torch.no_grad():
        for L in graph:
            inp, want = seen[L["i"]]
            t, p = L["type"], L["prefix"]
            if not isinstance(L["f"], int):         # list inputs: the outputs of the source layers (Detect overwrites its list in place)
                inp = [seen[L["i"] - 1 if j == -1 else j][1] for j in L["f"]]
            if t == "Conv":
                got = R.conv(inp, p, L["k"], L["s"])
            elif t == "ConvHCA":
                got = R.convhca(inp, p, L["k"], L["s"])
            elif t == "C3k2":
                got = R.c3k2(inp, p, L["n"], L["c3k"])
            elif t == "C3k2GC":
                got = R.c3k2gc(inp, p, L["n"])
            elif t == "C3x":
                got = R.c3x(inp, p)
            elif t == "SPPF":
                got = R.sppf(inp, p, L["k"])
            elif t == "C2PSA":
                got = R.c2psa(inp, p, L["n"])
            elif t == "DDWConv":
                got = R.ddwconv(inp, p, L["k"], L["s"], L["d"])
            elif t == "Fusion":
                got = R.fusion(list(inp), p)
            elif t == "Concat":
                got = torch.cat(list(inp), 1)
            elif t == "nn.Upsample":
                got = torch.nn.functional.interpolate(inp, scale_factor=L["scale"], mode=L["mode"])
            elif t == "Detect":
                raw = R.detect_raw(list(inp), p, L["nc"], legacy)
                got = yolo_ref.detect_decode(raw, (8.0, 16.0, 32.0), L["nc"])
                want = want[0]
            else:
                raise AssertionError(f"layer type {t} not covered")
            scale_ = max(1.0, float(want.abs().max()))
            err = float((got - want).abs().max())
            assert got.shape == want.shape and err <= 2e-5 * scale_, f"layer {L['i']} ({t}): max |diff| {err} at scale {scale_}"
            checked += 1
    assert checked == len(graph)

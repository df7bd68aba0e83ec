"""Sweeps the scheduling knobs of the lighting frame (ilb_option) on the C4 scene and prints one JSON line per setting:
kernel time per frame (CUDA events on the library's stream, 3 warm-up + N timed frames) and whether the lightmap is
bit-identical to the first concurrent setting's (scheduling must never change results).

    python profiles/microbench/light_sweep.py [frames] [scene]      scene: c4 (default) | c5 | c2
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import _abi, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2] if len(sys.argv) > 2 else "c4"
scene = {"c4": scenes.config_c4, "c5": scenes.config_c5_lighting, "c2": scenes.config_c2}[which]()
W, H = scene.width, scene.height
ctx = ib.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
r = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
df = scenes.make_distance_field(ctx, scene)
df.Rasterize(scene.obstructions)
r.DistanceField = df
r.SetGBuffer(scene.gbuffer)
packed = r.build_batches()
out = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")


def run(n):
    with torch.cuda.stream(stream):
        for _ in range(3):
            r.RenderLightingDevice(out.data_ptr(), rows=(0, H), packed=packed)
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            r.RenderLightingDevice(out.data_ptr(), rows=(0, H), packed=packed)
        b.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


settings = [(0, 0, 0, 0, 0)] + [(1,) + s for s in ((2, 2, 1, 3), (2, 2, 0, 0), (2, 2, 1, 0), (2, 2, 0, 3), (1, 3, 2, 2), (2, 1, 1, 4), (1, 2, 2, 3),
                                                   (3, 0, 0, 5), (0, 5, 3, 0), (2, 3, 1, 2), (3, 1, 0, 4))]
first = None
for conc, a, b, c, d in settings:
    for opt, v in zip(range(5), (conc, a, b, c, d)):
        ctx.set_option(opt, v)
    ms = run(frames)
    img = out.clone()
    same = None
    if conc:
        if first is None:
            first = img
        same = bool(torch.equal(first.view(torch.int16), img.view(torch.int16)))
    else:
        seq = img
    diff = None
    if conc and first is not None:
        diff = float(((img.float() - seq.float()).abs() / seq.float().abs().clamp_min(1e-3)).max().item())
    print(json.dumps({"scene": which, "concurrent": conc, "line_ctas": a, "other_ctas": b, "line_helpers": c, "other_helpers": d,
                      "ms_per_frame": round(ms, 4), "bit_identical_to_first_concurrent": same, "max_rel_diff_vs_sequential_half4": diff}), flush=True)

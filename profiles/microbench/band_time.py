"""Times the lighting kernels of each of N equal row bands of the C4 frame on ONE device (what each rank of an N-GPU run does),
with CUDA events on the library's stream: prints per-band ms, their sum and the whole-frame ms.  For A/B runs of kernel variants
(ILB_LIB selects the library) without paying for N GPUs.

    python profiles/microbench/band_time.py [bands=8] [reps=10]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import scenes, sharding  # noqa: E402

bands = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = ib.Context(0)
scene = scenes.config_c4()
r = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
df = scenes.make_distance_field(ctx, scene)
df.Rasterize(scene.obstructions)
r.DistanceField = df
r.SetGBuffer(scene.gbuffer)
packed = r.build_batches()
out = torch.empty((scene.height, scene.width, 4), dtype=torch.float16, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream, device=0)


def timed(rows):
    with torch.cuda.stream(stream):
        for _ in range(2):
            r.RenderLightingDevice(out.data_ptr(), rows=rows, packed=packed)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            r.RenderLightingDevice(out.data_ptr(), rows=rows, packed=packed)
        e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


per = [timed(sharding.row_band(k, bands, scene.height)) for k in range(bands)]
whole = timed((0, scene.height))
print(json.dumps({"bands": bands, "band_ms": [round(t, 4) for t in per], "sum_ms": round(sum(per), 4), "max_ms": round(max(per), 4),
                  "frame_ms": round(whole, 4)}))

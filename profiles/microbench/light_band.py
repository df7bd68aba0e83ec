"""Times the lighting kernels of the loaded library on the C4 scene: the whole 4K frame and a 270-row band (one rank's share at
8 GPUs), with and without programmatic dependent launch of the second pass.  One JSON line.  ILB_LIB selects the library
(e.g. a -DILB_TILE_H=8 build), see profiles/microbench/ab_variant.py.

    python profiles/microbench/light_band.py [frames]"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import _abi, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10
scene = scenes.config_c4()
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


def run(rows, n):
    with torch.cuda.stream(stream):
        for _ in range(3):
            r.RenderLightingDevice(out.data_ptr(), rows=rows, packed=packed)
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            r.RenderLightingDevice(out.data_ptr(), rows=rows, packed=packed)
        b.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {"lib": os.environ.get("ILB_LIB", "default")}
images = {}
for pdl in (0, 1):
    ctx.set_option(_abi.OPT_LIGHT_PDL, pdl)
    res[f"frame_ms_pdl{pdl}"] = round(run((0, H), frames), 4)
    images[pdl] = out.clone()
    bands = [run((k * 270, (k + 1) * 270), frames * 2) for k in range(8)]
    res[f"band_ms_pdl{pdl}"] = [round(b, 4) for b in bands]
    res[f"band_sum_ms_pdl{pdl}"] = round(sum(bands), 4)
res["pdl_bit_identical"] = bool(torch.equal(images[0].view(torch.int16), images[1].view(torch.int16)))
print(json.dumps(res), flush=True)

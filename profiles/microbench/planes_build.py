"""Builds the expanded distance-field planes of the C4 field twice -- with the TMA-staged kernel (default) and with the
one-thread-per-entry kernel (ILB_PLANES_TMA=0) -- and checks that a lightmap rendered from either is bit-identical.  Run under
ncu for the kernel times and DRAM bytes:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:df_planes_build \
        --csv --log-file gpurun_out/planes_build.csv python profiles/microbench/planes_build.py"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import scenes  # noqa: E402

scene = scenes.config_c4()
scene.environment.Lights = scene.environment.Lights[:6]
ctx = ib.Context(0)
r = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
r.SetGBuffer(scene.gbuffer)
images = {}
for tma in ("1", "0"):
    os.environ["ILB_PLANES_TMA"] = tma
    df = scenes.make_distance_field(ctx, scene)
    df.Rasterize(scene.obstructions)
    r.DistanceField = df
    images[tma] = r.RenderLighting(rows=(1000, 1064))          # the first sample of a new field builds its planes
print(json.dumps({"bit_identical": bool(np.array_equal(images["1"], images["0"]))}))

"""Runs a few launches of one hot kernel at the bench sizes, for ncu (never a timing source):

    ncu --set full --clock-control none --import-source on -k regex:light_accumulate -s 4 -c 2 -o gpurun_out/prof_light \
        python profiles/microbench/profile_hot.py light 3
    ncu --set full --clock-control none --import-source on -k regex:particle_step_kernel -s 6 -c 1 -o gpurun_out/prof_part \
        python profiles/microbench/profile_hot.py particles 8

light: C4 frames (each frame = the line-light launch + the sphere / directional launch), optionally rows [r0, r1) only:
`profile_hot.py light 2 0 270`;  particles: updates of the 8 M
particle C5 system (noise table launch + step launch each).
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import scenes  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "light"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = ib.Context(0)
if what == "light":
    scene = scenes.config_c4()
    r = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
    df = scenes.make_distance_field(ctx, scene)
    df.Rasterize(scene.obstructions)
    r.DistanceField = df
    r.SetGBuffer(scene.gbuffer)
    packed = r.build_batches()
    out = torch.empty((scene.height, scene.width, 4), dtype=torch.float16, device="cuda")
    rows = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, scene.height)   # a rank's row band
    for _ in range(count):
        r.RenderLightingDevice(out.data_ptr(), rows=rows, packed=packed)
    ctx.synchronize()
else:
    chunk, nchunks = 512, 32
    field_scene = scenes.lighting_scene(1, 1920, 1080, 0)
    pdf = scenes.make_distance_field(ctx, field_scene, resolution=0.25)
    pdf.Rasterize(field_scene.obstructions)
    ps = scenes.particle_scene(2, chunk * chunk * nchunks, chunk, 1920, 1080, steps_hint=1000, collision_field=pdf, spawn_rate=0.0)
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=0xB200))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=nchunks)
    system.Transforms = ps.transforms
    system.Spawn(ps.positions, ps.velocities, ps.attributes)
    now = 0.0
    for _ in range(count):
        now += ps.dt
        system.Update(now, ps.dt)
    ctx.synchronize()
print("done", what, count)

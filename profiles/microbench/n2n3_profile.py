"""Kernel-level timing of the N3 resolve and N2 rasteriser kernels at the bench sizes (4K frame, 8M particles), for ncu:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/n2n3_r1.csv python profiles/microbench/n2n3_profile.py

Also prints CUDA-event timings of back-to-back launches (8 per event pair, so the Python launch overhead that dominates a
single 4K resolve is amortised)."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import _abi, hdr, scenes  # noqa: E402

W, H = 3840, 2160
ctx = ib.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
out = {}


def timed(fn, reps, inner):
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(reps):
            a.record(stream)
            for _ in range(inner):
                fn()
            b.record(stream)
            ctx.synchronize()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / inner)
    return best


# ---- N3 resolve: half4 lightmap + Color albedo -> Color, tone-mapped; two buffer sets alternate (2 x 133 MB > L2)
cfg = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.2, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0))
rp = hdr.pack_resolve(W, H, _abi.FORMAT_HALF4, cfg, _abi.FORMAT_RGBA8, _abi.FORMAT_RGBA8)
lms = [(torch.rand((H, W, 4), device="cuda") * 2).half() for _ in range(2)]
als = [torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
outs = [torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
flip = [0]


def resolve():
    i = flip[0] = flip[0] ^ 1
    ctx.check(ctx.lib.ilb_resolve_lighting_device(ctx.handle, C.byref(rp), C.c_void_p(lms[i].data_ptr()), C.c_void_p(als[i].data_ptr()),
                                                  C.c_void_p(outs[i].data_ptr())))


ms = timed(resolve, 5, 8)
out["resolve_ms"] = ms
out["resolve_GBps_algorithmic"] = 16 * W * H / (ms * 1e-3) / 1e9
rp2 = hdr.pack_resolve(W, H, _abi.FORMAT_HALF4, None, _abi.FORMAT_RGBA8, _abi.FORMAT_RGBA8)


def resolve_plain():
    i = flip[0] = flip[0] ^ 1
    ctx.check(ctx.lib.ilb_resolve_lighting_device(ctx.handle, C.byref(rp2), C.c_void_p(lms[i].data_ptr()), None, C.c_void_p(outs[i].data_ptr())))


ms = timed(resolve_plain, 5, 8)
out["resolve_no_albedo_ms"] = ms
out["resolve_no_albedo_GBps_algorithmic"] = 12 * W * H / (ms * 1e-3) / 1e9
del lms, als, outs

# ---- N2 render: the bench's 8M particles (one update so that RenderColor / RenderData exist) -> 4K half4 target, additive
chunk, nchunks = 512, 32
count = chunk * chunk * nchunks
ps = scenes.particle_scene(2, count, chunk, 1920, 1080, steps_hint=1000, collision_field=None, spawn_rate=0.0)
engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=0xB200))
system = ib.ParticleSystem(engine, ps.configuration, maxChunks=nchunks)
system.Transforms = []
system.Spawn(ps.positions, ps.velocities, ps.attributes)
system.Update(ps.dt, ps.dt)
params = system.render_params(W, H, "Additive", ib.ParticleRenderParameters(Scale=(2.0, 2.0)), clearColor=(0, 0, 0, 0), target_format=_abi.FORMAT_HALF4)
target = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")


def render():
    ctx.check(ctx.lib.ilb_particles_render_device(system.handle, C.byref(params), None, C.c_void_p(target.data_ptr())))


ms = timed(render, 3, 1)
out["render_ms"] = ms
out["render_Mparticles_per_s"] = count / (ms * 1e-3) / 1e6
print(json.dumps(out))

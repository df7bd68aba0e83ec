#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: `python tests/golden/make_golden.py`).

The reference ships no golden vectors and cannot run here, so these fixtures are REGRESSION vectors of the oracle
(pinned by tests/test_oracle_kat.py), not reference outputs: they freeze the oracle's behaviour so that a change to
oracle/ or to the synthetic scenes is visible, and give the GPU tests inputs + expected outputs that travel to the
GPU box.  Inputs are regenerated from the seeds; only the expected outputs are stored (float16-safe sizes)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import illuminant_b200 as ib  # noqa: E402
from illuminant_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def golden_lighting():
    s = scenes.lighting_scene(77, 96, 64, 3, n_directional=1, n_line=1, n_probes=16, ramp=(40.0, 120.0), ao=True, float4_lightmap=True)
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, s.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField, r._gbuffer_shape = df, s.gbuffer.shape[:2]
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    lm = oracle.render_lighting(tex, s.gbuffer, frame, batches, nb, verts, nv)
    pos = np.array([list(p.Position) + [1.0] for p in s.probes], np.float32)
    nrm = np.array([list(p.Normal) + [1.0] for p in s.probes], np.float32)
    probes = oracle.update_light_probes(tex, frame, batches, nb, verts, nv, pos, nrm)
    return {"df_crc": np.array([int(tex.astype(np.uint64).sum())], np.uint64), "df_corner": tex[:8, :8].copy(), "lightmap": lm, "probes": probes}


def golden_particles():
    s = scenes.lighting_scene(78, 128, 96, 0)
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, s.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    ps = scenes.particle_scene(78, 3000, 64, 128, 96, steps_hint=20, collision_field=df, spawn_rate=12000.0)
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=64, RandomSeed=78))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=2)
    system.Transforms = ps.transforms
    system._chunk_next_offset = [64 * 64]      # chunk 0 is the user chunk holding the initial particles
    per = 64 * 64
    P, V, A = (np.zeros((per, 4), np.float32) for _ in range(3))
    P[:3000], V[:3000], A[:3000] = ps.positions, ps.velocities, ps.attributes
    now = 0.0
    for _ in range(6):
        now += ps.dt
        spawns, ops, u = system.plan_spawns(now, ps.dt), system.plan_ops(now), system.system_uniforms(ps.dt)
        live = system.LiveChunkCount
        if P.shape[0] < live * per:
            P, V, A = (np.concatenate([a, np.zeros((live * per - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        P, V, A, RC, RD = oracle.particles_step(P, V, A, 64, u, spawns, ops, engine.RandomnessTexture, tex, 1)
    return {"P": P, "V": V, "A": A, "RC": RC, "RD": RD}


if __name__ == "__main__":
    out = Path(__file__).resolve().parent
    np.savez_compressed(out / "lighting_96x64.npz", **golden_lighting())
    np.savez_compressed(out / "particles_64.npz", **golden_particles())
    print("wrote", [p.name for p in out.glob("*.npz")])

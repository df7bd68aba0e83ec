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


def golden_next_rows():
    """Regression vectors of the oracle for the "next" rows: N3 resolve + luminance, N2 render, N4 spawner sources."""
    from illuminant_b200 import _abi
    out = {}
    rs = np.random.RandomState(79)
    # N3: tone-mapped resolve with an sRGB Color albedo of a half4 lightmap, and level 1 of its luminance chain
    lm = (rs.rand(16, 32, 4) ** 2 * 3).astype(np.float16)
    al = rs.randint(0, 256, size=(16, 32, 4), dtype=np.uint8)
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.3, Gamma=0.9, InverseScaleFactor=0.5, AlbedoIsSRGB=True,
                              ToneMapping=ib.ToneMappingConfiguration(WhitePoint=2.8))
    out["resolve"] = oracle.resolve_lighting(ib.pack_resolve(32, 16, _abi.FORMAT_HALF4, hdr, _abi.FORMAT_RGBA8, _abi.FORMAT_FLOAT4), lm, al)
    out["luminance1"] = oracle.compute_luminance(lm, 1)
    # N2: 200 rotated, rounded, textured quads blended over a grey target
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16, RandomSeed=79))
    cfg = ib.ParticleSystemConfiguration()
    cfg.Size = (1.5, 1.0)
    tex = rs.randint(0, 256, size=(8, 16, 4), dtype=np.uint8)
    cfg.Appearance = ib.ParticleAppearance(Texture=tex, SizePx=(4, 4), AnimationRate=(1.5, 0.0), Rounded=True, RelativeSize=False)
    system = ib.ParticleSystem(engine, cfg)
    n = 200
    P = np.zeros((n, 4), np.float32)
    P[:, 0], P[:, 1], P[:, 3] = rs.rand(n) * 48, rs.rand(n) * 32, np.where(rs.rand(n) < 0.1, 0, rs.rand(n) * 3 + 0.1)
    RD = np.zeros((n, 4), np.float32)
    RD[:, 0], RD[:, 1], RD[:, 3] = rs.rand(n) * 4 + 1, rs.rand(n) * 12 - 2, rs.randint(0, 2, n)
    RC = rs.rand(n, 4).astype(np.float32)
    RC[:, :3] *= RC[:, 3:4]
    r = system.render_params(48, 32, "AlphaBlend")
    out["render"] = oracle.particles_render(P, RD, RC, r, texture=tex, target=np.full((32, 48, 4), 0.25, np.float32))
    # N4: one frame of a 7-position polygon spawner (position texture) and of a pattern spawner
    def fresh():
        c = ib.ParticleSystemConfiguration()
        c.Friction, c.LifeDecayPerSecond = 0.0, 0.0
        return ib.ParticleSystem(ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16, RandomSeed=79)), c, maxChunks=2)
    sysA = fresh()
    poly = ib.Spawner(MinRate=6000, MaxRate=6000, Seed=79, Position=ib.Formula(Constant=(1.0, 2.0, 3.0), RandomScale=(2, 2, 0), Type=ib.FormulaType.Spherical),
                      Velocity=ib.Formula(RandomScale=(10, 10, 2), Type=ib.FormulaType.Spherical), Life=(4.0, 1.0, 0),
                      AdditionalPositions=[(10.0 * i, 5.0 * (i % 3), float(i)) for i in range(1, 7)], PolygonRate=2.5, RatePerPosition=False,
                      VelocityAlongPolygon=(3.0, 1.0, 0.0), AlphaDiscardThreshold=0.0)
    sysA.Transforms = [poly]
    spawns = sysA.plan_spawns(1.0, 1 / 60.0)
    Z = np.zeros((256 * sysA.LiveChunkCount, 4), np.float32)
    out["spawn_polygon"] = np.concatenate(oracle.particles_step(Z, Z, Z, 16, sysA.system_uniforms(1 / 60.0), spawns, [], sysA.Engine.RandomnessTexture,
                                                                sources=sysA.last_sources)[:3], axis=1)
    sysB = fresh()
    pat = ib.PatternSpawner(MinRate=600, MaxRate=600, Seed=79, Texture=rs.randint(64, 256, size=(10, 12, 4), dtype=np.uint8), Divisor=2, WholeSpawn=True,
                            Position=ib.Formula(Constant=(50.0, 40.0, 0.0)), Velocity=ib.Formula(Type=ib.FormulaType.Linear), Life=(2.0, 0, 0),
                            AlphaDiscardThreshold=1.0)
    sysB.Transforms = [pat]
    spawns = sysB.plan_spawns(1.0, 1 / 60.0)      # 640 requested: two RunSpawner passes fill two chunks of 256
    Z = np.zeros((256 * sysB.LiveChunkCount, 4), np.float32)
    out["spawn_pattern"] = np.concatenate(oracle.particles_step(Z, Z, Z, 16, sysB.system_uniforms(1 / 60.0), spawns, [], sysB.Engine.RandomnessTexture,
                                                                sources=sysB.last_sources)[:3], axis=1)
    return out


if __name__ == "__main__":
    out = Path(__file__).resolve().parent
    np.savez_compressed(out / "lighting_96x64.npz", **golden_lighting())
    np.savez_compressed(out / "particles_64.npz", **golden_particles())
    np.savez_compressed(out / "next_rows.npz", **golden_next_rows())
    print("wrote", [p.name for p in out.glob("*.npz")])

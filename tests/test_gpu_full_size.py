"""Parity and size-independent properties at BASELINE.json's full configuration sizes (C2, C3, C4, C5 particle half)."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes, sharding
from helpers import LIGHTING_RTOL, check_particles, lighting_rel_err, make_renderer, oracle_lightmap

pytestmark = pytest.mark.gpu


def test_c2_1080p_32_sphere_lights_full_frame_parity(ctx, oracle):
    s = scenes.config_c2()
    s.configuration.Float4Lightmap = True
    r, tex = make_renderer(ctx, s)
    gpu = r.RenderLighting()
    ref = oracle_lightmap(oracle, r, tex, s)
    err = lighting_rel_err(gpu, ref)
    assert err.max() <= LIGHTING_RTOL, f"C2 max rel err {err.max():.3e} at {np.unravel_index(np.argmax(err), err.shape)}"
    lit = gpu[..., 3] > 0
    assert 0.3 < lit.mean() <= 1.0                        # the scene really lights a large part of the frame
    assert np.array_equal(gpu[..., 3], ref[..., 3])       # alpha = number of lights touching each pixel, exactly
    # HighQuality (half4) output is the rounding of the same fp32 values
    s.configuration.Float4Lightmap = False
    half = r.RenderLighting()
    d = np.abs(half.view(np.uint16).astype(np.int32) - oracle.float_to_half(ref).view(np.uint16).astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3


def test_c4_4k_128_mixed_lights_band_parity_and_sharding(ctx, oracle):
    s = scenes.config_c4()
    s.configuration.Float4Lightmap = True
    r, tex = make_renderer(ctx, s)
    full = r.RenderLighting()
    assert full.shape == (2160, 3840, 4) and np.isfinite(full).all()
    # parity on a strided sample of the WHOLE frame, all 128 lights: both rows at every tile edge (16 k - 1, 16 k) and the
    # row in the middle of every tile row (16 k + 7) -- 404 of the 2160 rows, 1.55 M pixels -- plus three 24-row bands
    worst, alpha_mismatch, rows_checked = 0.0, 0, 0
    bands = [(0, 24), (1068, 1092), (2136, 2160)]
    for k in range(0, 2160, 16):
        bands.append((max(k - 1, 0), k + 1))
        bands.append((k + 7, k + 8))
    for r0, r1 in bands:
        ref = oracle_lightmap(oracle, r, tex, s, rows=(r0, r1))
        err = lighting_rel_err(full[r0:r1], ref)
        worst = max(worst, float(err.max()))
        alpha_mismatch += int((full[r0:r1, :, 3] != ref[..., 3]).sum())
        rows_checked += r1 - r0
        assert err.max() <= LIGHTING_RTOL, f"C4 rows [{r0},{r1}): max rel err {err.max():.3e}"
    assert rows_checked >= 470 and alpha_mismatch == 0, (rows_checked, alpha_mismatch, worst)   # light counts per pixel: exact
    # the 8-GPU row bands of bench.py reproduce the full frame bit for bit
    bands = [r.RenderLighting(rows=sharding.row_band(k, 8, 2160)) for k in range(8)]
    assert np.array_equal(np.concatenate(bands, axis=0), full)
    # every pixel is touched by the 8 directional + 24 line lights; sphere lights add to the count
    assert full[..., 3].min() >= 8 and full[..., 3].max() <= 128
    # probes (config 4's GI probes) against the oracle
    gpu_p = r.UpdateLightProbes(float4=True)
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    pos = np.array([list(p.Position) + [1.0] for p in s.probes], np.float32)
    nrm = np.array([list(p.Normal) + [1.0] for p in s.probes], np.float32)
    assert lighting_rel_err(gpu_p, oracle.update_light_probes(tex, frame, batches, nb, verts, nv, pos, nrm)).max() <= LIGHTING_RTOL


def _particle_system(ctx, count, chunk, nchunks, field, seed, spawn_rate=60000.0):
    ps = scenes.particle_scene(seed, count, chunk, 1920, 1080, steps_hint=1000, collision_field=field, spawn_rate=spawn_rate)
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=0xB200))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=nchunks)
    system.Transforms = ps.transforms
    system.Spawn(ps.positions, ps.velocities, ps.attributes)
    return ps, engine, system


def test_c3_1m_particles_chain_parity_and_long_run_properties(ctx, oracle):
    fs = scenes.lighting_scene(1, 1920, 1080, 0)
    df = scenes.make_distance_field(ctx, fs, resolution=0.25)          # quarter-res collision field (SimpleParticles.cs:216-219)
    df.Rasterize(fs.obstructions)
    tex = df.Save()
    count, chunk = 1 << 20, 256
    ps, engine, system = _particle_system(ctx, count, chunk, 20, df, 2)
    per = chunk * chunk
    P, V, A = ps.positions.copy(), ps.velocities.copy(), ps.attributes.copy()
    now = 0.0
    for _ in range(8):                                                # parity over 8 full steps incl. the spawner
        now += ps.dt
        spawns, ops, u = system.plan_spawns(now, ps.dt), system.plan_ops(now), system.system_uniforms(ps.dt)
        live = system.LiveChunkCount
        if P.shape[0] < live * per:
            P, V, A = (np.concatenate([a, np.zeros((live * per - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        system.step_packed(u, spawns, ops, 1)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, chunk, u, spawns, ops, engine.RandomnessTexture, tex, 1)
    assert system.LiveChunkCount == 17
    for c in (0, 7, 15, 16):
        g = system.ReadChunk(c)
        sl = slice(c * per, (c + 1) * per)
        check_particles(g, (P[sl], V[sl], A[sl], RC[sl], RD[sl]), f"chunk {c}")
    assert system.LiveCount == int((P[:, 3] > 0).sum())
    # long run (C3 = 1000 steps; 300 here keeps the test short): life decays monotonically, nothing goes NaN, dead stays zero
    before = system.ReadChunk(3)[0][:, 3].copy()
    for _ in range(300):
        now += ps.dt
        system.Update(now, ps.dt)
    p, v, a, rc, rd = system.ReadChunk(3)
    assert np.isfinite(p).all() and np.isfinite(v).all() and np.isfinite(rc).all() and np.isfinite(rd).all()
    alive = p[:, 3] > 0
    assert (p[alive, 3] < before[alive]).all()
    assert (p[~alive] == 0).all() and (v[~alive] == 0).all() and (rc[~alive] == 0).all()
    assert np.allclose(p[alive, 3], before[alive] - 1.2 * 300 * ps.dt, atol=2e-3)      # no LifePenalty configured: exactly linear decay


def test_c5_8m_particles_chunk_sharding_is_exact(ctx):
    """C5's particle half at full size (8M = 32 chunks x 512^2): stepping chunk ranges separately (what each GPU does)
    gives bit-identical state to stepping all chunks in one system."""
    fs = scenes.lighting_scene(1, 1920, 1080, 0)
    df = scenes.make_distance_field(ctx, fs, resolution=0.25)
    df.Rasterize(fs.obstructions)
    count, chunk, nchunks = 1 << 23, 512, 32
    ps, engine, whole = _particle_system(ctx, count, chunk, nchunks, df, 2, spawn_rate=0.0)
    ops, u = whole.plan_ops(ps.dt), whole.system_uniforms(ps.dt)
    whole.step_packed(u, [], ops, 3)
    per = chunk * chunk
    c0, c1 = sharding.chunk_range(5, 8, nchunks)                      # rank 5 of 8 owns chunks [20, 24)
    part = ib.ParticleSystem(engine, ps.configuration, maxChunks=c1 - c0)
    sl = slice(c0 * per, c1 * per)
    part.Spawn(ps.positions[sl], ps.velocities[sl], ps.attributes[sl])
    part.step_packed(u, [], ops, 3)
    for k in range(c1 - c0):
        for a, b in zip(part.ReadChunk(k), whole.ReadChunk(c0 + k)):
            assert np.array_equal(a, b)
    assert whole.LiveCount == count


def test_c5_combined_frame_loop_matches_oracle_frame_by_frame(ctx, oracle):
    """Config C5's loop (TestGame/TestGame/Scenes/ParticleLights.cs:333-378) at a size the oracle finishes in seconds: every frame
    updates the particle system (Spawner + Gravity + Noise + FMA, collision against the LIGHTING field) and then renders the
    lighting, whose light list includes one sphere light per live particle (ParticleLightSource) read from the state the update
    just produced.  Both halves are compared with the oracle every frame, so a stale or mis-ordered hand-over would show."""
    from illuminant_b200._abi import LightBatch, LightVertex
    s = scenes.lighting_scene(61, 288, 176, 4, n_directional=1, n_line=1, ramp=(60.0, 180.0), float4_lightmap=True)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    tex = df.Save()
    chunk, n0 = 32, 600
    ps = scenes.particle_scene(61, n0, chunk, 288, 176, steps_hint=40, collision_field=df, spawn_rate=1800.0)
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=61))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=4)
    system.Transforms = ps.transforms
    system.Spawn(ps.positions, ps.velocities, ps.attributes)
    template = ib.SphereLightSource(Radius=2.0, RampLength=18.0, Color=(0.8, 0.6, 0.4, 0.5), CastsShadows=True)
    pls = ib.ParticleLightSource(Template=template, System=system)
    s.environment.Lights.append(pls)
    r = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r.DistanceField = df
    r.SetGBuffer(s.gbuffer)
    per = chunk * chunk
    P, V, A = (np.zeros((per, 4), np.float32) for _ in range(3))
    P[:n0], V[:n0], A[:n0] = ps.positions, ps.velocities, ps.attributes
    now = 0.0
    for frame_index in range(3):
        now += ps.dt
        spawns, ops, u = system.plan_spawns(now, ps.dt), system.plan_ops(now), system.system_uniforms(ps.dt)
        live = system.LiveChunkCount
        if P.shape[0] < live * per:
            P, V, A = (np.concatenate([a, np.zeros((live * per - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        system.step_packed(u, spawns, ops, 1)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, chunk, u, spawns, ops, engine.RandomnessTexture, tex, 1)
        gpu_state = [np.concatenate(x) for x in zip(*[system.ReadChunk(c) for c in range(live)])]
        check_particles(gpu_state, (P, V, A, RC, RD), f"frame {frame_index}")
        lightmap = r.RenderLighting()
        frame = r.build_frame()
        batches, nb, verts, nv = r.build_batches()
        pv = pls.light_vertices(P, A, True)                    # the oracle's particle state -> the oracle's particle lights
        assert len(pv) > 300
        allv = (LightVertex * (nv + len(pv)))(*([verts[i] for i in range(nv)] + pv))
        allb = (LightBatch * (nb + 1))(*[batches[i] for i in range(nb)])
        allb[nb].light_type, allb[nb].first_vertex, allb[nb].vertex_count = 3, nv, len(pv)
        allb[nb].df = r._df_uniforms(None)
        ref = oracle.render_lighting(tex, s.gbuffer, frame, allb, nb + 1, allv, nv + len(pv))
        err = lighting_rel_err(lightmap, ref)
        assert err.max() <= LIGHTING_RTOL, f"frame {frame_index}: max rel err {err.max():.3e}"
        assert np.array_equal(lightmap[..., 3], ref[..., 3])
    # another renderer of the same context without particle lights must not inherit them (they are registered per context)
    s.environment.Lights.remove(pls)
    r2 = ib.LightingRenderer(ctx, s.environment, s.configuration)
    r2.DistanceField = df
    r2.SetGBuffer(s.gbuffer)
    plain = r2.RenderLighting()
    ref = oracle_lightmap(oracle, r2, tex, s)
    assert lighting_rel_err(plain, ref).max() <= LIGHTING_RTOL and np.array_equal(plain[..., 3], ref[..., 3])

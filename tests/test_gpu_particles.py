"""GPU parity of the particle hot path (P1-P10) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from helpers import check_particles

pytestmark = pytest.mark.gpu


def _run_both(ctx, oracle, ps, chunk_size, steps, tex=None, seed=3, max_chunks=4, transforms=None, life_ramp=None):
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk_size, RandomSeed=seed))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=max_chunks)
    system.Transforms = ps.transforms if transforms is None else transforms
    per = chunk_size * chunk_size
    system.Spawn(ps.positions, ps.velocities, ps.attributes)
    n0 = system.LiveChunkCount * per
    P, V, A = (np.zeros((n0, 4), np.float32) for _ in range(3))
    P[:ps.count], V[:ps.count], A[:ps.count] = ps.positions, ps.velocities, ps.attributes
    RC = RD = None
    now = 0.0
    for _ in range(steps):
        now += ps.dt
        spawns, ops, u = system.plan_spawns(now, ps.dt), system.plan_ops(now), system.system_uniforms(ps.dt)
        live = system.LiveChunkCount
        if P.shape[0] < live * per:
            P, V, A = (np.concatenate([a, np.zeros((live * per - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        system.step_packed(u, spawns, ops, 1)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, chunk_size, u, spawns, ops, engine.RandomnessTexture, tex, 1, life_ramp=life_ramp)
    gpu = [np.concatenate(x) for x in zip(*[system.ReadChunk(c) for c in range(system.LiveChunkCount)])]
    return system, gpu, (P, V, A, RC, RD)


def _check(gpu, ref, what=""):
    check_particles(gpu, ref, what)


def test_ballistic_no_transforms_closed_form(ctx, oracle):
    ps = scenes.particle_scene(40, 5000, 128, 512, 512, steps_hint=16)
    ps.configuration.Friction = 0.0
    system, gpu, ref = _run_both(ctx, oracle, ps, 128, 10, transforms=[])
    _check(gpu, ref, "ballistic")
    # p(t) = p0 + v t, life linear (friction 0, no transforms, no field)
    t = 10 * ps.dt
    want = ps.positions[:, :3] + ps.velocities[:, :3] * np.float32(t)
    assert np.abs(gpu[0][:5000, :3] - want).max() < 2e-3
    assert np.allclose(gpu[0][:5000, 3], ps.positions[:, 3] - 1.2 * t, atol=1e-4)
    assert (gpu[0][5000:] == 0).all()


def test_full_chain_with_collision_and_spawner(ctx, oracle):
    s = scenes.lighting_scene(41, 384, 256, 0)
    df = scenes.make_distance_field(ctx, s, resolution=0.5)
    df.Rasterize(s.obstructions)
    tex = df.Save()
    ps = scenes.particle_scene(41, 30000, 128, 384, 256, steps_hint=40, collision_field=df, spawn_rate=90000.0)
    system, gpu, ref = _run_both(ctx, oracle, ps, 128, 12, tex=tex, max_chunks=6)
    assert system.LiveChunkCount >= 3          # the spawner opened at least one new chunk
    _check(gpu, ref, "chain")
    assert system.LiveCount == int((ref[0][:, 3] > 0).sum())


def test_full_field_addressing_extension(ctx, oracle):
    s = scenes.lighting_scene(42, 256, 256, 0)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    tex = df.Save()
    ps = scenes.particle_scene(42, 20000, 256, 256, 256, steps_hint=40, collision_field=df, spawn_rate=0.0)
    ps.configuration.Collision.FullFieldAddressing = True
    ps.configuration.Collision.LifePenalty = 0.05
    _, gpu, ref = _run_both(ctx, oracle, ps, 256, 8, tex=tex, max_chunks=1)
    _check(gpu, ref, "full-field")


def test_each_transform_alone_and_areas(ctx, oracle):
    base = scenes.particle_scene(43, 9000, 128, 400, 300, steps_hint=20)
    area = ib.TransformArea(Type=ib.AreaType.Box, Center=(200, 150, 0), Size=(120, 80, 50), Falloff=40.0, Rotation=0.3)
    variants = {
        "gravity": [base.transforms[1]],
        "noise_replace": [ib.Noise(VelocityScale=(30, 30, 30), SpeedScale=4.0, PositionScale=(3, 3, 0, 0), Seed=5, Area=area)],
        "noise_add": [ib.Noise(VelocityScale=(30, 30, 30), ReplaceOldVelocity=False, Interval=50.0, Seed=6)],
        "fma": [ib.FMA(PositionAdd=(1, 2, 0), VelocityMultiply=(0.9, 0.9, 1.0), CyclesPerSecond=None, Area=area)],
        "matrix": [ib.MatrixMultiply(Position=(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 2, 1, 0, 1),
                                     Velocity=(0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1), CategoryFilter=(0, 0),
                                     Area=ib.TransformArea(Type=ib.AreaType.Cylinder, Center=(200, 150, 0), Size=(90, 90, 40), Falloff=10, Rotation=1.0))],
        "octagon_spheroid": [ib.FMA(VelocityAdd=(3, 0, 0), Area=ib.TransformArea(Type=ib.AreaType.Octagon, Center=(200, 150, 0), Size=(100, 60, 30), Falloff=20, Rotation=0.7)),
                             ib.FMA(VelocityAdd=(0, 3, 0), Area=ib.TransformArea(Type=ib.AreaType.Spheroid, Center=(100, 100, 0), Size=(50, 80, 30), Falloff=20, Rotation=0.2))],
    }
    for name, tr in variants.items():
        _, gpu, ref = _run_both(ctx, oracle, base, 128, 5, transforms=tr)
        _check(gpu, ref, name)


def test_spawner_formulas_and_polygon(ctx, oracle):
    ps = scenes.particle_scene(44, 100, 128, 300, 300, steps_hint=30)
    for kind, sp in {
        "rect_towards": ib.Spawner(MinRate=20000, MaxRate=40000, Seed=9,
                                   Position=ib.Formula(Constant=(150, 150, 0), RandomScale=(40, 40, 0), Offset=(30, 20, 0), Type=ib.FormulaType.Rectangular),
                                   Velocity=ib.Formula(Constant=(150, 150, 0), RandomScale=(20, 20, 0), Offset=(5, 5, 0), Type=ib.FormulaType.Towards),
                                   Life=(2.0, 1.0, -0.5), Category=(0.0, 3.0, 0.0), ColorRandomScale=(0.5, 0.5, 0.5, 1.0), ColorOffset=(0, 0, 0, -0.4),
                                   AlphaDiscardThreshold=20.0),
        "polygon": ib.Spawner(MinRate=30000, MaxRate=30000, Seed=10, AdditionalPositions=[(250, 50, 0), (250, 250, 0)], PolygonRate=7.0,
                              PolygonLoop=True, VelocityAlongPolygon=(10.0, 5.0, 0.0), AlignVelocityAndPosition=True,
                              Position=ib.Formula(Constant=(50, 50, 0), RandomScale=(3, 3, 0), Offset=(0, 0, 0), Type=ib.FormulaType.Spherical),
                              Velocity=ib.Formula(RandomScale=(10, 10, 0), Offset=(2, 2, 0), Type=ib.FormulaType.Spherical), AxisMask=(1, 1, 0)),
    }.items():
        _, gpu, ref = _run_both(ctx, oracle, ps, 128, 6, transforms=[sp], max_chunks=3)
        _check(gpu, ref, kind)
        assert (gpu[0][:, 3] > 0).sum() > 1000


def test_render_outputs_off_and_beziers(ctx, oracle):
    ps = scenes.particle_scene(45, 8000, 128, 300, 300, steps_hint=20)
    ps.configuration.ColorFromLife = ib.Bezier4V(Count=4, MinValue=0.0, MaxValue=3.0, A=(1, 0, 0, 0), B=(1, 1, 0, 1), C=(0, 1, 1, 1), D=(0, 0, 1, 0.2))
    ps.configuration.SizeFromVelocity = ib.BezierF(Count=2, MinValue=0.0, MaxValue=60.0, A=0.5, B=3.0, Mode=1)
    ps.configuration.SizeFromLife = ib.BezierF(Count=3, MinValue=4.0, MaxValue=1.0, A=1.0, B=2.0, C=4.0)
    ps.configuration.ColorFromVelocity = ib.Bezier4V(Count=2, Mode=256 + 2, MinValue=0, MaxValue=20, A=(1, 1, 1, 1), B=(0.5, 0.5, 0.5, 1))
    _, gpu, ref = _run_both(ctx, oracle, ps, 128, 4)
    _check(gpu, ref, "beziers")
    ps.configuration.WriteRenderOutputs = False
    _, gpu2, ref2 = _run_both(ctx, oracle, ps, 128, 4)
    _check(gpu2[:3], ref2[:3], "no render outputs")
    assert (gpu2[3] == 0).all()


def test_steps_argument_equals_repeated_calls(ctx):
    ps = scenes.particle_scene(46, 4000, 128, 300, 300, steps_hint=20)
    out = []
    for mode in (0, 1):
        engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=128, RandomSeed=1))
        system = ib.ParticleSystem(engine, ps.configuration, maxChunks=1)
        system.Spawn(ps.positions, ps.velocities, ps.attributes)
        ops = [t.pack(system, 0.0) for t in ps.transforms[1:] if not isinstance(t, ib.Noise)]
        u = system.system_uniforms(ps.dt)
        if mode == 0:
            system.step_packed(u, [], ops, 5)
        else:
            for _ in range(5):
                system.step_packed(u, [], ops, 1)
        out.append(system.ReadChunk(0))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_errors(ctx):
    ps = scenes.particle_scene(47, 100, 64, 100, 100)
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=64))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=1)
    system.Spawn(ps.positions, ps.velocities, ps.attributes)
    ps.configuration.Collision = ib.ParticleCollision(DistanceField=None)
    u = system.system_uniforms(ps.dt)
    u.has_collision_field = 1
    with pytest.raises(ib.IlluminantError) as e:   # collision without a field (ParticleSystem.cs:834-836)
        system.step_packed(u, [], [], 1)
    assert e.value.code == -4
    g = ib.Gravity(Attractors=[ib.Attractor()] * 17)
    with pytest.raises(ib.IlluminantError):        # Transforms.cs:348-349
        g.pack(system, 0.0)
    with pytest.raises(ib.IlluminantError):
        system.Spawn(ps.positions, ps.velocities, ps.attributes)   # out of chunks


def test_tma_staged_kernel_is_bit_identical_to_direct_kernel(ctx, monkeypatch):
    """ILB_PARTICLE_TMA=1 routes the specialised chains through the persistent kernel that stages particle chunks with
    bulk async copies (cp.async.bulk + mbarrier) and stores results with bulk async stores: same bits as the direct kernel."""
    s = scenes.lighting_scene(48, 256, 256, 0)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    ps = scenes.particle_scene(48, 3 * 128 * 128 - 77, 128, 256, 256, steps_hint=40, collision_field=df, spawn_rate=0.0)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("ILB_PARTICLE_TMA", flag)
        engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=128, RandomSeed=1))
        system = ib.ParticleSystem(engine, ps.configuration, maxChunks=3)
        system.Transforms = ps.transforms
        system.Spawn(ps.positions, ps.velocities, ps.attributes)
        ops, u = system.plan_ops(ps.dt), system.system_uniforms(ps.dt)
        system.step_packed(u, [], ops, 7)
        u.has_collision_field = 0                      # also the no-collision specialisation
        system.step_packed(u, [], [], 2)
        out.append([system.ReadChunk(c) for c in range(3)])
    for a, b in zip(out[0], out[1]):
        for i, (x, y) in enumerate(zip(a, b)):
            if i < 3:
                assert np.array_equal(x, y)          # particle state (position, velocity, attributes): bit-identical
            else:
                # render colour / data are computed with contracted (FMA) arithmetic, which the compiler may schedule
                # differently in the two kernels: a few ulp
                assert np.allclose(x, y, rtol=2e-6, atol=1e-6)


def test_planes_match_atlas_bit_for_bit(ctx, monkeypatch):
    """Collision against the expanded planes (csrc/planes.cu) == collision against the Rgba64 atlas, bit for bit, for
    the reference's flat addressing and for the FullFieldAddressing extension."""
    s = scenes.lighting_scene(49, 256, 256, 0)
    for full in (False, True):
        out = []
        for flag in ("1", "0"):
            monkeypatch.setenv("ILB_NO_PLANES", flag)
            df = scenes.make_distance_field(ctx, s)
            df.Rasterize(s.obstructions)
            ps = scenes.particle_scene(49, 2 * 128 * 128, 128, 256, 256, steps_hint=40, collision_field=df, spawn_rate=0.0)
            ps.configuration.Collision.FullFieldAddressing = full
            engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=128, RandomSeed=1))
            system = ib.ParticleSystem(engine, ps.configuration, maxChunks=2)
            system.Transforms = ps.transforms
            system.Spawn(ps.positions, ps.velocities, ps.attributes)
            ops, u = system.plan_ops(ps.dt), system.system_uniforms(ps.dt)
            system.step_packed(u, [], ops, 9)
            out.append([system.ReadChunk(c) for c in range(2)])
        for a, b in zip(out[0], out[1]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def _check_nan_aware(gpu, ref):
    check_particles(gpu, ref, "degenerate", allow_nan=True)


def test_degenerate_vectors_take_the_ieee_fallback(ctx, oracle):
    """Operands outside the fast window of the guarded square roots / reciprocals of the fast chains -- squared lengths
    that are tiny but not zero, or overflow -- must come out of the IEEE re-evaluation like the oracle; exactly-zero
    vectors (a particle on an attractor, live particles at rest) stay on the fast path through the zero select.
    (A build with -DILB_BREAK_FALLBACK=1 fails this test: it does reach stepParticleExact.)"""
    s = scenes.lighting_scene(51, 256, 256, 0)
    df = scenes.make_distance_field(ctx, s)
    df.Rasterize(s.obstructions)
    tex = df.Save()
    # (a) the Gravity -> Noise -> FMA chain: an attractor at the origin with particles 1e-17 away from it, particles
    #     exactly on the other attractors, particles at rest
    ps = scenes.particle_scene(51, 6000, 128, 256, 256, steps_hint=40, collision_field=df, spawn_rate=0.0)
    gravity = ps.transforms[1]
    gravity.Attractors[0] = ib.Attractor(Position=(0.0, 0.0, 0.0), Radius=150.0, Strength=400.0, Type=ib.AttractorType.Linear)
    ps.positions[0:60, :3] = np.float32(1e-17)
    for k, a in enumerate(gravity.Attractors[1:], start=1):
        ps.positions[100 * k:100 * k + 40, :3] = np.asarray(a.Position, np.float32)
    ps.velocities[1000:1400, :3] = 0.0
    _, gpu, ref = _run_both(ctx, oracle, ps, 128, 3, tex=tex, max_chunks=1)
    _check_nan_aware(gpu, ref)
    # (b) the empty chain with collision: |v|^2 below / above the window
    ps = scenes.particle_scene(52, 6000, 128, 256, 256, steps_hint=40, collision_field=df, spawn_rate=0.0)
    ps.velocities[0:300, :3] = np.float32(1e-17)
    ps.velocities[300:600, :3] = np.float32(3e19)
    ps.velocities[600:900, :3] = 0.0
    _, gpu, ref = _run_both(ctx, oracle, ps, 128, 3, tex=tex, max_chunks=1, transforms=[])
    _check_nan_aware(gpu, ref)


def test_life_ramp_texture(ctx, oracle):
    """getRampedColorForLifeValueAndIndex with a LifeRampTexture (UpdateCommon.fxh:6-13,67-80): POINT sampled, U clamped,
    V wrapped by the particle index; normal and inverted, partial strength."""
    rs = np.random.RandomState(7)
    for invert, strength, shape in ((False, 1.0, (4, 16, 4)), (True, 0.6, (3, 7, 4))):
        ps = scenes.particle_scene(53, 9000, 128, 300, 300, steps_hint=20)
        ramp = rs.uniform(0.0, 1.0, shape).astype(np.float32)
        ps.configuration.LifeRamp = ib.ParticleColorLifeRamp(Minimum=0.5, Maximum=3.0, Strength=strength, Invert=invert, Texture=ramp)
        _, gpu, ref = _run_both(ctx, oracle, ps, 128, 4, life_ramp=ramp)
        _check(gpu, ref, f"life ramp invert={invert}")
        fresh = scenes.particle_scene(53, 9000, 128, 300, 300, steps_hint=20)   # same seed, fresh host-side spawner state
        _, plain, _ = _run_both(ctx, oracle, fresh, 128, 4)
        assert not np.allclose(gpu[3], plain[3])      # the ramp really changes renderColor
        assert np.array_equal(gpu[0], plain[0])       # ... and nothing else


def test_dead_chunks_are_reaped_so_a_continuous_spawner_never_runs_out(ctx):
    """chunks were only ever appended, so a continuous Spawner stopped for good after MaxChunks * ChunkSize^2
    spawns.  With the reference's liveness check every LivenessCheckInterval frames and reaping after DeadFrameThreshold empty
    results (ParticleLiveness.cs:14-129, ParticleSystem.cs:675) a short-lived stream keeps spawning for ever."""
    from illuminant_b200.particles import Formula, Spawner
    chunk = 16
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=5))
    cfg = ib.ParticleSystemConfiguration(LifeDecayPerSecond=1.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=4)           # capacity 4 * 256 = 1024 particles
    system.DeadFrameThreshold = 2
    system.Transforms = [Spawner(MinRate=3000.0, MaxRate=3000.0, Seed=1, Position=Formula(Constant=(50.0, 50.0, 0.0), RandomScale=(20.0, 20.0, 0.0)),
                                 Velocity=Formula(Constant=(1.0, 0.0, 0.0)), Life=(0.12, 0.0, 0.0), ColorConstant=(1.0, 1.0, 1.0, 1.0))]
    dt, now, peak = 1 / 60.0, 0.0, 0
    for frame in range(400):
        now += dt
        system.Update(now, dt)
        if frame % 7 == 0:
            system._poll_liveness(wait=True)                      # a slow consumer would see the counts a few frames late; either works
        peak = max(peak, system.LiveChunkCount)
    spawner = system.Transforms[0]
    assert spawner.TotalSpawned > 6 * 1024, spawner.TotalSpawned   # up to 50 per frame for 400 frames: far past the capacity of 1024
    assert system.ReapedChunkCount >= 15 and peak <= 4
    live = system.LiveCount
    assert 100 <= live <= 420, live                                # ~7 frames of life at up to 50 per frame stay alive
    # the reaped slots were zeroed and the survivors kept their order: every live chunk holds only live-or-zero texels
    for c in range(system.LiveChunkCount):
        p, v, a, rc, rd = system.ReadChunk(c)
        dead = p[:, 3] <= 0
        assert np.isfinite(p).all() and (p[dead][:, 3] <= 0).all()
    system.Clear()
    system.Update(now + dt, dt)
    assert system.LiveChunkCount <= 1 and system.LiveCount <= 60   # only this update's spawn survives the Clear


def test_removing_a_chunk_keeps_the_order_of_the_others(ctx):
    chunk = 16
    engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=5))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=5)
    per = chunk * chunk
    rs = np.random.RandomState(3)
    P = rs.uniform(1.0, 9.0, (4 * per, 4)).astype(np.float32)
    V = rs.uniform(-1.0, 1.0, (4 * per, 4)).astype(np.float32)
    A = rs.uniform(0.0, 1.0, (4 * per, 4)).astype(np.float32)
    system.Spawn(P, V, A)
    system._sync_chunk_lists()
    system._reap_chunk(1)
    assert system.LiveChunkCount == 3
    for new, old in enumerate((0, 2, 3)):
        p, v, a, _, _ = system.ReadChunk(new)
        sl = slice(old * per, (old + 1) * per)
        assert np.array_equal(p, P[sl]) and np.array_equal(v, V[sl]) and np.array_equal(a, A[sl])
    p, v, a, rc, rd = system.ReadChunk(3)                           # the vacated slot is zeroed for the next CreateChunk
    assert not p.any() and not v.any() and not a.any()
    assert system.LiveCount == 3 * per


def test_stale_handles_are_rejected_not_dereferenced():
    """ilb_destroy releases a context's fields and particle systems; a later call through any stale handle must come back as
    ILB_ERR_INVALID_ARGUMENT instead of touching freed memory (a Python ParticleSystem can outlive its Context)."""
    import ctypes as C
    from illuminant_b200 import _abi
    ctx2 = ib.Context(0)
    lib = ctx2.lib
    ps, df = C.c_void_p(), C.c_void_p()
    assert lib.ilb_particles_create(ctx2.handle, 16, 2, C.byref(ps)) == 0
    assert lib.ilb_df_create_empty(ctx2.handle, 32, 32, C.byref(df)) == 0
    stale_ctx = C.c_void_p(ctx2.handle.value)
    ctx2.close()
    count, n, v = C.c_int64(0), C.c_int(0), C.c_int(0)
    u = _abi.PsysUniforms()
    assert lib.ilb_particles_count_live(ps, C.byref(count)) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_particles_step(ps, C.byref(u), None, 0, None, 0, 1) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_particles_set_live_chunks(ps, 1) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_particles_request_chunk_liveness(ps) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_particles_remove_chunk(ps, 0) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_particles_device_buffer(ps, 0) is None
    buf = (C.c_uint16 * (32 * 32 * 4))()
    assert lib.ilb_df_download(df, buf, 32 * 32 * 8) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_synchronize(stale_ctx) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_set_option(stale_ctx, 0, 1) == _abi.ERR_INVALID_ARGUMENT
    assert lib.ilb_gbuffer_upload(stale_ctx, 4, 4, 0, buf) == _abi.ERR_INVALID_ARGUMENT
    lib.ilb_particles_destroy(ps)      # no-ops
    lib.ilb_df_destroy(df)
    lib.ilb_destroy(stale_ctx)

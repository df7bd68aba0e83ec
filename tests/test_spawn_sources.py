"""N4 spawn sources: SpawnParticlesFromPositionTexture and SpawnFeedbackParticles (SpawnParticles.fx:32-120).
CPU part: closed-form known answers for the oracle restatement.  GPU part (marked): the CUDA kernels against the oracle through
ilb_particles_step_sources."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi
from helpers import check_particles

f32 = np.float32
CS = 32          # chunk size
PER = CS * CS


def _engine(ctx=None):
    return ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=CS, RandomSeed=7))


def _system(engine, max_chunks=2):
    cfg = ib.ParticleSystemConfiguration()
    cfg.Friction = 0.0
    cfg.LifeDecayPerSecond = 0.0
    return ib.ParticleSystem(engine, cfg, maxChunks=max_chunks)


def _source_state(seed=11, live=700):
    rs = np.random.RandomState(seed)
    P = np.zeros((PER, 4), np.float32)
    V = np.zeros((PER, 4), np.float32)
    RC = np.zeros((PER, 4), np.float32)
    P[:live, :3] = rs.rand(live, 3) * 200
    P[:live, 3] = rs.rand(live) * 5 + 0.05
    P[50:60, 3] = 0                         # dead source particles: outside SourceLifeRange -> nothing spawned
    V[:live, :3] = rs.randn(live, 3) * 10
    V[:live, 3] = rs.randint(0, 3, live)
    RC[:live] = rs.rand(live, 4)
    return P, V, RC


def _polygon_spawner(n_positions, **kw):
    pts = [(10.0 * i, 5.0 * (i % 3), float(i)) for i in range(1, n_positions)]
    return ib.Spawner(MinRate=60 * 300, MaxRate=60 * 300, Position=ib.Formula(Constant=(1.0, 2.0, 3.0), RandomScale=(0, 0, 0)),
                      Velocity=ib.Formula(Constant=(0, 0, 0), RandomScale=(0, 0, 0), Type=ib.FormulaType.Linear), Life=(4.0, 0, 0),
                      AdditionalPositions=pts, RatePerPosition=False, AlphaDiscardThreshold=0.0, **kw)


def _plan(system, spawner, now=1.0, dt=1 / 60.0):
    system.Transforms = [spawner]
    spawns = system.plan_spawns(now, dt)
    return spawns, system.last_sources, system.system_uniforms(dt)


# ------------------------------------------------------------------------------------------------------------- CPU / oracle
def test_position_texture_polygon_known_answers(oracle):
    engine = _engine()
    system = _system(engine)
    sp = _polygon_spawner(7, PolygonRate=4.0, PolygonLoop=True, VelocityAlongPolygon=(3.0, 0, 0))
    spawns, sources, u = _plan(system, sp)
    assert len(spawns) == 1 and sources[0].kind == _abi.SPAWN_POSITION_TEXTURE and spawns[0].PositionConstantCount == 7
    n = int(spawns[0].ChunkSizeAndIndices.z) + 1
    assert n == 300
    Z = np.zeros((PER, 4), np.float32)
    P, V, A, _, _ = oracle.particles_step(Z, Z, Z, CS, u, spawns, [], engine.RandomnessTexture, sources=sources)
    pts = np.array([(1.0, 2.0, 3.0)] + list(sp.AdditionalPositions), np.float32)
    k = np.arange(n)
    i1, t = (k // 4) % 7, (k % 4) / 4.0            # PolygonRate 4: four particles per edge, index w = 0 (TotalSpawned 0)
    i2 = (i1 + 1) % 7
    want = pts[i1] + (pts[i2] - pts[i1]) * t[:, None].astype(np.float32)
    # life constant 4 minus nothing (LifeDecay 0); position advanced by one step of the along-polygon velocity
    d = pts[i2] - pts[i1]
    vel = 3.0 * d / np.sqrt((d * d).sum(1))[:, None]
    assert np.allclose(V[:n, :3], vel, rtol=1e-5, atol=1e-5)
    assert np.allclose(P[:n, :3], want + vel / 60.0, rtol=1e-5, atol=1e-4)
    assert np.all(P[:n, 3] == 4.0) and np.all(P[n:] == 0)
    # without a source array the same spawn is rejected (5+ positions do not fit the inline constants)
    with pytest.raises(RuntimeError):
        oracle.particles_step(Z, Z, Z, CS, u, spawns, [], engine.RandomnessTexture)


def test_feedback_spawn_known_answers(oracle):
    engine = _engine()
    source, target = _system(engine), _system(engine)
    source.handle = 1
    sP, sV, sRC = _source_state()
    source._chunk_next_offset = [700]
    source._sync_chunk_lists()
    fs = ib.FeedbackSpawner(MinRate=60 * 400, MaxRate=60 * 400, SourceSystem=source, InstanceMultiplier=2, SourceVelocityFactor=0.5,
                            MultiplyLife=True, MultiplyColorConstant=True, Position=ib.Formula(Constant=(1.0, -2.0, 0.5)),
                            Velocity=ib.Formula(Constant=(3.0, 0, 0), Type=ib.FormulaType.Linear), Life=(2.0, 0, 0), Category=(5.0, 0, 0),
                            ColorConstant=(0.5, 1.0, 0.25, 1.0), AlphaDiscardThreshold=0.0, SourceLifeRange=(0.0, 9999.0))
    spawns, sources, u = _plan(target, fs)
    n = int(spawns[0].ChunkSizeAndIndices.z) + 1
    assert n == 400 and sources[0].FeedbackSourceIndex == 0
    Z = np.zeros((PER, 4), np.float32)
    P, V, A, _, _ = oracle.particles_step(Z, Z, Z, CS, u, spawns, [], engine.RandomnessTexture, sources=sources,
                                          source_states=[(sP, sV, sRC, CS)])
    src = np.arange(n) // 2                          # InstanceMultiplier 2: two new particles per source particle
    alive = sP[src, 3] > 0
    assert alive.sum() == n - 20 and np.all(P[:n][~alive] == 0) and np.all(A[:n][~alive] == 0)
    vel = np.array([3.0, 0, 0], np.float32) + sV[src, :3] * f32(0.5)
    assert np.allclose(V[:n][alive][:, :3], vel[alive], rtol=1e-6, atol=1e-6)
    assert np.allclose(V[:n][alive][:, 3], 5.0 + sV[src, 3][alive] * 0.5)            # category constant + source w * factor
    want = np.array([1.0, -2.0, 0.5], np.float32) + sP[src, :3] + vel / f32(60.0)     # AlignPositionConstant + one Euler step
    assert np.allclose(P[:n][alive][:, :3], want[alive], rtol=1e-5, atol=1e-4)
    assert np.allclose(P[:n][alive][:, 3], 2.0 * sP[src, 3][alive], rtol=1e-6)        # MultiplyLife
    assert np.allclose(A[:n][alive], np.array([0.5, 1.0, 0.25, 1.0], np.float32) * sRC[src][alive], rtol=1e-6)   # x RenderColor
    # SourceLifeRange filters on the source particle's life
    sources[0].SourceLifeRange[:] = [1.0, 3.0]
    P2, _, _, _, _ = oracle.particles_step(Z, Z, Z, CS, u, spawns, [], engine.RandomnessTexture, sources=sources, source_states=[(sP, sV, sRC, CS)])
    inrange = (sP[src, 3] > 1.0) & (sP[src, 3] < 3.0)
    assert np.array_equal(P2[:n, 3] != 0, inrange)


# ------------------------------------------------------------------------------------------------------------------- GPU
def _check(gpu, ref, what):
    check_particles(gpu, ref, what)


@pytest.mark.gpu
@pytest.mark.parametrize("loop,rate", [(True, 4.0), (False, 2.5), (True, None)])
def test_gpu_position_texture_spawner(ctx, oracle, loop, rate):
    engine = _engine(ctx)
    system = _system(engine)
    sp = _polygon_spawner(9, PolygonRate=rate, PolygonLoop=loop, VelocityAlongPolygon=(3.0, 2.0, 0.5))
    sp.Position.RandomScale, sp.Position.Type = (4.0, 4.0, 1.0), ib.FormulaType.Spherical
    sp.Velocity = ib.Formula(Constant=(1, 0, 0), RandomScale=(20, 20, 5), Type=ib.FormulaType.Spherical)
    P = V = A = np.zeros((0, 4), np.float32)
    now = 0.0
    for _ in range(3):
        now += 1 / 60.0
        spawns, sources, u = _plan(system, sp, now)
        live = system.LiveChunkCount
        if P.shape[0] < live * PER:
            P, V, A = (np.concatenate([a, np.zeros((live * PER - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        system.step_packed(u, spawns, [], 1, sources)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, CS, u, spawns, [], engine.RandomnessTexture, sources=sources)
    gpu = [np.concatenate(x) for x in zip(*[system.ReadChunk(c) for c in range(system.LiveChunkCount)])]
    assert (gpu[0][:, 3] > 0).sum() == 900
    _check(gpu, (P, V, A, RC, RD), f"position texture loop={loop} rate={rate}")


@pytest.mark.gpu
@pytest.mark.parametrize("multiplier,entire", [(1, False), (3, False), (2, True)])
def test_gpu_feedback_spawner(ctx, oracle, multiplier, entire):
    engine = _engine(ctx)
    source, target = _system(engine), _system(engine)
    sP, sV, sRC = _source_state()
    sA = np.ones((PER, 4), np.float32)
    source.Spawn(sP[:700], sV[:700], sA[:700])
    u0 = source.system_uniforms(1 / 60.0)
    source.step_packed(u0, [], [], 1)                       # one update so that RenderColor (the attribute source) is written
    srcP, srcV, _, srcRC, _ = source.ReadChunk(0)
    fs = ib.FeedbackSpawner(MinRate=60 * 240, MaxRate=60 * 240, SourceSystem=source, InstanceMultiplier=multiplier,
                            SpawnFromEntireWindow=entire, SlidingWindowMargin=5, SourceVelocityFactor=0.75, MultiplyLife=True,
                            MultiplyColorConstant=True, Position=ib.Formula(Constant=(0, 0, 0), RandomScale=(6, 6, 2), Type=ib.FormulaType.Spherical),
                            Velocity=ib.Formula(Constant=(0, 0, 0), RandomScale=(30, 30, 30), Offset=(5, 5, 5), Type=ib.FormulaType.Towards),
                            Life=(1.5, 0.5, 0), ColorConstant=(0.9, 0.8, 0.7, 1.0), ColorRandomScale=(0.1, 0.1, 0.1, 0.0),
                            AlphaDiscardThreshold=1.0, SourceLifeRange=(0.5, 4.5))
    P = V = A = np.zeros((0, 4), np.float32)
    now = 0.0
    for _ in range(2):
        now += 1 / 60.0
        spawns, sources, u = _plan(target, fs, now)
        assert len(spawns) == 1 and sources[0].source_chunk == 0
        live = target.LiveChunkCount
        if P.shape[0] < live * PER:
            P, V, A = (np.concatenate([a, np.zeros((live * PER - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        target.step_packed(u, spawns, [], 1, sources)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, CS, u, spawns, [], engine.RandomnessTexture, sources=sources,
                                                source_states=[(srcP, srcV, srcRC, CS)])
    gpu = [np.concatenate(x) for x in zip(*[target.ReadChunk(c) for c in range(target.LiveChunkCount)])]
    assert 0 < (gpu[0][:, 3] > 0).sum() < 480            # some source particles fall outside SourceLifeRange
    _check(gpu, (P, V, A, RC, RD), f"feedback x{multiplier} entire={entire}")


@pytest.mark.gpu
def test_gpu_spawn_source_errors(ctx):
    engine = _engine(ctx)
    system, other = _system(engine), _system(engine)
    fs = ib.FeedbackSpawner(MinRate=6000, MaxRate=6000, SourceSystem=other)
    other.Spawn(*(np.ones((200, 4), np.float32) for _ in range(3)))
    spawns, sources, u = _plan(system, fs)
    assert len(spawns) == 1
    sources[0].source_chunk = 7
    with pytest.raises(ib.IlluminantError) as e:
        system.step_packed(u, spawns, [], 1, sources)
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT and "source chunk" in str(e.value)
    sources[0].source_chunk = 0
    sources[0].source_system = system.handle            # a system cannot feed itself
    with pytest.raises(ib.IlluminantError):
        system.step_packed(u, spawns, [], 1, sources)
    sp = _polygon_spawner(6)
    spawns, sources, u = _plan(system, sp, 2.0)
    with pytest.raises(ib.IlluminantError) as e:         # 6 positions without the position texture
        system.step_packed(u, spawns, [], 1, None)
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT


# ------------------------------------------------------------------------------------------------ PatternSpawner
def _pattern_texture(w, h, seed=5):
    rs = np.random.RandomState(seed)
    t = rs.randint(0, 256, size=(h, w, 4), dtype=np.uint8)
    t[..., 3] = np.where(rs.rand(h, w) < 0.2, 0, 255)      # transparent pixels spawn nothing (alpha below the discard threshold)
    return t


def test_pattern_spawner_known_answers(oracle):
    engine = _engine()
    system = _system(engine)
    tex = _pattern_texture(8, 4)
    sp = ib.PatternSpawner(MinRate=120, MaxRate=120, MaximumTotal=1, Texture=tex, WholeSpawn=True, Divisor=1,
                           Position=ib.Formula(Constant=(100.0, 50.0, 0.0)), Velocity=ib.Formula(Type=ib.FormulaType.Linear),
                           Life=(3.0, 0, 0), ColorConstant=(1.0, 0.5, 1.0, 1.0), AlphaDiscardThreshold=1.0)
    assert (sp.ParticlesPerRow, sp.RowsPerInstance, sp.CountScale) == (8, 4, 32)
    spawns, sources, u = _plan(system, sp)
    assert len(spawns) == 1 and sources[0].kind == _abi.SPAWN_PATTERN
    s, src = spawns[0], sources[0]
    assert (s.ChunkSizeAndIndices.y, s.ChunkSizeAndIndices.z) == (0, 31) and sp.RateError == 0      # the instant whole spawn (:148-166)
    assert src.StepWidthAndSizeScale.tuple() == (1, 8, 1 / 8, 1 / 4) and tuple(src.CenteringOffset) == (-4.0, -2.0)
    assert src.TexelOffsetAndMipBias.tuple() == (-0.5 / 8, -0.5 / 4, 0, -0.5) and src.YOffsetsAndCoordScale.tuple() == (0, 0, 1, 1)
    Z = np.zeros((PER, 4), np.float32)
    P, V, A, _, _ = oracle.particles_step(Z, Z, Z, CS, u, spawns, [], engine.RandomnessTexture, sources=sources)
    ix, iy = np.arange(32) % 8, np.arange(32) // 8
    # texCoord = index / size - half a texel: exactly the centre of texel (index - 1), CLAMP at the border
    texel = tex[np.maximum(iy - 1, 0), np.maximum(ix - 1, 0)].astype(np.float32) / np.float32(255)
    spawned = texel[:, 3] >= 1 / 255
    assert np.array_equal(P[:32, 3] > 0, spawned)
    want = np.stack([100.0 + ix - 4.0, 50.0 + iy - 2.0, 0 * ix], 1)
    assert np.allclose(P[:32][spawned][:, :3], want[spawned]) and np.all(P[:32][spawned][:, 3] == 3.0)
    assert np.allclose(A[:32][spawned], texel[spawned] * np.array([1.0, 0.5, 1.0, 1.0], np.float32), rtol=1e-6)
    assert system.plan_spawns(2.0, 1 / 60.0) == []             # MaximumTotal reached


def test_pattern_spawner_rows_and_rates():
    engine = _engine()
    system = _system(engine)
    tex = _pattern_texture(20, 12)
    sp = ib.PatternSpawner(MinRate=60, MaxRate=60, Texture=tex, Divisor=2, TextureTopLeftPx=(2, 2))
    assert sp.DirectTextureSize == (18.0, 10.0) and (sp.ParticlesPerRow, sp.RowsPerInstance) == (16, 8)
    system.Transforms = [sp]
    rows = []
    for k in range(10):
        spawns = system.plan_spawns(k / 60.0, 1 / 60.0)       # 60/s * CountScale 16 / 60 = one row per frame
        assert len(spawns) == 1 and spawns[0].ChunkSizeAndIndices.z - spawns[0].ChunkSizeAndIndices.y == 15
        src = system.last_sources[0]
        rows.append(int(src.YOffsetsAndCoordScale.x))
        assert src.TexelOffsetAndMipBias.w == pytest.approx(0.5) and src.StepWidthAndSizeScale.tuple() == (2, 16, pytest.approx(0.1), pytest.approx(1 / 6))
        assert src.TexelOffsetAndMipBias.x == pytest.approx(-0.5 / 20 + 2 / 20)
    assert rows == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]               # RowsSpawned % RowsPerInstance (:205-209)
    # a row never straddles a chunk: PartialSpawnAllowed is false (:137-141) -- 1024 / 16 = 64 rows fill chunk 0 exactly
    for k in range(10, 70):
        system.plan_spawns(k / 60.0, 1 / 60.0)
    assert system.LiveChunkCount == 2 and system._chunk_next_offset == [1024, 96]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,divisor,whole,topleft,multiply", [(8, 4, 1, True, None, True), (20, 12, 2, False, (2, 2), False),
                                                               (64, 48, 3, True, None, True), (31, 17, 4, False, (1, 0), True)])
def test_gpu_pattern_spawner(ctx, oracle, w, h, divisor, whole, topleft, multiply):
    engine = _engine(ctx)
    system = _system(engine, max_chunks=3)
    tex = _pattern_texture(w, h, seed=w + divisor)
    sp = ib.PatternSpawner(MinRate=600, MaxRate=600, Texture=tex, WholeSpawn=whole, Divisor=divisor, TextureTopLeftPx=topleft,
                           MultiplyColorConstant=multiply, Position=ib.Formula(Constant=(300.0, 200.0, 4.0), RandomScale=(1.5, 1.5, 0.0)),
                           Velocity=ib.Formula(Constant=(0, 0, 0), RandomScale=(12, 12, 3), Type=ib.FormulaType.Spherical),
                           Life=(2.0, 1.0, 0), ColorConstant=(0.8, 0.9, 1.0, 1.0) if multiply else (0.05, 0.0, 0.1, 0.0),
                           ColorRandomScale=(0.1, 0.1, 0.1, 0.0), AlphaDiscardThreshold=8.0)
    P = V = A = np.zeros((0, 4), np.float32)
    now, total = 0.0, 0
    for _ in range(4):
        now += 1 / 60.0
        spawns, sources, u = _plan(system, sp, now)
        total += len(spawns)
        live = system.LiveChunkCount
        if P.shape[0] < live * PER:
            P, V, A = (np.concatenate([a, np.zeros((live * PER - a.shape[0], 4), np.float32)]) for a in (P, V, A))
        system.step_packed(u, spawns, [], 1, sources)
        P, V, A, RC, RD = oracle.particles_step(P, V, A, CS, u, spawns, [], engine.RandomnessTexture, sources=sources)
    assert total >= 1
    gpu = [np.concatenate(x) for x in zip(*[system.ReadChunk(c) for c in range(system.LiveChunkCount)])]
    assert (gpu[0][:, 3] > 0).sum() == (P[:, 3] > 0).sum() > 0
    _check(gpu, (P, V, A, RC, RD), f"pattern {w}x{h} /{divisor}")


def test_pattern_spawner_rows_rebuild_the_image(oracle):
    """Row-by-row mode end to end on the CPU: after RowsPerInstance frames every texel of the pattern has produced one particle at
    its pixel's position (centred on the spawner's position constant), coloured by that texel."""
    engine = _engine()
    system = _system(engine, max_chunks=1)
    tex = _pattern_texture(8, 8, seed=9)
    tex[..., 3] = 255
    sp = ib.PatternSpawner(MinRate=60, MaxRate=60, Texture=tex, Divisor=1, Position=ib.Formula(Constant=(100.0, 50.0, 0.0)),
                           Velocity=ib.Formula(Type=ib.FormulaType.Linear), Life=(3.0, 0, 0), ColorConstant=(1, 1, 1, 1), AlphaDiscardThreshold=1.0)
    system.Transforms = [sp]
    P = V = A = np.zeros((PER, 4), np.float32)
    for k in range(8):                                     # 60/s * CountScale 8 / 60 = one row of 8 particles per frame
        spawns = system.plan_spawns((k + 1) / 60.0, 1 / 60.0)
        assert len(spawns) == 1 and int(system.last_sources[0].YOffsetsAndCoordScale.x) == k
        P, V, A, _, _ = oracle.particles_step(P, V, A, CS, system.system_uniforms(1 / 60.0), spawns, [], engine.RandomnessTexture,
                                              sources=system.last_sources)
    assert (P[:, 3] > 0).sum() == 64
    ix, iy = np.arange(64) % 8, np.arange(64) // 8
    assert np.allclose(P[:64, :2], np.stack([100.0 + ix - 4.0, 50.0 + iy - 4.0], 1))
    texel = tex[np.maximum(iy - 1, 0), np.maximum(ix - 1, 0)].astype(np.float32) / np.float32(255)     # half-texel offset: texel (index - 1)
    assert np.allclose(A[:64], texel, rtol=1e-6)


def test_feedback_spawner_follows_its_source_across_frames(oracle):
    """Two systems on the CPU: a source that keeps spawning and a feedback spawner that consumes its particles in order, one new
    particle per source particle, at the source particle's position."""
    engine = _engine()
    source, target = _system(engine, max_chunks=1), _system(engine, max_chunks=1)
    emitter = ib.Spawner(MinRate=600, MaxRate=600, Position=ib.Formula(Constant=(0.0, 0.0, 0.0), RandomScale=(100, 100, 0)),
                         Velocity=ib.Formula(Type=ib.FormulaType.Linear), Life=(5.0, 0, 0), AlphaDiscardThreshold=0.0)
    source.Transforms = [emitter]
    fs = ib.FeedbackSpawner(MinRate=300, MaxRate=300, SourceSystem=source, Position=ib.Formula(Constant=(0.0, 0.0, 1.0)),
                            Velocity=ib.Formula(Type=ib.FormulaType.Linear), Life=(1.0, 0, 0), AlphaDiscardThreshold=0.0)
    target.Transforms = [fs]
    sP = sV = sA = np.zeros((PER, 4), np.float32)
    tP = tV = tA = np.zeros((PER, 4), np.float32)
    sRC = np.zeros((PER, 4), np.float32)
    consumed = []
    for k in range(6):
        now = (k + 1) / 60.0
        spawns = source.plan_spawns(now, 1 / 60.0)                      # 10 new source particles per frame
        sP, sV, sA, sRC, _ = oracle.particles_step(sP, sV, sA, CS, source.system_uniforms(1 / 60.0), spawns, [], engine.RandomnessTexture)
        source.handle = 1                                               # what ilb_spawn_source.source_system carries (no device here)
        spawns = target.plan_spawns(now, 1 / 60.0)                      # 5 feedback particles per frame
        assert len(spawns) == 1
        src = target.last_sources[0]
        consumed.append(int(src.FeedbackSourceIndex))
        first = int(spawns[0].ChunkSizeAndIndices.y)
        tP, tV, tA, _, _ = oracle.particles_step(tP, tV, tA, CS, target.system_uniforms(1 / 60.0), spawns, [], engine.RandomnessTexture,
                                                 sources=target.last_sources, source_states=[(sP, sV, sRC, CS)])
        got = tP[first:first + 5, :3]
        want = sP[consumed[-1]:consumed[-1] + 5, :3] + np.array([0, 0, 1], np.float32)
        assert np.allclose(got, want, atol=1e-5), k
    assert consumed == [0, 5, 10, 15, 20, 25] and source.AvailableForFeedback(0) == 60 - 30

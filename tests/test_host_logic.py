"""Host-side mirror of the reference API: LightVertex packing, batching, frame uniforms, spawner accounting, sharding."""
import ctypes as C
import math

import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi, scenes, sharding
from illuminant_b200.lighting import pack_light_vertex

F = np.float32


def test_sphere_light_vertex_packing():    # LightingRenderer.cs:1193-1219 / SURVEY appendix A
    l = ib.SphereLightSource(Position=(1, 2, 3), Radius=4, RampLength=50, RampMode=ib.LightSourceRampMode.Exponential, Color=(0.1, 0.2, 0.3, 0.5),
                             Opacity=0.5, SpecularColor=(0.9, 0.8, 0.7), SpecularPower=3, AmbientOcclusionRadius=6, AmbientOcclusionOpacity=0.25,
                             FalloffYFactor=2, ShadowFilter=ib.ShadowFilter.Shadowed)
    v = pack_light_vertex(l, 2.0, True)
    assert v.LightPosition1.tuple() == v.LightPosition2.tuple() == v.LightPosition3.tuple() == (1, 2, 3, 0)
    assert v.LightProperties.tuple() == (4, 50, 1, 1)
    assert v.MoreLightProperties.tuple() == (6, -99999, 2, 0.25)
    assert v.Color1.tuple() == pytest.approx((0.1, 0.2, 0.3, 0.5 * 0.5 * 2.0))
    assert v.Color2.tuple() == pytest.approx((0.9, 0.8, 0.7, 3))
    assert v.EvenMoreLightProperties.x == 1 and v.EvenMoreLightProperties.z == pytest.approx(-math.pi) and v.EvenMoreLightProperties.w == pytest.approx(1 / (2 * math.pi))
    assert pack_light_vertex(l, 1.0, False).LightProperties.w == 0          # no distance field -> no shadows
    l.Opacity = 0
    assert pack_light_vertex(l, 1.0, True) is None                          # skipped on the host (:1195)


def test_directional_and_line_light_vertex_packing():    # :1256-1307, :1309-1337
    d = ib.DirectionalLightSource(Color=(1, 1, 1, 0.4), ShadowDistanceFalloff=12.0)
    d.Direction = (0, 3, -4)
    v = pack_light_vertex(d, 1.0, True)
    assert v.LightPosition1.tuple() == (-99999, -99999, 0, 0) and v.LightPosition2.tuple() == (99999, 99999, 0, 0)
    assert v.Color2.tuple() == pytest.approx((0, 0.6, -0.8, 1.0))
    assert v.LightProperties.tuple() == (1, 256, 12, 0.5) and v.MoreLightProperties.tuple() == (0, 12, 0, 1)
    d.Direction, d.Bounds = None, ((10, 20), (30, 40))
    v = pack_light_vertex(d, 1.0, True)
    assert v.Color2.tuple() == (0, 0, 0, 0) and v.LightPosition1.tuple() == (10, 20, 0, 0) and v.LightPosition2.tuple() == (30, 40, 0, 0)
    ln = ib.LineLightSource(StartPosition=(0, 0, 5), EndPosition=(100, 0, 5), Radius=8, StartColor=(1, 0, 0, 1), EndColor=(0, 0, 1, 0.5), Opacity=0.5)
    v = pack_light_vertex(ln, 1.0, True)
    assert v.LightProperties.tuple() == (8, 0, 0, 1) and v.Color1.w == 0.5 and v.Color2.w == 0.25 and v.EvenMoreLightProperties.x == 0


def test_batches_group_by_type_and_quality_in_draw_order():
    s = scenes.lighting_scene(0, 64, 64, 3, n_directional=1, n_line=1)
    q2 = ib.RendererQualitySettings(MinStepSize=1.0, LongStepFactor=0.5, OcclusionToOpacityPower=0.7)
    s.environment.Lights[1].Quality = q2
    s.environment.Lights[2].Enabled = False
    s.environment.Lights[0].SortKey = 5                       # sorted after the others (stable)
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    batches, nb, verts, nv = r.build_batches()
    assert (nb, nv) == (4, 4)
    kinds = [(batches[i].light_type, batches[i].first_vertex, batches[i].vertex_count) for i in range(nb)]
    assert kinds == [(1, 0, 1), (2, 1, 1), (4, 2, 1), (1, 3, 1)]
    assert batches[0].df.StepAndMisc2.y == 1.0 and batches[0].df.StepAndMisc2.x == 64      # q2 quality on the first sphere batch
    assert batches[0].df.Extent.x == 0                        # no field bound: `_DistanceField == null` uniforms (:1906-1916)
    assert verts[3].LightPosition1.tuple()[:3] == pytest.approx(s.environment.Lights[0].Position)


def test_frame_uniforms_scale_compensation_and_clear_colour():
    s = scenes.lighting_scene(0, 100, 60, 0)
    s.configuration.RenderScale = (0.5, 0.25)
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r._gbuffer_shape = (15, 50)
    r.ViewportPosition = (3.0, 4.0)
    f = r.build_frame(2.0, rows=(2, 9))
    assert (f.width, f.height, f.row_begin, f.row_end) == (50, 15, 2, 9)
    assert tuple(f.ViewportPosition) == (3.0 + 1.0, 4.0 + 2.0)                        # + 0.5 / RenderScale (:714-720)
    assert f.GBufferTexelSizeAndMisc.tuple() == pytest.approx((1 / 50, 1 / 15, 1, 1))
    assert f.EnvironmentZToY.tuple() == (0, 0, 0, 0)                                  # TwoPointFiveD off -> ZToY 0 (:695-697)
    assert f.ClearColor.tuple() == pytest.approx((0.1, 0.1, 0.16, 0.0))               # ambient * intensity, alpha zeroed (fullbright mode)
    s.configuration.TwoPointFiveD = True
    s.environment.ZToYMultiplier = 2.0
    assert r.build_frame().EnvironmentZToY.tuple() == (2.0, 0.5, 0, 0)
    assert r.lightmap_format == _abi.FORMAT_HALF4
    s.configuration.HighQuality = False
    assert r.lightmap_format == _abi.FORMAT_RGBA8


def test_spawner_rate_accounting_and_indices():    # ParticleSpawner.cs:152-194, ParticleSpawning.cs:115-197
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=3)
    sp = ib.Spawner(MinRate=100, MaxRate=100, Seed=1)
    system.Transforms = [sp]
    total = 0
    for step in range(40):
        spawns = system.plan_spawns(step / 60.0, 1 / 60.0)
        for s in spawns:
            first, last = int(s.ChunkSizeAndIndices.y), int(s.ChunkSizeAndIndices.z)
            assert 0 <= first <= last < 256 and s.ChunkSizeAndIndices.x == 16
            total += last - first + 1
    assert total == sp.TotalSpawned == int(100 * 40 / 60.0)          # the fractional part is carried in RateError
    assert 0 <= sp.RateError < 1
    # a request larger than the free space of the chunk is split: second RunSpawner pass into a new chunk
    system2 = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=3)
    big = ib.Spawner(MinRate=60 * 200, MaxRate=60 * 200, Seed=2)
    system2.Transforms = [big]
    a = system2.plan_spawns(0.0, 1 / 60.0)
    b = system2.plan_spawns(1 / 60.0, 1 / 60.0)
    assert [s.chunk for s in a] == [0] and (a[0].ChunkSizeAndIndices.y, a[0].ChunkSizeAndIndices.z) == (0, 199)
    assert [s.chunk for s in b] == [0, 1] and (b[0].ChunkSizeAndIndices.y, b[0].ChunkSizeAndIndices.z) == (200, 255)
    assert system2.LiveChunkCount == 2
    # MaximumTotal caps the spawner
    capped = ib.Spawner(MinRate=6000, MaxRate=6000, MaximumTotal=150, Seed=3)
    system3 = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=3)
    system3.Transforms = [capped]
    for step in range(5):
        system3.plan_spawns(step / 60.0, 1 / 60.0)
    assert capped.TotalSpawned == 150


def test_spawner_uniform_packing():    # ParticleSpawner.cs:200-256, :361-403
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=32))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=1)
    sp = ib.Spawner(Position=ib.Formula(Constant=(1, 2, 3), RandomScale=(4, 5, 6), Offset=(7, 8, 9), Type=ib.FormulaType.Spherical),
                    Velocity=ib.Formula(Constant=(10, 11, 12), RandomScale=(13, 14, 15), Offset=(16, 17, 18), Type=ib.FormulaType.Linear),
                    Life=(2.0, 0.5, -0.25), Category=(1.0, 2.0, 3.0), AlphaDiscardThreshold=51.0, AdditionalPositions=[(9, 9, 9)],
                    AlignVelocityAndPosition=True, PolygonRate=4.0, PolygonLoop=False)
    sp.Indices = (5, 20)
    sp.TotalSpawned = 10
    s = sp.pack(system, 0.0, 0)
    assert s.Configuration[0].tuple() == (4, 5, 6, 0.5) and s.Configuration[1].tuple() == (7, 8, 9, -0.25)
    assert s.Configuration[2].tuple() == (10, 11, 12, 1) and s.Configuration[4].tuple() == (16, 17, 18, 3)
    assert s.InlinePositionConstants[0].tuple() == (1, 2, 3, 2.0) and s.InlinePositionConstants[1].tuple() == (9, 9, 9, 2.0)
    assert s.FormulaTypes.tuple() == (1, 0, 0, 0) and s.PositionConstantCount == 2
    assert s.AlignVelocityAndPosition == 0.0                         # only if BOTH formulas are circular (:240-242)
    assert s.AttributeDiscardThreshold == pytest.approx(0.2)
    assert s.ChunkSizeAndIndices.tuple() == (32, 5, 20, pytest.approx((10 / 4.0) % 1))   # !PolygonLoop: count - 1 = 1
    assert 0 <= s.RandomnessOffset[0] < 253 and 0 <= s.RandomnessOffset[1] < 127
    # more than MaxInlinePositions = 4 positions: the SpawnParticlesFromPositionTexture material and its PositionBuffer
    # ((count + 127) / 128 * 128 texels of (position, life), ParticleSpawner.cs:306-352, :376-384)
    big = ib.Spawner(Position=ib.Formula(Constant=(1, 2, 3)), Life=(9.0, 0, 0), AdditionalPositions=[(i, 0, -i) for i in range(1, 5)])
    s = big.pack(system, 0.0, 0)
    src = big._source
    assert s.PositionConstantCount == 5 and src.kind == _abi.SPAWN_POSITION_TEXTURE and src.position_count == 128
    assert np.array_equal(big._position_buffer[:6], np.array([[1, 2, 3, 9], [1, 0, -1, 9], [2, 0, -2, 9], [3, 0, -3, 9], [4, 0, -4, 9], [0, 0, 0, 0]], np.float32))
    assert src.positions == big._position_buffer.ctypes.data
    assert ib.Spawner(AdditionalPositions=[(0, 0, 0)] * 3).pack(system, 0.0, 0) is not None and ib.Spawner()._rng is not None


def test_feedback_spawner_bookkeeping():
    """FeedbackSpawner.BeginTick / RunSpawner bookkeeping (SpecialSpawners.cs:325-403, ParticleSpawning.cs:115-197, :246-264)."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    source = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=2)
    target = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=2)
    source.handle = 12345     # packed into ilb_spawn_source.source_system; no device here
    source._chunk_next_offset = [100]      # a chunk with 100 spawned particles that is still the spawn target
    source._sync_chunk_lists()
    source._spawn_target = 0
    fs = ib.FeedbackSpawner(MinRate=600, MaxRate=600, SourceSystem=source, InstanceMultiplier=3, SlidingWindowMargin=10,
                            SourceVelocityFactor=0.5, MultiplyLife=True, SourceLifeRange=(0.1, 50.0))
    target.Transforms = [fs]
    spawns = target.plan_spawns(1.0, 1 / 60.0)           # 600/s * 1/60 = 10 -> 3 instances x 3, one left as rate error
    assert len(spawns) == 1 and target.last_sources is not None
    s, src = spawns[0], target.last_sources[0]
    assert (s.ChunkSizeAndIndices.y, s.ChunkSizeAndIndices.z, s.ChunkSizeAndIndices.w) == (0, 8, 0)
    assert src.kind == _abi.SPAWN_FEEDBACK and src.source_system == 12345 and src.source_chunk == 0
    assert (src.FeedbackSourceIndex, src.InstanceMultiplier, src.SourceVelocityFactor) == (0.0, 3.0, 0.5)
    assert (src.AlignPositionConstant, src.MultiplyLife, src.MultiplyAttributeConstant) == (1.0, 1.0, 0.0)
    assert tuple(src.SourceLifeRange) == (pytest.approx(0.1), 50.0)
    assert fs.RateError == pytest.approx(1.0) and fs.TotalSpawned == 9
    assert source._chunk_consumed[0] == 3 and source.AvailableForFeedback(0) == 97       # consumed 9 / 3 source particles
    assert target._chunk_is_feedback == [True] and target._feedback_spawn_target == 0 and target._spawn_target == -1
    spawns = target.plan_spawns(1.0 + 1 / 60.0, 1 / 60.0)   # 10 + 1 carried = 11 -> 3 instances again, FeedbackSourceIndex advanced
    assert target.last_sources[0].FeedbackSourceIndex == 3.0 and spawns[0].ChunkSizeAndIndices.y == 9
    # sliding window: only the newest 20 source particles are eligible -> the older ones are skipped
    fs.SlidingWindowSize = 20
    target.plan_spawns(1.0 + 2 / 60.0, 1 / 60.0)
    assert target.last_sources[0].FeedbackSourceIndex == 80.0          # 100 - 20
    # a system cannot feed itself (SpecialSpawners.cs:333-335)
    fs.SourceSystem = target
    assert target.plan_spawns(2.0, 1 / 60.0) == [] and target.last_sources is None


def test_transform_packing_and_noise_uv_cycle():
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=32))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=1)
    f = ib.FMA(CyclesPerSecond=None).pack(system, 0.0).u.fma
    assert f.TimeDivisor == -1 and f.PositionMultiply.w == 1 and f.PositionAdd.w == 0 and tuple(f.area.CategoryFilter) == (-9999, 9999)
    assert ib.FMA().pack(system, 0.0).u.fma.TimeDivisor == 100.0            # 1000 / CyclesPerSecond (Transforms.cs:40)
    n = ib.Noise(Interval=500.0, Seed=4)
    u0 = (n.CurrentU, n.NextU)
    a = n.pack(system, 0.25).u.noise
    assert a.FrequencyLerp == pytest.approx(0.5) and tuple(a.RandomnessTexel) == pytest.approx((1 / 807, 1 / 653))
    assert a.RandomnessOffset[0] == pytest.approx(u0[0] * 253, rel=1e-6) and a.NextRandomnessOffset[0] == pytest.approx(u0[1] * 253, rel=1e-6)
    b = n.pack(system, 0.55).u.noise                                         # past the interval: UVs cycle
    assert b.RandomnessOffset[0] == pytest.approx(u0[1] * 253, rel=1e-6) and b.FrequencyLerp == pytest.approx(0.1, abs=1e-6)
    g = ib.Gravity(Attractors=[ib.Attractor(Position=(1, 2, 3), Radius=4, Strength=5, Type=ib.AttractorType.Exponential)])
    gp = g.pack(system, 0.0).u.gravity
    assert gp.AttractorCount == 1 and gp.AttractorRadiusesAndStrengths[0].tuple() == (4, 5, 2, 0) and tuple(gp.CategoryFilter) == (0, 0)
    assert not ib.Gravity().IsValid and system.plan_ops(0.0) == []
    area = ib.TransformArea(Type=ib.AreaType.Box, Falloff=0.2)
    assert ib.FMA(Area=area).pack(system, 0.0).u.fma.area.AreaFalloff == 1.0   # Math.Max(1, falloff) (ParticleTransform.cs:305)


def test_system_uniforms():    # Uniforms.cs:208-235, ParticleSystem.cs:547-575
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=64))
    cfg = ib.ParticleSystemConfiguration(Friction=0.2, MaximumVelocity=99, LifeDecayPerSecond=1.5, RotationFromLife=90.0, OpacityFromLife=4.0,
                                         Collision=ib.ParticleCollision(EscapeVelocity=7, BounceVelocityMultiplier=0.5, Distance=1.5, LifePenalty=0.1))
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    u = system.system_uniforms(1 / 60.0)
    assert u.GlobalSettings.tuple() == pytest.approx((1000 / 60.0, 0.2, 99, 1.5)) and u.CollisionSettings.tuple() == pytest.approx((7, 0.5, 1.5, 0.1))
    assert u.TexelAndSize.x == 1 / 64 and u.has_collision_field == 0 and u.write_render_outputs == 1
    assert u.ColorFromLife.RangeAndCount.tuple() == (0, 0.25, 2, 0) and u.ColorFromLife.A.tuple() == (1, 1, 1, 0)   # OpacityFromLife ramp
    assert u.RotationFromLifeAndIndex[0] == pytest.approx(math.pi / 2)
    df = ib.DistanceField(None, 128, 128, 64.0, 6)
    cfg.Collision.DistanceField = df
    with pytest.raises(ib.IlluminantError):          # "If a distance field is active, you must set DistanceFieldMaximumZ"
        system.system_uniforms(1 / 60.0)
    cfg.Collision.DistanceFieldMaximumZ = 64.0
    u = system.system_uniforms(1 / 60.0)
    assert u.has_collision_field == 1 and u.CollisionField.Packed1.tuple() == (0, 0, 0, 0)     # reference quirk: never set
    cfg.Collision.FullFieldAddressing = True
    assert system.system_uniforms(1 / 60.0).CollisionField.Packed1.y > 0


def test_sharding_partitions():
    for height in (2160, 1080, 7, 1):
        for world in (1, 2, 3, 4, 8):
            bands = [sharding.row_band(r, world, height) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == height
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(b - a <= sharding.band_height(height, world) for a, b in bands)
    for chunks in (32, 5, 1):
        for world in (1, 2, 4, 8):
            rs = [sharding.chunk_range(r, world, chunks) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == chunks and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_randomness_texture_is_seeded():
    a = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(RandomSeed=5)).RandomnessTexture
    b = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(RandomSeed=5)).RandomnessTexture
    assert a.shape == (653, 807, 4) and a.dtype == np.float32 and np.array_equal(a, b) and 0 <= a.min() and a.max() < 1


def test_rebalance_rows_equalises_measured_cost():
    from illuminant_b200 import sharding
    H, n = 2160, 8
    rng = np.random.RandomState(3)
    density = 1.0 + 2.0 * np.exp(-((np.arange(H) - 700) / 200.0) ** 2)      # a cluster of lights around row 700
    bounds = [min(k * sharding.band_height(H, n), H) for k in range(n)] + [H]
    for _ in range(4):
        times = [float(density[bounds[k]:bounds[k + 1]].sum()) for k in range(n)]
        bounds = sharding.rebalance_rows(bounds, times, H)
        assert bounds[0] == 0 and bounds[-1] == H and all(b % 16 == 0 for b in bounds[1:-1])
        assert all(bounds[k] <= bounds[k + 1] for k in range(n))
    times = np.array([density[bounds[k]:bounds[k + 1]].sum() for k in range(n)])
    assert times.max() / times.mean() < 1.10            # equal-height bands of this profile: 1.5 (16-row quantisation limits the fit)
    assert sharding.rebalance_rows([0, 1080, 2160], [0.0, 0.0], 2160) == [0, 1080, 2160]


def test_particle_light_source_uniforms_and_vertices():    # LightingRenderer.cs:769-789, ParticleLight.fx:16-82
    t = ib.SphereLightSource(Radius=3, RampLength=20, RampMode=ib.LightSourceRampMode.Exponential, Color=(0.5, 1.0, 2.0, 0.5),
                             SpecularColor=(0.1, 0.2, 0.3), SpecularPower=5, AmbientOcclusionRadius=7, AmbientOcclusionOpacity=1.5, FalloffYFactor=2)
    pls = ib.ParticleLightSource(Template=t)
    props, more, color, spec = pls.uniforms(True)
    assert props.tuple() == (3, 20, 1, 1) and pls.uniforms(False)[0].w == 0
    assert more.tuple() == (7, -99999, 2, 1.0)                      # AO opacity saturated
    assert color.tuple() == (0.5, 1.0, 2.0, 0.5) and spec.tuple() == pytest.approx((0.1, 0.2, 0.3, 5))
    t.AmbientOcclusionOpacity = 0.0005
    assert pls.uniforms(True)[1].x == 0                             # AO radius dropped when its opacity is <= 0.001
    P = np.array([[1, 2, 3, 1.0], [4, 5, 6, 0.0], [7, 8, 9, 2.0], [1, 1, 1, 1.0]], np.float32)
    A = np.array([[0.2, 0.4, 0.1, 0.5], [1, 1, 1, 1], [0.3, 0.3, 0.3, 0.0], [0.5, 0.25, 0.125, 1.0]], np.float32)
    vs = pls.light_vertices(P, A, True)
    assert len(vs) == 2                                             # particle 1 is dead, particle 2 has alpha 0
    assert vs[0].LightPosition1.tuple() == (1, 2, 3, 0) and vs[0].EvenMoreLightProperties.x == -1
    assert vs[0].Color1.tuple() == pytest.approx((0.4 * 0.5, 0.8 * 1.0, 0.2 * 2.0, 0.5 * 0.5))   # unpremultiplied, times LightColor
    assert vs[1].Color1.tuple() == pytest.approx((0.25, 0.25, 0.25, 0.5))
    # a particle light source never enters the host batches
    s = scenes.lighting_scene(0, 64, 64, 1)
    s.environment.Lights.append(pls)
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    _, nb, _, nv = r.build_batches()
    assert (nb, nv) == (1, 1)


def test_life_ramp_settings():                             # MaybeSetLifeRampParameters ParticleSystem.cs:911-941
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration()
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    assert system.system_uniforms(1 / 60).LifeRampSettings.tuple() == (0, 0, 1, 1)
    cfg.LifeRamp = ib.ParticleColorLifeRamp(Minimum=1.0, Maximum=1.0002, Strength=0.75, Invert=True, Texture=np.zeros((5, 9, 4), np.float32))
    u = system.system_uniforms(1 / 60)
    assert u.LifeRampSettings.x == -0.75 and u.LifeRampSettings.y == 1.0 and u.LifeRampSettings.w == 5
    assert u.LifeRampSettings.z == pytest.approx(0.001)             # max(range, 0.001)


def test_dynamic_distance_field_is_a_distance_field():
    df = ib.DynamicDistanceField(None, 320, 200, 128.0, 8)
    assert (df.SliceCount, df.PhysicalSliceCount, df.TextureWidth, df.TextureHeight) == (9, 3, 640, 400)
    assert df.static_handle is None


def test_chunk_reaping_bookkeeping_without_a_device():
    """ParticleLiveness.cs:46-129 on the host mirror: a chunk whose count stays 0 for DeadFrameThreshold liveness results is
    reaped at the top of the next update; the chunk lists and the spawn-target indices follow the shift."""
    import illuminant_b200 as ib
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    system = ib.ParticleSystem(engine, ib.ParticleSystemConfiguration(), maxChunks=4)
    for _ in range(3):
        system._create_chunk()
    system._chunk_next_offset[:] = [256, 256, 40]
    system._chunk_total_spawned[:] = [256, 256, 40]
    system._spawn_target = 2
    system._feedback_source = 1
    for k in range(system.DeadFrameThreshold - 1):
        system._process_liveness([5, 0, 7])
    assert system._chunk_reap == [False, False, False] and system._chunk_dead_frames == [0, system.DeadFrameThreshold - 1, 0]
    system._process_liveness([5, 3, 7])                      # one live result resets the count
    assert system._chunk_dead_frames[1] == 0
    for k in range(system.DeadFrameThreshold):
        system._process_liveness([5, 0, 7])
    assert system._chunk_reap == [False, True, False]
    system._update_live_count_and_reap()
    assert system.LiveChunkCount == 2 and system._chunk_next_offset == [256, 40] and system._chunk_live_count == [5, 7]
    assert system._spawn_target == 1 and system._feedback_source == -1 and system.ReapedChunkCount == 1
    assert system._create_chunk() == 2 and system._create_chunk() == 3 and system._create_chunk() == -1   # a slot was freed
    system.TotalSpawnCount = 99
    system.Clear()
    system._update_live_count_and_reap()
    assert system.LiveChunkCount == 0 and system.TotalSpawnCount == 0 and not system.IsClearPending and system._spawn_target == -1

"""Known-answer tests that pin the CPU oracle.  The reference ships no tests, golden images or fixtures for these paths
 and cannot run here, so the oracle is pinned by (1) an independent transcription of the
reference's own C# mirror of Bezier.fxh (Bezier.cs:461-492, :759-832), (2) closed forms listed in SURVEY.md section 8c and
(3) hand-computed vectors.  ("Parity unpinned" by reference outputs -- see oracle/README.md.)"""
import math

import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi, scenes
from illuminant_b200._abi import Float4
from illuminant_b200.particles import clamped_bezier1, clamped_bezier4

F = np.float32


# ---- (1) Bezier: transcription of ClampedBezier4.tForScaledBezier / Evaluate from Bezier.cs, in float32 ----------
def cs_wrap_exclusive(v, lo, hi):   # Squared.Util Arithmetic.WrapExclusive for positive ranges
    d = hi - lo
    return F(v - d * math.floor((v - lo) / d))


def cs_t(range_and_count, value):
    minValue, invDivisor, count, w = (F(x) for x in range_and_count)
    mode = int(w)
    repeating, bouncing = mode > 255, mode > 511
    t = F(F(F(value) - minValue) * F(abs(invDivisor)))
    if bouncing:
        t = F(t * F(2))
        t = F(2 - cs_wrap_exclusive(t, 0, 2)) if invDivisor < 0 else cs_wrap_exclusive(t, 0, 2)
        if t > 1:
            t = F(1 - F(t - 1))
    elif repeating:
        t = F(1 - cs_wrap_exclusive(t, 0, 1)) if invDivisor < 0 else cs_wrap_exclusive(t, 0, 1)
    else:
        s = F(min(max(t, F(0)), F(1)))
        t = F(1 - s) if invDivisor < 0 else s
    m = mode % 256
    if m == 1:
        t = F(math.sin(float(t) * math.pi * 0.5))
    elif m == 2:
        t = F(t * t)
    return int(count), t


def cs_lerp(a, b, t):
    return F(a + F(F(b - a) * t))    # Arithmetic.Lerp(a, b, x) = a + (b - a) * x


def cs_evaluate(range_and_count, a, b, c, d, value):
    count, t = cs_t(range_and_count, value)
    if count <= 1.5:
        return a
    ab = cs_lerp(a, b, t)
    if count <= 2.5:
        return ab
    if count <= 3.5:
        return a if t <= 0 else (c if t >= 1 else b)
    bc, cd = cs_lerp(b, c, t), cs_lerp(c, d, t)
    return cs_lerp(cs_lerp(ab, bc, t), cs_lerp(bc, cd, t), t)


@pytest.mark.parametrize("count", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", [0, 1, 2, 256, 257, 512, 514])
@pytest.mark.parametrize("rng", [(0.0, 1.0), (2.0, 5.0), (4.0, 1.0)])
def test_bezier_matches_reference_cpu_mirror(oracle, count, mode, rng):
    src = ib.BezierF(Count=count, Mode=mode, MinValue=rng[0], MaxValue=rng[1], A=0.25, B=2.0, C=-1.5, D=0.75)
    b1 = clamped_bezier1(src)
    src4 = ib.Bezier4V(Count=count, Mode=mode, MinValue=rng[0], MaxValue=rng[1], A=(0.25, 1, 0, 0.5), B=(2, 0, 1, 1), C=(-1.5, 0.5, 0.25, 0), D=(0.75, 0.1, 0.9, 1))
    b4 = clamped_bezier4(src4)
    rc = b1.RangeAndCount.tuple()
    for value in np.linspace(-1.0, 7.0, 41, dtype=np.float32):
        # Below the range start the two reference implementations disagree in the repeating / bouncing modes: the
        # shader's `t % 1` keeps the sign of t (HLSL fmod) while Bezier.cs wraps into [0, 1).  The oracle follows
        # the shader (that is what runs on the GPU); compare with the C# mirror only where both are defined alike.
        if mode > 255 and value < min(rng):
            continue
        want = cs_evaluate(rc, F(0.25), F(2.0), F(-1.5), F(0.75), float(value))
        got = oracle.bezier1(b1, float(value))
        assert got == pytest.approx(float(want), abs=2e-6), (count, mode, rng, value)
        got4 = oracle.bezier4(b4, float(value))
        for k, (a, b, c, d) in enumerate(zip(src4.A, src4.B, src4.C, src4.D)):
            assert got4[k] == pytest.approx(float(cs_evaluate(rc, F(a), F(b), F(c), F(d), float(value))), abs=2e-6)


def test_clamped_bezier_constructor_matches_bezier_cs():
    b = clamped_bezier1(ib.BezierF(Count=2, MinValue=3.0, MaxValue=1.0, A=1, B=2))       # Bezier.cs:444-458
    assert b.RangeAndCount.tuple() == (1.0, -0.5, 2.0, 0.0)
    assert clamped_bezier1(None).RangeAndCount.tuple() == (0.0, 1.0, 1.0, 0.0)           # ClampedBezier1.One
    assert clamped_bezier1(ib.BezierF(Count=1, MinValue=0, MaxValue=9)).RangeAndCount.y == 1.0   # range forced to 1 when Count <= 1


# ---- (2) closed forms, lighting -----------------------------------------------------------------------------------
def _field(width=64, height=48, depth=128.0, slices=8):
    df = ib.DistanceField(None, width, height, depth, slices)
    df.ValidSliceCount = df.SliceCount
    return df


def test_distance_field_descriptor_arithmetic():
    c2 = ib.DistanceField(None, 1920, 1080, 128.0, 8)        # 8 requested -> 9 virtual / 3 physical, 2x2 atlas
    assert (c2.SliceCount, c2.PhysicalSliceCount, c2.ColumnCount, c2.RowCount) == (9, 3, 2, 2)
    assert (c2.TextureWidth, c2.TextureHeight) == (3840, 2160)
    c4 = ib.DistanceField(None, 3840, 2160, 128.0, 8)
    assert (c4.TextureWidth, c4.TextureHeight, c4.ColumnCount, c4.RowCount) == (7680, 4320, 2, 2)
    q = ib.DistanceField(None, 1920, 1080, 128.0, 9, 0.25)   # SimpleParticles.cs:216-219
    assert (q.SliceWidth, q.SliceHeight, q.Resolution) == (480, 270, 0.25)
    assert ib.DistanceField(None, 100, 100, 64.0, 1).SliceCount == 3
    many = ib.DistanceField(None, 4096, 4096, 256.0, 64)     # atlas capped at 8192^2: 2x2 cells x 3 slices
    assert many.SliceCount == 12
    c2.ValidSliceCount = 9
    u = c2.uniforms()
    assert u.Packed1.x == float(F(F(1) / F(2)) * (F(1) / F(3)))
    assert u.Packed1.y == float(F(1) / F(128) * F(9)) and u.Packed1.z == 128.0 and u.Packed1.w == 3.0
    assert u.TextureSliceAndTexelSize.tuple() == (0.5, 0.5, float(F(1) / F(3840)), float(F(1) / F(2160)))


def test_cleared_texel_decodes_to_192_over_255_of_max(oracle):
    df = _field()
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    d = oracle.sample_distance_field(tex, df.uniforms(), 10.0, 10.0, 5.0)
    assert d == pytest.approx(192.0 / 255.0 * 128.0, rel=1e-6)          # DistanceFieldCommon.fxh:8,268-270 -> 96.376
    # outside the volume the Euclidean distance to the box is added (:320-321, :352)
    assert oracle.sample_distance_field(tex, df.uniforms(), -3.0, 10.0, 5.0) == pytest.approx(96.37647 + 3.0, rel=1e-6)
    assert oracle.sample_distance_field(tex, df.uniforms(), 64.0 + 3.0, 48.0 + 4.0, 5.0) == pytest.approx(96.37647 + 5.0, rel=1e-6)


def test_sampler_interpolates_z_slices_and_bilinear(oracle):
    df = _field(8, 8, 90.0, 9)          # 9 slices over depth 90: one slice per 10 z units
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    u = df.uniforms()
    enc = lambda dist: int(round((192.0 / 255.0 - dist / 128.0) * 65535))
    # physical slice 0 holds z-slices 0..3 in r,g,b,a: distance = 10 * (slice index + 1) everywhere
    tex[:8, :8] = [enc(10), enc(20), enc(30), enc(40)]
    for z, want in [(0.0, 10.0), (5.0, 15.0), (10.0, 20.0), (17.5, 27.5), (25.0, 35.0)]:
        assert oracle.sample_distance_field(tex, u, 4.0, 4.0, z) == pytest.approx(want, abs=2e-3)
    # x gradient: texel i holds distance 10 + i at slice 0; texel centres sit at i + 0.5
    for i in range(8):
        tex[:8, i, 0] = enc(10 + i)
    assert oracle.sample_distance_field(tex, u, 3.5, 4.0, 0.0) == pytest.approx(13.0, abs=2e-3)
    assert oracle.sample_distance_field(tex, u, 4.0, 4.0, 0.0) == pytest.approx(13.5, abs=2e-3)


def test_cone_trace_closed_forms(oracle):
    df = _field()
    u = df.uniforms()
    # no field (Extent.x <= 0): coneTrace == 1 (ConeTrace.fxh:159,190)
    v, steps = oracle.cone_trace(None, ib.DistanceField.empty_uniforms(128.0), (30, 30, 20), 8.0, 100.0, (5, 5, 0))
    assert (v, steps) == (1.0, 0)
    # disabled trace returns 1
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    assert oracle.cone_trace(tex, u, (30, 30, 20), 8.0, 100.0, (5, 5, 0), enable=False)[0] == 1.0
    # empty (cleared) field: every sample is 96.376 -> visibility stays 1, steps = ceil((len - 0.5) / 96.376)
    v, steps = oracle.cone_trace(tex, u, (60, 40, 20), 8.0, 100.0, (5, 5, 0))
    length = math.dist((60, 40, 20), (5, 5, 0)) - 8.0
    assert v == pytest.approx(1.0, abs=1e-6) and steps == math.ceil((length - 0.5) / 96.37647)
    # fully blocked: a field whose distance is -20 everywhere drives visibility to 0 in one step
    blocked = np.full_like(tex, int(round((192.0 / 255.0 + 20.0 / 128.0) * 65535)))
    v, steps = oracle.cone_trace(blocked, u, (60, 40, 20), 8.0, 100.0, (5, 5, 0))
    assert v == 0.0 and steps == 1
    # step cap: MaxStepCount 2 with MinStepSize 3 on a field of distance 0 -> stepsRemaining hits 0 -> visibility 0
    zero = np.full_like(tex, int(round(192.0 / 255.0 * 65535)))
    q = ib.RendererQualitySettings(MaxStepCount=2)
    v, steps = oracle.cone_trace(zero, df.uniforms(q), (60, 40, 20), 8.0, 100.0, (5, 5, 0))
    assert steps == 2 and v == 0.0


def _cone_trace_constant_field_f64(d0, start, end, light_radius, ramp_length, q):
    """coneTrace (ConeTrace.fxh:13-71, 120-191) in float64 over a field whose distance is d0 everywhere, written from the shader
    text with its own constants; returns (result, steps taken)."""
    MIN_CONE_RADIUS, WINDOW, T0, DARK, LIT, HACK = 0.33, 2.0, 0.5, 0.075, 0.95, 1.5
    length = math.dist(start, end)
    t, limit, vis = T0, max(length - light_radius, 1.0), 1.0
    max_radius = min(max(light_radius, MIN_CONE_RADIUS), q.MaxConeRadius)
    growth = max_radius / max(ramp_length, 16.0) * 1.0
    min_step = max(1.0, q.MinStepSize)
    steps_remaining, live, steps = float(q.MaxStepCount), 1.0, 0
    sat = lambda x: min(max(x, 0.0), 1.0)     # noqa: E731
    while live > 0:
        steps_remaining -= 1
        steps += 1
        vis = min(vis, (d0 + HACK) / min(growth * t + MIN_CONE_RADIUS, max_radius))
        t += max(abs(d0) * q.LongStepFactor, min_step)
        live = steps_remaining * sat(vis - DARK) * sat(limit - t)
    v = min(vis, steps_remaining / WINDOW)
    return sat(sat(v - DARK) / (LIT - DARK)) ** q.OcclusionToOpacityPower, steps


def test_cone_trace_penumbra_over_a_constant_field(oracle):
    """Partial shadow: over a field that reads d0 everywhere the visibility is (d0 + 1.5) / cone radius at the last sample, the
    radius growing by maxRadius / rampLength per pixel from 0.33 up to the light's radius (capped at MaxConeRadius); the march
    advances by max(|d0| * LongStepFactor, MinStepSize).  The float64 restatement above must give the oracle's value and step
    count for light radii below / above MaxConeRadius, a non-default quality, a negative distance and a small step budget."""
    df = _field(128, 96, 128.0, 9)
    tex0 = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    start, end = (10.0, 12.0, 2.0), (100.0, 70.0, 30.0)
    cases = [(3.0, 8.0, 100.0, ib.RendererQualitySettings()), (6.0, 40.0, 60.0, ib.RendererQualitySettings()),
             (1.0, 12.0, 10.0, ib.RendererQualitySettings(MinStepSize=5.0, LongStepFactor=0.5, MaxStepCount=40, MaxConeRadius=16, OcclusionToOpacityPower=2.0)),
             (-0.5, 8.0, 100.0, ib.RendererQualitySettings()), (2.0, 4.0, 200.0, ib.RendererQualitySettings(MinStepSize=1.0, MaxStepCount=12))]
    seen = []
    for d0, radius, ramp, q in cases:
        code = int(round((192.0 / 255.0 - d0 / 128.0) * 65535))
        tex = np.full_like(tex0, code)
        d_stored = (192.0 / 255.0 - code / 65535.0) * 128.0          # what the texel decodes to (quantised to 1/512 px)
        want, want_steps = _cone_trace_constant_field_f64(d_stored, start, end, radius, ramp, q)
        got, steps = oracle.cone_trace(tex, df.uniforms(q), end, radius, ramp, start)
        assert steps == want_steps, (d0, radius, steps, want_steps)
        assert got == pytest.approx(want, abs=2e-5), (d0, radius, got, want)
        seen.append(want)
    assert 0.05 < seen[0] < 0.95 and 0.05 < seen[1] < 0.95 and 0.0 < seen[3] < 0.1    # real penumbrae; just inside a surface is nearly dark


def test_sphere_light_opacity_closed_forms(oracle):
    s = scenes.lighting_scene(0, 32, 32, 0)
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    f = r.build_frame()
    up = (0, 0, 1)
    assert oracle.sphere_light_opacity(f, (10, 10, 0), up, (10, 12, 3), (16, 100, 0, 1)) == 1.0     # inside Radius -> 1 (LightCommon.fxh:209)
    # linear ramp, light straight above the point: normal factor 1, falloff 1 - (d - R) / ramp
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 0, 60), (10, 100, 0, 1)) == pytest.approx(0.5, abs=1e-6)
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 0, 60), (10, 100, 1, 1)) == pytest.approx(0.25, abs=1e-6)   # exponential
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 0, 200), (10, 100, 0, 1)) == 0.0       # beyond R + ramp
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 0, -60), (10, 100, 0, 1)) == 0.0       # light behind the surface
    assert oracle.sphere_light_opacity(f, (0, 0, 0), (0, 0, 0), (0, 0, -60), (10, 100, 0, 1)) == pytest.approx(0.5, abs=1e-6)  # no normal
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 30, 0.001), (10, 100, 2, 1)) == 0.0    # RampMode None: 1 - sat(d - R)
    assert oracle.sphere_light_opacity(f, (0, 0, 0), up, (0, 30, 40), (10, 100, 0, 1), 2.0) == pytest.approx(
        (1 - (math.hypot(60, 40) - 10) / 100) * min(max((40 / math.hypot(60, 40) + 0.15) / 0.15, 0), 1) ** 0.85, abs=1e-6)   # FalloffYFactor


def test_gbuffer_encoding_closed_forms(oracle):
    g = oracle.encode_gbuffer_sample((0, 0, 1), 0.0, 0.0)
    assert tuple(g) == (0.5, 1.0, 0.0, 1.0)                  # ground plane z = 0 (GBufferShaderCommon.fxh:22-33)
    assert tuple(oracle.encode_gbuffer_sample((0, 0, 1), 0.0, 0.0, dead=True)) == (0.0, 0.0, -99999.0, -99999.0)
    assert oracle.encode_gbuffer_sample((0, 0, 1), 0.0, 32.0, enable_shadows=False)[3] == pytest.approx(-(32 + 1024) / 1024 - 1)
    assert oracle.encode_gbuffer_sample((0, 0, 1), 0.0, 0.0, fullbright=True)[3] == 99999.0
    # the vectorised product-side encoder agrees with the oracle's scalar one
    n = np.array([[0, 0, 1], [0.6, 0, 0.8], [0, 1, 0], [0, 0, 0]], np.float32)
    z = np.array([0, 12.5, 64, 3], np.float32)
    enc = ib.encode_gbuffer(n, np.zeros(4, np.float32), z, np.array([True, True, False, True]), np.array([False, False, False, True]))
    for i in range(4):
        want = oracle.encode_gbuffer_sample(n[i], 0.0, float(z[i]), enable_shadows=bool([True, True, False, True][i]), fullbright=(i == 3))
        assert np.allclose(enc[i], want, atol=1e-6)
    # decode(encode(x)) round trip through sampleGBuffer
    s = scenes.lighting_scene(0, 4, 1, 0)
    s.gbuffer = enc.reshape(1, 4, 4)
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r._gbuffer_shape = (1, 4)
    f = r.build_frame()
    wp, nn, es, fb = oracle.decode_gbuffer(f, s.gbuffer, 1, 0)
    assert np.allclose(wp, [1.5, 0.5, 12.5], atol=1e-4) and np.allclose(nn, [0.6, 0, 0.8], atol=1e-6) and es and not fb
    wp, nn, es, fb = oracle.decode_gbuffer(f, s.gbuffer, 2, 0)
    assert wp[2] == pytest.approx(64.0, abs=1e-3) and not es and not fb
    assert oracle.decode_gbuffer(f, s.gbuffer, 3, 0)[3]      # fullbright


def test_unobstructed_light_is_ambient_plus_falloff(oracle):
    """No obstructions, no G-buffer: lightmap = ambient + color.rgb * color.a * computeSphereLightOpacity."""
    s = scenes.lighting_scene(0, 48, 32, 0, float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    light = ib.SphereLightSource(Position=(20.0, 12.0, 30.0), Radius=6.0, RampLength=40.0, Color=(1.0, 0.5, 0.25, 0.8), CastsShadows=True)
    s.environment.Lights = [light]
    df = scenes.make_distance_field(None, s)
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)    # cleared field: nothing occludes
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    lm = oracle.render_lighting(tex, None, frame, batches, nb, verts, nv)
    for (x, y) in [(20, 12), (30, 20), (5, 5), (47, 31), (0, 0)]:
        pos = (x + 0.5, y + 0.5, 0.0)         # ScaleCompensation: pixel centres
        op = oracle.sphere_light_opacity(frame, pos, (0, 0, 1), light.Position, (6.0, 40.0, 0, 1))
        want = np.array(s.environment.Ambient[:3]) + np.array([1.0, 0.5, 0.25]) * 0.8 * op
        assert np.allclose(lm[y, x, :3], want, atol=2e-6), (x, y)
        assert lm[y, x, 3] == s.environment.Ambient[3] + (1.0 if op > 0 else 0.0)      # additive blend counts touching lights


def _triangle_solid_angle(a, b, c):
    """Van Oosterom & Strackee (1983): tan(omega / 2) = a . (b x c) / (|a||b||c| + (a.b)|c| + (a.c)|b| + (b.c)|a|) -- not the
    dihedral-angle sum the shader uses, so agreement pins rectangleSolidAngle (FBPBR.fxh:33-51) from outside."""
    la, lb, lc = np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(c)
    num = float(np.dot(a, np.cross(b, c)))
    den = la * lb * lc + np.dot(a, b) * lc + np.dot(a, c) * lb + np.dot(b, c) * la
    return abs(2.0 * np.arctan2(num, den))


def _line_light_opacity_f64(wp, n, P0, P1, radius):
    """computeLineLightOpacity (FBPBR.fxh:53-101) in float64, the solid angle from two triangles."""
    wp, n, P0, P1 = (np.asarray(v, np.float64) for v in (wp, n, P0, P1))
    ab = P1 - P0
    u = min(max(np.dot(wp - P0, ab) / np.dot(ab, ab), 0.0), 1.0)      # closestPointOnLineSegment3, DistanceFieldCommon.fxh:151-155
    sphere = P0 + u * ab
    left = ab / np.linalg.norm(ab)
    forward = (sphere - wp) / np.linalg.norm(sphere - wp)
    up = np.cross(left, forward)
    p = [P0 + radius * up, P0 - radius * up, P1 - radius * up, P1 + radius * up]
    v = [q - wp for q in p]
    omega = _triangle_solid_angle(v[0], v[1], v[2]) + _triangle_solid_angle(v[0], v[2], v[3])
    sat = lambda x: min(max(x, 0.0), 1.0)     # noqa: E731
    total = sum(sat(np.dot(q / np.linalg.norm(q), n)) for q in v) + sat(np.dot((0.5 * (P0 + P1) - wp) / np.linalg.norm(0.5 * (P0 + P1) - wp), n))
    illum = omega * 0.2 * total
    d = sphere - wp
    illum += np.pi * sat(np.dot(d / np.linalg.norm(d), n)) * (radius * radius) / np.dot(d, d)
    return sat(illum), u


def test_unobstructed_line_light_matches_an_independent_solid_angle(oracle):
    """No obstructions, no G-buffer (flat ground, normal +z): lightmap = ambient + lerp(StartColor, EndColor, u).rgb * a *
    computeLineLightOpacity, the latter evaluated in float64 with the rectangle's solid angle taken from the Van Oosterom-Strackee
    triangle formula instead of the shader's sum of four dihedral angles.  The four arc-cosines nearly cancel in fp32, hence
    the tolerance (3e-4 of the opacity, far below a wrong formula, which is off by tens of per cent)."""
    s = scenes.lighting_scene(0, 64, 40, 0, float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    P0, P1, radius = (14.0, 12.0, 22.0), (49.0, 27.0, 31.0), 5.0
    light = ib.LineLightSource(StartPosition=P0, EndPosition=P1, Radius=radius, StartColor=(1.0, 0.5, 0.25, 0.8), EndColor=(0.2, 0.9, 0.6, 0.5),
                               CastsShadows=True)
    s.environment.Lights = [light]
    df = scenes.make_distance_field(None, s)
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)    # cleared field: nothing occludes
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    lm = oracle.render_lighting(tex, None, frame, batches, nb, verts, nv)
    c0, c1 = np.array(light.StartColor, np.float64), np.array(light.EndColor, np.float64)
    seen = []
    for (x, y) in [(14, 12), (31, 19), (49, 27), (5, 35), (60, 3), (0, 0), (63, 39), (30, 5)]:
        op, u = _line_light_opacity_f64((x + 0.5, y + 0.5, 0.0), (0.0, 0.0, 1.0), P0, P1, radius)
        col = c0 + (c1 - c0) * u
        want = np.array(s.environment.Ambient[:3], np.float64) + col[:3] * col[3] * op
        got = lm[y, x, :3].astype(np.float64)
        assert np.allclose(got, want, rtol=0, atol=3e-4 * max(op, 1e-3) + 2e-6), (x, y, got, want, op)
        assert lm[y, x, 3] == s.environment.Ambient[3] + (1.0 if op > 0 else 0.0)
        seen.append(op)
    assert max(seen) > 0.4 and 0.0 < min(seen) < 0.15          # under the light and far from it


def test_ambient_occlusion_closed_form(oracle):
    """computeAO (AOCommon.fxh:1-20) by hand: on flat ground next to a tall box the sample at p + (0, 0, n.z * radius) is D px from
    the box's face (as the field stores it), so the light's opacity is multiplied by (1 - o) + o * (1 - (1 - D / radius)^2); beyond
    the radius by 1.
    The light casts no shadows, so nothing else of the field enters."""
    s = scenes.lighting_scene(0, 96, 64, 0, float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    box = ib.LightObstruction(ib.LightObstructionType.Box, (70.0, 32.0, 0.0), (10.0, 60.0, 400.0))     # face at x = 60, spans every slice
    light = ib.SphereLightSource(Position=(30.0, 32.0, 40.0), Radius=10.0, RampLength=120.0, Color=(1.0, 0.8, 0.6, 1.0), CastsShadows=False,
                                 AmbientOcclusionRadius=20.0, AmbientOcclusionOpacity=0.6)
    s.environment.Lights = [light]
    s.obstructions = [box]
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, [box])
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    lm = oracle.render_lighting(tex, None, frame, batches, nb, verts, nv)
    factors = []
    for x in (20, 39, 45, 52, 57):
        pos = (x + 0.5, 32.5, 0.0)
        D = 60.0 - x      # the sample at the pixel centre reads field texel x, which holds the distance at x (test_distance_field_generation_box_slice)
        ao = 1.0 if D >= 20.0 else (1.0 - 0.6) + 0.6 * (1.0 - (1.0 - D / 20.0) ** 2)
        op = oracle.sphere_light_opacity(frame, pos, (0, 0, 1), light.Position, (10.0, 120.0, 0, 0))
        want = np.array(s.environment.Ambient[:3], np.float64) + np.array([1.0, 0.8, 0.6]) * op * ao
        assert np.allclose(lm[32, x, :3], want, rtol=0, atol=2e-3 * op), (x, lm[32, x], want, ao)
        factors.append(ao)
    assert factors[0] == 1.0 and factors[-1] < 0.6          # out of reach, and deep in the corner


def test_sphere_light_specular_closed_form(oracle):
    """SphereLightPixelShader (SphereLight.fx:7-46) + CalcSphereLightSpecularity (LightCommon.fxh:212-222) by hand, without a
    G-buffer: the camera sits straight above the pixel at MaximumZ + 0.01, h = normalize(normalize(camera - p) - (p - light)) --
    the light direction is NOT normalised in the reference -- and the light adds specular.rgb * pow(saturate(dot(h, n)), power)
    * opacity on top of color.rgb * color.a * opacity."""
    s = scenes.lighting_scene(0, 48, 32, 0, float4_lightmap=True)
    s.configuration.EnableGBuffer = False
    light = ib.SphereLightSource(Position=(20.0, 12.0, 30.0), Radius=6.0, RampLength=60.0, Color=(1.0, 0.5, 0.25, 0.8), CastsShadows=False,
                                 SpecularColor=(0.3, 0.6, 0.9), SpecularPower=6.0)
    s.environment.Lights = [light]
    df = scenes.make_distance_field(None, s)
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    lm = oracle.render_lighting(tex, None, frame, batches, nb, verts, nv)
    seen = []
    for (x, y) in [(20, 12), (28, 18), (8, 5), (40, 28)]:
        p = np.array([x + 0.5, y + 0.5, 0.0])
        cam = np.array([x + 0.5, y + 0.5, s.environment.MaximumZ + 0.01])
        to_cam = (cam - p) / np.linalg.norm(cam - p)
        h = to_cam - (p - np.array(light.Position, np.float64))
        h /= np.linalg.norm(h)
        spec = min(max(h[2], 0.0), 1.0) ** 6.0
        op = oracle.sphere_light_opacity(frame, tuple(p), (0, 0, 1), light.Position, (6.0, 60.0, 0, 0))
        want = np.array(s.environment.Ambient[:3], np.float64) + (np.array([1.0, 0.5, 0.25]) * 0.8 + np.array([0.3, 0.6, 0.9]) * spec) * op
        assert np.allclose(lm[y, x, :3], want, rtol=0, atol=2e-4), (x, y, lm[y, x], want, spec)
        seen.append(spec)
    assert seen[0] > 0.99 and 0.0 < min(seen) < 0.9        # under the light the half vector is the normal; it falls off sideways


def test_light_probes_closed_form(oracle):
    """SphereLightProbePixelShader (SphereLightProbe.fx:19-44) by hand in a cleared field: a probe is cleared to 0 (no ambient),
    then every light adds color.rgb * color.a * probe opacity * computeSphereLightOpacity at the probe's position and normal
    (ambient occlusion switched off for probes); two lights add up, alpha counts them."""
    s = scenes.lighting_scene(0, 64, 48, 0, float4_lightmap=True)
    lights = [ib.SphereLightSource(Position=(20.0, 12.0, 30.0), Radius=6.0, RampLength=80.0, Color=(1.0, 0.5, 0.25, 0.8), CastsShadows=True,
                                   AmbientOcclusionRadius=12.0),
              ib.SphereLightSource(Position=(50.0, 40.0, 10.0), Radius=4.0, RampLength=70.0, Color=(0.2, 0.9, 0.4, 0.5), CastsShadows=False)]
    s.environment.Lights = lights
    df = scenes.make_distance_field(None, s)
    tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField = df
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    positions = np.array([[22.0, 14.0, 5.0, 1.0], [40.0, 30.0, 0.0, 1.0], [60.0, 5.0, 20.0, 1.0]], np.float32)     # xyz, opacity
    normals = np.array([[0.0, 0.0, 1.0, 1.0], [0.6, 0.0, 0.8, 1.0], [0.0, -1.0, 0.0, 0.0]], np.float32)             # xyz, enable shadows
    out = oracle.update_light_probes(tex, frame, batches, nb, verts, nv, positions, normals)
    for k in range(3):
        want, touched = np.zeros(3), 0
        for L in lights:
            op = oracle.sphere_light_opacity(frame, tuple(positions[k, :3]), tuple(normals[k, :3]), L.Position, (L.Radius, L.RampLength, 0, 0))
            want += np.array(L.Color[:3], np.float64) * L.Color[3] * op
            touched += 1 if op > 0 else 0
        assert np.allclose(out[k, :3], want, rtol=0, atol=3e-6), (k, out[k], want)
        assert out[k, 3] == touched
    assert out[0, :3].max() > 0.3


def test_unobstructed_directional_light_is_ambient_plus_normal_factor(oracle):
    """No obstructions, no G-buffer (normal +z): every pixel = ambient + color.rgb * color.a * pow(saturate((dot(-dir, n) + 0.35) /
    0.35), 0.85) (computeDirectionalLightOpacity / computeNormalFactorEx, LightCommon.fxh:154-165, :224-231), evaluated in
    float64 -- light from above (factor 1), grazing from below the horizon (inside the ramp) and from straight below (0)."""
    for direction, inside in (((0.3, 0.2, -0.9327379), False), ((0.6, 0.742294, 0.2987), True), ((0.8, 0.5244044, 0.2915), True), ((0.0, 0.0, 1.0), False)):
        s = scenes.lighting_scene(0, 32, 24, 0, float4_lightmap=True)
        s.configuration.EnableGBuffer = False
        light = ib.DirectionalLightSource(Color=(0.9, 0.6, 0.3, 0.7), CastsShadows=True)
        light.Direction = direction
        s.environment.Lights = [light]
        df = scenes.make_distance_field(None, s)
        tex = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
        df.ValidSliceCount, df.handle = df.SliceCount, 1
        r = ib.LightingRenderer(None, s.environment, s.configuration)
        r.DistanceField = df
        batches, nb, verts, nv = r.build_batches()
        lm = oracle.render_lighting(tex, None, r.build_frame(), batches, nb, verts, nv)
        d = np.array(light.Direction, np.float64)
        d = d / np.linalg.norm(d)
        factor = min(max((-d[2] + 0.35) / 0.35, 0.0), 1.0) ** 0.85
        assert (0.0 < factor < 1.0) == inside
        want = np.array(s.environment.Ambient[:3], np.float64) + np.array([0.9, 0.6, 0.3]) * 0.7 * factor
        assert np.allclose(lm[..., :3].reshape(-1, 3), want, rtol=0, atol=2e-5), (direction, lm[0, 0], want)
        assert np.all(lm[..., 3] == s.environment.Ambient[3] + 1.0)      # a directional light never discards a visible pixel


# ---- closed forms, particles --------------------------------------------------------------------------------------
def test_ballistic_particles_closed_form(oracle):
    """friction 0, no transforms, no field: p(t) = p0 + v t, life linear, dead particles become zeros."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=2.0, MaximumVelocity=1000.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[:4] = [[1, 2, 3, 1.0], [0, 0, 0, 0.05], [5, 5, 5, 0.0], [9, 9, 9, 10.0]]
    V[:4] = [[10, -20, 5, 0], [1, 1, 1, 0], [1, 1, 1, 0], [0, 0, 0, 2]]
    u = system.system_uniforms(1 / 50.0)
    P2, V2, A2, RC, RD = oracle.particles_step(P, V, A, 16, u, [], [], engine.RandomnessTexture, None, 5)
    t = 5 / 50.0
    assert np.allclose(P2[0], [1 + 10 * t, 2 - 20 * t, 3 + 5 * t, 1.0 - 2.0 * t], atol=1e-5)
    assert (P2[1] == 0).all() and (V2[1] == 0).all() and (RC[1] == 0).all()          # died: life 0.05 < 2 * 0.1
    assert (P2[2] == 0).all() and (V2[2] == 0).all()                                 # dead on entry: discarded -> cleared zeros
    assert np.allclose(P2[3], [9, 9, 9, 10 - 2 * t], atol=1e-5) and V2[3, 3] == 2.0  # |v| <= 0.001 -> velocity 0, category kept
    assert np.allclose(RD[0], [1.0, 0.0, math.sqrt(100 + 400 + 25), 0.0], atol=1e-4)  # size 1, rotation 0, |v|, category
    assert np.allclose(RC[0], [1, 1, 1, 1])


def test_friction_and_maximum_velocity(oracle):
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.5, LifeDecayPerSecond=0.0, MaximumVelocity=50.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[0, 3] = 1.0
    V[0] = [300, 400, 0, 0]        # |v| = 500 -> clamped to 50, then l -= l * friction * dt
    u = system.system_uniforms(0.1)
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, u, [], [], engine.RandomnessTexture, None, 1)
    speed = 50 - 50 * 0.5 * 0.1
    assert np.allclose(V2[0, :3], [0.6 * speed, 0.8 * speed, 0], atol=1e-4)
    assert np.allclose(P2[0, :3], [0.6 * speed * 0.1, 0.8 * speed * 0.1, 0], atol=1e-5)


def test_gravity_single_linear_attractor(oracle):
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=1000.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    system.Transforms = [ib.Gravity(MaximumAcceleration=1000.0, Attractors=[ib.Attractor(Position=(10, 0, 0), Radius=40.0, Strength=6.0, Type=ib.AttractorType.Linear)])]
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[0] = [0, 0, 0, 1]; P[1] = [0, 0, 0, 1]; V[1, 3] = 3.0     # particle 1 is in bounce delay: Gravity's (0,0) category filter skips it
    dt = 0.02
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), engine.RandomnessTexture, None, 1)
    assert V2[0, 0] == pytest.approx((1 - 10 / 40.0) * dt * 6.0, rel=1e-5) and V2[0, 1] == 0        # Gravity.fx:36-48
    assert V2[1, 0] == 0.0 and V2[1, 3] == 3.0       # untouched (only UpdateWithDistanceField counts the delay down)


def test_fma_transform_closed_form(oracle):
    """FMA.fx:15-51 + PS_Update (UpdateParticleSystem.fx:9-38) in float64: with no area the weight is Strength, the lerp
    parameter weight * getDeltaTime() / TimeDivisor = Strength * dt * CyclesPerSecond (both carry VelocityConstantScale), the
    position then advances by the NEW velocity times dt; a dead particle and one outside the category filter pass through."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=1000.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    fma = ib.FMA(CyclesPerSecond=10, PositionAdd=(1.0, 2.0, 3.0), PositionMultiply=(2.0, 1.0, 0.5), VelocityAdd=(0.0, 5.0, 0.0),
                 VelocityMultiply=(0.5, 0.5, 0.5), Strength=0.8, CategoryFilter=(0.0, 1.0))
    system.Transforms = [fma]
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[0] = [3, 4, 5, 1]; V[0] = [10, -20, 5, 0]
    P[1] = [3, 4, 5, 1]; V[1] = [10, -20, 5, 2]      # category 2: outside the filter, only the update tail moves it
    dt = 0.02
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), engine.RandomnessTexture, None, 1)
    t = 0.8 * dt * 10.0
    p0, v0 = np.array([3.0, 4.0, 5.0]), np.array([10.0, -20.0, 5.0])
    p1 = p0 + t * (p0 * np.array([2.0, 1.0, 0.5]) + np.array([1.0, 2.0, 3.0]) - p0)
    v1 = v0 + t * (v0 * 0.5 + np.array([0.0, 5.0, 0.0]) - v0)
    assert np.allclose(V2[0, :3], v1, rtol=2e-6) and V2[0, 3] == 0.0
    assert np.allclose(P2[0, :3], p1 + v1 * dt, rtol=2e-6) and P2[0, 3] == 1.0
    assert np.allclose(V2[1, :3], v0, rtol=2e-6) and V2[1, 3] == 2.0
    assert np.allclose(P2[1, :3], p0 + v0 * dt, rtol=2e-6)
    assert not P2[2:].any() and not V2[2:].any()       # dead particles stay cleared


def test_collision_tail_closed_forms(oracle):
    """PS_Update of UpdateParticleSystemWithDistanceField.fx:29-147 by hand.  In a cleared field (every sample decodes to
    96.4 px: nothing within reach) a particle flies p + v * dt and its bounce delay counts down by one; a particle that flies
    straight at the face of a box stops where the sampled distance first drops below CollisionDistance -- never inside the
    box -- and leaves with a reflected velocity of |v| * BounceVelocityMultiplier, velocity.w = 3 (the bounce delay) and its
    life reduced by LifePenalty."""
    df = _field(128, 128, 64.0, 9)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=10000.0)
    cfg.Collision = ib.ParticleCollision(DistanceField=df, DistanceFieldMaximumZ=64.0, EscapeVelocity=50.0, BounceVelocityMultiplier=0.5,
                                         Distance=1.0, LifePenalty=0.25)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    system.Transforms = []
    dt = 0.02
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[0] = [20, 30, 8, 1]; V[0] = [100, -50, 25, 2]
    cleared = np.zeros((df.TextureHeight, df.TextureWidth, 4), np.uint16)
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), engine.RandomnessTexture, cleared, 1)
    assert np.allclose(P2[0, :3], P[0, :3] + V[0, :3] * dt, rtol=2e-6) and P2[0, 3] == 1.0
    assert np.allclose(V2[0, :3], V[0, :3], rtol=2e-6) and V2[0, 3] == 1.0            # bounce delay 2 -> 1
    # a wall: box of half size 10 centred at x = 80; its -x face is the plane x = 70
    box = ib.LightObstruction(ib.LightObstructionType.Box, (80.0, 64.0, 0.0), (10.0, 40.0, 200.0))
    tex = oracle.generate_distance_field(df, [box])
    P[0] = [60, 64, 8, 1]; V[0] = [1000, 0, 0, 0]        # 20 px per step, 10 px in front of the face, category 0 = bounces
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), engine.RandomnessTexture, tex, 1)
    assert 60.0 <= P2[0, 0] <= 70.0 and abs(P2[0, 1] - 64.0) < 1e-3 and abs(P2[0, 2] - 8.0) < 1e-3     # stopped in front of the face
    assert P2[0, 0] >= 70.0 - 1.0 - 16.0                                                              # after at most one back-off
    assert V2[0, 0] == pytest.approx(-1000.0 * 0.5, rel=2e-3) and abs(V2[0, 1]) < 2.0 and abs(V2[0, 2]) < 2.0   # reflected off the x face
    assert V2[0, 3] == 3.0 and P2[0, 3] == pytest.approx(1.0 - 0.25, abs=1e-6)


@pytest.mark.parametrize("replace", [True, False])
def test_noise_transform_closed_form(oracle, replace):
    """PS_Noise (Noise.fx:28-72) + PS_Update in float64 over a CONSTANT randomness table (every lookup returns the same texel,
    so the four random vectors are known): delta = sign(r + offset) * max(|r + offset|, minimum) * scale; the position is
    lerped by t = Strength * dt * CyclesPerSecond; the velocity is lerped towards the delta by the WEIGHT (ReplaceOldVelocity)
    or towards old + delta by t, plus normalize(old velocity) * the speed delta."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=10000.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    noise = ib.Noise(CyclesPerSecond=5, PositionOffset=(-0.5, -0.5, -0.5, -0.5), PositionMinimum=(0.0, 0.4, 0.0, 0.0), PositionScale=(8.0, 4.0, 2.0, 0.0),
                     VelocityOffset=(-0.5, 0.25, -0.5), VelocityMinimum=(0.0, 0.0, 0.5), VelocityScale=(100.0, 50.0, 20.0),
                     SpeedOffset=0.25, SpeedMinimum=0.0, SpeedScale=10.0, ReplaceOldVelocity=replace, Strength=0.5)
    system.Transforms = [noise]
    texel = np.array([0.75, 0.25, 0.625, 0.125], np.float32)
    table = np.broadcast_to(texel, engine.RandomnessTexture.shape).copy()
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
    P[0] = [30, 40, 5, 1]; V[0] = [30, 0, 40, 0]
    dt = 0.02
    P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), table, None, 1)
    r = texel.astype(np.float64)

    def delta(offset, minimum, scale):
        d = r + np.array(offset, np.float64)
        return np.sign(d) * np.maximum(np.abs(d), np.array(minimum, np.float64)) * np.array(scale, np.float64)
    dp = delta((-0.5, -0.5, -0.5, -0.5), (0.0, 0.4, 0.0, 0.0), (8.0, 4.0, 2.0, 0.0))
    dv = delta((-0.5, 0.25, -0.5, 0.25), (0.0, 0.0, 0.5, 0.0), (100.0, 50.0, 20.0, 10.0))
    assert dp[1] == pytest.approx(-0.4 * 4.0) and dv[2] == pytest.approx(0.5 * 20.0)      # both minimums engage
    weight, t = 0.5, 0.5 * dt * 5.0
    p0, v0 = np.array([30.0, 40.0, 5.0]), np.array([30.0, 0.0, 40.0])
    p1 = p0 + t * dp[:3]
    v1 = (v0 + weight * (dv[:3] - v0)) if replace else (v0 + t * dv[:3])
    v1 = v1 + v0 / np.linalg.norm(v0) * dv[3]
    assert np.allclose(V2[0, :3], v1, rtol=3e-6, atol=1e-5) and V2[0, 3] == 0.0
    assert np.allclose(P2[0, :3], p1 + v1 * dt, rtol=3e-6) and P2[0, 3] == pytest.approx(1.0 + t * dp[3])


def test_render_data_closed_form(oracle):
    """computeRenderData (UpdateCommon.fxh:97-117) by hand with the default ramps (all 1) and OpacityFromLife: renderColor =
    attributes * (1, 1, 1, life / OpacityFromLife) with alpha saturated and rgb premultiplied; renderData = (size 1, atan2 of the
    velocity brought into [0, 2 pi) + life * RotationFromLife + index * RotationFromIndex, max(|v|, 1e-4), velocity.w), index =
    x + y * 256 (the reference's own "FIXME" constant, whatever the chunk size)."""
    import math
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=10000.0, RotationFromVelocity=True,
                                         RotationFromLife=30.0, RotationFromIndex=2.0, OpacityFromLife=4.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    system.Transforms = []
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.zeros((256, 4), np.float32)
    i = 3 + 2 * 16                                   # texel (3, 2) of the 16 x 16 chunk
    P[i] = [10, 20, 0, 2.0]; V[i] = [30, -40, 0, 1.0]; A[i] = [0.8, 0.6, 0.4, 0.9]
    P[5] = [1, 1, 0, 8.0]; V[5] = [0.001, 0.002, 0, 0]; A[5] = [1, 1, 1, 1]     # slow (no rotation from velocity), alpha saturates
    P2, V2, A2, RC, RD = oracle.particles_step(P, V, A, 16, system.system_uniforms(0.02), [], system.plan_ops(0.0), engine.RandomnessTexture, None, 1)
    alpha = min(max(0.9 * (2.0 / 4.0), 0.0), 1.0)
    assert np.allclose(RC[i], [0.8 * alpha, 0.6 * alpha, 0.4 * alpha, alpha], rtol=2e-6)
    angle = math.atan2(-40.0, 30.0) + 2.0 * math.pi
    index = 3 + 2 * 256
    assert RD[i, 0] == pytest.approx(1.0, rel=1e-6)
    assert RD[i, 1] == pytest.approx(angle + 2.0 * math.radians(30.0) + index * math.radians(2.0), rel=3e-6)
    assert RD[i, 2] == pytest.approx(50.0, rel=1e-6) and RD[i, 3] == 1.0
    assert np.allclose(RC[5], [1, 1, 1, 1]) and RD[5, 1] == pytest.approx(8.0 * math.radians(30.0) + 5 * math.radians(2.0), rel=3e-6)
    assert RD[5, 2] == pytest.approx(math.hypot(0.001, 0.002), rel=1e-5)
    assert not RC[0].any() and not RD[0].any()       # dead particle


def test_gravity_mixed_attractors_and_the_acceleration_clamp(oracle):
    """Gravity.fx:12-61 in float64: a Physical attractor (1 / max(|d|^2 - radius, 0.001), not time-scaled), a Linear and an
    Exponential one (time-scaled falloffs) are summed, the sum is clamped to MaximumAcceleration * dt, and the new velocity is
    capped per component at MaximumVelocity."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    attractors = [ib.Attractor(Position=(40, 10, 0), Radius=50.0, Strength=300.0, Type=ib.AttractorType.Physical),
                  ib.Attractor(Position=(0, 60, 20), Radius=120.0, Strength=900.0, Type=ib.AttractorType.Linear),
                  ib.Attractor(Position=(-30, -30, 10), Radius=90.0, Strength=1500.0, Type=ib.AttractorType.Exponential)]
    dt = 0.02
    p0, v0 = np.array([5.0, 8.0, 2.0]), np.array([3.0, -2.0, 1.0])

    def expected(max_accel, max_velocity):
        acc = np.zeros(3)
        for a in attractors:
            to = np.array(a.Position, np.float64) - p0
            dist = np.linalg.norm(to)
            if a.Type == ib.AttractorType.Physical:
                att = 1.0 / max(np.dot(to, to) - a.Radius, 0.001)
            else:
                att = 1.0 - min(max(dist / a.Radius, 0.0), 1.0)
                if a.Type == ib.AttractorType.Exponential:
                    att *= att
                att = att * dt          # getDeltaTime() / VelocityConstantScale
            acc += to / dist * att * a.Strength
        cap = max_accel * dt
        if np.linalg.norm(acc) > cap:
            acc = acc / np.linalg.norm(acc) * cap
        return np.minimum(max_velocity, v0 + acc), np.linalg.norm(acc)

    for max_accel, max_velocity, clamped in ((1.0e6, 1000.0, False), (200.0, 1000.0, True), (1.0e6, 4.0, False)):
        cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=max_velocity)
        system = ib.ParticleSystem(engine, cfg, maxChunks=1)
        system.Transforms = [ib.Gravity(MaximumAcceleration=max_accel, Attractors=attractors)]
        P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.ones((256, 4), np.float32)
        P[0] = [*p0, 1]; V[0] = [*v0, 0]
        P2, V2, *_ = oracle.particles_step(P, V, A, 16, system.system_uniforms(dt), [], system.plan_ops(0.0), engine.RandomnessTexture, None, 1)
        want, acc_len = expected(max_accel, max_velocity)
        assert (acc_len == pytest.approx(max_accel * dt)) == clamped
        if max_velocity > 100.0:      # (the update tail rescales a velocity longer than MaximumVelocity: checked on the op's output otherwise)
            assert np.allclose(V2[0, :3], want, rtol=3e-6), (max_accel, V2[0], want)
            assert np.allclose(P2[0, :3], p0 + want * dt, rtol=3e-6)
        else:
            capped = want / np.linalg.norm(want) * min(np.linalg.norm(want), max_velocity)     # applyFrictionAndMaximum, UpdateCommon.fxh:20-35
            assert np.any(want == max_velocity) and np.allclose(V2[0, :3], capped, rtol=3e-6), (V2[0], want, capped)


def test_spawner_linear_formulas_closed_form(oracle):
    """PS_Spawn (SpawnParticles.fx:10-30) -> Spawn_Stage1 / Spawn_Stage2 / evaluateFormula (SpawnerCommon.fxh:62-70, 119-188) by
    hand over a CONSTANT randomness table (random1 = random2 = random3 = the texel): every spawned particle gets position =
    constant + (r + offset) * scale with the life formula in .w, velocity likewise with the category in .w, attributes =
    ColorConstant + (r + ColorOffset) * ColorRandomScale; the update pass of the same tick then advances it by velocity * dt.
    Exactly rate * dt particles appear, in index order from the start of the chunk; the rest of the chunk stays dead."""
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=16))
    cfg = ib.ParticleSystemConfiguration(Friction=0.0, LifeDecayPerSecond=0.0, MaximumVelocity=10000.0)
    system = ib.ParticleSystem(engine, cfg, maxChunks=1)
    spawner = ib.Spawner(MinRate=500.0, MaxRate=500.0,
                         Position=ib.Formula(Constant=(10.0, 20.0, 5.0), RandomScale=(4.0, 2.0, 1.0), Offset=(-0.5, -0.5, 0.0), Type=ib.FormulaType.Linear),
                         Velocity=ib.Formula(Constant=(1.0, 2.0, 3.0), RandomScale=(10.0, 20.0, 40.0), Offset=(0.0, -1.0, 0.25), Type=ib.FormulaType.Linear),
                         Life=(2.0, 1.0, -0.5), Category=(0.0, 0.0, 0.0), ColorConstant=(0.5, 0.5, 0.5, 1.0), ColorRandomScale=(0.1, 0.2, 0.4, 0.0))
    system.Transforms = [spawner]
    texel = np.array([0.75, 0.25, 0.625, 0.125], np.float32)
    table = np.broadcast_to(texel, engine.RandomnessTexture.shape).copy()
    dt = 0.02
    spawns, ops, u = system.plan_spawns(dt, dt), system.plan_ops(dt), system.system_uniforms(dt)
    assert system.LiveChunkCount == 1
    P = np.zeros((256, 4), np.float32); V = np.zeros((256, 4), np.float32); A = np.zeros((256, 4), np.float32)
    P2, V2, A2, *_ = oracle.particles_step(P, V, A, 16, u, spawns, ops, table, None, 1)
    r = texel.astype(np.float64)
    pos = np.array([10.0, 20.0, 5.0]) + (r[:3] + np.array([-0.5, -0.5, 0.0])) * np.array([4.0, 2.0, 1.0])
    vel = np.array([1.0, 2.0, 3.0]) + (r[:3] + np.array([0.0, -1.0, 0.25])) * np.array([10.0, 20.0, 40.0])
    life = 2.0 + (r[3] - 0.5) * 1.0
    col = np.array([0.5, 0.5, 0.5, 1.0]) + r * np.array([0.1, 0.2, 0.4, 0.0])
    n = int(round(500.0 * dt))
    live = P2[:, 3] > 0
    assert live.sum() == n and live[:n].all()
    for i in range(n):
        assert np.allclose(V2[i], [*vel, 0.0], rtol=3e-6), (i, V2[i])
        assert np.allclose(P2[i], [*(pos + vel * dt), life], rtol=3e-6), (i, P2[i])
        assert np.allclose(A2[i], col, rtol=3e-6)
    assert not P2[n:].any() and not V2[n:].any()


def test_area_weight_quirk_scalar_rotation(oracle):
    """AreaRotation is a scalar broadcast into a quaternion (FMA.fx:11,17): rotation 0 collapses the local position to 0,
    so the weight is `Strength` everywhere (distance = -min size); a unit quaternion would be the identity."""
    assert oracle.evaluate_by_type_id(2, (100, 0, 0), (0, 0, 0), (5, 6, 7), (0, 0, 0, 0)) == pytest.approx(-5.0)
    assert oracle.evaluate_by_type_id(2, (100, 0, 0), (0, 0, 0), (5, 6, 7), (0, 0, 0, 1)) == pytest.approx(95.0)
    assert oracle.evaluate_by_type_id(1, (0, 0, 0), (0, 0, 0), (5, 6, 7)) == pytest.approx(-5.0)     # ellipsoid centre
    assert oracle.evaluate_by_type_id(3, (0, 0, 30), (0, 0, 0), (3, 4, 10)) == pytest.approx(20.0)   # capped cylinder above the cap
    assert oracle.evaluate_by_type_id(0, (1, 2, 3), (0, 0, 0), (1, 1, 1)) == 0.0                     # AreaType None


def test_distance_field_generation_box_slice(oracle):
    df = _field(64, 64, 90.0, 9)
    box = ib.LightObstruction(ib.LightObstructionType.Box, (32.0, 32.0, 0.0), (8.0, 8.0, 200.0))
    tex = oracle.generate_distance_field(df, [box])
    dec = lambda c: (192.0 / 255.0 - c / 65535.0) * 128.0
    assert dec(tex[32, 32, 0]) == pytest.approx(-8.0, abs=2e-3)        # centre of the box: -half size
    assert dec(tex[32, 50, 0]) == pytest.approx(10.0, abs=2e-3)        # 18 px right of centre, 8 px half size
    assert dec(tex[32, 50, 3]) == pytest.approx(10.0, abs=2e-3)        # slice 3 (z = 30) still inside the tall box's span
    # texel (r,g,b,a) = z-slices 3p..3p+3: channel a of physical slice 0 == channel r of physical slice 1
    ox = df.SliceWidth
    assert np.array_equal(tex[:64, :64, 3], tex[:64, ox:ox + 64, 0])

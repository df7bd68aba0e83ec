"""SphereLightWithDistanceRamp (Shaders/SphereLight.fx:48-87, SphereLightCore.fxh:99-119, :160-199): a sphere light whose
LightSource.RampTexture is set takes its rgb from RampTexture(preTraceOpacity, (atan2(dy, dx) + offset) * rate) * coneOpacity.

CPU: known answers for the oracle (a constant ramp scales the light, an identity ramp reproduces the plain sphere light, a
1x1 texture means "no ramp", V wraps with the angle).  GPU: the CUDA path against the oracle through the C-ABI."""
import math

import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from helpers import LIGHTING_RTOL, lighting_rel_err, make_renderer, oracle_lightmap


def _scene(seed=41, w=96, h=64, lights=3):
    s = scenes.lighting_scene(seed, w, h, lights, ramp=(40.0, 120.0), float4_lightmap=True)
    for l in s.environment.Lights:
        l.Color = (0.9, 0.7, 0.5, 1.0)
    return s


def _cpu_lightmap(oracle, s):
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, s.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField, r._gbuffer_shape = df, s.gbuffer.shape[:2]
    frame = r.build_frame()
    batches, nb, verts, nv = r.build_batches()
    oracle.set_ramp_textures(r.ramp_textures)
    try:
        return oracle.render_lighting(tex, s.gbuffer, frame, batches, nb, verts, nv), nb
    finally:
        oracle.set_ramp_textures(None)


def _identity_ramp(w=256, h=4):
    t = np.zeros((h, w, 4), np.float32)
    t[..., :3] = ((np.arange(w) + 0.5) / w)[None, :, None]      # texel centres: RampTexture(u).rgb == u under LINEAR filtering
    t[..., 3] = 1
    return t


def test_constant_ramp_scales_the_cone_opacity(oracle):
    plain = _scene()
    for l in plain.environment.Lights:
        l.CastsShadows = False                                   # coneOpacity == 1: the plain light is color * preTraceOpacity
    ref, nb0 = _cpu_lightmap(oracle, plain)
    ramped = _scene()
    const = np.zeros((3, 5, 4), np.float32)
    const[..., :3] = (0.25, 0.5, 1.0)
    for l in ramped.environment.Lights:
        l.CastsShadows = False
        l.RampTexture = const
    out, nb1 = _cpu_lightmap(oracle, ramped)
    assert nb0 == nb1 == 1                                       # one texture, one render state
    amb = np.array(plain.environment.Ambient[:3], np.float32)
    lit = out[..., 3] >= 1                                       # alpha counts the light passes that did not discard (cleared to 0: fullbright mode)
    assert lit.any() and np.array_equal(out[..., 3], ref[..., 3])     # discards are decided by the distance falloff alone
    # every lit fragment contributes color * ramp(…) = color * (0.25, 0.5, 1) regardless of its falloff
    n = out[..., 3][..., None]
    want = amb + n * np.array([0.9 * 0.25, 0.7 * 0.5, 0.5 * 1.0], np.float32)
    assert np.allclose(out[..., :3][lit], want[lit], rtol=2e-6, atol=1e-6)


def test_identity_ramp_reproduces_the_plain_light_and_1x1_means_none(oracle):
    plain, ramped, tiny = _scene(), _scene(), _scene()
    ident = _identity_ramp()
    for l in ramped.environment.Lights:
        l.RampTexture = ident
    for l in tiny.environment.Lights:
        l.RampTexture = np.full((1, 1, 4), 0.3, np.float32)
    ref, _ = _cpu_lightmap(oracle, plain)
    out, _ = _cpu_lightmap(oracle, ramped)
    one, _ = _cpu_lightmap(oracle, tiny)
    assert np.array_equal(one, ref)                              # LightingRenderer.cs:819-827
    # RampTexture(u) = u up to the clamp at the first / last half texel (|error| <= 0.5 / 256) times colour <= 0.9, three lights
    assert np.abs(out[..., :3] - ref[..., :3]).max() <= 3 * 0.9 * 0.5 / 256 + 1e-5
    assert np.array_equal(out[..., 3], ref[..., 3])


def test_ramp_v_follows_the_angle_and_wraps(oracle):
    s = _scene(lights=0)
    rows = np.zeros((4, 2, 4), np.float32)                       # four angular sectors, constant along u
    rows[0, :, 0], rows[1, :, 1], rows[2, :, 2], rows[3, :, :3] = 1, 1, 1, 1
    # above the raised box of the synthetic G-buffer, so that every pixel faces the light
    l = ib.SphereLightSource(Position=(48.0, 32.0, 100.0), Radius=10.0, RampLength=200.0, CastsShadows=False, Color=(1.0, 1.0, 1.0, 1.0))
    l.RampTexture = rows
    s.environment.Lights = [l]
    s.environment.Ambient = (0.0, 0.0, 0.0, 1.0)
    out, _ = _cpu_lightmap(oracle, s)
    # v = (atan2(dy, dx) - pi) / (2 pi) in [-1, 0]: sector k covers v in [k / 4 - 1, (k + 1) / 4 - 1); sample the sector centres
    for k, channel in enumerate([(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)]):
        ang = -math.pi + (k + 0.5) * math.pi / 2
        x, y = int(48 + 20 * math.cos(ang)), int(32 + 20 * math.sin(ang))
        px = out[y, x, :3]
        assert np.allclose(px / max(px.max(), 1e-9), channel, atol=0.02), (k, px)


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_ramp_textured_sphere_lights_match_the_oracle(ctx, oracle):
    s = scenes.lighting_scene(43, 160, 96, 6, n_directional=1, n_line=1, ramp=(50.0, 140.0), ao=True, float4_lightmap=True)
    rs = np.random.RandomState(9)
    smooth = np.zeros((8, 64, 4), np.float32)
    u, v = np.meshgrid((np.arange(64) + 0.5) / 64, (np.arange(8) + 0.5) / 8)
    smooth[..., 0], smooth[..., 1], smooth[..., 2] = u, u * (0.6 + 0.4 * np.cos(2 * np.pi * v)), u * u
    bytes_ramp = np.floor(np.clip(rs.rand(2, 16, 4), 0, 1) * 255 + 0.5).astype(np.uint8)
    spheres = [l for l in s.environment.Lights if isinstance(l, ib.SphereLightSource)]
    for i, l in enumerate(spheres):
        l.RampTexture = [smooth, None, bytes_ramp][i % 3]
        l.RampOffsetAndRate = (0.3 * i, 1.0 + (i % 2))
        if i == 0:
            l.SpecularColor, l.SpecularPower = (0.3, 0.2, 0.1), 3.0
    r, tex = make_renderer(ctx, s)
    gpu = r.RenderLighting()
    batches, nb, _, _ = r.build_batches()
    assert sorted(batches[b].ramp_texture for b in range(nb))[-1] == 2 and len(r.ramp_textures) == 2
    oracle.set_ramp_textures(r.ramp_textures)
    try:
        ref = oracle_lightmap(oracle, r, tex, s)
    finally:
        oracle.set_ramp_textures(None)
    err = lighting_rel_err(gpu, ref)
    assert err.max() <= LIGHTING_RTOL, f"max rel err {err.max():.3e}"
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    # the ramp changes the picture (this is not the plain path)
    for l in spheres:
        l.RampTexture = None
    r2, _ = make_renderer(ctx, s)
    assert np.abs(r2.RenderLighting()[..., :3] - gpu[..., :3]).max() > 0.05


@pytest.mark.gpu
def test_ramp_texture_api_errors(ctx):
    import ctypes as C
    from illuminant_b200 import _abi
    rid = C.c_int32(0)
    t = np.zeros((2, 2, 4), np.float32)
    assert ctx.lib.ilb_ramp_texture_create(ctx.handle, 2, 2, _abi.FORMAT_HALF4, t.ctypes.data_as(C.c_void_p), C.byref(rid)) == _abi.ERR_INVALID_ARGUMENT
    ctx.check(ctx.lib.ilb_ramp_texture_create(ctx.handle, 2, 2, _abi.FORMAT_FLOAT4, t.ctypes.data_as(C.c_void_p), C.byref(rid)))
    assert rid.value >= 1
    ctx.check(ctx.lib.ilb_ramp_texture_destroy(ctx.handle, rid.value))
    assert ctx.lib.ilb_ramp_texture_destroy(ctx.handle, rid.value) == _abi.ERR_INVALID_ARGUMENT
    # a batch that names a destroyed texture is rejected; so is a ramp on a directional light
    s = scenes.lighting_scene(44, 64, 48, 1, n_directional=1, float4_lightmap=True)
    r, _ = make_renderer(ctx, s)
    batches, nb, verts, nv = r.build_batches()
    frame = r.build_frame()
    out = np.empty((48, 64, 4), np.float32)
    batches[0].ramp_texture = rid.value
    rc = ctx.lib.ilb_render_lighting(ctx.handle, r.DistanceField.handle, C.byref(frame), C.cast(batches, C.c_void_p), nb, C.cast(verts, C.c_void_p), nv,
                                     out.ctypes.data_as(C.c_void_p))
    assert rc == _abi.ERR_INVALID_ARGUMENT

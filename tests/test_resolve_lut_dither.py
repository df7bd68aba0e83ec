"""N3 remainder: the LUT-blended resolve (Shaders/LUTResolve.fx:57-135) and ApplyDither.

CPU part: known answers for the oracle restatement -- identity LUTs make the LUT resolve `saturate(albedo) * light * 2`, the
dark / neutral / bright bands select what LUTResolve.fx:93-112 says they select, PerChannel blends per channel, LUTOnly drops the
light factor; the dither convention (ilb_dithering) quantises to multiples of 1 / Unit with the documented 17-periodic threshold.
GPU part: ilb_resolve_lighting_lut / ilb_set_dithering against the oracle through the C-ABI.

ReadLUT and ApplyDither live in the un-vendored sq/Fracture: what is pinned here is the blending logic of LUTResolve.fx (which IS
in the reference) and the conventions stated in include/illuminant_b200.h for the two helpers."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi, hdr as H

f32 = np.float32


def _params(w, h, hdr=None, lm=_abi.FORMAT_FLOAT4, al=_abi.FORMAT_FLOAT4):
    return H.pack_resolve(w, h, lm, hdr, al, _abi.FORMAT_FLOAT4)


def _tinted(res, gain, rows=1):
    """An identity table scaled per channel (a colour grade whose effect is easy to predict)."""
    t = ib.ColorLUT.Identity(res).Texture.astype(np.float32)
    t[..., :3] = np.clip(t[..., :3] * np.asarray(gain, np.float32), 0, 255)
    t = np.round(t).astype(np.uint8)
    return ib.ColorLUT(np.ascontiguousarray(np.tile(t, (rows, 1, 1))), res, rows)


def _scene(seed, w, h):
    rs = np.random.RandomState(seed)
    lm = (rs.rand(h, w, 4) * 1.2).astype(np.float32)
    lm[..., 3] = np.floor(rs.rand(h, w) * 3)
    al = rs.rand(h, w, 4).astype(np.float32)
    return lm, al


def test_identity_luts_reduce_to_albedo_times_light(oracle):
    lm, al = _scene(1, 9, 7)
    cfg = ib.LUTBlendingConfiguration(ib.ColorLUT.Identity(16), ib.ColorLUT.Identity(16), DarkLevel=0.2, BrightLevel=0.9)
    out = oracle.resolve_lighting_lut(_params(9, 7), cfg.pack(), cfg.DarkLUT.Texture, cfg.BrightLUT.Texture, lm, al)
    want = np.clip(al[..., :3], 0, 1) * (lm[..., :3] * f32(2))
    assert np.abs(out[..., :3] - want).max() <= 2.5 / 255          # the 8-bit table quantises the identity to 1 / 255 per entry
    assert np.array_equal(out[..., 3], al[..., 3])                  # alpha = albedo.a (LUTResolve.fx:115)
    # LUTOnly drops the light factor (:115)
    cfg.LUTOnly = True
    out = oracle.resolve_lighting_lut(_params(9, 7), cfg.pack(), cfg.DarkLUT.Texture, cfg.BrightLUT.Texture, lm, al)
    assert np.abs(out[..., :3] - np.clip(al[..., :3], 0, 1)).max() <= 2.5 / 255


def test_bands_select_dark_neutral_and_bright(oracle):
    """weight = dot(light * 2, (0.299, 0.587, 0.144)); below DarkLevel the dark table, above BrightLevel the bright one, and with a
    neutral band the plain albedo in between (LUTResolve.fx:74-112)."""
    dark, bright = _tinted(8, (0.5, 0.5, 0.5)), _tinted(8, (1.0, 0.25, 0.25))
    al = np.zeros((1, 3, 4), np.float32)
    al[..., :3] = (4 / 7, 2 / 7, 6 / 7)       # on the table's lattice: ReadLUT returns table entries exactly
    al[..., 3] = 1
    lm = np.zeros((1, 3, 4), np.float32)
    gray = lambda v: v / (2 * (0.299 + 0.587 + 0.144))
    lm[0, 0, :3], lm[0, 1, :3], lm[0, 2, :3] = gray(0.05), gray(0.5), gray(1.4)
    cfg = ib.LUTBlendingConfiguration(dark, bright, LUTOnly=True, DarkLevel=0.2, NeutralBandSize=0.3, BrightLevel=0.9)
    out = oracle.resolve_lighting_lut(_params(3, 1), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    a = al[0, 0, :3]
    q = lambda v: np.round(np.asarray(v) * 255) / 255
    assert np.allclose(out[0, 0, :3], q(q(a) * 0.5), atol=1.5 / 255)             # dark table
    assert np.allclose(out[0, 1, :3], a, atol=1e-6)                               # neutral band: the albedo itself (:95-98)
    assert np.allclose(out[0, 2, :3], q(q(a) * (1.0, 0.25, 0.25)), atol=1.5 / 255)  # bright table
    # without a neutral band: a straight ramp between the two tables over [DarkLevel, BrightLevel] (:99-112)
    cfg.NeutralBandSize = 0.0
    lm[0, 1, :3] = gray(0.55)     # the middle of [0.2, 0.9]
    out = oracle.resolve_lighting_lut(_params(3, 1), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    mid = 0.5 * (q(q(a) * 0.5) + q(q(a) * (1.0, 0.25, 0.25)))
    assert np.allclose(out[0, 1, :3], mid, atol=2.0 / 255)
    # BrightLevel <= DarkLevel: the "HACK" branch, weight = saturate(weight - DarkLevel) (:105-108)
    cfg.BrightLevel, cfg.DarkLevel = 0.1, 0.1
    lm[0, 1, :3] = gray(0.6)      # weight - DarkLevel = 0.5
    out = oracle.resolve_lighting_lut(_params(3, 1), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    assert np.allclose(out[0, 1, :3], mid, atol=2.0 / 255)


def test_per_channel_weights(oracle):
    dark, bright = _tinted(8, (0.0, 0.0, 0.0)), ib.ColorLUT.Identity(8)
    al = np.ones((1, 1, 4), np.float32)
    lm = np.zeros((1, 1, 4), np.float32)
    lm[0, 0, :3] = (0.0, 0.25, 0.5)           # x2 -> per-channel weights 0, 0.5, 1 over [0, 1]
    cfg = ib.LUTBlendingConfiguration(dark, bright, PerChannel=True, LUTOnly=True, DarkLevel=0.0, BrightLevel=1.0)
    out = oracle.resolve_lighting_lut(_params(1, 1), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    assert np.allclose(out[0, 0, :3], (0.0, 0.5, 1.0), atol=1e-6)
    cfg.PerChannel = False                    # normalised: one gray weight for the three channels (:80-83)
    out = oracle.resolve_lighting_lut(_params(1, 1), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    w = 0.5 * 0.587 + 1.0 * 0.144
    assert np.allclose(out[0, 0, :3], (w, w, w), atol=1e-6)


def test_lut_resolve_rejects_hdr_modes(oracle):
    lm, al = _scene(2, 4, 4)
    cfg = ib.LUTBlendingConfiguration(ib.ColorLUT.Identity(4), ib.ColorLUT.Identity(4))
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=2.0))
    with pytest.raises(RuntimeError):   # LightingRenderer.cs:1593-1594
        oracle.resolve_lighting_lut(_params(4, 4, hdr), cfg.pack(), cfg.DarkLUT.Texture, cfg.BrightLUT.Texture, lm, al)


def _dither(strength=1.0, unit=255.0, frame=0.0, band=1.0, lo=0.0, hi=1.0):
    return ib.DitheringSettings(Unit=unit, Strength=strength, FrameIndex=frame, BandSize=band, RangeMin=lo, RangeMax=hi)


def test_dither_convention_known_answers(oracle):
    w, h = 34, 5
    lm = np.zeros((h, w, 4), np.float32)
    lm[..., :3] = (10.5 / 255, 3.25 / 255, 200.0 / 255)
    try:
        oracle.set_dithering(_dither().pack())
        out = oracle.resolve_lighting(_params(w, h), lm)
    finally:
        oracle.set_dithering(None)
    x, y = np.meshgrid(np.arange(w), np.arange(h))
    ph = f32(23.0) * f32(0.5) / f32(17.0)                              # frac(23 * ((FrameIndex mod 4) + 0.5) / 17) in fp32
    s = ((2 * x + 7 * y) % 17).astype(np.float32) * (f32(1.0) / f32(17.0)) + (ph - np.floor(ph))
    t = s - np.floor(s)
    assert np.array_equal(np.round(out[..., 0] * 255) == 11, f32(0.5) >= t)      # rounds up exactly where the fraction reaches the threshold
    assert np.array_equal(np.round(out[..., 1] * 255) == 4, f32(0.25) >= t)
    levels = np.round(out[..., :3] * 255)
    assert np.abs(out[..., :3] * 255 - levels).max() < 1e-4          # every value sits on a multiple of 1 / Unit
    assert set(np.unique(levels[..., 0])) == {10.0, 11.0} and set(np.unique(levels[..., 1])) == {3.0, 4.0}
    assert np.all(levels[..., 2] == 200)                              # already on the lattice: untouched
    # the mean over a 17-periodic row recovers the value (what ordered dithering is for)
    assert abs(levels[0, :17, 0].mean() - 10.5) <= 0.5 / 17 + 1e-6 and abs(levels[0, :17, 1].mean() - 3.25) <= 1.0 / 17 + 1e-6
    # Strength 0 (the handler's default) is the identity; Strength 0.5 lands half way; values outside the range are kept
    assert np.array_equal(oracle.resolve_lighting(_params(w, h), lm), np.concatenate([lm[..., :3], np.ones((h, w, 1), np.float32)], -1))
    try:
        oracle.set_dithering(_dither(strength=0.5).pack())
        half = oracle.resolve_lighting(_params(w, h), lm)
        oracle.set_dithering(_dither(lo=0.5).pack())
        ranged = oracle.resolve_lighting(_params(w, h), lm)
        oracle.set_dithering(_dither(frame=1.0).pack())
        moved = oracle.resolve_lighting(_params(w, h), lm)
    finally:
        oracle.set_dithering(None)
    assert np.allclose(half[..., :3], 0.5 * (lm[..., :3] + out[..., :3]), atol=1e-7)
    assert np.array_equal(ranged[..., :2], lm[..., :2])               # below RangeMin: untouched
    assert not np.array_equal(moved[..., 0], out[..., 0])             # the pattern moves with the frame index


# ------------------------------------------------------------------------------------------------------------------- GPU
def _rendered(ctx, w, h, fmt=_abi.FORMAT_HALF4):
    r = ib.LightingRenderer(ctx, ib.LightingEnvironment(), ib.RendererConfiguration((w, h)))
    return ib.RenderedLighting(r, w, h, fmt, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("per_channel,lut_only,neutral", [(False, False, 0.0), (True, False, 0.0), (False, True, 0.25), (True, True, 0.0)])
@pytest.mark.parametrize("w,h,lm_dtype,al_dtype", [(64, 40, np.float16, np.uint8), (37, 23, np.float32, np.float32)])
def test_lut_resolve_matches_oracle(ctx, oracle, per_channel, lut_only, neutral, w, h, lm_dtype, al_dtype):
    rs = np.random.RandomState(w + 2 * int(per_channel) + int(lut_only))
    lm = (rs.rand(h, w, 4) * 1.1).astype(lm_dtype)
    al = rs.rand(h, w, 4).astype(np.float32)
    al = np.floor(al * 255 + 0.5).astype(np.uint8) if al_dtype == np.uint8 else al
    dark, bright = _tinted(16, (0.6, 0.5, 0.9)), _tinted(8, (1.0, 0.8, 0.4), rows=2)
    cfg = ib.LUTBlendingConfiguration(dark, bright, PerChannel=per_channel, LUTOnly=lut_only, DarkLevel=0.15, NeutralBandSize=neutral, BrightLevel=0.95)
    hdr = ib.HDRConfiguration(InverseScaleFactor=0.9, Offset=-0.01, Exposure=1.3, Gamma=1.0, ResolveToSRGB=bool(per_channel))
    lm_fmt = _abi.FORMAT_HALF4 if lm_dtype == np.float16 else _abi.FORMAT_FLOAT4
    al_fmt = _abi.FORMAT_RGBA8 if al_dtype == np.uint8 else _abi.FORMAT_FLOAT4
    ref = oracle.resolve_lighting_lut(ib.pack_resolve(w, h, lm_fmt, hdr, al_fmt, _abi.FORMAT_FLOAT4), cfg.pack(), dark.Texture, bright.Texture, lm, al)
    gpu = _rendered(ctx, w, h).Resolve(al, hdr, float4=True, lightmap=lm, lutBlending=cfg)
    err = np.abs(gpu.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1 / 255)
    assert err.max() <= 1e-4, err.max()
    rgba = _rendered(ctx, w, h).Resolve(al, hdr, lightmap=lm, lutBlending=cfg)
    want = np.floor(np.clip(ref, 0, 1) * 255 + 0.5).astype(np.int32)
    diff = np.abs(rgba.astype(np.int32) - want)
    assert diff.max() <= 1 and (diff != 0).mean() < 5e-3
    with pytest.raises(_abi.IlluminantError):      # LUT blending with a tone-mapped resolve throws (LightingRenderer.cs:1593-1594)
        _rendered(ctx, w, h).Resolve(al, ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, ToneMapping=ib.ToneMappingConfiguration(2.0)), lightmap=lm, lutBlending=cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("frame,unit,strength", [(0.0, 255.0, 1.0), (3.0, 31.0, 1.0), (6.0, 255.0, 0.5)])
def test_dither_is_bit_exact_on_exact_inputs(ctx, oracle, frame, unit, strength):
    """Plain resolve of a float lightmap at the effect defaults is the identity in both implementations, so the dither decisions see
    the same bits and the comparison is exact (Strength 1) or one FMA rounding apart (the blend at Strength 0.5)."""
    w, h = 130, 37
    lm = np.random.RandomState(int(frame) + 11).rand(h, w, 4).astype(np.float32)
    ds = _dither(strength=strength, unit=unit, frame=frame, band=0.9, lo=0.05, hi=0.97)
    hdr = ib.HDRConfiguration(Dithering=ds)
    try:
        oracle.set_dithering(ds.pack())
        ref = oracle.resolve_lighting(ib.pack_resolve(w, h, _abi.FORMAT_FLOAT4, hdr, _abi.FORMAT_RGBA8, _abi.FORMAT_FLOAT4), lm)
    finally:
        oracle.set_dithering(None)
    rl = _rendered(ctx, w, h)
    gpu = rl.Resolve(None, hdr, float4=True, lightmap=lm)
    if strength == 1.0:
        assert np.array_equal(gpu, ref)
    else:
        assert np.abs(gpu - ref).max() <= 1.2e-7
    assert not np.array_equal(gpu[..., :3], lm[..., :3])
    # the settings are per call: a resolve without a Dithering configuration is back to the identity
    assert np.array_equal(rl.Resolve(None, None, float4=True, lightmap=lm)[..., :3], lm[..., :3])


@pytest.mark.gpu
def test_dithered_tone_mapped_resolve_with_albedo(ctx, oracle):
    """Dithering behind the full pixel shader: the two implementations differ by a few ulp before quantisation, so a few pixels
    that sit on a threshold may land on the neighbouring level -- never farther than 1 / Unit, and rarely."""
    w, h = 96, 64
    rs = np.random.RandomState(5)
    lm = (rs.rand(h, w, 4) * 2).astype(np.float16)
    al = rs.randint(0, 256, size=(h, w, 4), dtype=np.uint8)
    ds = _dither(frame=2.0)
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.1, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0), Dithering=ds)
    try:
        oracle.set_dithering(ds.pack())
        ref = oracle.resolve_lighting(ib.pack_resolve(w, h, _abi.FORMAT_HALF4, hdr, _abi.FORMAT_RGBA8, _abi.FORMAT_FLOAT4), lm, al)
    finally:
        oracle.set_dithering(None)
    gpu = _rendered(ctx, w, h).Resolve(al, hdr, float4=True, lightmap=lm)
    d = np.abs(gpu[..., :3] - ref[..., :3])
    assert d.max() <= 1 / 255 + 1e-6 and (d > 1e-6).mean() < 2e-3
    # placed resolve: the pattern follows the target's pixel grid
    target = np.zeros((h + 10, w + 6, 4), np.float32)
    placed = _rendered(ctx, w, h).ResolvePlaced(target, position=(3.0, 5.0), albedo=al, hdr=hdr, lightmap=lm)
    inner = placed[5:5 + h, 3:3 + w, :3]
    inside = inner <= 1.0                                                 # values above RangeMax are kept as they are
    assert inside.mean() > 0.5 and np.abs(np.round(inner * 255) - inner * 255)[inside].max() < 1e-3      # quantised to the lattice

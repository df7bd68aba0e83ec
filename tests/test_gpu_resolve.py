"""GPU parity of the "next" row N3 -- lightmap resolve (Resolve.fx / HDR.fxh), luminance buffer and its mip chain -- against
the CPU oracle, through the C-ABI (ilb_resolve_lighting, ilb_compute_luminance)."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi, scenes
from helpers import make_renderer

pytestmark = pytest.mark.gpu

RESOLVE_RTOL = 1e-4      # same bar as the lightmap itself (north_star: 1e-4 relative per channel)
RESOLVE_FLOOR = 1 / 255  # |ref| floor of the relative error: one LSB of the Color backbuffer the values are resolved into, i.e. an
                         # absolute tolerance of 1e-4 LSB.  (Uncharted2Tonemap subtracts two nearly equal numbers near black --
                         # q - kE/kF with q ~ 0.0667 -- so values below the floor carry the absolute rounding error of q.)


def _rendered(ctx, w, h, fmt=_abi.FORMAT_HALF4):
    r = ib.LightingRenderer(ctx, ib.LightingEnvironment(), ib.RendererConfiguration((w, h)))
    return ib.RenderedLighting(r, w, h, fmt, 1.0)


def _inputs(seed, w, h, lm_dtype, al_dtype):
    rs = np.random.RandomState(seed)
    lm = (rs.rand(h, w, 4) ** 2 * 3 + 0.001).astype(np.float32)   # not within 1e-8 of 0: the tone-map curve crosses 0 there
    lm[..., 3] = np.floor(rs.rand(h, w) * 4)            # light count in alpha; 0 = untouched pixels
    lm[rs.rand(h, w) < 0.05] = 0                         # black texels (0 / 0 in GammaCompress)
    al = rs.rand(h, w, 4).astype(np.float32)
    al[..., 3] = np.where(rs.rand(h, w) < 0.1, 0.0, np.maximum(al[..., 3], 0.25))
    al[..., :3] *= al[..., 3:4]
    if lm_dtype == np.uint8:
        lm = np.floor(np.clip(lm / 3, 0, 1) * 255 + 0.5).astype(np.uint8)
    else:
        lm = lm.astype(lm_dtype)
    al = np.floor(al * 255 + 0.5).astype(np.uint8) if al_dtype == np.uint8 else al.astype(al_dtype)
    return lm, al


def _check(gpu, ref, what):
    assert np.array_equal(np.isnan(gpu), np.isnan(ref)), what
    g, r = np.nan_to_num(gpu.astype(np.float64)), np.nan_to_num(ref.astype(np.float64))
    err = np.abs(g - r) / np.maximum(np.abs(r), RESOLVE_FLOOR)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= RESOLVE_RTOL, f"{what}: max rel err {err.max():.3e} at {worst}: gpu {gpu[worst]} ref {ref[worst]}"


HDRS = {
    "none": None,
    "exposure-gamma": ib.HDRConfiguration(InverseScaleFactor=0.5, Offset=-0.02, Exposure=1.7, Gamma=2.2),
    "gamma-compress": ib.HDRConfiguration(Mode=ib.HDRMode.GammaCompress, Offset=0.01, InverseScaleFactor=0.8,
                                          GammaCompression=ib.GammaCompressionConfiguration(0.6, 0.45, 2.5)),
    "tone-map": ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.3, Gamma=0.9, Offset=0.0,
                                    ToneMapping=ib.ToneMappingConfiguration(WhitePoint=2.8)),
    "tone-map-srgb": ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=0.9, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=4.0),
                                         ResolveToSRGB=True, AlbedoIsSRGB=True),
}


@pytest.mark.parametrize("name", list(HDRS))
@pytest.mark.parametrize("with_albedo", [False, True])
@pytest.mark.parametrize("w,h,lm_dtype,al_dtype", [(64, 48, np.float16, np.uint8), (37, 23, np.float32, np.float32),
                                                     (130, 3, np.uint8, np.uint8)])
def test_resolve_matches_oracle(ctx, oracle, name, with_albedo, w, h, lm_dtype, al_dtype):
    hdr = HDRS[name]
    lm, al = _inputs(sum(map(ord, name)) + w, w, h, lm_dtype, al_dtype)
    rl = _rendered(ctx, w, h)
    albedo = al if with_albedo else None
    lm_fmt = {np.float32: _abi.FORMAT_FLOAT4, np.float16: _abi.FORMAT_HALF4, np.uint8: _abi.FORMAT_RGBA8}[lm_dtype]
    al_fmt = _abi.FORMAT_RGBA8 if al_dtype == np.uint8 else _abi.FORMAT_FLOAT4
    ref = oracle.resolve_lighting(ib.pack_resolve(w, h, lm_fmt, hdr, al_fmt, _abi.FORMAT_FLOAT4), lm, albedo)
    gpu = rl.Resolve(albedo, hdr, float4=True, lightmap=lm)
    _check(gpu, ref, f"{name} albedo={with_albedo} {w}x{h}")
    # backbuffer (SurfaceFormat.Color): round-to-nearest UNORM8 of the same values, NaN -> 0
    rgba = rl.Resolve(albedo, hdr, lightmap=lm)
    assert rgba.dtype == np.uint8 and rgba.shape == (h, w, 4)
    want = np.floor(np.clip(np.nan_to_num(ref, nan=0.0), 0, 1) * 255 + 0.5).astype(np.int32)
    diff = np.abs(rgba.astype(np.int32) - want)
    assert diff.max() <= 1 and (diff != 0).mean() < 5e-3


def test_identity_resolve_is_bit_exact_at_4k(ctx):
    """Size-independent property at BASELINE.json's frame size: with default parameters and no albedo the resolve of a Color
    lightmap returns its rgb unchanged (c / 255 -> round(x * 255) == c) with alpha 255 (Resolve.fx:41)."""
    w, h = 3840, 2160
    rs = np.random.RandomState(7)
    lm = rs.randint(0, 256, size=(h, w, 4), dtype=np.uint8)
    out = _rendered(ctx, w, h).Resolve(lightmap=lm)
    assert np.array_equal(out[..., :3], lm[..., :3]) and np.all(out[..., 3] == 255)
    # and with a white, fully lit lightmap at 0.5 (x2 == 1) the albedo passes through unchanged
    half = np.zeros((h, w, 4), np.float16)
    half[..., :3] = 0.5
    half[..., 3] = 1
    al = rs.randint(0, 256, size=(h, w, 4), dtype=np.uint8)
    assert np.array_equal(_rendered(ctx, w, h).Resolve(al, lightmap=half), al)


def test_resolve_reads_the_resident_lightmap(ctx, oracle):
    s = scenes.lighting_scene(31, 192, 128, 4, ramp=(60.0, 200.0))
    r, _ = make_renderer(ctx, s)
    lightmap = r.RenderLighting(intensityScale=0.5)
    assert lightmap.dtype == np.float16 and r.LastRendered.InverseScaleFactor == 2.0
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, InverseScaleFactor=r.LastRendered.InverseScaleFactor, Exposure=1.2,
                              ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0))
    albedo = np.random.RandomState(8).randint(0, 256, size=(128, 192, 4), dtype=np.uint8)
    resident = r.LastRendered.Resolve(albedo, hdr)
    assert np.array_equal(resident, r.LastRendered.Resolve(albedo, hdr, lightmap=lightmap))   # same texels, staged from the host
    ref = oracle.resolve_lighting(ib.pack_resolve(192, 128, _abi.FORMAT_HALF4, hdr), lightmap, albedo)
    want = np.floor(np.clip(ref, 0, 1) * 255 + 0.5).astype(np.int32)
    assert np.abs(resident.astype(np.int32) - want).max() <= 1
    # a pipelined host-to-host frame leaves its lightmap resident too
    lm2 = r.RenderLightingFrame(s.gbuffer)
    assert np.array_equal(r.LastRendered.Resolve(albedo), r.LastRendered.Resolve(albedo, lightmap=lm2))
    assert np.array_equal(r.LastRendered.ComputeLuminance(2), oracle.compute_luminance(lm2, 2))


@pytest.mark.parametrize("w,h,dtype", [(256, 128, np.float16), (100, 60, np.float32), (66, 34, np.uint8), (3840, 2160, np.float16)])
def test_luminance_chain_is_bit_exact(ctx, oracle, w, h, dtype):
    lm, _ = _inputs(w, w, h, dtype, np.uint8)
    rl = _rendered(ctx, w, h)
    for level in range(4):
        gpu = rl.ComputeLuminance(level, lightmap=lm)
        assert np.array_equal(gpu, oracle.compute_luminance(lm, level)), (w, h, level)


def test_histogram_of_a_rendered_frame(ctx, oracle):
    s = scenes.lighting_scene(32, 256, 256, 5, ramp=(60.0, 200.0))
    r, _ = make_renderer(ctx, s)
    lightmap = r.RenderLighting(intensityScale=0.25)
    h = ib.Histogram(4.0, 2.0)
    done = []
    assert r.LastRendered.TryComputeHistogram(h, done.append, accuracyFactor=3)
    ref = ib.Histogram(4.0, 2.0)
    buf = oracle.compute_luminance(lightmap, 3)
    assert buf.shape == (16, 16)
    ref.Add(buf.reshape(-1), None, 4.0)
    assert done == [h] and h.SampleCount == 256
    assert (h.Mean, h.Median, h.Min, h.Max) == (ref.Mean, ref.Median, ref.Min, ref.Max) and np.array_equal(h._count, ref._count)


def test_resolve_error_behaviour(ctx):
    rl = _rendered(ctx, 16, 16)
    lm = np.zeros((16, 16, 4), np.float16)
    with pytest.raises(ib.IlluminantError) as e:
        rl.Resolve(lightmap=lm, uvOffset=(0.5, 0.0))
    assert e.value.code == _abi.ERR_UNSUPPORTED
    with pytest.raises(ib.IlluminantError) as e:     # LUT blending needs HDRMode.None (LightingRenderer.cs:1593-1594)
        lut = ib.LUTBlendingConfiguration(ib.ColorLUT.Identity(4), ib.ColorLUT.Identity(4))
        rl.Resolve(np.zeros((16, 16, 4), np.uint8), lightmap=lm, lutBlending=lut,
                   hdr=ib.HDRConfiguration(Mode=ib.HDRMode.GammaCompress, GammaCompression=ib.GammaCompressionConfiguration(0.5, 0.5, 2.0)))
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT and "LUT blending" in str(e.value)
    with pytest.raises(ib.IlluminantError) as e:
        rl.Resolve(np.zeros((8, 8, 4), np.uint8), lightmap=lm)
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT
    with pytest.raises(ib.IlluminantError) as e:
        rl.ComputeLuminance(9, lightmap=lm)
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT
    c2 = ib.Context(0)
    try:
        fresh = ib.RenderedLighting(ib.LightingRenderer(c2, ib.LightingEnvironment(), ib.RendererConfiguration((16, 16))), 16, 16, _abi.FORMAT_HALF4, 1.0)
        with pytest.raises(ib.IlluminantError) as e:
            fresh.Resolve()     # nothing rendered yet on this context
        assert e.value.code == _abi.ERR_INVALID_OPERATION
    finally:
        c2.close()


@pytest.mark.parametrize("name", ["none", "tone-map-srgb"])
@pytest.mark.parametrize("with_albedo", [False, True])
def test_placed_resolve_matches_oracle(ctx, oracle, name, with_albedo):
    """Scaled / offset resolve (ResolveLighting as a quad, LightingRenderer.cs:1537-1645) against the oracle: magnified and
    minified, off the target's edges, an albedo sub-region, a LightmapUVOffset; pixels outside the quad are untouched."""
    from illuminant_b200 import hdr as H
    w, h, tw, th = 96, 54, 200, 120
    lm, al = _inputs(31, w, h, np.float16, np.uint8)
    al = al[:40, :64].copy() if with_albedo else None
    rl = _rendered(ctx, w, h)
    target = np.random.RandomState(2).rand(th, tw, 4).astype(np.float32)
    for position, scale, region, uv in (((10.5, 7.25), (1.9, 2.1), (0.0, 0.0, 1.0, 1.0), (0.0, 0.0)),
                                        ((-9.0, -4.5), (0.6, 0.45), (0.25, 0.125, 0.875, 1.0), (0.01, -0.02)),
                                        ((150.0, 90.0), (3.0, 3.0), (0.0, 0.0, 0.5, 0.5), (0.0, 0.0))):
        gpu = rl.ResolvePlaced(target, position, scale, al, region, HDRS[name], lightmap=lm, uvOffset=uv)
        params = H.pack_resolve(w, h, _abi.FORMAT_FLOAT4, HDRS[name], _abi.FORMAT_FLOAT4, _abi.FORMAT_FLOAT4, uvOffset=uv)
        pl = _abi.ResolvePlacement()
        pl.target_width, pl.target_height = tw, th
        pl.Position[:], pl.Scale[:], pl.AlbedoRegion[:] = position, scale, region
        if al is not None:
            pl.albedo_width, pl.albedo_height = al.shape[1], al.shape[0]
        ref = oracle.resolve_lighting_placed(params, pl, lm.astype(np.float32), al.astype(np.float32) / 255 if al is not None else None, target)
        _check(gpu, ref, f"placed {name} albedo={with_albedo} at {position}")
        untouched = ref == target
        assert np.array_equal(gpu[untouched], target[untouched]) and 0.002 < 1.0 - untouched.all(axis=2).mean() < 0.9

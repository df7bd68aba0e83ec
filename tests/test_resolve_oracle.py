"""N3 (lightmap resolve / luminance / histogram) on the CPU: known answers for the oracle restatement of Resolve.fx + HDR.fxh,
the host mirror of LightingResolveHandler / SetToneMappingParameters packing, and Histogram.cs against an independent scalar
transcription.  No GPU."""
import math

import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import _abi, hdr as H

f32 = np.float32


def _params(w, h, hdr=None, lm=_abi.FORMAT_FLOAT4, al=_abi.FORMAT_FLOAT4):
    return H.pack_resolve(w, h, lm, hdr, al, _abi.FORMAT_FLOAT4)


def test_resolve_defaults_are_identity_with_alpha_one(oracle):
    rs = np.random.RandomState(1)
    lm = rs.rand(5, 7, 4).astype(np.float32) * 2
    out = oracle.resolve_lighting(_params(7, 5), lm)
    assert np.array_equal(out[..., :3], lm[..., :3])        # Offset 0, Exposure 1, pow(x, 1)
    assert np.all(out[..., 3] == 1.0)                        # Resolve.fx:41
    out = oracle.resolve_lighting(_params(7, 5, ib.HDRConfiguration(InverseScaleFactor=0.25, Offset=-0.1, Exposure=2.0)), lm)
    want = np.maximum(lm[..., :3] * f32(0.25) + f32(-0.1), 0) * f32(2.0)
    assert np.allclose(out[..., :3], want, rtol=1e-6, atol=0)


def test_resolve_with_albedo_known_answers(oracle):
    rs = np.random.RandomState(2)
    al = rs.rand(4, 4, 4).astype(np.float32)
    lm = np.zeros((4, 4, 4), np.float32)
    lm[..., :3] = 0.5
    lm[..., 3] = 1.0     # one light touched the pixel (additive alpha, LightingRenderer.cs:206)
    out = oracle.resolve_lighting(_params(4, 4), lm, al)
    assert np.allclose(out, al, rtol=1e-6)                    # light * 2 == 1 -> albedo unchanged; alpha = albedo.a (Resolve.fx:63-66)
    lm[..., 3] = 0.0     # untouched pixel: lerp weight saturate(light.a) = 0 -> plain albedo even with bright light
    lm[..., :3] = 3.0
    assert np.array_equal(oracle.resolve_lighting(_params(4, 4), lm, al), al)
    lm[..., 3] = 5.0
    lm[..., :3] = 1.0    # x2
    out = oracle.resolve_lighting(_params(4, 4), lm, al)
    assert np.allclose(out[..., :3], al[..., :3] * 2, rtol=1e-6)


def test_tone_map_and_gamma_compress_closed_forms(oracle):
    wp = 3.5
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.0, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=wp))
    lm = np.zeros((1, 3, 4), np.float32)
    lm[0, 0, :3] = wp      # a value at the white point maps to exactly 1 (Resolve.fx:131)
    lm[0, 1, :3] = 0.0     # black stays (about) black
    lm[0, 2, :3] = [0.1, 1.0, 10.0]
    out = oracle.resolve_lighting(_params(3, 1, hdr), lm)
    assert np.allclose(out[0, 0, :3], 1.0, atol=1e-6)
    assert np.allclose(out[0, 1, :3], 0.0, atol=1e-6)

    def u2(x):
        A, B, C_, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
        return ((x * (A * x + C_ * B) + D * E) / (x * (A * x + B) + D * F)) - E / F
    assert np.allclose(out[0, 2, :3], [u2(v) / u2(wp) for v in (0.1, 1.0, 10.0)], rtol=1e-5)
    assert np.all(np.diff(out[0, 2, :3]) > 0)

    g = ib.GammaCompressionConfiguration(MiddleGray=0.6, AverageLuminance=0.4, MaximumLuminance=2.0)
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.GammaCompress, GammaCompression=g, Offset=0.05)
    rgb = np.array([0.3, 0.8, 0.2])
    lm[0, 0, :3] = rgb
    out = oracle.resolve_lighting(_params(3, 1, hdr), lm)
    c = rgb + 0.05
    L = c @ np.array([0.299, 0.587, 0.114])
    s = L * 0.6 / 0.4
    comp = s * (1 + s / 4.0) / (1 + s)
    assert np.allclose(out[0, 0, :3], c * comp / L, rtol=1e-5)
    assert np.isnan(oracle.resolve_lighting(_params(3, 1, ib.HDRConfiguration(Mode=ib.HDRMode.GammaCompress, GammaCompression=g)), lm)[0, 1, :3]).all()  # 0 / 0 like the shader


def test_srgb_flags_round_trip(oracle):
    rs = np.random.RandomState(3)
    al = rs.rand(6, 6, 4).astype(np.float32)
    al[..., 3] = np.maximum(al[..., 3], 0.2)
    al[..., :3] *= al[..., 3:4]      # premultiplied
    lm = np.full((6, 6, 4), 0.5, np.float32)
    lm[..., 3] = 1
    both = oracle.resolve_lighting(_params(6, 6, ib.HDRConfiguration(AlbedoIsSRGB=True, ResolveToSRGB=True)), lm, al)
    assert np.allclose(both, al, rtol=2e-5, atol=1e-6)       # decode, x1, encode
    lin = oracle.resolve_lighting(_params(6, 6, ib.HDRConfiguration(AlbedoIsSRGB=True)), lm, al)
    s = (al[..., :3] / al[..., 3:4]).astype(np.float64)
    want = np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4) * al[..., 3:4]
    assert np.allclose(lin[..., :3], want, rtol=1e-5, atol=1e-7)


def test_pack_resolve_clamps_like_the_reference():
    hdr = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=0.0, Gamma=9.0, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=1e9))
    p = H.pack_resolve(8, 8, _abi.FORMAT_HALF4, hdr)
    assert p.ExposureMinusOne == f32(f32(1 / 256.0) - f32(1)) and p.GammaMinusOne == f32(3.0) and p.WhitePoint == f32(99999.0)
    p = H.pack_resolve(8, 8, _abi.FORMAT_HALF4, ib.HDRConfiguration(Mode=ib.HDRMode.None_, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=7)))
    assert p.WhitePoint == 1.0 and p.InverseScaleFactor == 1.0            # LightingRenderer.cs:1516-1521, :1469-1473
    g = ib.GammaCompressionConfiguration(MiddleGray=-1, AverageLuminance=0, MaximumLuminance=3)
    p = H.pack_resolve(8, 8, _abi.FORMAT_HALF4, ib.HDRConfiguration(Mode=ib.HDRMode.GammaCompress, GammaCompression=g))
    assert p.MiddleGray == 0 and p.AverageLuminance == f32(1 / 256.0) and p.MaximumLuminanceSquared == 9.0
    assert p.ExposureMinusOne == 0 and p.GammaMinusOne == 0
    p = H.pack_resolve(8, 8, _abi.FORMAT_HALF4, None)
    assert (p.hdr_mode, p.InverseScaleFactor, p.Offset, p.DitheringStrength) == (0, 1.0, 0.0, 0.0)


def test_luminance_known_answers(oracle):
    lm = np.zeros((16, 24, 4), np.float32)
    lm[..., :3] = [0.5, 0.25, 2.0]
    want = f32(f32(f32(0.5) * f32(0.299)) + f32(f32(0.25) * f32(0.587))) + f32(f32(2.0) * f32(0.144))   # 0.144: Resolve.fx:15
    for level, shape in ((0, (8, 12)), (1, (4, 6)), (2, (2, 3))):
        out = oracle.compute_luminance(lm, level)
        assert out.shape == shape and np.all(out == want)
    # level 0 point-samples texel (2x+1, 2y+1); each mip level is a 2x2 box filter
    rs = np.random.RandomState(4)
    lm = rs.rand(8, 8, 4).astype(np.float32)
    l0 = oracle.compute_luminance(lm, 0)
    t = lm[1::2, 1::2]
    assert np.array_equal(l0, (t[..., 0] * f32(0.299) + t[..., 1] * f32(0.587)) + t[..., 2] * f32(0.144))
    l1 = oracle.compute_luminance(lm, 1)
    assert np.array_equal(l1, ((l0[0::2, 0::2] + l0[0::2, 1::2]) + (l0[1::2, 0::2] + l0[1::2, 1::2])) * f32(0.25))


class _ScalarHistogram:
    """Histogram.cs:61-229 transcribed statement by statement (scalar loops), independent of the vectorised product class."""

    def __init__(self, maxValue, power, bucketCount=64, ignoreZeroes=False):
        self.n = bucketCount
        self.ignore = ignoreZeroes
        lg = math.log(float(f32(1) + f32(maxValue))) / math.log(power)
        self.maxv = [f32(f32(math.pow(power, (lg / bucketCount) * (i + 1))) - f32(1)) for i in range(bucketCount)]
        self.count = [0] * bucketCount
        self.sum = [f32(0)] * bucketCount
        self.min = [np.finfo(np.float32).max] * bucketCount
        self.max = [f32(0)] * bucketCount
        self.SampleCount, self.Sum = 0, f32(0)

    def pick(self, v):
        if v < self.maxv[0]:
            return 0
        if v >= self.maxv[self.n - 2]:
            return self.n - 1
        i, mx = 0, self.n - 1
        while i <= mx:
            pivot = i + ((mx - i) >> 1)
            if self.maxv[pivot] <= v:
                i = pivot + 1
            else:
                mx = pivot - 1
        return i

    def add(self, buf, scale):
        buf = sorted(f32(v) for v in buf)
        count = len(buf)
        off = 0
        if self.ignore:
            off = max([i for i, v in enumerate(buf) if v == 0], default=-1)
        mi = min(max(int((count - off) / 2) + off, 0), count - 1)
        self.Median = f32(buf[mi] * f32(scale))
        added = 0
        for raw in buf:
            if self.ignore and raw <= 0:
                continue
            v = f32(raw * f32(scale))
            self.Sum = f32(self.Sum + v)
            added += 1
            j = self.pick(v)
            self.count[j] += 1
            self.sum[j] = f32(self.sum[j] + v)
            self.min[j] = min(self.min[j], v)
            self.max[j] = max(self.max[j], v)
        self.SampleCount += added
        self.Mean = f32(self.Sum / f32(self.SampleCount)) if self.SampleCount else f32(0)


@pytest.mark.parametrize("ignore", [False, True])
def test_histogram_matches_scalar_transcription(ignore):
    rs = np.random.RandomState(5)
    x = (rs.rand(3000) ** 3 * 5).astype(np.float32)
    x[rs.rand(3000) < 0.1] = 0
    h = ib.Histogram(4.0, 2.0, ignoreZeroes=ignore)
    s = _ScalarHistogram(4.0, 2.0, ignoreZeroes=ignore)
    assert np.array_equal(h.BucketMaxValues, np.array(s.maxv, np.float32)) and h.BucketMaxValues[-1] == 4.0
    for part, scale in ((x[:2000], 0.5), (x[2000:], 1.25)):   # two Adds accumulate (Histogram.cs:186-208)
        h.Add(part, None, scale)
        s.add(part, scale)
        assert h.SampleCount == s.SampleCount and h.Sum == s.Sum and h.Mean == s.Mean and h.Median == s.Median
        assert np.array_equal(h._count, s.count) and np.array_equal(h._sum, np.array(s.sum, np.float32))
        assert np.array_equal(h._min, np.array(s.min, np.float32)) and np.array_equal(h._max, np.array(s.max, np.float32))
    assert [int(h.PickBucketForValue(v)) for v in (0.0, 0.01, 1.0, 3.99, 4.0, 100.0)] == [s.pick(f32(v)) for v in (0.0, 0.01, 1.0, 3.99, 4.0, 100.0)]
    found, bucket, value = h.GetPercentile(50)
    assert found and h.BucketMaxValues[bucket - 1] <= value <= h.BucketMaxValues[bucket]
    assert h.GetPercentile(101)[0] is False
    assert sum(b["Count"] for b in h.Buckets) == h.SampleCount
    h.Clear()
    assert h.SampleCount == 0 and h.GetPercentile(50) == (False, 0, 0)


def test_unorm8_decode_by_one_fma_is_within_an_ulp_and_round_trips():
    """resolve.cu decodes UNORM8 texels as one fused multiply-add, RN((8388608 + c) * r - 8388608 * r) = RN(c * r) with
    r = fl(1 / 255): for all 256 inputs the result is within one ulp of the correctly rounded c / 255 (what the oracle decodes),
    and packing it again (floor(x * 255 + 0.5), the kernel's truncating add) returns c -- the exact pass-through properties of the
    GPU tests rest on that."""
    r = np.float32(1.0) / np.float32(255.0)
    worst = 0.0
    for c in range(256):
        x = np.float32(np.float64(c) * np.float64(r))        # one rounding of the exact product (24 x 8 bits fit a double)
        ref = np.float32(c) / np.float32(255.0)
        if c:
            worst = max(worst, abs(float(x) - float(ref)) / float(np.spacing(ref)))
        assert int(np.floor(np.float32(np.float32(x * np.float32(255.0)) + np.float32(0.5)))) == c
        assert np.float64(8388608.0) * np.float64(r) == np.float64(np.float32(np.float32(8388608.0) * r))   # the addend is exact
    assert worst <= 1.0


# ---- scaled / offset resolve (ResolveLighting drawn as a quad, LightingRenderer.cs:1537-1645) -------------------------------
def _placement(tw, th, position=(0.0, 0.0), scale=(1.0, 1.0), region=(0.0, 0.0, 1.0, 1.0), albedo=None):
    pl = _abi.ResolvePlacement()
    pl.target_width, pl.target_height = tw, th
    pl.Position[:] = position
    pl.Scale[:] = scale
    pl.AlbedoRegion[:] = region
    if albedo is not None:
        pl.albedo_width, pl.albedo_height = albedo.shape[1], albedo.shape[0]
    return pl


def test_placed_resolve_known_answers(oracle):
    rs = np.random.RandomState(4)
    lm = (rs.rand(6, 8, 4) * 2).astype(np.float32)
    # 1:1 placement: the plain resolve up to the rounding of the texture coordinates
    same = oracle.resolve_lighting_placed(_params(8, 6), _placement(8, 6), lm, None, np.full((6, 8, 4), 9.0, np.float32))
    assert np.allclose(same, oracle.resolve_lighting(_params(8, 6), lm), rtol=0, atol=2e-6)
    # a constant lightmap stays constant under any scale; pixels outside the quad keep the target's contents
    const = np.full((6, 8, 4), 0.25, np.float32)
    out = oracle.resolve_lighting_placed(_params(8, 6), _placement(40, 30, position=(4.0, 3.0), scale=(2.5, 3.0)), const, None,
                                         np.full((30, 40, 4), -1.0, np.float32))
    inside = np.zeros((30, 40), bool)
    inside[3:21, 4:24] = True                                    # quad: 8 * 2.5 = 20 wide, 6 * 3 = 18 tall at (4, 3)
    assert np.allclose(out[inside][:, :3], 0.25, atol=1e-7) and np.all(out[inside][:, 3] == 1.0) and np.all(out[~inside] == -1.0)
    # a horizontal ramp magnified 4x: LINEAR sampling interpolates between texel centres and clamps at the edges
    ramp = np.zeros((1, 4, 4), np.float32)
    ramp[0, :, 0] = [0.0, 1.0, 2.0, 3.0]
    out = oracle.resolve_lighting_placed(_params(4, 1), _placement(16, 1, scale=(4.0, 1.0)), ramp, None, np.zeros((1, 16, 4), np.float32))
    centres = (np.arange(16) + 0.5) / 4.0 - 0.5                  # texel-space coordinate of every output pixel centre
    assert np.allclose(out[0, :, 0], np.clip(centres, 0.0, 3.0), atol=1e-6)
    # LightmapUVOffset shifts the lightmap fetch by whole texels when it is k / width
    p = H.pack_resolve(4, 1, _abi.FORMAT_FLOAT4, None, _abi.FORMAT_FLOAT4, _abi.FORMAT_FLOAT4, uvOffset=(0.25, 0.0))
    out = oracle.resolve_lighting_placed(p, _placement(4, 1), ramp, None, np.zeros((1, 4, 4), np.float32))
    assert np.allclose(out[0, :, 0], [1.0, 2.0, 3.0, 3.0], atol=1e-6)
    # with albedo the quad is the albedo REGION in texels times the scale, and both textures stretch over it
    al = np.zeros((4, 4, 4), np.float32)
    al[..., 3] = 1.0
    al[:, 2:, :3] = 0.5                                          # right half of the sheet is grey
    light = np.full((2, 2, 4), 0.5, np.float32)                  # x 2 = 1: albedo passes through
    light[..., 3] = 1.0
    pl = _placement(12, 8, position=(2.0, 1.0), scale=(3.0, 1.5), region=(0.5, 0.0, 1.0, 1.0), albedo=al)
    out = oracle.resolve_lighting_placed(_params(2, 2), pl, light, al, np.zeros((8, 12, 4), np.float32))
    assert np.allclose(out[1:7, 3:8, :3], 0.5, atol=1e-6) and not out[:, 8:].any() and not out[7:].any()   # 2 texels * 3 = 6 wide, 4 * 1.5 = 6 tall
    assert np.allclose(out[1:7, 2, :3], 1.0 / 3.0, atol=1e-6)   # the bilinear footprint of the first column reaches the black texel left of the region

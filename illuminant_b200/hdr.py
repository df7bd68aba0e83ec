"""Host-side mirror of the reference's resolve / HDR surface ("next" row N3 of SURVEY.md section 8f):
`HDRConfiguration`, `RenderedLighting.Resolve`, `RenderedLighting.TryComputeHistogram` (LightingRenderer.HDR.cs) and
`Histogram` (Histogram.cs).  The per-pixel work (Resolve.fx, the luminance buffer and its mip chain) runs in
libilluminant_b200.so through `ilb_resolve_lighting` / `ilb_compute_luminance`; the histogram is host code in the reference
(HistogramUpdateTask sorts the read-back luminance on a worker thread, LightingRenderer.HDR.cs:21-57) and is host code here.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _abi
from ._abi import FORMAT_FLOAT4, FORMAT_HALF4, FORMAT_RGBA8, HDR_GAMMA_COMPRESS, HDR_NONE, HDR_TONE_MAP, Resolve

f32 = np.float32


class HDRMode:  # LightingRenderer.HDR.cs:269-273
    None_, GammaCompress, ToneMap = HDR_NONE, HDR_GAMMA_COMPRESS, HDR_TONE_MAP


@dataclass
class GammaCompressionConfiguration:  # LightingRenderer.HDR.cs:216-218
    MiddleGray: float = 0.0
    AverageLuminance: float = 0.0
    MaximumLuminance: float = 0.0


@dataclass
class ToneMappingConfiguration:  # :220-222
    WhitePoint: float = 0.0


@dataclass
class DitheringSettings:  # Squared.Render (un-vendored sq/Fracture); Unit / Strength / FrameIndex are what the resolve handler sets (:1489-1497)
    Unit: float = 255.0
    Strength: float = 0.0
    FrameIndex: float = 0.0   # the handler overwrites it with DeviceManager.FrameIndex
    BandSize: float = 1.0
    RangeMin: float = 0.0
    RangeMax: float = 1.0

    def pack(self) -> "_abi.Dithering":
        d = _abi.Dithering()
        d.Strength, d.Unit, d.FrameIndex = float(self.Strength), float(self.Unit), float(self.FrameIndex)
        d.BandSize, d.RangeMin, d.RangeMax = float(self.BandSize), float(self.RangeMin), float(self.RangeMax)
        return d


@dataclass
class ColorLUT:  # Squared.Render.ColorLUT (un-vendored): a SurfaceFormat.Color strip texture, (Resolution^2) x (Resolution * RowCount) texels
    Texture: np.ndarray          # uint8 [Resolution * RowCount, Resolution * Resolution, 4]
    Resolution: int
    RowCount: int = 1

    @staticmethod
    def Identity(resolution: int = 16) -> "ColorLUT":
        """The neutral table: texel (b * res + r, g) = (r, g, b) / (res - 1)."""
        res = int(resolution)
        t = np.zeros((res, res * res, 4), np.uint8)
        ramp = np.round(np.arange(res) * 255.0 / (res - 1)).astype(np.uint8)
        for b in range(res):
            t[:, b * res:(b + 1) * res, 0] = ramp[None, :]
            t[:, b * res:(b + 1) * res, 1] = ramp[:, None]
            t[:, b * res:(b + 1) * res, 2] = ramp[b]
        t[..., 3] = 255
        return ColorLUT(t, res, 1)


@dataclass
class LUTBlendingConfiguration:  # LightingRenderer.HDR.cs:260-273
    DarkLUT: ColorLUT = None
    BrightLUT: ColorLUT = None
    PerChannel: bool = False
    LUTOnly: bool = False
    DarkLevel: float = 0.0
    NeutralBandSize: float = 0.0
    BrightLevel: float = 1.0     # stored as _BrightLevelMinus1 so that default(LUTBlendingConfiguration) means 1

    def pack(self) -> "_abi.LutBlending":  # IlluminantMaterials.SetLUTBlending, IlluminantMaterials.cs:139-149
        b = _abi.LutBlending()
        b.dark_resolution, b.bright_resolution = int(self.DarkLUT.Resolution), int(self.BrightLUT.Resolution)
        b.dark_row_count, b.bright_row_count = int(self.DarkLUT.RowCount), int(self.BrightLUT.RowCount)
        b.DarkLevel, b.NeutralBandSize, b.BrightLevel = float(self.DarkLevel), float(self.NeutralBandSize), float(self.BrightLevel)
        b.PerChannel, b.LUTOnly = (1.0 if self.PerChannel else 0.0), (1.0 if self.LUTOnly else 0.0)
        return b


@dataclass
class HDRConfiguration:  # LightingRenderer.HDR.cs:215-267
    Mode: int = HDRMode.None_
    InverseScaleFactor: float = 0.0
    Offset: float = 0.0
    GammaCompression: GammaCompressionConfiguration = field(default_factory=GammaCompressionConfiguration)
    ToneMapping: ToneMappingConfiguration = field(default_factory=ToneMappingConfiguration)
    Dithering: Optional[DitheringSettings] = None
    ResolveToSRGB: bool = False
    AlbedoIsSRGB: bool = False
    Exposure: float = 1.0   # stored as ExposureMinusOne so that default(HDRConfiguration) means 1 (:228, :248-255)
    Gamma: float = 1.0      # GammaMinusOne (:257-265)


def _clamp(v: float, lo: float, hi: float) -> np.float32:  # MathHelper.Clamp on floats
    return f32(min(max(f32(v), f32(lo)), f32(hi)))


def pack_resolve(width: int, height: int, lightmap_format: int, hdr: Optional[HDRConfiguration], albedo_format: int = FORMAT_RGBA8,
                 output_format: int = FORMAT_RGBA8, uvOffset: Tuple[float, float] = (0.0, 0.0)) -> Resolve:
    """What LightingResolveHandler._Before sets on the resolve material (LightingRenderer.cs:1464-1523) with the clamps of
    SetGammaCompressionParameters / SetToneMappingParameters (IlluminantMaterials.cs:81-137).  Without `hdr` the effect
    keeps its defaults: Offset 0, Exposure 1, Gamma 1."""
    r = Resolve()
    r.width, r.height = int(width), int(height)
    r.lightmap_format, r.albedo_format, r.output_format = lightmap_format, albedo_format, output_format
    r.hdr_mode = hdr.Mode if hdr is not None else HDRMode.None_
    r.InverseScaleFactor = (hdr.InverseScaleFactor if hdr.InverseScaleFactor != 0 else 1.0) if hdr is not None else 1.0
    r.AlbedoIsSRGB = 1.0 if (hdr is not None and hdr.AlbedoIsSRGB) else 0.0
    r.ResolveToSRGB = 1.0 if (hdr is not None and hdr.ResolveToSRGB) else 0.0
    r.LightmapUVOffset[0], r.LightmapUVOffset[1] = float(uvOffset[0]), float(uvOffset[1])
    r.DitheringStrength = hdr.Dithering.Strength if (hdr is not None and hdr.Dithering is not None) else 0.0
    r.WhitePoint = 1.0
    r.MiddleGray, r.AverageLuminance, r.MaximumLuminanceSquared = 0.0, 1.0, 1.0
    if hdr is not None:
        lo, hi = 1.0 / 256.0, 99999.0
        r.Offset = hdr.Offset
        if hdr.Mode == HDRMode.GammaCompress:
            g = hdr.GammaCompression
            r.MiddleGray = _clamp(g.MiddleGray, 0.0, hi)
            r.AverageLuminance = _clamp(g.AverageLuminance, lo, hi)
            m = _clamp(g.MaximumLuminance, lo, hi)
            r.MaximumLuminanceSquared = f32(m * m)
        else:
            r.ExposureMinusOne = f32(_clamp(hdr.Exposure, lo, hi) - f32(1))
            r.GammaMinusOne = f32(_clamp(hdr.Gamma, 0.1, 4.0) - f32(1))
            r.WhitePoint = _clamp(hdr.ToneMapping.WhitePoint, lo, hi) if hdr.Mode == HDRMode.ToneMap else 1.0
    return r


_NP = {FORMAT_FLOAT4: np.float32, FORMAT_HALF4: np.float16, FORMAT_RGBA8: np.uint8}


def _texels(a: np.ndarray, what: str) -> Tuple[np.ndarray, int]:
    a = np.asarray(a)
    for fmt, dt in _NP.items():
        if a.dtype == dt:
            return np.ascontiguousarray(a), fmt
    raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, f"{what} must be float32, float16 or uint8 texels, got {a.dtype}")


class Histogram:  # Histogram.cs:17-258
    """Luminance histogram with power-spaced buckets; `Add` mirrors Histogram.Add (sort, median, sequential fp32 sums in
    sorted order)."""

    def __init__(self, maxValue: float, power: float, bucketCount: int = 64, ignoreZeroes: bool = False):
        self.BucketCount = int(bucketCount)
        self.MaxInputValue = float(maxValue)
        self.IgnoreZeroes = bool(ignoreZeroes)
        maxValuePlusOneLog = math.log(float(f32(1) + f32(maxValue))) / math.log(power)   # Math.Log(1 + maxValue, power) :67
        self.BucketMaxValues = np.empty(self.BucketCount, dtype=np.float32)
        for i in range(self.BucketCount):
            valueLog = (maxValuePlusOneLog / self.BucketCount) * (i + 1)
            self.BucketMaxValues[i] = f32(f32(math.pow(power, valueLog)) - f32(1))             # :69-73
        self.FirstBucketMaxValue = self.BucketMaxValues[0]
        self.LastBucketMinValue = self.BucketMaxValues[self.BucketCount - 2]
        self.Clear()

    def Clear(self) -> None:  # :98-113
        self.SampleCount = 0
        self.Min = f32(0)
        self.Max = f32(0)
        self.Sum = f32(0)
        self.Mean = f32(0)
        self.Median = f32(0)
        n = self.BucketCount
        self._count = np.zeros(n, dtype=np.int64)
        self._sum = np.zeros(n, dtype=np.float32)
        self._min = np.full(n, np.finfo(np.float32).max, dtype=np.float32)
        self._max = np.zeros(n, dtype=np.float32)

    def PickBucketForValue(self, value) -> np.ndarray:  # :115-135 (binary search for the first bucket whose max exceeds value)
        v = np.asarray(value, dtype=np.float32)
        return np.minimum(np.searchsorted(self.BucketMaxValues, v, side="right"), self.BucketCount - 1)

    def Add(self, buffer: np.ndarray, count: Optional[int] = None, scaleFactor: float = 1.0) -> None:  # :168-229
        buf = np.asarray(buffer, dtype=np.float32).reshape(-1)
        count = buf.size if count is None else int(count)
        if count > buf.size:
            raise ValueError("count")
        buf = np.sort(buf[:count], kind="stable")
        if count == 0:
            return
        medianOffset = 0
        if self.IgnoreZeroes:
            z = np.nonzero(buf == 0)[0]
            medianOffset = int(z[-1]) if z.size else -1   # Array.LastIndexOf(buffer, 0)
        medianIndex = min(max(((count - medianOffset) // 2) + medianOffset, 0), count - 1)
        scale = f32(scaleFactor)
        self.Median = f32(buf[medianIndex] * scale)
        vals = buf[buf > 0] if self.IgnoreZeroes else buf
        vals = (vals * scale).astype(np.float32)
        if vals.size:
            self.Sum = np.cumsum(np.concatenate([[self.Sum], vals]), dtype=np.float32)[-1]
            buckets = self.PickBucketForValue(vals)
            # the values are sorted and bucket indices are monotonic in the value, so each bucket is one contiguous run
            starts = np.searchsorted(buckets, np.arange(self.BucketCount), side="left")
            ends = np.searchsorted(buckets, np.arange(self.BucketCount), side="right")
            for j in range(self.BucketCount):
                run = vals[starts[j]:ends[j]]
                if run.size == 0:
                    continue
                self._count[j] += run.size
                self._sum[j] = np.cumsum(np.concatenate([[self._sum[j]], run]), dtype=np.float32)[-1]
                self._min[j] = min(self._min[j], run.min())
                self._max[j] = max(self._max[j], run.max())
        self.SampleCount += int(vals.size)
        self.Mean = f32(self.Sum / f32(self.SampleCount)) if self.SampleCount > 0 else f32(0)
        mn = min(np.finfo(np.float32).max, self._min.min())
        self.Min = f32(mn) if self.SampleCount > 0 else f32(0)
        self.Max = f32(max(0.0, self._max.max()))

    def GetPercentile(self, percent: float):  # :137-166 -> (found, bucketIndex, value)
        if self.SampleCount < 1 or percent < 0 or percent > 100:
            return False, 0, f32(0)
        sampleIndex = int(f32(f32(self.SampleCount) * f32(percent)) / f32(100))
        first = 0
        for i in range(self.BucketCount):
            c = int(self._count[i])
            local = sampleIndex - first
            if 0 <= local < c:
                lo = self.BucketMaxValues[i - 1] if i > 0 else f32(0)
                hi = self.BucketMaxValues[i]
                return True, i, f32(lo + (hi - lo) * f32(f32(local) / f32(c)))   # Arithmetic.Lerp
            first += c
        raise RuntimeError("percentile outside every bucket")   # the reference throws here too (:165)

    @property
    def Buckets(self) -> List[dict]:  # :231-257
        out = []
        for i in range(self.BucketCount):
            c = int(self._count[i])
            out.append(dict(BucketStart=self.BucketMaxValues[i - 1] if i > 0 else f32(0), BucketEnd=self.BucketMaxValues[i], Count=c,
                            Min=self._min[i] if c > 0 else f32(0), Max=self._max[i], Mean=f32(self._sum[i] / f32(c)) if c > 0 else f32(0)))
        return out


class RenderedLighting:  # LightingRenderer.HDR.cs:69-212
    """What RenderLighting hands back in the reference: the lightmap that stays on the GPU plus InverseScaleFactor.  Created by
    `LightingRenderer.RenderLighting` / `RenderLightingFrame` for whole frames (`renderer.LastRendered`)."""

    def __init__(self, renderer, width: int, height: int, lightmap_format: int, inverseScaleFactor: float):
        self.Renderer = renderer
        self.Width, self.Height = int(width), int(height)
        self.LightmapFormat = lightmap_format
        self.InverseScaleFactor = float(inverseScaleFactor)

    @property
    def IsValid(self) -> bool:
        return self.Renderer is not None and self.Renderer.ctx is not None

    def _bind_dithering(self, ctx, hdr: Optional[HDRConfiguration]) -> None:
        """LightingResolveHandler._Before (:1489-1497): the configuration's DitheringSettings or {Unit 255, Strength 0}."""
        ds = hdr.Dithering if (hdr is not None and hdr.Dithering is not None) else None
        ctx.check(ctx.lib.ilb_set_dithering(ctx.handle, C.byref(ds.pack()) if ds is not None else None))

    def Resolve(self, albedo: Optional[np.ndarray] = None, hdr: Optional[HDRConfiguration] = None, float4: bool = False,
                lightmap: Optional[np.ndarray] = None, uvOffset: Tuple[float, float] = (0.0, 0.0),
                lutBlending: Optional[LUTBlendingConfiguration] = None) -> np.ndarray:
        """RenderedLighting.Resolve at 1:1 (position 0, scale 1): uint8 [H, W, 4] backbuffer texels (float32 with float4=True).
        `lightmap` = None resolves the lightmap resident on the device from the last frame; an array resolves those texels.
        `lutBlending` selects the LUT-blended material like ResolveLighting (LightingRenderer.cs:1558-1561): it needs an albedo and
        HDRMode.None."""
        if not self.IsValid:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "Invalid")
        ctx = self.Renderer.ctx
        self._bind_dithering(ctx, hdr)
        lm_ptr, lm_fmt, w, h = None, self.LightmapFormat, self.Width, self.Height
        if lightmap is not None:
            lightmap, lm_fmt = _texels(lightmap, "lightmap")
            h, w = lightmap.shape[0], lightmap.shape[1]
            lm_ptr = lightmap.ctypes.data_as(C.c_void_p)
        al_ptr, al_fmt = None, FORMAT_RGBA8
        if albedo is not None:
            albedo, al_fmt = _texels(albedo, "albedo")
            if al_fmt == FORMAT_HALF4 or albedo.shape[0] != h or albedo.shape[1] != w:
                raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "albedo must be uint8 or float32 texels of the lightmap's size")
            al_ptr = albedo.ctypes.data_as(C.c_void_p)
        out_fmt = FORMAT_FLOAT4 if float4 else FORMAT_RGBA8
        params = pack_resolve(w, h, lm_fmt, hdr, al_fmt, out_fmt, uvOffset)
        out = np.empty((h, w, 4), dtype=_NP[out_fmt])
        if lutBlending is not None and albedo is not None and params.hdr_mode == HDRMode.None_:
            lut = lutBlending.pack()
            dark, bright = (np.ascontiguousarray(t.Texture, dtype=np.uint8) for t in (lutBlending.DarkLUT, lutBlending.BrightLUT))
            for t, l in ((dark, lutBlending.DarkLUT), (bright, lutBlending.BrightLUT)):
                if t.shape != (l.Resolution * l.RowCount, l.Resolution * l.Resolution, 4):
                    raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "a ColorLUT texture is [Resolution * RowCount, Resolution^2, 4] uint8 texels")
            ctx.check(ctx.lib.ilb_resolve_lighting_lut(ctx.handle, C.byref(params), C.byref(lut), dark.ctypes.data_as(C.c_void_p),
                                                       bright.ctypes.data_as(C.c_void_p), lm_ptr, al_ptr, out.ctypes.data_as(C.c_void_p)))
            return out
        if lutBlending is not None and albedo is not None:  # LightingRenderer.cs:1593-1594
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "LUT blending is not compatible with this type of lighting resolve.")
        ctx.check(ctx.lib.ilb_resolve_lighting(ctx.handle, C.byref(params), lm_ptr, al_ptr, out.ctypes.data_as(C.c_void_p)))
        return out

    def ResolvePlaced(self, target: np.ndarray, position=(0.0, 0.0), scale=(1.0, 1.0), albedo: Optional[np.ndarray] = None,
                      albedoRegion=(0.0, 0.0, 1.0, 1.0), hdr: Optional[HDRConfiguration] = None, lightmap: Optional[np.ndarray] = None,
                      uvOffset: Tuple[float, float] = (0.0, 0.0)) -> np.ndarray:
        """RenderedLighting.Resolve with a position, a scale and an albedo region (LightingRenderer.HDR.cs:101-152 ->
        ResolveLighting, LightingRenderer.cs:1537-1645): the resolve drawn as a quad into `target` (uint8 or float32 [H, W, 4] of
        any size); returns the updated copy.  Without albedo pass scale / RenderScale like the reference (:1635)."""
        if not self.IsValid:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "Invalid")
        ctx = self.Renderer.ctx
        lm_ptr, lm_fmt, w, h = None, self.LightmapFormat, self.Width, self.Height
        if lightmap is not None:
            lightmap, lm_fmt = _texels(lightmap, "lightmap")
            h, w = lightmap.shape[0], lightmap.shape[1]
            lm_ptr = lightmap.ctypes.data_as(C.c_void_p)
        target, out_fmt = _texels(np.array(target, copy=True, order="C"), "target")
        if out_fmt == FORMAT_HALF4:
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "target must be uint8 or float32 texels")
        pl = _abi.ResolvePlacement()
        pl.target_width, pl.target_height = target.shape[1], target.shape[0]
        pl.Position[:] = [float(position[0]), float(position[1])]
        pl.Scale[:] = [float(scale[0]), float(scale[1])]
        pl.AlbedoRegion[:] = [float(v) for v in albedoRegion]
        al_ptr, al_fmt = None, FORMAT_RGBA8
        if albedo is not None:
            albedo, al_fmt = _texels(albedo, "albedo")
            if al_fmt == FORMAT_HALF4:
                raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "albedo must be uint8 or float32 texels")
            pl.albedo_width, pl.albedo_height = albedo.shape[1], albedo.shape[0]
            al_ptr = albedo.ctypes.data_as(C.c_void_p)
        params = pack_resolve(w, h, lm_fmt, hdr, al_fmt, out_fmt, uvOffset)
        self._bind_dithering(ctx, hdr)
        ctx.check(ctx.lib.ilb_resolve_lighting_placed(ctx.handle, C.byref(params), C.byref(pl), lm_ptr, al_ptr, target.ctypes.data_as(C.c_void_p)))
        return target

    def ComputeLuminance(self, level: int, lightmap: Optional[np.ndarray] = None) -> np.ndarray:
        """Level `level` of the luminance buffer (UpdateLuminanceBuffer + mips): float32 [(H/2) >> level, (W/2) >> level]."""
        ctx = self.Renderer.ctx
        lm_ptr, lm_fmt, w, h = None, self.LightmapFormat, self.Width, self.Height
        if lightmap is not None:
            lightmap, lm_fmt = _texels(lightmap, "lightmap")
            h, w = lightmap.shape[0], lightmap.shape[1]
            lm_ptr = lightmap.ctypes.data_as(C.c_void_p)
        out = np.empty((max((h // 2) >> level, 0), max((w // 2) >> level, 0)), dtype=np.float32)
        ctx.check(ctx.lib.ilb_compute_luminance(ctx.handle, w, h, lm_fmt, lm_ptr, int(level), out.ctypes.data_as(C.c_void_p)))
        return out

    def TryComputeHistogram(self, histogram: Histogram, onComplete=None, accuracyFactor: int = 3,
                            lightmap: Optional[np.ndarray] = None) -> bool:
        """LightingRenderer.HDR.cs:154-186 + HistogramUpdateTask (:21-57).  The reference analyses the PREVIOUS frame's lightmap
        (LightingRenderer.cs:989-1001: it avoids stalling on the frame in flight); here the luminance buffer is computed from the
        resident lightmap when the call is made."""
        if not self.IsValid:
            return False
        lw, lh = self.Width // 2, self.Height // 2
        if lightmap is not None:
            lh, lw = np.asarray(lightmap).shape[0] // 2, np.asarray(lightmap).shape[1] // 2
        if lw <= 0 or lh <= 0:
            return False
        levelCount = int(math.floor(math.log2(max(lw, lh)))) + 1
        levelIndex = min(int(accuracyFactor), levelCount - 1)
        buf = self.ComputeLuminance(levelIndex, lightmap)
        histogram.Clear()
        histogram.Add(buf.reshape(-1), buf.size, self.InverseScaleFactor)
        if onComplete is not None:
            onComplete(histogram)
        return True

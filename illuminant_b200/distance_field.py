"""Host-side mirror of Squared.Illuminant.DistanceField (reference: Illuminant/SDF/DistanceField.cs:18-122) and of the
uniform block the renderers derive from it (Illuminant/Uniforms.cs:79-108, Lighting/LightingRenderer.cs:1894-1940).

The atlas itself lives in HBM behind an `ilb_df` handle; this class owns the descriptor arithmetic (slice packing,
atlas tiling) and reproduces it in float32 exactly as the C# code does, because the shaders' addressing depends on
the rounding of these uniforms (see the FIXME at LightingRenderer.cs:1934-1936).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import DFUniforms, Float4, Obstruction

F = np.float32
MaxSurfaceSize = 8192            # DistanceField.cs:19
DefaultMaximumEncodedDistance = 128  # DistanceField.cs:20
PackedSliceCount = 3             # LightingRenderer.cs:313


def _round_half_even(x: float) -> int:
    # Math.Round(double) in .NET is banker's rounding; Python's round() matches
    return int(round(x))


@dataclass
class RendererQualitySettings:
    """Lighting/LightingRenderer.Configuration.cs:254-313 (defaults identical)."""
    MinStepSize: float = 3.0
    LongStepFactor: float = 1.0
    MaxStepCount: int = 64
    MaxConeRadius: float = 24
    ConeGrowthFactor: float = 1.0   # dead uniform in the shaders (DistanceFieldCommon.fxh:234-237)
    OcclusionToOpacityPower: float = 1


class LightObstructionType:
    """Illuminant/Lighting/LightObstruction.cs (type ids consumed by DistanceFunctionCommon.fxh:167-186)."""
    Ellipsoid, Box, Cylinder, Spheroid, Octagon = 1, 2, 3, 4, 5


@dataclass
class LightObstruction:
    Type: int
    Center: tuple
    Size: tuple
    Rotation: tuple = (0.0, 0.0, 0.0, 1.0)  # identity quaternion
    IsDynamic: bool = False


class DistanceField:
    """DistanceField(coordinator, virtualWidth, virtualHeight, virtualDepth, requestedSliceCount,
    requestedResolution = 1, maximumEncodedDistance = 128) -- DistanceField.cs:43-122."""

    def __init__(self, ctx: _abi.Context, virtualWidth: int, virtualHeight: int, virtualDepth: float,
                 requestedSliceCount: int, requestedResolution: float = 1.0,
                 maximumEncodedDistance: int = DefaultMaximumEncodedDistance):
        self.ctx = ctx
        self.VirtualWidth, self.VirtualHeight = int(virtualWidth), int(virtualHeight)
        self.VirtualDepth = float(virtualDepth)
        self.MaximumEncodedDistance = int(maximumEncodedDistance)
        self.RequestedResolution = requestedResolution
        requestedResolution = min(max(requestedResolution, 0.05), 1.0)
        cw = _round_half_even(self.VirtualWidth * requestedResolution)
        ch = _round_half_even(self.VirtualHeight * requestedResolution)
        frac = (self.VirtualWidth / cw + self.VirtualHeight / ch) / 2
        self.Resolution = min(max(round(1.0 / frac, 3), 0.05), 1.0)
        self.SliceWidth = _round_half_even(self.VirtualWidth * self.Resolution)
        self.SliceHeight = _round_half_even(self.VirtualHeight * self.Resolution)
        maxSlicesX, maxSlicesY = MaxSurfaceSize // self.SliceWidth, MaxSurfaceSize // self.SliceHeight
        if maxSlicesX < 1 or maxSlicesY < 1:
            raise ValueError("distance field slice larger than the 8192x8192 surface limit")
        maxSlices = maxSlicesX * maxSlicesY * PackedSliceCount
        sliceCount = max(3, int(requestedSliceCount))
        sliceCount = ((sliceCount + 2) // 3) * 3
        self.SliceCount = min(sliceCount, maxSlices)
        self.PhysicalSliceCount = int(math.ceil(self.SliceCount / float(PackedSliceCount)))
        self.ColumnCount = min(maxSlicesX, self.PhysicalSliceCount)
        self.RowCount = min(maxSlicesY, max(int(math.ceil(self.PhysicalSliceCount / float(maxSlicesX))), 1))
        while self.RowCount < self.ColumnCount and self.RowCount < maxSlicesY:   # rebalance, :96-109
            newRow = self.RowCount + 1
            newCol = int(math.ceil(self.PhysicalSliceCount / float(newRow)))
            newRow = min(newRow, maxSlicesX)
            newCol = min(newCol, maxSlicesY)
            if newRow * newCol < self.PhysicalSliceCount:
                break
            self.RowCount, self.ColumnCount = newRow, newCol
        self.TextureWidth = self.SliceWidth * self.ColumnCount
        self.TextureHeight = self.SliceHeight * self.RowCount
        self.ZOffset = 0.0
        self.ValidSliceCount = 0     # SliceInfo.ValidSliceCount
        self.handle = None

    # ---- uniforms ---------------------------------------------------------------------------------------------
    def uniforms(self, quality: RendererQualitySettings | None = None) -> DFUniforms:
        """Uniforms.DistanceField(df) (Uniforms.cs:88-108) + SetDistanceFieldParameters (LightingRenderer.cs:1915-1939)."""
        q = quality or RendererQualitySettings()
        u = DFUniforms()
        sliceZSize = F(self.VirtualDepth) / F(self.SliceCount)
        tsc_z = F(min(self.ValidSliceCount, self.SliceCount)) * sliceZSize
        u.Extent = Float4(self.VirtualWidth, self.VirtualHeight, self.VirtualDepth, self.MaximumEncodedDistance)
        u.TextureSliceCount = Float4(self.ColumnCount, self.RowCount, tsc_z, self.SliceCount)
        u.TextureSliceAndTexelSize = Float4(F(1) / F(self.ColumnCount), F(1) / F(self.RowCount),
                                            F(1) / F(self.VirtualWidth * self.ColumnCount),
                                            F(1) / F(self.VirtualHeight * self.RowCount))
        u.ConeAndMisc = Float4(q.MaxConeRadius, self.ZOffset, q.OcclusionToOpacityPower,
                               F(float(self.VirtualWidth) / self.SliceWidth))
        u.StepAndMisc2 = Float4(float(int(q.MaxStepCount)), q.MinStepSize, q.LongStepFactor,
                                F(float(self.VirtualHeight) / self.SliceHeight))
        px = (F(1) / max(F(0.0001), F(u.TextureSliceCount.x))) * (F(1) / F(3))
        py = (F(1) / max(F(0.0001), F(u.Extent.z))) * F(u.TextureSliceCount.w)
        u.Packed1 = Float4(px, py, u.TextureSliceCount.z, q.MinStepSize)
        return u

    @staticmethod
    def empty_uniforms(maximumZ: float, quality: RendererQualitySettings | None = None) -> DFUniforms:
        """The `_DistanceField == null` branch (LightingRenderer.cs:1906-1916): Extent.x = 0 disables trace and AO."""
        q = quality or RendererQualitySettings()
        u = DFUniforms()
        u.ConeAndMisc = Float4(0, 0, 0, 1)
        u.StepAndMisc2 = Float4(float(int(q.MaxStepCount)), q.MinStepSize, 0, 1)
        u.Extent = Float4(0, 0, maximumZ, 0)
        return u

    # ---- storage ----------------------------------------------------------------------------------------------
    def _release(self):
        if self.handle:
            self.ctx.lib.ilb_df_destroy(self.handle)
            self.handle = None

    def Load(self, data) -> None:
        """DistanceField.Load (DistanceField.cs:195-213): raw Rgba64 atlas, 8*W*H bytes, no header."""
        arr = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint16) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint16)
        size = 8 * self.TextureWidth * self.TextureHeight
        if arr.nbytes != size:
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "Truncated file")
        self._release()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.ilb_df_create(self.ctx.handle, self.TextureWidth, self.TextureHeight,
                                                  arr.ctypes.data_as(C.c_void_p), arr.nbytes, C.byref(h)))
        self.handle = h
        self.ValidSliceCount = ((self.SliceCount + 2) // 3) * 3

    def LoadDevice(self, device_ptr: int) -> None:
        """Same from a device pointer on this context's GPU (used to replicate the field across ranks)."""
        self._release()
        h = C.c_void_p()
        size = 8 * self.TextureWidth * self.TextureHeight
        self.ctx.check(self.ctx.lib.ilb_df_create_device(self.ctx.handle, self.TextureWidth, self.TextureHeight,
                                                         C.c_void_p(device_ptr), size, C.byref(h)))
        self.handle = h
        self.ValidSliceCount = ((self.SliceCount + 2) // 3) * 3

    def Save(self) -> np.ndarray:
        """DistanceField.Save (DistanceField.cs:178-193) -> uint16 array [TextureHeight, TextureWidth, 4]."""
        if self.handle is None or self.ValidSliceCount < self.SliceCount:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "The distance field must be fully valid")
        out = np.empty((self.TextureHeight, self.TextureWidth, 4), dtype=np.uint16)
        self.ctx.check(self.ctx.lib.ilb_df_download(self.handle, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def Rasterize(self, obstructions) -> None:
        """RenderDistanceField for analytic obstructions (LightingRenderer.DistanceField.cs:347-400), all slices at once."""
        obs = pack_obstructions(obstructions)
        self.ValidSliceCount = self.SliceCount
        u = self.uniforms()
        self._release()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.ilb_df_generate(self.ctx.handle, self.TextureWidth, self.TextureHeight, self.SliceWidth,
                                                    self.SliceHeight, self.SliceCount, C.byref(u),
                                                    C.cast(obs, C.c_void_p) if len(obs) else None, len(obs), C.byref(h)))
        self.handle = h

    def Dispose(self):
        self._release()


class DynamicDistanceField(DistanceField):
    """DynamicDistanceField (SDF/DistanceField.cs:248-310): keeps a static field (obstructions with IsDynamic == false)
    and derives the sampled field from it every time the dynamic obstructions move: the slices are cleared to the static
    texture and only the dynamic obstructions are rasterised on top (LightingRenderer.DistanceField.cs:99-118)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.static_handle = None

    def _release(self):
        super()._release()
        if getattr(self, "static_handle", None):
            self.ctx.lib.ilb_df_destroy(self.static_handle)
            self.static_handle = None

    def _generate(self, obstructions):
        obs = pack_obstructions(obstructions)
        self.ValidSliceCount = self.SliceCount
        u = self.uniforms()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.ilb_df_generate(self.ctx.handle, self.TextureWidth, self.TextureHeight, self.SliceWidth,
                                                    self.SliceHeight, self.SliceCount, C.byref(u),
                                                    C.cast(obs, C.c_void_p) if len(obs) else None, len(obs), C.byref(h)))
        return h

    def Rasterize(self, obstructions) -> None:
        """Invalidate(invalidateStatic = true) + a full update: static field from the static obstructions, then the
        sampled field from it and the dynamic ones."""
        static = [o for o in obstructions if not o.IsDynamic]
        self._release()
        self.static_handle = self._generate(static)
        self.handle = self._generate(static)            # a second atlas of the same size to hold the sampled field
        self.RasterizeDynamic([o for o in obstructions if o.IsDynamic])

    def RasterizeDynamic(self, dynamic_obstructions) -> None:
        """Invalidate(invalidateStatic = false) + update: per-frame path, rewrites the sampled field in place."""
        if self.static_handle is None or self.handle is None:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "Rasterize() the static field first")
        obs = pack_obstructions(dynamic_obstructions)
        u = self.uniforms()
        self.ctx.check(self.ctx.lib.ilb_df_update_dynamic(self.handle, self.static_handle, self.SliceWidth, self.SliceHeight, self.SliceCount,
                                                          C.byref(u), C.cast(obs, C.c_void_p) if len(obs) else None, len(obs)))

    def SaveStatic(self) -> np.ndarray:
        out = np.empty((self.TextureHeight, self.TextureWidth, 4), dtype=np.uint16)
        self.ctx.check(self.ctx.lib.ilb_df_download(self.static_handle, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def pack_obstructions(obstructions):
    arr = (Obstruction * max(len(obstructions), 1))()
    for i, o in enumerate(obstructions):
        arr[i].type = int(o.Type)
        arr[i].center[:] = [float(v) for v in o.Center]
        arr[i].size[:] = [float(v) for v in o.Size]
        arr[i].rotation[:] = [float(v) for v in o.Rotation]
    return arr if len(obstructions) else (Obstruction * 0)()

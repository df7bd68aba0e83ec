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


@dataclass
class SimpleHeightVolume:
    """SimpleHeightVolume(polygon, zBase, height) (SDF/HeightVolume.cs:14-140): an extruded polygon; only what the distance field
    reads -- Polygon, ZBase, Height, IsDynamic, Bounds (the G-buffer meshes are outside the hot-path scope)."""
    Polygon: list                      # [(x, y), ...]
    ZBase: float = 0.0
    Height: float = 0.0
    IsDynamic: bool = True             # HeightVolume.cs:23
    IsObstruction: bool = True

    @property
    def Bounds(self):
        xs, ys = [float(p[0]) for p in self.Polygon], [float(p[1]) for p in self.Polygon]
        return (min(xs), min(ys), max(xs), max(ys))


def pack_height_volumes(volumes):
    """(ilb_height_volume[], float4 edges[]) -- the VertexDataTexture of LightingRenderer.DistanceField.cs:228-240: edge j is
    (p[j], p[wrap(j + 1)])."""
    vols = [v for v in volumes if getattr(v, "IsObstruction", True)]
    nedges = sum(len(v.Polygon) for v in vols)
    arr = (_abi.HeightVolumeStruct * max(len(vols), 1))()
    edges = (Float4 * max(nedges, 1))()
    at = 0
    for i, v in enumerate(vols):
        n = len(v.Polygon)
        arr[i].first_edge, arr[i].edge_count = at, n
        arr[i].z_base, arr[i].height = float(v.ZBase), float(v.Height)
        arr[i].bounds[:] = [float(F(b)) for b in v.Bounds]
        for j in range(n):
            a, b = v.Polygon[j], v.Polygon[(j + 1) % n]
            edges[at + j] = Float4(a[0], a[1], b[0], b[1])
        at += n
    return arr, len(vols), edges, nedges


class SliceInfo:
    """SDF/DistanceField.cs:13-16."""

    def __init__(self):
        self.ValidSliceCount = 0
        self.InvalidSlices: list = []


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
        self.SliceInfo = SliceInfo()
        self.handle = None
        self.Invalidate()            # DistanceField.cs:121

    # ---- slice validity (DistanceField.cs:124-227) ----------------------------------------------------------------
    @property
    def ValidSliceCount(self) -> int:
        return self.SliceInfo.ValidSliceCount

    @ValidSliceCount.setter
    def ValidSliceCount(self, v: int) -> None:
        self.SliceInfo.ValidSliceCount = int(v)

    def Invalidate(self) -> None:
        for i in range(self.SliceCount):
            if i not in self.SliceInfo.InvalidSlices:
                self.SliceInfo.InvalidSlices.append(i)

    def ValidateSlice(self, index: int) -> None:
        if index in self.SliceInfo.InvalidSlices:
            self.SliceInfo.InvalidSlices.remove(index)

    def MarkValidSlice(self, index: int) -> None:
        self.SliceInfo.ValidSliceCount = max(self.SliceInfo.ValidSliceCount, index)

    @property
    def IsFullyGenerated(self) -> bool:
        return self.SliceInfo.ValidSliceCount >= self.SliceCount and not self.SliceInfo.InvalidSlices

    @property
    def NeedsRasterize(self) -> bool:
        return bool(self.SliceInfo.InvalidSlices)

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
        self.SliceInfo.InvalidSlices.clear()

    def LoadDevice(self, device_ptr: int) -> None:
        """Same from a device pointer on this context's GPU (used to replicate the field across ranks)."""
        self._release()
        h = C.c_void_p()
        size = 8 * self.TextureWidth * self.TextureHeight
        self.ctx.check(self.ctx.lib.ilb_df_create_device(self.ctx.handle, self.TextureWidth, self.TextureHeight,
                                                         C.c_void_p(device_ptr), size, C.byref(h)))
        self.handle = h
        self.ValidSliceCount = ((self.SliceCount + 2) // 3) * 3
        self.SliceInfo.InvalidSlices.clear()

    def Save(self) -> np.ndarray:
        """DistanceField.Save (DistanceField.cs:178-193) -> uint16 array [TextureHeight, TextureWidth, 4]."""
        if self.handle is None or self.ValidSliceCount < self.SliceCount:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "The distance field must be fully valid")
        out = np.empty((self.TextureHeight, self.TextureWidth, 4), dtype=np.uint16)
        self.ctx.check(self.ctx.lib.ilb_df_download(self.handle, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def Rasterize(self, obstructions, heightVolumes=()) -> None:
        """RenderDistanceField (LightingRenderer.DistanceField.cs:19-31, :347-464) with no limit on the slices per frame: every
        slice at once."""
        self.Invalidate()
        self.RenderDistanceField(obstructions, heightVolumes, maximumFieldUpdatesPerFrame=1 << 30)

    def _ensure_atlas(self) -> None:
        if self.handle is None:   # NeedClear (LightingRenderer.DistanceField.cs:51-55): a fresh target is cleared once
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.ilb_df_create_empty(self.ctx.handle, self.TextureWidth, self.TextureHeight, C.byref(h)))
            self.handle = h

    def _render_slices(self, handle, static_handle, first_physical: int, count: int, obstructions, heightVolumes) -> None:
        obs = pack_obstructions(obstructions)
        vols, nv, edges, ne = pack_height_volumes(heightVolumes)
        u = self.uniforms()
        self.ctx.check(self.ctx.lib.ilb_df_update_slices(handle, static_handle, self.SliceWidth, self.SliceHeight, self.SliceCount, C.byref(u),
                                                         C.cast(obs, C.c_void_p) if len(obs) else None, len(obs),
                                                         C.cast(vols, C.c_void_p) if nv else None, nv, C.cast(edges, C.c_void_p) if ne else None, ne,
                                                         first_physical, count))

    def RenderDistanceField(self, obstructions, heightVolumes=(), maximumFieldUpdatesPerFrame: int = 1) -> int:
        """RenderDistanceFieldPartition (LightingRenderer.DistanceField.cs:415-464) for a plain field: rasterises the slice triplets
        of at most `maximumFieldUpdatesPerFrame` invalid slices (Configuration.MaximumFieldUpdatesPerFrame, default 1) and
        validates them; returns the number of triplets rendered.  Call once per frame until NeedsRasterize is false."""
        info = self.SliceInfo
        slicesToUpdate = min(int(maximumFieldUpdatesPerFrame), len(info.InvalidSlices))
        if slicesToUpdate <= 0:
            return 0
        self._ensure_atlas()
        runs = []                                # consecutive physical slices are rendered by one launch
        while slicesToUpdate > 0 and info.InvalidSlices:
            s = info.InvalidSlices[0]
            physical = s // PackedSliceCount
            if runs and runs[-1][0] + runs[-1][1] == physical:
                runs[-1][1] += 1
            else:
                runs.append([physical, 1])
            for i in range(s, s + 3):            # RenderDistanceFieldSliceTriplet :137-147
                self.ValidateSlice(i)
            self.MarkValidSlice(s + 3)
            slicesToUpdate -= 3
        for first, n in runs:
            self._render_slices(self.handle, None, first, n, obstructions, heightVolumes)
        return sum(n for _, n in runs)

    def Dispose(self):
        self._release()


class DynamicDistanceField(DistanceField):
    """DynamicDistanceField (SDF/DistanceField.cs:248-310): keeps a static field (obstructions and height volumes with
    IsDynamic == false) and derives the sampled field from it: a slice of the sampled field is cleared to the static texture and
    only the dynamic items are rasterised on top (LightingRenderer.DistanceField.cs:99-118)."""

    def __init__(self, *args, **kwargs):
        self.StaticSliceInfo = SliceInfo()
        super().__init__(*args, **kwargs)
        self.static_handle = None

    def _release(self):
        super()._release()
        if getattr(self, "static_handle", None):
            self.ctx.lib.ilb_df_destroy(self.static_handle)
            self.static_handle = None

    # ---- slice validity (DistanceField.cs:265-300) ----------------------------------------------------------------
    def Invalidate(self, invalidateStatic: bool = True) -> None:
        for i in range(self.SliceCount):
            if i not in self.SliceInfo.InvalidSlices:
                self.SliceInfo.InvalidSlices.append(i)
            if invalidateStatic and i not in self.StaticSliceInfo.InvalidSlices:
                self.StaticSliceInfo.InvalidSlices.append(i)

    def _validate(self, index: int, dynamic: bool) -> None:
        if dynamic:
            if index not in self.StaticSliceInfo.InvalidSlices and index in self.SliceInfo.InvalidSlices:
                self.SliceInfo.InvalidSlices.remove(index)
        elif index in self.StaticSliceInfo.InvalidSlices:
            self.StaticSliceInfo.InvalidSlices.remove(index)

    def _mark_valid(self, index: int, dynamic: bool) -> None:
        if dynamic:
            self.SliceInfo.ValidSliceCount = min(max(self.SliceInfo.ValidSliceCount, index), self.StaticSliceInfo.ValidSliceCount)
        else:
            self.StaticSliceInfo.ValidSliceCount = max(self.StaticSliceInfo.ValidSliceCount, index)

    def _partition(self, dynamic: bool, obstructions, heightVolumes, budget: int) -> int:
        """RenderDistanceFieldPartition (:415-464) with dynamicFlagFilter = `dynamic`."""
        info = self.SliceInfo if dynamic else self.StaticSliceInfo
        slicesToUpdate = min(int(budget), len(info.InvalidSlices))
        if slicesToUpdate <= 0:
            return 0
        obs = [o for o in obstructions if bool(o.IsDynamic) == dynamic]
        vols = [v for v in heightVolumes if bool(v.IsDynamic) == dynamic]
        rendered = 0
        while slicesToUpdate > 0 and info.InvalidSlices:
            s = info.InvalidSlices[0]
            before = len(info.InvalidSlices)
            self._render_slices(self.handle if dynamic else self.static_handle, self.static_handle if dynamic else None, s // PackedSliceCount, 1,
                                obs, vols)
            for i in range(s, s + 3):
                self._validate(i, dynamic)
            self._mark_valid(s + 3, dynamic)
            rendered += 1
            slicesToUpdate -= 3
            if len(info.InvalidSlices) == before:   # a dynamic slice cannot become valid before its static slice (:24-26)
                break
        return rendered

    def RenderDistanceField(self, obstructions, heightVolumes=(), maximumFieldUpdatesPerFrame: int = 1) -> int:
        """RenderDistanceField (:19-31): the static partition, then the dynamic one, each with the per-frame slice budget."""
        if self.static_handle is None:
            for name in ("static_handle", "handle"):
                h = C.c_void_p()
                self.ctx.check(self.ctx.lib.ilb_df_create_empty(self.ctx.handle, self.TextureWidth, self.TextureHeight, C.byref(h)))
                setattr(self, name, h)
        n = self._partition(False, obstructions, heightVolumes, maximumFieldUpdatesPerFrame)
        return n + self._partition(True, obstructions, heightVolumes, maximumFieldUpdatesPerFrame)

    def Rasterize(self, obstructions, heightVolumes=()) -> None:
        """Invalidate(invalidateStatic = true) + a full update: static field from the static items, then the sampled field from it
        and the dynamic ones."""
        self.Invalidate(True)
        self.RenderDistanceField(obstructions, heightVolumes, 1 << 30)

    def RasterizeDynamic(self, dynamic_obstructions, dynamicHeightVolumes=()) -> None:
        """Invalidate(invalidateStatic = false) + a full update of the sampled field: the per-frame path for moving obstructions."""
        if self.static_handle is None or self.handle is None:
            raise _abi.IlluminantError(_abi.ERR_INVALID_OPERATION, "Rasterize() the static field first")
        self.Invalidate(False)
        self._partition(True, [o for o in dynamic_obstructions if o.IsDynamic], [v for v in dynamicHeightVolumes if v.IsDynamic], 1 << 30)

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

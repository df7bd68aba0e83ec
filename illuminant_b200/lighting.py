"""Host-side mirror of the reference lighting API for the hot path: LightingEnvironment, the three light-source types
the path covers, RendererConfiguration and LightingRenderer.RenderLighting / UpdateLightProbes.

Names, defaults and packing follow the reference (Illuminant/Lighting/*.cs, cited per member); the work the
reference enqueues as instanced draws (LightingRenderer.cs:1112-1168) becomes one call into the C-ABI
(`ilb_render_lighting*`).  Only evaluated values cross the boundary.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _abi
from ._abi import (FORMAT_FLOAT4, FORMAT_HALF4, FORMAT_RGBA8, LIGHT_DIRECTIONAL, LIGHT_LINE, LIGHT_PARTICLE, LIGHT_SPHERE, DFUniforms, Float4,
                   LightBatch, LightingFrame, LightVertex)
from .distance_field import DistanceField, LightObstruction, RendererQualitySettings

F = np.float32


class LightSourceRampMode:  # LightSource.cs:622-629
    Linear, Exponential, None_ = 0, 1, 2


class ShadowFilter:  # LightSource.cs:23-27
    None_, Unshadowed, Shadowed = -1, 0, 1


@dataclass
class LightSource:  # LightSource.cs:58-82
    Opacity: float = 1.0
    CastsShadows: bool = True
    ShadowDistanceFalloff: Optional[float] = None
    AmbientOcclusionRadius: float = 0.0
    AmbientOcclusionOpacity: float = 1.0
    FalloffYFactor: float = 1.0
    RampOffsetAndRate: Tuple[float, float] = (0.0, 1.0)
    Quality: Optional[RendererQualitySettings] = None
    Enabled: bool = True
    SortKey: int = 0
    RampTexture: Optional[np.ndarray] = None   # LightSource.cs:182-193: uint8 / float32 [H, W, 4]; 1x1 means "no ramp" (LightingRenderer.cs:819-827)

    @property
    def RampOffsetForGPU(self) -> float:
        return float(F(-math.pi) + F(self.RampOffsetAndRate[0]))

    @property
    def RampRateForGPU(self) -> float:
        return float(F(1.0 / (math.pi * 2) * self.RampOffsetAndRate[1]))


@dataclass
class SphereLightSource(LightSource):  # LightSource.cs:171-250
    Position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Radius: float = 0.0
    RampLength: float = 1.0
    RampMode: int = LightSourceRampMode.Linear
    Color: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    SpecularColor: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    SpecularPower: float = 2.0
    ShadowFilter: int = ShadowFilter.None_
    TypeID = LIGHT_SPHERE


@dataclass
class DirectionalLightSource(LightSource):  # LightSource.cs:84-169
    _Direction: Optional[Tuple[float, float, float]] = None
    Bounds: Optional[Tuple[Tuple[float, float], Tuple[float, float]]] = None  # (TopLeft, BottomRight)
    ShadowTraceLength: float = 256
    ShadowSoftness: float = 12
    ShadowRampRate: float = 0.5
    Color: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    ShadowFilter: int = ShadowFilter.None_
    TypeID = LIGHT_DIRECTIONAL

    @property
    def Direction(self):
        return self._Direction

    @Direction.setter
    def Direction(self, value):
        if value is None:
            self._Direction = None
            return
        v = np.asarray(value, dtype=F)
        n = F(np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
        self._Direction = tuple(float(c) for c in (v / n))  # Vector3.Normalize


@dataclass
class LineLightSource(LightSource):  # LightSource.cs:252-
    StartPosition: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    EndPosition: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Radius: float = 0.0
    RampMode: int = LightSourceRampMode.Linear
    StartColor: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    EndColor: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    TypeID = LIGHT_LINE


@dataclass
class ParticleLightSource(LightSource):  # LightSource.cs:466-500
    """Every live particle of `System` is a sphere light with `Template`'s properties (ParticleLight.fx); the particle
    state is read on the device when the frame is rendered."""
    Template: SphereLightSource = field(default_factory=SphereLightSource)
    System: object = None          # illuminant_b200.ParticleSystem
    IsActive: bool = True
    TypeID = LIGHT_PARTICLE

    def uniforms(self, hasDistanceField: bool):
        """_ParticleLightBatchSetup (LightingRenderer.cs:769-789): (LightProperties, MoreLightProperties, LightColor, LightSpecularColor)."""
        t = self.Template
        falloff = -99999.0 if t.ShadowDistanceFalloff is None else t.ShadowDistanceFalloff
        props = Float4(t.Radius, t.RampLength, int(t.RampMode), 1.0 if (t.CastsShadows and hasDistanceField) else 0.0)
        more = Float4(t.AmbientOcclusionRadius if t.AmbientOcclusionOpacity > 0.001 else 0.0, falloff, t.FalloffYFactor,
                      min(max(t.AmbientOcclusionOpacity, 0.0), 1.0))
        return props, more, _vec4(list(t.Color)), _vec4(list(t.SpecularColor) + [t.SpecularPower])

    def light_vertices(self, positions: np.ndarray, attributes: np.ndarray, hasDistanceField: bool):
        """What ParticleLightVertexShader (ParticleLight.fx:16-82) hands to the pixel shader for particle state
        [n,4] / [n,4] (float32 arithmetic, StippleFactor 1): a LightVertex per live particle whose light colour has
        alpha > 0.  Used to drive the oracle with the same lights the device builds."""
        props, more, color, spec = self.uniforms(hasDistanceField)
        lc = np.array([color.x, color.y, color.z, color.w], dtype=F)
        out = []
        for p, a in zip(np.asarray(positions, dtype=F), np.asarray(attributes, dtype=F)):
            if not (p[3] > 0):
                continue
            c = a.copy()
            if c[3] > 0:
                c[:3] = c[:3] / c[3]
            c = c * lc
            if not (c[3] > 0):
                continue
            v = LightVertex()
            v.LightPosition1 = v.LightPosition2 = v.LightPosition3 = Float4(float(p[0]), float(p[1]), float(p[2]), 0.0)
            v.LightProperties, v.MoreLightProperties = props, more
            v.EvenMoreLightProperties = Float4(-1, 0, 0, 0)
            v.Color1 = Float4(*[float(x) for x in c])
            v.Color2 = spec
            out.append(v)
        return out


@dataclass
class LightProbe:  # Lighting/LightProbe.cs
    Position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Normal: Optional[Tuple[float, float, float]] = None
    EnableShadows: bool = True


@dataclass
class LightingEnvironment:  # Lighting/LightingEnvironment.cs
    Lights: List[LightSource] = field(default_factory=list)
    Obstructions: List[LightObstruction] = field(default_factory=list)
    HeightVolumes: list = field(default_factory=list)   # SimpleHeightVolume: only their distance-field footprint is in scope
    GroundZ: float = 0.0
    MaximumZ: float = 128.0
    ZToYMultiplier: float = 1.0
    Ambient: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)


@dataclass
class RendererConfiguration:  # Lighting/LightingRenderer.Configuration.cs:13-252
    MaximumRenderSize: Tuple[int, int] = (1920, 1080)
    MaximumFieldUpdatesPerFrame: int = 1   # LightingRenderer.Configuration.cs:88-91: distance-field slices re-rasterised per frame
    HighQuality: bool = True            # HalfVector4 lightmap (LightingRenderer.cs:477-479)
    HighQualityGBuffer: bool = True     # Vector4 G-buffer (GBuffer.cs:31-39)
    StencilCulling: bool = False
    EnableGBuffer: bool = True
    GBufferViewportRelative: bool = False
    TwoPointFiveD: bool = False
    ScaleCompensation: bool = True
    AllowFullbright: bool = True
    LightOcclusion: float = 0.0
    RenderScale: Tuple[float, float] = (1.0, 1.0)
    MaximumLightProbeCount: int = 256
    DefaultQuality: RendererQualitySettings = field(default_factory=RendererQualitySettings)
    Float4Lightmap: bool = False        # not in the reference: fp32 lightmap for parity tests (no half rounding)


def _vec4(v, w=None) -> Float4:
    v = list(v)
    if w is not None:
        v = v[:3] + [w]
    return Float4(*v)


def pack_light_vertex(light: LightSource, intensityScale: float, hasDistanceField: bool) -> Optional[LightVertex]:
    """Render{Sphere,Directional,Line}LightSource (LightingRenderer.cs:1193-1219, :1256-1307, :1309-1337)."""
    if light.Opacity <= 0:
        return None
    v = LightVertex()
    falloff = -99999.0 if light.ShadowDistanceFalloff is None else light.ShadowDistanceFalloff
    opacity = F(light.Opacity) * F(intensityScale)
    if isinstance(light, SphereLightSource):
        v.LightPosition1 = v.LightPosition2 = v.LightPosition3 = _vec4(list(light.Position) + [0])
        color = list(light.Color)
        color[3] = float(F(color[3]) * opacity)
        v.Color1 = _vec4(color)
        v.Color2 = _vec4(list(light.SpecularColor) + [light.SpecularPower])
        v.LightProperties = Float4(light.Radius, light.RampLength, int(light.RampMode), 1.0 if (light.CastsShadows and hasDistanceField) else 0.0)
        v.MoreLightProperties = Float4(light.AmbientOcclusionRadius, falloff, light.FalloffYFactor, light.AmbientOcclusionOpacity)
        v.EvenMoreLightProperties = Float4(int(light.ShadowFilter), 0, light.RampOffsetForGPU, light.RampRateForGPU)
    elif isinstance(light, DirectionalLightSource):
        if light.Bounds is not None:
            v.LightPosition1 = Float4(light.Bounds[0][0], light.Bounds[0][1], 0, 0)
            v.LightPosition2 = Float4(light.Bounds[1][0], light.Bounds[1][1], 0, 0)
        else:
            v.LightPosition1 = Float4(-99999, -99999, 0, 0)
            v.LightPosition2 = Float4(99999, 99999, 0, 0)
        color = list(light.Color)
        color[3] = float(F(color[3]) * opacity)
        v.Color1 = _vec4(color)
        v.Color2 = _vec4(list(light._Direction) + [1.0]) if light._Direction is not None else Float4()
        v.LightProperties = Float4(1.0 if light.CastsShadows else 0.0, light.ShadowTraceLength, light.ShadowSoftness, light.ShadowRampRate)
        v.MoreLightProperties = Float4(light.AmbientOcclusionRadius, falloff, 0, light.AmbientOcclusionOpacity)
        v.EvenMoreLightProperties = Float4(int(light.ShadowFilter), 0, light.RampOffsetForGPU, light.RampRateForGPU)
    elif isinstance(light, LineLightSource):
        v.LightPosition1 = _vec4(list(light.StartPosition) + [0])
        v.LightPosition2 = _vec4(list(light.EndPosition) + [0])
        c1, c2 = list(light.StartColor), list(light.EndColor)
        c1[3] = float(F(c1[3]) * opacity)
        c2[3] = float(F(c2[3]) * opacity)
        v.Color1, v.Color2 = _vec4(c1), _vec4(c2)
        v.LightProperties = Float4(light.Radius, 0, int(light.RampMode), 1.0 if (light.CastsShadows and hasDistanceField) else 0.0)
        v.MoreLightProperties = Float4(light.AmbientOcclusionRadius, falloff, light.FalloffYFactor, light.AmbientOcclusionOpacity)
        v.EvenMoreLightProperties = Float4(0, 0, light.RampOffsetForGPU, light.RampRateForGPU)
    else:
        raise NotImplementedError(type(light).__name__)  # LightingRenderer.cs:203
    return v


class LightingRenderer:
    """LightingRenderer(content, coordinator, materials, environment, configuration) -- the parts on the hot path."""

    def __init__(self, ctx: Optional[_abi.Context], environment: LightingEnvironment, configuration: RendererConfiguration):
        self.ctx = ctx
        self.Environment = environment
        self.Configuration = configuration
        self.DistanceField: Optional[DistanceField] = None
        self.Probes: List[LightProbe] = []
        self.ViewportPosition = (0.0, 0.0)   # Materials.ViewportPosition
        self.ViewportScale = (1.0, 1.0)      # Materials.ViewportScale
        self._gbuffer_shape = None
        self.LastRendered = None             # hdr.RenderedLighting of the most recent host-output frame

    # ---- G-buffer ---------------------------------------------------------------------------------------------
    def SetGBuffer(self, texels: Optional[np.ndarray]) -> None:
        """Uploads the G-buffer (encoding GBufferShaderCommon.fxh:10-35): float32 [H,W,4] when HighQualityGBuffer,
        else float16.  None disables it (Configuration.EnableGBuffer = false)."""
        if texels is None:
            self.ctx.check(self.ctx.lib.ilb_gbuffer_upload(self.ctx.handle, 0, 0, FORMAT_FLOAT4, None))
            self._gbuffer_shape = None
            return
        fmt = FORMAT_FLOAT4 if self.Configuration.HighQualityGBuffer else FORMAT_HALF4
        arr = np.ascontiguousarray(texels, dtype=np.float32 if fmt == FORMAT_FLOAT4 else np.float16)
        h, w = arr.shape[0], arr.shape[1]
        self.ctx.check(self.ctx.lib.ilb_gbuffer_upload(self.ctx.handle, w, h, fmt, arr.ctypes.data_as(C.c_void_p)))
        self._gbuffer_shape = (h, w)

    def UpdateFields(self) -> int:
        """The distance-field part of LightingRenderer.UpdateFields (LightingRenderer.cs:1949-1975 -> RenderDistanceField,
        LightingRenderer.DistanceField.cs:19-31): re-rasterises at most Configuration.MaximumFieldUpdatesPerFrame invalid slices of
        the field from Environment.Obstructions and Environment.HeightVolumes.  Returns the number of slice triplets rendered."""
        df = self.DistanceField
        if df is None or not df.NeedsRasterize:
            return 0
        return df.RenderDistanceField(self.Environment.Obstructions, self.Environment.HeightVolumes, self.Configuration.MaximumFieldUpdatesPerFrame)

    def SetGBufferDevice(self, device_ptr: int, width: int, height: int) -> None:
        fmt = FORMAT_FLOAT4 if self.Configuration.HighQualityGBuffer else FORMAT_HALF4
        self.ctx.check(self.ctx.lib.ilb_gbuffer_upload_device(self.ctx.handle, width, height, fmt, C.c_void_p(device_ptr)))
        self._gbuffer_shape = (height, width)

    # ---- frame packing ----------------------------------------------------------------------------------------
    @property
    def lightmap_format(self) -> int:
        if self.Configuration.Float4Lightmap:
            return FORMAT_FLOAT4
        return FORMAT_HALF4 if self.Configuration.HighQuality else FORMAT_RGBA8

    def _df_uniforms(self, quality: Optional[RendererQualitySettings]) -> DFUniforms:
        q = quality or self.Configuration.DefaultQuality
        if self.DistanceField is None or self.DistanceField.handle is None:
            return DistanceField.empty_uniforms(self.Environment.MaximumZ, q)
        return self.DistanceField.uniforms(q)

    def build_frame(self, intensityScale: float = 1.0, rows: Optional[Tuple[int, int]] = None) -> LightingFrame:
        cfg, env = self.Configuration, self.Environment
        f = LightingFrame()
        f.width = int(cfg.MaximumRenderSize[0] * cfg.RenderScale[0])   # _BeginLightPass LightingRenderer.cs:729-730
        f.height = int(cfg.MaximumRenderSize[1] * cfg.RenderScale[1])
        f.lightmap_format = self.lightmap_format
        f.row_begin, f.row_end = rows if rows is not None else (0, f.height)
        f.stencil_culling = 1 if (cfg.StencilCulling and cfg.EnableGBuffer) else 0
        # ComputeUniforms LightingRenderer.cs:691-701, Uniforms.cs:14-77
        ztoy = env.ZToYMultiplier if cfg.TwoPointFiveD else 0.0
        inv = 0.0 if abs(ztoy) <= 0.0001 else float(F(1.0) / F(ztoy))
        f.EnvironmentZAndScale = Float4(env.GroundZ, env.MaximumZ, cfg.RenderScale[0], cfg.RenderScale[1])
        f.EnvironmentZToY = Float4(ztoy, inv, cfg.LightOcclusion, 0)
        # SetGBufferParameters LightingRenderer.GBuffer.cs:520-534
        if cfg.EnableGBuffer and self._gbuffer_shape is not None:
            gh, gw = self._gbuffer_shape
            f.GBufferTexelSizeAndMisc = Float4(F(1) / F(gw), F(1) / F(gh), self.ViewportScale[0], self.ViewportScale[1])
        else:
            f.GBufferTexelSizeAndMisc = Float4(0, 0, self.ViewportScale[0], self.ViewportScale[1])
        f.GBufferViewportRelative = 1.0 if cfg.GBufferViewportRelative else 0.0
        # PushLightingViewTransform LightingRenderer.cs:711-724
        vx, vy = F(self.ViewportPosition[0]), F(self.ViewportPosition[1])
        if cfg.ScaleCompensation:
            vx = vx + (F(1.0) / F(cfg.RenderScale[0])) * F(0.5)
            vy = vy + (F(1.0) / F(cfg.RenderScale[1])) * F(0.5)
        f.ViewportPosition[0], f.ViewportPosition[1] = float(vx), float(vy)
        # clear colour LightingRenderer.cs:1013-1016
        amb = [float(F(c) * F(intensityScale)) for c in env.Ambient]
        if cfg.AllowFullbright and cfg.EnableGBuffer:
            amb[3] = 0.0
        f.ClearColor = Float4(*amb)
        return f

    def build_batches(self, intensityScale: float = 1.0):
        """Groups enabled lights into LightTypeRenderStates keyed on (type, quality) in first-use order after the
        stable SortKey sort (LightingRenderer.cs:1050-1110, :801-837) and packs their LightVertex arrays."""
        hasDF = self.DistanceField is not None and self.DistanceField.handle is not None
        lights = sorted((l for l in self.Environment.Lights if l.Enabled and l.TypeID != LIGHT_PARTICLE), key=lambda l: l.SortKey)
        self._sync_particle_lights()
        groups = {}
        for l in lights:
            v = pack_light_vertex(l, intensityScale, hasDF)
            if v is None:
                continue
            q = l.Quality or self.Configuration.DefaultQuality
            ramp = self._ramp_texture_id(getattr(l, "RampTexture", None))
            key = (l.TypeID, id(q), ramp)      # LightTypeRenderStateKey: type, ramp texture, quality (LightingRenderer.cs:44-90)
            if key not in groups:
                groups[key] = (l.TypeID, q, [], ramp)
            groups[key][2].append(v)
        n = sum(len(g[2]) for g in groups.values())
        verts = (LightVertex * max(n, 1))()
        batches = (LightBatch * max(len(groups), 1))()
        i = 0
        for b, (typ, q, vs, ramp) in enumerate(groups.values()):
            batches[b].light_type = typ
            batches[b].ramp_texture = ramp
            batches[b].first_vertex = i
            batches[b].vertex_count = len(vs)
            batches[b].df = self._df_uniforms(q)
            for v in vs:
                verts[i] = v
                i += 1
        return batches, len(groups), verts, n

    def _ramp_texture_id(self, texture) -> int:
        """The library id of a light's ramp texture (uploaded once per array object); 0 for none and for 1x1 textures.  Without a
        context (CPU-side packing for the oracle) ids are assigned in first-use order; `ramp_textures` lists the arrays by id."""
        if texture is None:
            return 0
        t = np.asarray(texture)
        if t.shape[0] == 1 and t.shape[1] == 1:
            return 0
        cache = self.__dict__.setdefault("_ramp_ids", {})
        if id(texture) in cache:
            return cache[id(texture)][0]
        if t.dtype == np.uint8:
            arr, fmt = np.ascontiguousarray(t), FORMAT_RGBA8
        else:
            arr, fmt = np.ascontiguousarray(t, dtype=np.float32), FORMAT_FLOAT4
        if self.ctx is not None:
            rid = C.c_int32(0)
            self.ctx.check(self.ctx.lib.ilb_ramp_texture_create(self.ctx.handle, arr.shape[1], arr.shape[0], fmt, arr.ctypes.data_as(C.c_void_p), C.byref(rid)))
            rid = int(rid.value)
        else:
            rid = len(cache) + 1
        cache[id(texture)] = (rid, texture)   # keeps the array alive, so its id() stays unique
        return rid

    @property
    def ramp_textures(self):
        """[(id, array)] of the ramp textures seen by build_batches so far."""
        return sorted(self.__dict__.get("_ramp_ids", {}).values(), key=lambda p: p[0])

    def _sync_particle_lights(self) -> None:
        """Hands the enabled, active ParticleLightSources to the library (LightingRenderer.cs:1126-1144)."""
        if self.ctx is None:
            return
        hasDF = self.DistanceField is not None and self.DistanceField.handle is not None
        srcs = [l for l in self.Environment.Lights if l.Enabled and l.TypeID == LIGHT_PARTICLE and l.IsActive and l.System is not None]
        # the library keeps the sources on the CONTEXT until they are replaced: a renderer without particle lights must clear
        # what another renderer of the same context registered, or its frames would be lit by that renderer's particles
        if not srcs and not getattr(self.ctx, "_particle_lights_registered", False):
            return
        arr = (_abi.ParticleLightSourceStruct * max(len(srcs), 1))()
        for i, l in enumerate(srcs):
            arr[i].system = l.System.handle
            arr[i].LightProperties, arr[i].MoreLightProperties, arr[i].LightColor, arr[i].LightSpecularColor = l.uniforms(hasDF)
            arr[i].df = self._df_uniforms(l.Template.Quality)
        self.ctx.check(self.ctx.lib.ilb_lighting_set_particle_lights(self.ctx.handle, C.cast(arr, C.c_void_p) if srcs else None, len(srcs)))
        self.ctx._particle_lights_registered = bool(srcs)

    # ---- the hot path -----------------------------------------------------------------------------------------
    def RenderLighting(self, intensityScale: float = 1.0, rows: Optional[Tuple[int, int]] = None) -> np.ndarray:
        """LightingRenderer.RenderLighting (LightingRenderer.cs:917): returns the lightmap [rows, W, 4]
        (float16 when HighQuality, uint8 otherwise, float32 with Float4Lightmap)."""
        frame = self.build_frame(intensityScale, rows)
        batches, nb, verts, nv = self.build_batches(intensityScale)
        h = frame.row_end - frame.row_begin
        dtype = {FORMAT_FLOAT4: np.float32, FORMAT_HALF4: np.float16, FORMAT_RGBA8: np.uint8}[frame.lightmap_format]
        out = np.empty((h, frame.width, 4), dtype=dtype)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        self.ctx.check(self.ctx.lib.ilb_render_lighting(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                        C.cast(verts, C.c_void_p), nv, out.ctypes.data_as(C.c_void_p)))
        self._set_last_rendered(frame, intensityScale)
        return out

    def _set_last_rendered(self, frame: LightingFrame, intensityScale: float) -> None:
        """RenderLighting returns a RenderedLighting in the reference (LightingRenderer.cs:951-954: lightmap + 1 / intensityScale);
        here the texels are the return value and the handle to the device-resident lightmap is `LastRendered`."""
        from .hdr import RenderedLighting
        self.LastRendered = RenderedLighting(self, frame.width, frame.row_end - frame.row_begin, frame.lightmap_format,
                                             1.0 / intensityScale)

    def RenderLightingFrame(self, gbuffer: np.ndarray, intensityScale: float = 1.0, rows: Optional[Tuple[int, int]] = None,
                            out: Optional[np.ndarray] = None) -> np.ndarray:
        """One host-to-host frame through `ilb_render_lighting_frame`: this frame's G-buffer goes up, the lightmap comes
        down, both pipelined behind the kernels over row bands.  Same result as SetGBuffer + RenderLighting."""
        fmt = FORMAT_FLOAT4 if self.Configuration.HighQualityGBuffer else FORMAT_HALF4
        arr = np.ascontiguousarray(gbuffer, dtype=np.float32 if fmt == FORMAT_FLOAT4 else np.float16)
        gh, gw = arr.shape[0], arr.shape[1]
        self._gbuffer_shape = (gh, gw)
        frame = self.build_frame(intensityScale, rows)
        batches, nb, verts, nv = self.build_batches(intensityScale)
        h = frame.row_end - frame.row_begin
        dtype = {FORMAT_FLOAT4: np.float32, FORMAT_HALF4: np.float16, FORMAT_RGBA8: np.uint8}[frame.lightmap_format]
        if out is None:
            out = np.empty((h, frame.width, 4), dtype=dtype)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        self.ctx.check(self.ctx.lib.ilb_render_lighting_frame(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                              C.cast(verts, C.c_void_p), nv, gw, gh, fmt, arr.ctypes.data_as(C.c_void_p),
                                                              out.ctypes.data_as(C.c_void_p)))
        self._set_last_rendered(frame, intensityScale)
        return out

    def RenderLightingFrameAsync(self, gbuffer: np.ndarray, out: np.ndarray, intensityScale: float = 1.0,
                                 rows: Optional[Tuple[int, int]] = None, packed=None) -> int:
        """`ilb_render_lighting_frame_async`: queues the frame and returns its ticket; `WaitLightingFrame(ticket)` returns when
        `out` is complete.  `gbuffer` (a C-contiguous array of the G-buffer's dtype) and `out` must stay untouched until then
        and be page-locked (`Context.host_register`) for the call to return before the copies have run."""
        fmt = FORMAT_FLOAT4 if self.Configuration.HighQualityGBuffer else FORMAT_HALF4
        want = np.float32 if fmt == FORMAT_FLOAT4 else np.float16
        if gbuffer.dtype != want or not gbuffer.flags["C_CONTIGUOUS"] or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("asynchronous frames need C-contiguous arrays of the G-buffer's own dtype (no staging copy can outlive the call)")
        gh, gw = gbuffer.shape[0], gbuffer.shape[1]
        self._gbuffer_shape = (gh, gw)
        frame = self.build_frame(intensityScale, rows)
        batches, nb, verts, nv = packed if packed is not None else self.build_batches(intensityScale)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        ticket = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.ilb_render_lighting_frame_async(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                                    C.cast(verts, C.c_void_p), nv, gw, gh, fmt, gbuffer.ctypes.data_as(C.c_void_p),
                                                                    out.ctypes.data_as(C.c_void_p), C.byref(ticket)))
        self._set_last_rendered(frame, intensityScale)
        return int(ticket.value)

    def WaitLightingFrame(self, ticket: int) -> None:
        self.ctx.check(self.ctx.lib.ilb_render_lighting_frame_wait(self.ctx.handle, C.c_uint64(int(ticket))))

    def RenderLightingDevice(self, device_ptr: int, intensityScale: float = 1.0, rows: Optional[Tuple[int, int]] = None,
                             packed=None) -> None:
        """Asynchronous variant writing rows [row_begin,row_end) to a device buffer that starts at row_begin."""
        frame = self.build_frame(intensityScale, rows)
        batches, nb, verts, nv = packed if packed is not None else self.build_batches(intensityScale)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        self.ctx.check(self.ctx.lib.ilb_render_lighting_device(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                               C.cast(verts, C.c_void_p), nv, C.c_void_p(device_ptr)))

    def RenderLightingPeers(self, peer_ptrs, intensityScale: float = 1.0, rows: Optional[Tuple[int, int]] = None, packed=None) -> None:
        """Band render that stores every texel into each full-frame buffer of `peer_ptrs` (in-kernel all-gather)."""
        frame = self.build_frame(intensityScale, rows)
        batches, nb, verts, nv = packed if packed is not None else self.build_batches(intensityScale)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(p) for p in peer_ptrs])
        self.ctx.check(self.ctx.lib.ilb_render_lighting_peers(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                              C.cast(verts, C.c_void_p), nv, arr, len(peer_ptrs)))

    def UpdateLightProbes(self, intensityScale: float = 1.0, float4: bool = False) -> np.ndarray:
        """UpdateLightProbes (LightingRenderer.LightProbes.cs:49-110): probe values [N,4] (half4 like the reference's
        HalfVector4 target, or float32 for tests)."""
        n = len(self.Probes)
        if n > self.Configuration.MaximumLightProbeCount:
            raise _abi.IlluminantError(_abi.ERR_INVALID_ARGUMENT, "too many light probes")
        pos = np.zeros((max(n, 1), 4), dtype=np.float32)
        nrm = np.zeros((max(n, 1), 4), dtype=np.float32)
        for i, p in enumerate(self.Probes):   # UpdateLightProbeTexture :88-110
            pos[i] = list(p.Position) + [1.0]
            nrm[i] = (list(p.Normal) if p.Normal is not None else [0, 0, 0]) + [1.0 if p.EnableShadows else 0.0]
        frame = self.build_frame(intensityScale)
        batches, nb, verts, nv = self.build_batches(intensityScale)
        fmt = FORMAT_FLOAT4 if float4 else FORMAT_HALF4
        out = np.zeros((n, 4), dtype=np.float32 if float4 else np.float16)
        if n == 0:
            return out
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        self.ctx.check(self.ctx.lib.ilb_update_light_probes(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                            C.cast(verts, C.c_void_p), nv, pos.ctypes.data_as(C.c_void_p),
                                                            nrm.ctypes.data_as(C.c_void_p), n, fmt, out.ctypes.data_as(C.c_void_p)))
        return out


    def pack_probes(self):
        """Probe positions / normals as UpdateLightProbeTexture uploads them (LightingRenderer.LightProbes.cs:88-110)."""
        n = len(self.Probes)
        pos = np.zeros((max(n, 1), 4), dtype=np.float32)
        nrm = np.zeros((max(n, 1), 4), dtype=np.float32)
        for i, p in enumerate(self.Probes):
            pos[i] = list(p.Position) + [1.0]
            nrm[i] = (list(p.Normal) if p.Normal is not None else [0, 0, 0]) + [1.0 if p.EnableShadows else 0.0]
        return pos, nrm, n

    def UpdateLightProbesDevice(self, d_out: int, packed=None, probes=None, intensityScale: float = 1.0, float4: bool = False):
        """Asynchronous UpdateLightProbes: half4 (or float4) texels into the device buffer `d_out`, in stream order."""
        pos, nrm, n = probes if probes is not None else self.pack_probes()
        if n == 0:
            return
        frame = self.build_frame(intensityScale)
        batches, nb, verts, nv = packed if packed is not None else self.build_batches(intensityScale)
        df = self.DistanceField.handle if (self.DistanceField is not None and self.DistanceField.handle) else None
        self.ctx.check(self.ctx.lib.ilb_update_light_probes_device(self.ctx.handle, df, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                                   C.cast(verts, C.c_void_p), nv, pos.ctypes.data_as(C.c_void_p),
                                                                   nrm.ctypes.data_as(C.c_void_p), n, FORMAT_FLOAT4 if float4 else FORMAT_HALF4,
                                                                   C.c_void_p(int(d_out))))


def encode_gbuffer(normal: np.ndarray, relativeY: np.ndarray, z: np.ndarray, enableShadows: np.ndarray | bool = True,
                   fullbright: np.ndarray | bool = False, dead: np.ndarray | bool = False) -> np.ndarray:
    """Vectorised encodeGBufferSample (Shaders/GBufferShaderCommon.fxh:10-35, EnvironmentCommon.fxh:34-40) -> float32 [...,4].
    Used to synthesise benchmark scenes; the production G-buffer comes from the caller."""
    normal = np.asarray(normal, dtype=F)
    z = np.asarray(z, dtype=F)
    shape = z.shape
    out = np.zeros(shape + (4,), dtype=F)
    nx = normal[..., 0].copy()
    has_n = (normal != 0).any(axis=-1)
    nx = np.where(np.abs(nx) < F(0.0001), F(0.0001), nx)
    ex = ((np.arctan2(normal[..., 1], nx).astype(F) / F(math.pi)) + F(1.0)) * F(0.5)
    ey = (normal[..., 2] + F(1.0)) * F(0.5)
    out[..., 0] = np.where(has_n, ex, F(0))
    out[..., 1] = np.where(has_n, ey, F(0))
    out[..., 2] = np.asarray(relativeY, dtype=F)
    es = np.broadcast_to(np.asarray(enableShadows, dtype=bool), shape)
    w = ((z + F(1024)) / F(1024)) * np.where(es, F(1), F(-1)) + np.where(es, F(0), F(-1))
    out[..., 3] = np.where(np.broadcast_to(np.asarray(fullbright, dtype=bool), shape), F(99999), w)
    d = np.broadcast_to(np.asarray(dead, dtype=bool), shape)
    out[d] = np.array([0, 0, -99999, -99999], dtype=F)
    return out

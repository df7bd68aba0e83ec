"""ctypes mirror of include/illuminant_b200.h and the loader of libilluminant_b200.so.

There is no CPU fallback: importing the package works without the library (host-side packing logic is pure
Python), but any call that needs the device raises IlluminantError if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["ILB_LIB"]) if os.environ.get("ILB_LIB") else PKG / "libilluminant_b200.so"  # ILB_LIB: A/B builds of kernel variants

ILB_OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_INVALID_OPERATION, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
FORMAT_FLOAT4, FORMAT_HALF4, FORMAT_RGBA8 = 0, 1, 2
LIGHT_SPHERE, LIGHT_DIRECTIONAL, LIGHT_PARTICLE, LIGHT_LINE = 1, 2, 3, 4
HDR_NONE, HDR_GAMMA_COMPRESS, HDR_TONE_MAP = 0, 1, 2
SPAWN_INLINE, SPAWN_POSITION_TEXTURE, SPAWN_FEEDBACK, SPAWN_PATTERN = 0, 1, 2, 3
BLEND_ALPHA, BLEND_ADDITIVE, BLEND_OPAQUE = 0, 1, 2
TEXTURE_NONE, TEXTURE_POINT, TEXTURE_LINEAR = 0, 1, 2
OP_GRAVITY, OP_NOISE, OP_FMA, OP_MATRIX_MULTIPLY = 1, 2, 3, 4
MAX_ATTRACTORS = 16
FORMAT_BYTES = {FORMAT_FLOAT4: 16, FORMAT_HALF4: 8, FORMAT_RGBA8: 4}


class Float4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]

    def __init__(self, x=0.0, y=0.0, z=0.0, w=0.0):
        super().__init__(float(x), float(y), float(z), float(w))

    def tuple(self):
        return (self.x, self.y, self.z, self.w)


def f4(v) -> Float4:
    if isinstance(v, Float4):
        return Float4(v.x, v.y, v.z, v.w)
    v = list(v)
    return Float4(*v)


class DFUniforms(C.Structure):
    _fields_ = [(n, Float4) for n in ("ConeAndMisc", "TextureSliceAndTexelSize", "StepAndMisc2", "TextureSliceCount", "Extent", "Packed1")]


class Obstruction(C.Structure):
    _fields_ = [("type", C.c_int32), ("center", C.c_float * 3), ("size", C.c_float * 3), ("rotation", C.c_float * 4)]


class LightVertex(C.Structure):
    _fields_ = [(n, Float4) for n in ("LightPosition1", "LightPosition2", "LightPosition3", "LightProperties",
                                      "MoreLightProperties", "EvenMoreLightProperties", "Color1", "Color2")]


class LightBatch(C.Structure):
    _fields_ = [("light_type", C.c_int32), ("first_vertex", C.c_int32), ("vertex_count", C.c_int32), ("ramp_texture", C.c_int32),
                ("df", DFUniforms)]


class ParticleLightSourceStruct(C.Structure):  # ilb_particle_light_source
    _fields_ = [("system", C.c_void_p), ("LightProperties", Float4), ("MoreLightProperties", Float4), ("LightColor", Float4),
                ("LightSpecularColor", Float4), ("df", DFUniforms)]


class LightingFrame(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("lightmap_format", C.c_int32), ("row_begin", C.c_int32),
                ("row_end", C.c_int32), ("stencil_culling", C.c_int32),
                ("EnvironmentZAndScale", Float4), ("EnvironmentZToY", Float4), ("GBufferTexelSizeAndMisc", Float4),
                ("GBufferViewportRelative", C.c_float), ("ViewportPosition", C.c_float * 2), ("reserved2", C.c_float),
                ("ClearColor", Float4)]


class Resolve(C.Structure):  # ilb_resolve
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("lightmap_format", C.c_int32), ("albedo_format", C.c_int32),
                ("output_format", C.c_int32), ("hdr_mode", C.c_int32), ("InverseScaleFactor", C.c_float),
                ("AlbedoIsSRGB", C.c_float), ("ResolveToSRGB", C.c_float), ("Offset", C.c_float), ("ExposureMinusOne", C.c_float),
                ("GammaMinusOne", C.c_float), ("MiddleGray", C.c_float), ("AverageLuminance", C.c_float),
                ("MaximumLuminanceSquared", C.c_float), ("WhitePoint", C.c_float), ("LightmapUVOffset", C.c_float * 2),
                ("DitheringStrength", C.c_float), ("reserved", C.c_float)]


class Bezier1(C.Structure):
    _fields_ = [("RangeAndCount", Float4), ("ABCD", Float4)]


class Bezier4(C.Structure):
    _fields_ = [("RangeAndCount", Float4), ("A", Float4), ("B", Float4), ("C", Float4), ("D", Float4)]


class PsysUniforms(C.Structure):
    _fields_ = [("GlobalSettings", Float4), ("CollisionSettings", Float4), ("TexelAndSize", Float4),
                ("AnimationRateAndRotationAndZToY", Float4), ("ColorFromLife", Bezier4), ("ColorFromVelocity", Bezier4),
                ("SizeFromLife", Bezier1), ("SizeFromVelocity", Bezier1), ("LifeRampSettings", Float4),
                ("RotationFromLifeAndIndex", C.c_float * 2), ("has_collision_field", C.c_int32),
                ("write_render_outputs", C.c_int32), ("CollisionField", DFUniforms)]


class Area(C.Structure):
    _fields_ = [("AreaType", C.c_int32), ("AreaCenter", C.c_float * 3), ("AreaSize", C.c_float * 3), ("AreaFalloff", C.c_float),
                ("AreaRotation", C.c_float), ("Strength", C.c_float), ("CategoryFilter", C.c_float * 2)]


class GravityOp(C.Structure):
    _fields_ = [("AttractorCount", C.c_int32), ("MaximumAcceleration", C.c_float), ("CategoryFilter", C.c_float * 2),
                ("AttractorPositions", Float4 * MAX_ATTRACTORS), ("AttractorRadiusesAndStrengths", Float4 * MAX_ATTRACTORS)]


class NoiseOp(C.Structure):
    _fields_ = [("area", Area), ("TimeDivisor", C.c_float), ("FrequencyLerp", C.c_float), ("ReplaceOldVelocity", C.c_float),
                ("reserved", C.c_float), ("RandomnessOffset", C.c_float * 2), ("NextRandomnessOffset", C.c_float * 2),
                ("RandomnessTexel", C.c_float * 2), ("PositionOffset", Float4), ("PositionMinimum", Float4), ("PositionScale", Float4),
                ("VelocityOffset", Float4), ("VelocityMinimum", Float4), ("VelocityScale", Float4)]


class FMAOp(C.Structure):
    _fields_ = [("area", Area), ("TimeDivisor", C.c_float), ("reserved", C.c_float * 3), ("PositionAdd", Float4),
                ("PositionMultiply", Float4), ("VelocityAdd", Float4), ("VelocityMultiply", Float4)]


class MatrixOp(C.Structure):
    _fields_ = [("area", Area), ("TimeDivisor", C.c_float), ("reserved", C.c_float * 3), ("PositionMatrix", C.c_float * 16),
                ("VelocityMatrix", C.c_float * 16)]


class _OpUnion(C.Union):
    _fields_ = [("gravity", GravityOp), ("noise", NoiseOp), ("fma", FMAOp), ("matrix", MatrixOp)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32 * 3), ("u", _OpUnion)]


class Spawn(C.Structure):
    _fields_ = [("chunk", C.c_int32), ("reserved", C.c_int32 * 3), ("ChunkSizeAndIndices", Float4), ("Configuration", Float4 * 9),
                ("FormulaTypes", Float4), ("InlinePositionConstants", Float4 * 4), ("PositionMatrix", C.c_float * 16),
                ("VelocityMatrix", C.c_float * 16), ("RandomnessOffset", C.c_float * 2), ("RandomnessTexel", C.c_float * 2),
                ("AxisMask", C.c_float * 3), ("AlignVelocityAndPosition", C.c_float), ("PositionConstantCount", C.c_float),
                ("PolygonRate", C.c_float), ("PolygonLoop", C.c_float), ("AttributeDiscardThreshold", C.c_float)]


class SpawnSource(C.Structure):  # ilb_spawn_source
    _fields_ = [("kind", C.c_int32), ("position_count", C.c_int32), ("positions", C.c_void_p), ("source_system", C.c_void_p),
                ("source_chunk", C.c_int32), ("FeedbackSourceIndex", C.c_float), ("InstanceMultiplier", C.c_float),
                ("SourceVelocityFactor", C.c_float), ("AlignPositionConstant", C.c_float), ("MultiplyLife", C.c_float),
                ("MultiplyAttributeConstant", C.c_float), ("SourceLifeRange", C.c_float * 2), ("reserved", C.c_int32),
                ("pattern_texels", C.c_void_p), ("pattern_width", C.c_int32), ("pattern_height", C.c_int32),
                ("StepWidthAndSizeScale", Float4), ("YOffsetsAndCoordScale", Float4), ("TexelOffsetAndMipBias", Float4),
                ("CenteringOffset", C.c_float * 2), ("reserved2", C.c_float * 2)]


class ParticleRender(C.Structure):  # ilb_particle_render
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("target_format", C.c_int32), ("blend", C.c_int32),
                ("texture_filter", C.c_int32), ("texture_width", C.c_int32), ("texture_height", C.c_int32), ("clear", C.c_int32),
                ("ClearColor", Float4), ("GlobalColor", Float4), ("BitmapTextureRegion", Float4), ("SizeFactorAndPosition", Float4),
                ("Scale", Float4), ("ZFormula", Float4), ("ZConfiguration", Float4), ("RoundingPowerFromLife", Bezier1),
                ("RenderingOptions", Float4), ("TexelAndSize", Float4), ("AnimationRateAndRotationAndZToY", Float4),
                ("ViewportPosition", C.c_float * 2), ("ViewportScale", C.c_float * 2), ("StippleFactor", C.c_float),
                ("reserved", C.c_float * 3)]


class ResolvePlacement(C.Structure):  # ilb_resolve_placement
    _fields_ = [("target_width", C.c_int32), ("target_height", C.c_int32), ("Position", C.c_float * 2), ("Scale", C.c_float * 2),
                ("AlbedoRegion", C.c_float * 4), ("albedo_width", C.c_int32), ("albedo_height", C.c_int32)]


class Dithering(C.Structure):  # ilb_dithering
    _fields_ = [("Strength", C.c_float), ("Unit", C.c_float), ("FrameIndex", C.c_float), ("BandSize", C.c_float), ("RangeMin", C.c_float),
                ("RangeMax", C.c_float)]


class LutBlending(C.Structure):  # ilb_lut_blending
    _fields_ = [("dark_resolution", C.c_int32), ("bright_resolution", C.c_int32), ("dark_row_count", C.c_int32), ("bright_row_count", C.c_int32),
                ("DarkLevel", C.c_float), ("NeutralBandSize", C.c_float), ("BrightLevel", C.c_float), ("PerChannel", C.c_float),
                ("LUTOnly", C.c_float), ("LUTOffsets", C.c_float * 4), ("reserved", C.c_float)]


class HeightVolumeStruct(C.Structure):  # ilb_height_volume
    _fields_ = [("first_edge", C.c_int32), ("edge_count", C.c_int32), ("z_base", C.c_float), ("height", C.c_float), ("bounds", C.c_float * 4)]


class IlluminantError(RuntimeError):
    """Raised for every non-zero ilb_status; `.code` carries the status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[ilb_status {code}] {message}")
        self.code = code


# ilb_option
(OPT_LIGHT_CONCURRENT, OPT_LIGHT_LINE_CTAS, OPT_LIGHT_OTHER_CTAS, OPT_LIGHT_LINE_HELPERS, OPT_LIGHT_OTHER_HELPERS, OPT_LIGHT_PDL,
 OPT_LIGHT_CONST_BANK, OPT_LIGHT_SPLIT_BAND, OPT_LIGHT_TILE_ORDER) = range(9)

# every symbol include/illuminant_b200.h declares: (name, restype, argtypes)
P = C.c_void_p
_PROTOTYPES = [
    ("ilb_abi_version", C.c_int, []),
    ("ilb_create", C.c_int, [C.c_int, C.POINTER(P)]),
    ("ilb_destroy", None, [P]),
    ("ilb_last_error", C.c_char_p, [P]),
    ("ilb_synchronize", C.c_int, [P]),
    ("ilb_stream", P, [P]),
    ("ilb_host_register", C.c_int, [P, P, C.c_size_t]),
    ("ilb_host_unregister", C.c_int, [P, P]),
    ("ilb_launch_count", C.c_uint64, [P]),
    ("ilb_set_option", C.c_int, [P, C.c_int, C.c_int]),
    ("ilb_get_option", C.c_int, [P, C.c_int, C.POINTER(C.c_int)]),
    ("ilb_debug_detmath", C.c_int, [P, C.c_int, P, P, C.c_int]),
    ("ilb_df_create", C.c_int, [P, C.c_int, C.c_int, P, C.c_size_t, C.POINTER(P)]),
    ("ilb_df_create_device", C.c_int, [P, C.c_int, C.c_int, P, C.c_size_t, C.POINTER(P)]),
    ("ilb_df_download", C.c_int, [P, P, C.c_size_t]),
    ("ilb_df_destroy", None, [P]),
    ("ilb_df_update_dynamic", C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_int]),
    ("ilb_df_create_empty", C.c_int, [P, C.c_int, C.c_int, C.POINTER(P)]),
    ("ilb_df_update_slices", C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_int, P, C.c_int, P, C.c_int, C.c_int, C.c_int]),
    ("ilb_df_generate", C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_int, C.POINTER(P)]),
    ("ilb_gbuffer_upload", C.c_int, [P, C.c_int, C.c_int, C.c_int, P]),
    ("ilb_gbuffer_upload_device", C.c_int, [P, C.c_int, C.c_int, C.c_int, P]),
    ("ilb_gbuffer_upload_rows", C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P]),
    ("ilb_render_lighting", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, P]),
    ("ilb_render_lighting_frame", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int, P, P]),
    ("ilb_render_lighting_frame_async", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int, P, P,
                                                  C.POINTER(C.c_uint64)]),
    ("ilb_render_lighting_frame_wait", C.c_int, [P, C.c_uint64]),
    ("ilb_lighting_set_particle_lights", C.c_int, [P, P, C.c_int]),
    ("ilb_render_lighting_device", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, P]),
    ("ilb_render_lighting_peers", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, C.POINTER(P), C.c_int]),
    ("ilb_update_light_probes", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, P, P, C.c_int, C.c_int, P]),
    ("ilb_update_light_probes_device", C.c_int, [P, P, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, P, P, C.c_int, C.c_int, P]),
    ("ilb_resolve_lighting", C.c_int, [P, C.POINTER(Resolve), P, P, P]),
    ("ilb_resolve_lighting_device", C.c_int, [P, C.POINTER(Resolve), P, P, P]),
    ("ilb_resolve_lighting_placed", C.c_int, [P, C.POINTER(Resolve), C.POINTER(ResolvePlacement), P, P, P]),
    ("ilb_resolve_lighting_placed_device", C.c_int, [P, C.POINTER(Resolve), C.POINTER(ResolvePlacement), P, P, P]),
    ("ilb_ramp_texture_create", C.c_int, [P, C.c_int, C.c_int, C.c_int, P, C.POINTER(C.c_int32)]),
    ("ilb_ramp_texture_destroy", C.c_int, [P, C.c_int32]),
    ("ilb_set_dithering", C.c_int, [P, C.POINTER(Dithering)]),
    ("ilb_resolve_lighting_lut", C.c_int, [P, C.POINTER(Resolve), C.POINTER(LutBlending), P, P, P, P, P]),
    ("ilb_resolve_lighting_lut_device", C.c_int, [P, C.POINTER(Resolve), C.POINTER(LutBlending), P, P, P, P, P]),
    ("ilb_compute_luminance", C.c_int, [P, C.c_int, C.c_int, C.c_int, P, C.c_int, P]),
    ("ilb_particles_create", C.c_int, [P, C.c_int, C.c_int, C.POINTER(P)]),
    ("ilb_particles_destroy", None, [P]),
    ("ilb_particles_set_randomness", C.c_int, [P, P, C.c_int, C.c_int]),
    ("ilb_particles_set_life_ramp", C.c_int, [P, P, C.c_int, C.c_int]),
    ("ilb_particles_set_collision_field", C.c_int, [P, P]),
    ("ilb_particles_upload_chunk", C.c_int, [P, C.c_int, P, P, P]),
    ("ilb_particles_upload_buffer", C.c_int, [P, C.c_int, C.c_int, P]),
    ("ilb_particles_download_chunk", C.c_int, [P, C.c_int, P, P, P, P, P]),
    ("ilb_particles_set_live_chunks", C.c_int, [P, C.c_int]),
    ("ilb_particles_step", C.c_int, [P, C.POINTER(PsysUniforms), P, C.c_int, P, C.c_int, C.c_int]),
    ("ilb_particles_step_sources", C.c_int, [P, C.POINTER(PsysUniforms), P, P, C.c_int, P, C.c_int, C.c_int]),
    ("ilb_particles_render", C.c_int, [P, C.POINTER(ParticleRender), P, P]),
    ("ilb_particles_render_device", C.c_int, [P, C.POINTER(ParticleRender), P, P]),
    ("ilb_particles_composite_layers", C.c_int, [P, C.POINTER(P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, C.POINTER(P), C.c_int]),
    ("ilb_particles_device_buffer", P, [P, C.c_int]),
    ("ilb_particles_count_live", C.c_int, [P, C.POINTER(C.c_int64)]),
    ("ilb_particles_request_chunk_liveness", C.c_int, [P]),
    ("ilb_particles_poll_chunk_liveness", C.c_int, [P, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int), C.c_int]),
    ("ilb_particles_remove_chunk", C.c_int, [P, C.c_int]),
]
EXPORTED_SYMBOLS = [p[0] for p in _PROTOTYPES]

_lib = None


def load_library():
    """Loads libilluminant_b200.so and binds every prototype. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise IlluminantError(ERR_NO_DEVICE, f"{LIB_PATH} is missing: build it with `python -m illuminant_b200.build` "
                                             "(the CUDA library is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, restype, argtypes in _PROTOTYPES:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(ctx_handle, code: int):
    if code != ILB_OK:
        lib = load_library()
        msg = lib.ilb_last_error(ctx_handle)
        raise IlluminantError(code, msg.decode("utf-8", "replace") if msg else "")


class Context:
    """One CUDA device + stream (ilb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = P()
        code = self.lib.ilb_create(int(device), C.byref(h))
        if code != ILB_OK:
            msg = self.lib.ilb_last_error(None)
            raise IlluminantError(code, msg.decode("utf-8", "replace") if msg else "")
        self.handle = h
        self.device = int(device)

    def check(self, code: int):
        check(self.handle, code)

    def synchronize(self):
        self.check(self.lib.ilb_synchronize(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.ilb_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ilb_launch_count(self.handle))

    def host_register(self, ptr: int, nbytes: int):
        """Page-locks caller memory for asynchronous copies (ilb_host_register)."""
        self.check(self.lib.ilb_host_register(self.handle, C.c_void_p(int(ptr)), C.c_size_t(int(nbytes))))

    def host_unregister(self, ptr: int):
        self.check(self.lib.ilb_host_unregister(self.handle, C.c_void_p(int(ptr))))

    def set_option(self, option: int, value: int):
        """Scheduling knobs (ilb_option): never change results."""
        self.check(self.lib.ilb_set_option(self.handle, int(option), int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int(0)
        self.check(self.lib.ilb_get_option(self.handle, int(option), C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ilb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

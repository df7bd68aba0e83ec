"""TEST INFRASTRUCTURE -- ctypes binding of liboracle.so (the CPU restatement of the reference algorithm).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It reuses the boundary struct definitions of the product package (`illuminant_b200._abi`) because both sides
consume the same C structs from include/illuminant_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from illuminant_b200 import _abi
from illuminant_b200._abi import Bezier1, Bezier4, DFUniforms, LightingFrame, PsysUniforms

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None
P = C.c_void_p


def build(force: bool = False) -> Path:
    srcs = [HERE / n for n in ("oracle_lighting.cpp", "oracle_particles.cpp", "oracle_inputs.cpp", "oracle_resolve.cpp", "hlsl.hpp", "oracle.h",
                               "Makefile")] + [HERE.parent / "include" / "illuminant_b200.h", HERE.parent / "include" / "ilb_detmath.h"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        res = subprocess.run(["make", "-C", str(HERE), "-B", "liboracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.orc_render_lighting.restype = C.c_int
        L.orc_render_lighting.argtypes = [P, C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_int, C.POINTER(LightingFrame), P, C.c_int, P,
                                          C.c_int, P, C.c_int]
        L.orc_render_lighting_strided.restype = C.c_int
        L.orc_render_lighting_strided.argtypes = [P, C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_int, C.POINTER(LightingFrame), P, C.c_int, P,
                                                  C.c_int, C.c_int, P, C.c_int]
        L.orc_update_light_probes.restype = C.c_int
        L.orc_update_light_probes.argtypes = [P, C.c_int, C.c_int, C.POINTER(LightingFrame), P, C.c_int, P, C.c_int, P, P, C.c_int, P]
        L.orc_sample_distance_field.restype = C.c_float
        L.orc_sample_distance_field.argtypes = [P, C.c_int, C.c_int, C.POINTER(DFUniforms), C.c_float, C.c_float, C.c_float]
        L.orc_cone_trace.restype = C.c_float
        L.orc_cone_trace.argtypes = [P, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_float, C.c_float, C.c_float, C.c_float, P,
                                     C.c_int, C.POINTER(C.c_int)]
        L.orc_sphere_light_opacity.restype = C.c_float
        L.orc_sphere_light_opacity.argtypes = [C.POINTER(LightingFrame), P, P, P, P, C.c_float]
        L.orc_decode_gbuffer.restype = None
        L.orc_decode_gbuffer.argtypes = [C.POINTER(LightingFrame), P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_evaluate_by_type_id.restype = C.c_float
        L.orc_evaluate_by_type_id.argtypes = [C.c_int, P, P, P, P]
        L.orc_set_life_ramp.restype = None
        L.orc_set_life_ramp.argtypes = [P, C.c_int, C.c_int]
        L.orc_bezier1.restype = C.c_float
        L.orc_bezier1.argtypes = [C.POINTER(Bezier1), C.c_float]
        L.orc_bezier4.restype = None
        L.orc_bezier4.argtypes = [C.POINTER(Bezier4), C.c_float, P]
        L.orc_particles_step.restype = C.c_int
        L.orc_particles_step.argtypes = [P, P, P, P, P, C.c_int, C.c_int, C.POINTER(PsysUniforms), P, C.c_int, P, C.c_int, P, C.c_int,
                                         C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_particles_step_sources.restype = C.c_int
        L.orc_particles_step_sources.argtypes = [P, P, P, P, P, C.c_int, C.c_int, C.POINTER(PsysUniforms), P, P, P, C.c_int, P, C.c_int, P,
                                                 C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_particles_render.restype = C.c_int
        L.orc_particles_render.argtypes = [P, P, P, C.c_long, C.POINTER(_abi.ParticleRender), P, P]
        L.orc_generate_distance_field.restype = C.c_int
        L.orc_generate_distance_field.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_int, C.c_int]
        L.orc_encode_gbuffer_sample.restype = None
        L.orc_encode_gbuffer_sample.argtypes = [P, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, P]
        L.orc_resolve_lighting.restype = C.c_int
        L.orc_resolve_lighting.argtypes = [C.POINTER(_abi.Resolve), P, P, P]
        L.orc_compute_luminance.restype = C.c_int
        L.orc_compute_luminance.argtypes = [P, C.c_int, C.c_int, C.c_int, P]
        L.orc_detmath.restype = None
        L.orc_detmath.argtypes = [C.c_int, P, P, C.c_long]
        L.orc_float_to_half.restype = None
        L.orc_float_to_half.argtypes = [P, P, C.c_long]
        L.orc_half_to_float.restype = None
        L.orc_half_to_float.argtypes = [P, P, C.c_long]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(P) if a is not None else None


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


def threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def render_lighting(df_tex, gbuffer, frame: LightingFrame, batches, nb, verts, nv, nthreads: int = 0, row_stride: int = 1) -> np.ndarray:
    """fp32 lightmap [rows, W, 4] of the multi-pass reference algorithm. df_tex: uint16 [TH,TW,4] or None;
    gbuffer: float32/float16 [H,W,4] or None.  row_stride > 1: only rows row_begin, row_begin + stride, ... (packed)."""
    tw = th = 0
    if df_tex is not None:
        df_tex = np.ascontiguousarray(df_tex, dtype=np.uint16)
        th, tw = df_tex.shape[0], df_tex.shape[1]
    gw = gh = gfmt = 0
    if gbuffer is not None:
        gfmt = _abi.FORMAT_HALF4 if gbuffer.dtype == np.float16 else _abi.FORMAT_FLOAT4
        gbuffer = np.ascontiguousarray(gbuffer)
        gh, gw = gbuffer.shape[0], gbuffer.shape[1]
    out = np.empty(((frame.row_end - frame.row_begin + row_stride - 1) // row_stride, frame.width, 4), dtype=np.float32)
    rc = lib().orc_render_lighting_strided(_ptr(df_tex), tw, th, _ptr(gbuffer), gw, gh, gfmt, C.byref(frame), C.cast(batches, P), nb,
                                           C.cast(verts, P), nv, row_stride, _ptr(out), nthreads or threads())
    if rc != 0:
        raise RuntimeError(f"orc_render_lighting failed: {rc}")
    return out


_ramp_keepalive = []


def set_ramp_textures(textures) -> None:
    """The ramp textures ilb_light_batch.ramp_texture refers to: a list of (id, array) with ids 1..n (LightingRenderer.ramp_textures),
    or [] / None.  uint8 texels decode as c / 255."""
    global _ramp_keepalive
    textures = sorted(textures or [], key=lambda p: p[0])
    arrays = []
    for k, (rid, t) in enumerate(textures):
        assert rid == k + 1, "ramp texture ids must be 1..n"
        t = np.asarray(t)
        arrays.append(np.ascontiguousarray(t.astype(np.float32) / np.float32(255.0)) if t.dtype == np.uint8 else np.ascontiguousarray(t, dtype=np.float32))
    n = len(arrays)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrays])
    ws = (C.c_int * max(n, 1))(*[a.shape[1] for a in arrays])
    hs = (C.c_int * max(n, 1))(*[a.shape[0] for a in arrays])
    L = lib()
    L.orc_set_ramp_textures.restype = None
    L.orc_set_ramp_textures.argtypes = [C.c_int, P, P, P]
    L.orc_set_ramp_textures(n, C.cast(ptrs, P), C.cast(ws, P), C.cast(hs, P))
    _ramp_keepalive = arrays     # the oracle keeps the texels by reference


def update_light_probes(df_tex, frame, batches, nb, verts, nv, positions, normals) -> np.ndarray:
    tw = th = 0
    if df_tex is not None:
        df_tex = np.ascontiguousarray(df_tex, dtype=np.uint16)
        th, tw = df_tex.shape[0], df_tex.shape[1]
    positions, normals = np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(normals, np.float32)
    n = positions.shape[0]
    out = np.zeros((n, 4), dtype=np.float32)
    rc = lib().orc_update_light_probes(_ptr(df_tex), tw, th, C.byref(frame), C.cast(batches, P), nb, C.cast(verts, P), nv,
                                       _ptr(positions), _ptr(normals), n, _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_update_light_probes failed: {rc}")
    return out


def sample_distance_field(df_tex, u: DFUniforms, x, y, z) -> float:
    df_tex = np.ascontiguousarray(df_tex, dtype=np.uint16)
    return float(lib().orc_sample_distance_field(_ptr(df_tex), df_tex.shape[1], df_tex.shape[0], C.byref(u), x, y, z))


def cone_trace(df_tex, u: DFUniforms, light_center, radius, ramp_length, shaded, growth=1.0, falloff=-99999.0, enable=True):
    tw = th = 0
    if df_tex is not None:
        df_tex = np.ascontiguousarray(df_tex, dtype=np.uint16)
        th, tw = df_tex.shape[0], df_tex.shape[1]
    steps = C.c_int(0)
    lc, sp = _f3(light_center), _f3(shaded)
    v = lib().orc_cone_trace(_ptr(df_tex), tw, th, C.byref(u), _ptr(lc), radius, ramp_length, growth, falloff, _ptr(sp),
                             1 if enable else 0, C.byref(steps))
    return float(v), steps.value


def sphere_light_opacity(frame, pos, normal, center, props, y_factor=1.0) -> float:
    a, b, c, d = _f3(pos), _f3(normal), _f3(center), _f3(props)
    return float(lib().orc_sphere_light_opacity(C.byref(frame), _ptr(a), _ptr(b), _ptr(c), _ptr(d), y_factor))


def decode_gbuffer(frame, gbuffer, x, y):
    gfmt = _abi.FORMAT_HALF4 if gbuffer.dtype == np.float16 else _abi.FORMAT_FLOAT4
    gbuffer = np.ascontiguousarray(gbuffer)
    wp, n = np.zeros(3, np.float32), np.zeros(3, np.float32)
    es, fb = C.c_int(0), C.c_int(0)
    lib().orc_decode_gbuffer(C.byref(frame), _ptr(gbuffer), gbuffer.shape[1], gbuffer.shape[0], gfmt, x, y, _ptr(wp), _ptr(n),
                             C.byref(es), C.byref(fb))
    return wp, n, bool(es.value), bool(fb.value)


def evaluate_by_type_id(type_id, world_pos, center, size, rotation=(0, 0, 0, 1)) -> float:
    a, b, c, d = _f3(world_pos), _f3(center), _f3(size), _f3(rotation)
    return float(lib().orc_evaluate_by_type_id(int(type_id), _ptr(a), _ptr(b), _ptr(c), _ptr(d)))


def bezier1(b: Bezier1, value: float) -> float:
    return float(lib().orc_bezier1(C.byref(b), value))


def bezier4(b: Bezier4, value: float) -> np.ndarray:
    out = np.zeros(4, np.float32)
    lib().orc_bezier4(C.byref(b), value, _ptr(out))
    return out


class SourceState(C.Structure):  # orc_source_state
    _fields_ = [("P", C.c_void_p), ("V", C.c_void_p), ("RC", C.c_void_p), ("chunk_size", C.c_int)]


def particles_step(P_, V_, A_, chunk_size, u: PsysUniforms, spawns, ops, rng_table, df_tex=None, steps=1, nthreads=0, life_ramp=None,
                   sources=None, source_states=None):
    """In-place multi-pass update of [chunks*chunk_size^2, 4] float32 state. Returns (P, V, A, RC, RD).
    life_ramp: float32 [H, W, 4] LifeRampTexture or None.
    sources: None or one ilb_spawn_source / None per spawn; source_states: for FEEDBACK entries a tuple
    (P, V, RC, chunk_size) of float32 [chunk_size^2, 4] arrays holding the SOURCE CHUNK's state."""
    if life_ramp is not None:
        life_ramp = np.ascontiguousarray(life_ramp, dtype=np.float32)
        lib().orc_set_life_ramp(_ptr(life_ramp), life_ramp.shape[1], life_ramp.shape[0])
    else:
        lib().orc_set_life_ramp(None, 0, 0)
    P_, V_, A_ = (np.ascontiguousarray(a, dtype=np.float32).copy() for a in (P_, V_, A_))
    per = chunk_size * chunk_size
    live = P_.shape[0] // per
    RC, RD = np.zeros_like(P_), np.zeros_like(P_)
    tw = th = 0
    if df_tex is not None:
        df_tex = np.ascontiguousarray(df_tex, dtype=np.uint16)
        th, tw = df_tex.shape[0], df_tex.shape[1]
    rw = rh = 0
    if rng_table is not None:
        rng_table = np.ascontiguousarray(rng_table, dtype=np.float32)
        rh, rw = rng_table.shape[0], rng_table.shape[1]
    sp = (_abi.Spawn * max(len(spawns), 1))(*spawns)
    opa = (_abi.Op * max(len(ops), 1))(*ops)
    if sources is not None:
        n = max(len(spawns), 1)
        src = (_abi.SpawnSource * n)(*[x if x is not None else _abi.SpawnSource() for x in sources])
        keep, st = [], (SourceState * n)()
        for i, ss in enumerate(source_states or []):
            if ss is None:
                continue
            arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in ss[:3]]
            keep.append(arrs)
            st[i].P, st[i].V, st[i].RC, st[i].chunk_size = arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, int(ss[3])
        rc = lib().orc_particles_step_sources(_ptr(P_), _ptr(V_), _ptr(A_), _ptr(RC), _ptr(RD), chunk_size, live, C.byref(u), C.cast(sp, P),
                                              C.cast(src, P), C.cast(st, P), len(spawns), C.cast(opa, P), len(ops), _ptr(rng_table), rw, rh,
                                              _ptr(df_tex), tw, th, steps, nthreads or threads())
    else:
        rc = lib().orc_particles_step(_ptr(P_), _ptr(V_), _ptr(A_), _ptr(RC), _ptr(RD), chunk_size, live, C.byref(u), C.cast(sp, P),
                                      len(spawns), C.cast(opa, P), len(ops), _ptr(rng_table), rw, rh, _ptr(df_tex), tw, th, steps,
                                      nthreads or threads())
    if rc != 0:
        raise RuntimeError(f"orc_particles_step failed: {rc}")
    return P_, V_, A_, RC, RD


def generate_distance_field(df, obstructions, nthreads=0, base=None) -> np.ndarray:
    """df: illuminant_b200.DistanceField descriptor (host arithmetic only). Returns uint16 [TH, TW, 4].
    base: the static field (uint16 [TH, TW, 4]) of a DynamicDistanceField, or None."""
    from illuminant_b200.distance_field import pack_obstructions
    obs = pack_obstructions(obstructions)
    saved = df.ValidSliceCount
    df.ValidSliceCount = df.SliceCount
    u = df.uniforms()
    df.ValidSliceCount = saved
    out = np.zeros((df.TextureHeight, df.TextureWidth, 4), dtype=np.uint16)
    if base is not None:
        base = np.ascontiguousarray(base, dtype=np.uint16)
    rc = lib().orc_generate_distance_field(_ptr(out), _ptr(base), df.TextureWidth, df.TextureHeight, df.SliceWidth, df.SliceHeight, df.SliceCount,
                                           C.byref(u), C.cast(obs, P) if len(obstructions) else None, len(obstructions),
                                           nthreads or threads())
    if rc != 0:
        raise RuntimeError(f"orc_generate_distance_field failed: {rc}")
    return out


def update_distance_field_slices(tex, df, obstructions, height_volumes, first_physical: int, physical_count: int, base=None, nthreads=0) -> np.ndarray:
    """RenderDistanceFieldSliceTriplet for physical slices [first, first + count) IN PLACE on `tex` (uint16 [TH, TW, 4]), with
    height volumes; `base`: the static field of a DynamicDistanceField or None."""
    from illuminant_b200.distance_field import pack_height_volumes, pack_obstructions
    obs = pack_obstructions(obstructions)
    vols, nv, edges, ne = pack_height_volumes(height_volumes)
    saved = df.ValidSliceCount
    df.ValidSliceCount = df.SliceCount
    u = df.uniforms()
    df.ValidSliceCount = saved
    assert tex.dtype == np.uint16 and tex.flags["C_CONTIGUOUS"] and tex.shape == (df.TextureHeight, df.TextureWidth, 4)
    if base is not None:
        base = np.ascontiguousarray(base, dtype=np.uint16)
    L = lib()
    L.orc_update_distance_field_slices.restype = C.c_int
    L.orc_update_distance_field_slices.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DFUniforms), P, C.c_int, P, C.c_int, P,
                                                   C.c_int, C.c_int, C.c_int, C.c_int]
    rc = L.orc_update_distance_field_slices(_ptr(tex), _ptr(base), df.TextureWidth, df.TextureHeight, df.SliceWidth, df.SliceHeight, df.SliceCount,
                                            C.byref(u), C.cast(obs, P) if len(obstructions) else None, len(obstructions),
                                            C.cast(vols, P) if nv else None, nv, C.cast(edges, P) if ne else None, ne, first_physical,
                                            physical_count, nthreads or threads())
    if rc != 0:
        raise RuntimeError(f"orc_update_distance_field_slices failed: {rc}")
    return tex


def height_volume_distance(volume, x: float, y: float, z: float) -> float:
    """finalEval(z, zRange, sdPolygon(xy)) of Shaders/DistanceField.fx for one SimpleHeightVolume."""
    from illuminant_b200.distance_field import pack_height_volumes
    vols, nv, edges, ne = pack_height_volumes([volume])
    L = lib()
    L.orc_height_volume_distance.restype = C.c_float
    L.orc_height_volume_distance.argtypes = [P, P, C.c_float, C.c_float, C.c_float]
    return float(L.orc_height_volume_distance(C.cast(vols, P), C.cast(edges, P), x, y, z))


def encode_gbuffer_sample(normal, relative_y, z, dead=False, enable_shadows=True, fullbright=False) -> np.ndarray:
    n = _f3(normal)
    out = np.zeros(4, np.float32)
    lib().orc_encode_gbuffer_sample(_ptr(n), relative_y, z, int(dead), int(enable_shadows), int(fullbright), _ptr(out))
    return out


def detmath(function: str, x: np.ndarray) -> np.ndarray:
    """The host build of include/ilb_detmath.h: function in ("sin", "cos", "acos")."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_detmath({"sin": 0, "cos": 1, "acos": 2}[function], _ptr(x), _ptr(out), x.size)
    return out


def float_to_half(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().orc_float_to_half(_ptr(a), _ptr(out), a.size)
    return out.view(np.float16)


def resolve_lighting(params, lightmap, albedo=None) -> np.ndarray:
    """Resolve.fx on fp32-decoded texels: lightmap [H, W, 4] (any float dtype, or uint8 UNORM), albedo likewise or None.
    Returns the un-quantised float32 [H, W, 4] pixel-shader output."""
    def decode(a):
        if a is None:
            return None
        a = np.asarray(a)
        if a.dtype == np.uint8:
            return np.ascontiguousarray(a.astype(np.float32) / np.float32(255.0))
        return np.ascontiguousarray(a, dtype=np.float32)
    lm, al = decode(lightmap), decode(albedo)
    out = np.empty((params.height, params.width, 4), dtype=np.float32)
    rc = lib().orc_resolve_lighting(C.byref(params), _ptr(lm), _ptr(al), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_resolve_lighting failed: {rc}")
    return out


def set_dithering(settings=None) -> None:
    """ApplyDither settings (an _abi.Dithering, or None for the handler's default) of every later resolve_* call."""
    L = lib()
    L.orc_set_dithering.restype = None
    L.orc_set_dithering.argtypes = [P]
    L.orc_set_dithering(C.byref(settings) if settings is not None else None)


def _decode_texels(a):
    a = np.asarray(a)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a.astype(np.float32) / np.float32(255.0))
    return np.ascontiguousarray(a, dtype=np.float32)


def resolve_lighting_lut(params, lut, dark, bright, lightmap, albedo) -> np.ndarray:
    """LUTResolve.fx on fp32-decoded texels; dark / bright: the two ColorLUT textures [res * rows, res * res, 4]."""
    dk, br, lm, al = (_decode_texels(a) for a in (dark, bright, lightmap, albedo))
    out = np.empty((params.height, params.width, 4), dtype=np.float32)
    L = lib()
    L.orc_resolve_lighting_lut.restype = C.c_int
    L.orc_resolve_lighting_lut.argtypes = [C.POINTER(_abi.Resolve), C.POINTER(_abi.LutBlending), P, P, P, P, P]
    rc = L.orc_resolve_lighting_lut(C.byref(params), C.byref(lut), _ptr(dk), _ptr(br), _ptr(lm), _ptr(al), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_resolve_lighting_lut failed: {rc}")
    return out


def resolve_lighting_placed(params, placement, lightmap, albedo, target) -> np.ndarray:
    """ResolveLighting as a quad at placement.Position with placement.Scale: float32 arrays; returns the updated copy of `target`."""
    lm = np.ascontiguousarray(lightmap, dtype=np.float32)
    al = np.ascontiguousarray(albedo, dtype=np.float32) if albedo is not None else None
    out = np.array(target, dtype=np.float32, copy=True, order="C")
    L = lib()
    L.orc_resolve_lighting_placed.restype = C.c_int
    L.orc_resolve_lighting_placed.argtypes = [C.POINTER(_abi.Resolve), C.POINTER(_abi.ResolvePlacement), P, P, P]
    rc = L.orc_resolve_lighting_placed(C.byref(params), C.byref(placement), _ptr(lm), _ptr(al), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_resolve_lighting_placed failed: {rc}")
    return out


def compute_luminance(lightmap, level: int) -> np.ndarray:
    lm = np.asarray(lightmap)
    lm = np.ascontiguousarray(lm.astype(np.float32) / np.float32(255.0)) if lm.dtype == np.uint8 else np.ascontiguousarray(lm, dtype=np.float32)
    h, w = lm.shape[0], lm.shape[1]
    out = np.empty(((h // 2) >> level, (w // 2) >> level), dtype=np.float32)
    rc = lib().orc_compute_luminance(_ptr(lm), w, h, level, _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_compute_luminance failed: {rc}")
    return out


def particles_render(P_, RD, RC, params, texture=None, target=None) -> np.ndarray:
    """ParticleSystem.Render in the reference's form (one quad per particle in draw order).  P_, RD, RC: float32 [n, 4];
    texture: uint8 [H, W, 4] or None; target: float32 [H, W, 4] blended over when params.clear == 0.  Returns float32 [H, W, 4]."""
    P_, RD, RC = (np.ascontiguousarray(a, dtype=np.float32) for a in (P_, RD, RC))
    out = np.zeros((params.height, params.width, 4), dtype=np.float32) if target is None else np.array(target, dtype=np.float32, copy=True, order="C")
    if texture is not None:
        texture = np.ascontiguousarray(texture, dtype=np.uint8)
    rc = lib().orc_particles_render(_ptr(P_), _ptr(RD), _ptr(RC), P_.shape[0], C.byref(params), _ptr(texture), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_particles_render failed: {rc}")
    return out

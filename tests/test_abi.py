"""The C-ABI library loads and exports exactly what include/illuminant_b200.h declares; ctypes mirrors match the C
struct layouts.  No compute calls (no GPU here)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "illuminant_b200.h"


@pytest.fixture(scope="module")
def lib():
    from illuminant_b200 import _abi, build
    build.build()
    return _abi.load_library()


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"ILB_API\s+[\w\s\*]+?\b(ilb_\w+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported(lib):
    from illuminant_b200 import _abi
    declared = declared_symbols()
    assert len(declared) >= 25
    assert sorted(_abi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (ilb_\w+)", out)))
    assert exported == declared     # nothing else leaks (built with -fvisibility=hidden)


def test_abi_version_and_error_without_device(lib):
    from illuminant_b200 import _abi
    assert lib.ilb_abi_version() == 1
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = lib.ilb_create(0, C.byref(h))
        assert rc == _abi.ERR_NO_DEVICE and not h.value      # fails loudly: no CPU fallback
        assert b"no CPU fallback" in lib.ilb_last_error(None)
        with pytest.raises(_abi.IlluminantError):
            _abi.Context(0)


def test_struct_layouts_match_the_header(tmp_path):
    from illuminant_b200 import _abi
    names = {"ilb_float4": _abi.Float4, "ilb_df_uniforms": _abi.DFUniforms, "ilb_obstruction": _abi.Obstruction,
             "ilb_light_vertex": _abi.LightVertex, "ilb_light_batch": _abi.LightBatch, "ilb_lighting_frame": _abi.LightingFrame,
             "ilb_bezier1": _abi.Bezier1, "ilb_bezier4": _abi.Bezier4, "ilb_psys_uniforms": _abi.PsysUniforms, "ilb_area": _abi.Area,
             "ilb_gravity": _abi.GravityOp, "ilb_noise": _abi.NoiseOp, "ilb_fma": _abi.FMAOp, "ilb_matrix_multiply": _abi.MatrixOp,
             "ilb_op": _abi.Op, "ilb_spawn": _abi.Spawn, "ilb_resolve": _abi.Resolve, "ilb_spawn_source": _abi.SpawnSource, "ilb_particle_render": _abi.ParticleRender,
             "ilb_dithering": _abi.Dithering, "ilb_lut_blending": _abi.LutBlending}
    probes = {"ilb_lighting_frame": ["ClearColor", "ViewportPosition", "stencil_culling"], "ilb_psys_uniforms": ["CollisionField", "has_collision_field"],
              "ilb_spawn": ["AttributeDiscardThreshold", "PositionMatrix"], "ilb_noise": ["VelocityScale", "RandomnessTexel"], "ilb_op": ["u"],
              "ilb_light_batch": ["df"], "ilb_spawn_source": ["positions", "source_system", "source_chunk", "SourceLifeRange", "pattern_texels", "StepWidthAndSizeScale", "CenteringOffset"], "ilb_particle_render": ["ClearColor", "RoundingPowerFromLife", "ViewportScale", "StippleFactor"], "ilb_resolve": ["InverseScaleFactor", "WhitePoint", "DitheringStrength"], "ilb_gravity": ["AttractorRadiusesAndStrengths"],
              "ilb_dithering": ["RangeMax"], "ilb_lut_blending": ["BrightLevel", "LUTOffsets"]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for n in names:
        src.append(f'printf("{n} %zu\\n", sizeof({n}));')
        for f in probes.get(n, []):
            src.append(f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));')
    src.append('return 0;}')
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", str(c), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert C.sizeof(_abi.LightVertex) == 128       # Vertices.cs:10-39
    for n, cls in names.items():
        assert int(out[n]) == C.sizeof(cls), n
        for f in probes.get(n, []):
            assert int(out[f"{n}.{f}"]) == getattr(cls, f).offset, f"{n}.{f}"


def test_cuda_library_is_sm100a_with_lineinfo():
    from illuminant_b200 import _abi
    out = subprocess.run(["cuobjdump", "-lelf", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_csharp_shim_binds_every_symbol_with_matching_struct_sizes():
    """csharp/IlluminantB200.cs cannot be compiled here (no .NET), so it is checked textually: every symbol the header declares
    has a DllImport, and every blittable struct it declares has the size of the matching C struct (summing the field sizes of
    the Sequential, Pack = 4 layouts: int / float 4, Vector2 8, Vector3 12, Vector4 16, Matrix 64, long 8, `fixed float X[n]`)."""
    from illuminant_b200 import _abi
    cs = (ROOT / "csharp" / "IlluminantB200.cs").read_text()
    imported = set(re.findall(r"public static extern \w+\*? (ilb_\w+) ?\(", cs))
    assert imported == set(declared_symbols()), sorted(set(declared_symbols()) ^ imported)
    sizes = {"int": 4, "float": 4, "long": 8, "Vector2": 8, "Vector3": 12, "Vector4": 16, "Matrix": 64, "IntPtr": 8,
             "Uniforms.ClampedBezier1": C.sizeof(_abi.Bezier1), "Uniforms.ClampedBezier4": C.sizeof(_abi.Bezier4)}
    mirrors = {"IlbDFUniforms": _abi.DFUniforms, "IlbLightBatch": _abi.LightBatch, "IlbLightingFrame": _abi.LightingFrame,
               "IlbObstruction": _abi.Obstruction, "IlbPsysUniforms": _abi.PsysUniforms, "IlbArea": _abi.Area, "IlbGravity": _abi.GravityOp,
               "IlbNoise": _abi.NoiseOp, "IlbFMA": _abi.FMAOp, "IlbMatrixMultiply": _abi.MatrixOp, "IlbSpawn": _abi.Spawn,
               "IlbHeightVolume": _abi.HeightVolumeStruct, "IlbResolvePlacement": _abi.ResolvePlacement,
               "IlbDithering": _abi.Dithering, "IlbLutBlending": _abi.LutBlending}
    checked = 0
    for name, body in re.findall(r"public (?:unsafe )?struct (\w+) \{(.*?)\n    \}|public (?:unsafe )?struct (\w+) \{(.*?)\n        \}", cs, flags=re.S) and \
            [(m[0] or m[2], m[1] or m[3]) for m in re.findall(r"public (?:unsafe )?struct (\w+) \{(.*?)\n    \}|public (?:unsafe )?struct (\w+) \{(.*?)\n        \}", cs, flags=re.S)]:
        if name not in mirrors:
            continue
        total = 0
        for decl in re.findall(r"public ([^;()]+);", body):
            decl = decl.split("//")[0].strip()
            m = re.match(r"fixed (\w+) \w+\[([^\]]+)\]", decl)
            if m:
                total += sizes[m.group(1)] * eval(m.group(2))
                continue
            typ, names = decl.split(" ", 1)
            if typ in sizes or typ in mirrors:
                unit = sizes[typ] if typ in sizes else C.sizeof(mirrors[typ])
                total += unit * len([n for n in names.split(",") if n.strip()])
            else:
                raise AssertionError(f"{name}: unknown field type {typ!r}")
        assert total == C.sizeof(mirrors[name]), f"{name}: C# fields sum to {total} bytes, the C struct has {C.sizeof(mirrors[name])}"
        checked += 1
    assert checked >= 11, checked

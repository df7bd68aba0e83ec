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
             "ilb_op": _abi.Op, "ilb_spawn": _abi.Spawn, "ilb_resolve": _abi.Resolve, "ilb_spawn_source": _abi.SpawnSource, "ilb_particle_render": _abi.ParticleRender}
    probes = {"ilb_lighting_frame": ["ClearColor", "ViewportPosition", "stencil_culling"], "ilb_psys_uniforms": ["CollisionField", "has_collision_field"],
              "ilb_spawn": ["AttributeDiscardThreshold", "PositionMatrix"], "ilb_noise": ["VelocityScale", "RandomnessTexel"], "ilb_op": ["u"],
              "ilb_light_batch": ["df"], "ilb_spawn_source": ["positions", "source_system", "source_chunk", "SourceLifeRange", "pattern_texels", "StepWidthAndSizeScale", "CenteringOffset"], "ilb_particle_render": ["ClearColor", "RoundingPowerFromLife", "ViewportScale", "StippleFactor"], "ilb_resolve": ["InverseScaleFactor", "WhitePoint", "DitheringStrength"], "ilb_gravity": ["AttractorRadiusesAndStrengths"]}
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

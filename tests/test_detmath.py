"""include/ilb_detmath.h sits on BOTH sides of every parity comparison (the kernels and the CPU oracle include the same
header), so a bug in it would be invisible to the parity tests.  These tests pin it independently:

  * the host build against float64 libm over the ranges the paths use (G-buffer normal decode: [-pi, pi]; spawner
    Spherical formulas: theta in [0, pi], phi in [0, 2 pi]; collision escape vector: [0, 47.2]; bezier sine mode:
    [0, pi / 2]; line-light solid angle: acos on [-1, 1]) with the error bounds the header documents;
  * known values and symmetries;
  * (GPU) the device build, through ilb_debug_detmath, bit for bit against the host build -- the two sides compared
    directly and not only "by construction".
"""
import ctypes as C

import numpy as np
import pytest

SIN_COS_ABS = 1.0e-7     # |dm_sinf(x) - sin(x)|, |dm_cosf(x) - cos(x)| for |x| <= 8192 (measured 7.7e-8)
ACOS_ABS = 5.0e-7        # |dm_acosf(x) - acos(x)| on [-1, 1] (measured 4.3e-7 = 1.8 ulp of pi)
ACOS_ULP = 3.0


def sweep(lo, hi, n, seed):
    rs = np.random.RandomState(seed)
    x = rs.uniform(lo, hi, n).astype(np.float32)
    edges = np.array([lo, hi, 0.0, -0.0, np.float32(np.pi), np.float32(np.pi / 2), np.float32(np.pi / 4), 1e-8, -1e-8, 1.0, -1.0], np.float32)
    return np.concatenate([x, edges[(edges >= lo) & (edges <= hi)], np.linspace(lo, hi, 4097, dtype=np.float32)])


@pytest.mark.parametrize("lo,hi", [(-np.pi, np.pi), (0.0, 2 * np.pi), (0.0, 47.2), (-64.0, 64.0), (-8192.0, 8192.0)])
def test_sin_cos_against_float64_libm(oracle, lo, hi):
    x = sweep(np.float32(lo), np.float32(hi), 400_000, 1)
    x64 = x.astype(np.float64)
    s, c = oracle.detmath("sin", x).astype(np.float64), oracle.detmath("cos", x).astype(np.float64)
    assert np.abs(s - np.sin(x64)).max() <= SIN_COS_ABS
    assert np.abs(c - np.cos(x64)).max() <= SIN_COS_ABS
    assert np.abs(s * s + c * c - 1.0).max() <= 4e-7          # the pair stays on the unit circle


def test_acos_against_float64_libm(oracle):
    x = sweep(np.float32(-1.0), np.float32(1.0), 1_000_000, 2)
    a = oracle.detmath("acos", x).astype(np.float64)
    ref = np.arccos(x.astype(np.float64))
    err = np.abs(a - ref)
    assert err.max() <= ACOS_ABS
    assert (err / np.spacing(np.maximum(ref, 1e-3).astype(np.float32)).astype(np.float64)).max() <= ACOS_ULP
    assert np.isnan(oracle.detmath("acos", np.array([1.0000001, -1.5], np.float32))).all()       # like acos(): NaN outside [-1, 1]


def test_known_values_and_symmetries(oracle):
    f = lambda name, v: float(oracle.detmath(name, np.array([v], np.float32))[0])
    assert f("sin", 0.0) == 0.0 and f("cos", 0.0) == 1.0
    assert f("acos", 1.0) == 0.0
    assert abs(f("acos", -1.0) - np.pi) <= 2.4e-7 and abs(f("acos", 0.0) - np.pi / 2) <= 1.2e-7
    assert abs(f("sin", np.float32(np.pi / 2)) - 1.0) <= 6e-8 and abs(f("cos", np.float32(np.pi))) - 1.0 <= 6e-8
    x = sweep(np.float32(0.0), np.float32(100.0), 100_000, 3)
    assert np.array_equal(oracle.detmath("sin", -x), -oracle.detmath("sin", x))      # odd, exactly
    assert np.array_equal(oracle.detmath("cos", -x), oracle.detmath("cos", x))       # even, exactly
    xa = sweep(np.float32(0.0), np.float32(1.0), 100_000, 4)
    # acos(-x) = pi - acos(x) is how the reflection is computed: one rounding apart at most
    d = np.abs(oracle.detmath("acos", -xa).astype(np.float64) - (np.float64(np.float32(np.pi)) - oracle.detmath("acos", xa).astype(np.float64)))
    assert d.max() <= 2.4e-7


@pytest.mark.gpu
@pytest.mark.parametrize("name,fn,lo,hi", [("sin", 0, -8192.0, 8192.0), ("cos", 1, -8192.0, 8192.0), ("sin", 0, -7.0, 7.0), ("cos", 1, -7.0, 7.0),
                                            ("acos", 2, -1.0, 1.0)])
def test_device_build_is_bit_identical_to_host_build(ctx, oracle, name, fn, lo, hi):
    x = sweep(np.float32(lo), np.float32(hi), 2_000_000, 5 + fn)
    if name != "acos":   # large arguments too: both builds must agree even where the function is inaccurate.  (Beyond 1.6e9 the
        # octant index overflows an int -- undefined in C, saturating on the device -- and the kernels flush denormals; neither
        # range is reachable on the paths: angles are bounded by 2 pi, 47.2 and pi / 2, see the module docstring.)
        x = np.concatenate([x, np.array([1e5, -3e6, 1e9, -1.5e9], np.float32)])
    else:
        x = np.concatenate([x, np.array([1.5, -1.5, np.nan], np.float32)])
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    ctx.check(ctx.lib.ilb_debug_detmath(ctx.handle, fn, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), x.size))
    host = oracle.detmath(name, x)
    same = (out.view(np.uint32) == host.view(np.uint32)) | (np.isnan(out) & np.isnan(host))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, f"{name}: {bad.size} of {x.size} differ, first x={x[bad[0]]!r}: device {out[bad[0]]!r} host {host[bad[0]]!r}"

#!/usr/bin/env python
"""Compares two raw float32 lightmaps (width x height x 4): the oracle's and the one CrossCheckScene.cs dumped from the reference.

    python tools/crosscheck/compare.py oracle_lightmap.f32 reference_lightmap.f32 W H [rtol]"""
import sys

import numpy as np

a_path, b_path, w, h = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rtol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-4
a = np.fromfile(a_path, dtype=np.float32).reshape(h, w, 4)[..., :3].astype(np.float64)
b = np.fromfile(b_path, dtype=np.float32).reshape(h, w, 4)[..., :3].astype(np.float64)
err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
worst = np.unravel_index(np.argmax(err), err.shape)
print(f"max relative error {err.max():.3e} at (y, x, channel) = {worst}: {a[worst]:.6f} against {b[worst]:.6f}")
print(f"pixels beyond {rtol:g}: {(err.max(axis=2) > rtol).sum()} of {w * h} ({100.0 * (err.max(axis=2) > rtol).mean():.4f} %)")
print(f"percentiles of the per-pixel maximum error: 50 % {np.percentile(err.max(axis=2), 50):.2e}, 99 % {np.percentile(err.max(axis=2), 99):.2e}, "
      f"99.99 % {np.percentile(err.max(axis=2), 99.99):.2e}")
sys.exit(0 if err.max() <= rtol else 1)

"""Shared helpers of the parity tests: build a renderer from a synthetic scene and run both the CUDA path (through the
C-ABI) and the CPU oracle on identical inputs."""
from __future__ import annotations

import numpy as np

import illuminant_b200 as ib
from illuminant_b200 import scenes

LIGHTING_RTOL = 1e-4      # north_star: 1e-4 relative per channel
LIGHTING_FLOOR = 1e-3     # |ref| floor of the relative error (values are O(0.05..4))
PARTICLE_ATOL = 1e-5      # north_star: 1e-5 absolute on position / velocity (scaled by magnitude, see rel_err_particles)


def lighting_rel_err(gpu: np.ndarray, ref: np.ndarray) -> np.ndarray:
    return np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / np.maximum(np.abs(ref.astype(np.float64)), LIGHTING_FLOOR)


def particle_err(gpu: np.ndarray, ref: np.ndarray) -> float:
    """Absolute error in units of 1e-5 at unit scale: positions are O(1e3) px where one fp32 ulp is 6e-5, so the
    tolerance scales with the magnitude of the compared field (1e-5 * max(1, |ref|))."""
    scale = np.maximum(1.0, np.abs(ref.astype(np.float64)))
    return float((np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / scale).max())


def make_renderer(ctx, scene, oracle=None, generate_on_gpu=True):
    """Returns (renderer, df_texels) with the distance field rasterised and the G-buffer uploaded."""
    df = scenes.make_distance_field(ctx, scene)
    if generate_on_gpu:
        df.Rasterize(scene.obstructions)
        tex = df.Save()
    else:
        tex = oracle.generate_distance_field(df, scene.obstructions)
        df.Load(tex)
    r = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
    r.DistanceField = df
    r.Probes = scene.probes
    r.SetGBuffer(scene.gbuffer if scene.configuration.EnableGBuffer else None)
    return r, tex


def oracle_lightmap(oracle, renderer, tex, scene, rows=None, intensity=1.0):
    frame = renderer.build_frame(intensity, rows)
    batches, nb, verts, nv = renderer.build_batches(intensity)
    gb = scene.gbuffer if scene.configuration.EnableGBuffer else None
    if gb is not None and not scene.configuration.HighQualityGBuffer:
        gb = gb.astype(np.float16)
    return oracle.render_lighting(tex if renderer.DistanceField is not None else None, gb, frame, batches, nb, verts, nv)

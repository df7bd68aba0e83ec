"""Shared helpers of the parity tests: build a renderer from a synthetic scene and run both the CUDA path (through the
C-ABI) and the CPU oracle on identical inputs."""
from __future__ import annotations

import numpy as np

import illuminant_b200 as ib
from illuminant_b200 import scenes

LIGHTING_RTOL = 1e-4      # north_star: 1e-4 relative per channel
LIGHTING_FLOOR = 1e-3     # |ref| floor of the relative error (values are O(0.05..4))
PARTICLE_ATOL = 1e-5      # render outputs (renderColor / renderData): 1e-5, relative to max(1, |ref|) -- see particle_err


def lighting_rel_err(gpu: np.ndarray, ref: np.ndarray) -> np.ndarray:
    return np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / np.maximum(np.abs(ref.astype(np.float64)), LIGHTING_FLOOR)


def particle_err(gpu: np.ndarray, ref: np.ndarray) -> float:
    """For the RENDER OUTPUTS only (renderColor, renderData: plain fp32 arithmetic with FMA contraction, library pow / atan2,
    never fed back into particle state).  renderData.y carries `index * RotationFromIndex` (up to 1e5), where one fp32 ulp is
    8e-3, so the 1e-5 bound is taken relative to max(1, |ref|).  Particle STATE is held to exact equality instead, see
    assert_particle_state_equal."""
    scale = np.maximum(1.0, np.abs(ref.astype(np.float64)))
    return float((np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / scale).max())


def assert_particle_state_equal(gpu: np.ndarray, ref: np.ndarray, what: str = ""):
    """Particle state (PositionAndLife, Velocity, Attributes): north_star allows 1e-5 absolute, but positions are O(1e3) px
    where one fp32 ulp is 6e-5 -- the bound can only be met by identical values, and the state chain is built from IEEE
    operations in the oracle's order precisely so that it is.  Asserted as exact equality (NaN == NaN, -0 == +0); a failure
    reports how many values differ and the largest absolute difference."""
    g, r = np.asarray(gpu), np.asarray(ref)
    assert g.shape == r.shape, f"{what}: shape {g.shape} vs {r.shape}"
    same = (g == r) | (np.isnan(g) & np.isnan(r))
    if not same.all():
        d = np.abs(g.astype(np.float64) - r.astype(np.float64))
        d = d[~same & np.isfinite(d)]
        raise AssertionError(f"{what}: {int((~same).sum())} of {g.size} values differ from the oracle, max abs difference "
                             f"{(d.max() if d.size else float('nan')):.3e}")


def check_particles(gpu, ref, what: str = "", allow_nan: bool = False):
    """(P, V, A, RC, RD) of the CUDA path against the oracle: state exactly, render outputs within PARTICLE_ATOL."""
    for g, r, n in zip(gpu, ref, ("position", "velocity", "attributes", "renderColor", "renderData")):
        if not allow_nan:
            assert not np.isnan(g).any(), f"{what} {n} NaN"
        if n in ("position", "velocity", "attributes"):
            assert_particle_state_equal(g, r, f"{what} {n}")
        else:
            assert np.array_equal(np.isfinite(g), np.isfinite(r)), f"{what} {n}"
            ok = np.isfinite(r)
            e = particle_err(g[ok], r[ok]) if ok.any() else 0.0
            assert e <= PARTICLE_ATOL, f"{what} {n}: err {e:.3e}"


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

"""Randomised parity of the lighting path: many small scenes whose every parameter is drawn at random -- light types and
their properties, quality settings, G-buffer contents (tilted / flat / down-facing / missing normals, heights, shadow flags,
fullbright and dead texels), 2.5D, viewport position and scale, render scale, light occlusion, stencil culling, half or
Vector4 G-buffer, distance field present or not, obstructions that cover lights -- rendered by the CUDA path through the C-ABI
and by the CPU oracle.  The fast paths of the kernels (short sampler inside the volume, two-loop march, floors through the
round-down adder, flat-normal forms, constant-bank light records, deferred range guards with the IEEE fallback) all have to
land on the oracle's values: 1e-4 relative per channel, light counts (alpha) exactly."""
import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from helpers import LIGHTING_RTOL, lighting_rel_err, make_renderer, oracle_lightmap

pytestmark = pytest.mark.gpu


def _random_scene(seed):
    rs = np.random.RandomState(1000 + seed)
    w, h = int(rs.choice([48, 64, 80, 112])), int(rs.choice([32, 48, 64]))
    s = scenes.lighting_scene(200 + seed, w, h, 0, float4_lightmap=True)
    cfg, env = s.configuration, s.environment
    cfg.TwoPointFiveD = bool(rs.rand() < 0.3)
    env.ZToYMultiplier = float(rs.uniform(0.5, 2.5))
    cfg.LightOcclusion = float(rs.choice([0.0, 0.0, 25.0]))
    cfg.StencilCulling = bool(rs.rand() < 0.3)
    cfg.HighQualityGBuffer = bool(rs.rand() < 0.7)
    cfg.ScaleCompensation = bool(rs.rand() < 0.7)
    if rs.rand() < 0.3:
        cfg.DefaultQuality = ib.RendererQualitySettings(MinStepSize=float(rs.uniform(1.0, 4.0)), LongStepFactor=float(rs.uniform(0.4, 1.0)),
                                                        MaxStepCount=int(rs.choice([6, 24, 64])), MaxConeRadius=float(rs.uniform(4, 24)),
                                                        OcclusionToOpacityPower=float(rs.choice([1.0, 0.7, 1.6])))
    # G-buffer
    z = rs.uniform(0, 50, (h, w)).astype(np.float32) * (rs.rand(h, w) < 0.5)
    n = np.zeros((h, w, 3), np.float32)
    n[..., 2] = 1
    v = rs.normal(size=(h, w, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    tilt = rs.rand(h, w) < rs.uniform(0.0, 0.6)
    n[tilt] = v[tilt]
    n[rs.rand(h, w) < 0.05] = np.array([0, 0, -1], np.float32)
    n[rs.rand(h, w) < 0.05] = 0
    es = rs.rand(h, w) > 0.2
    fb = rs.rand(h, w) < 0.05
    dead = rs.rand(h, w) < 0.03
    rel_y = (rs.uniform(-3, 3, (h, w)) * (rs.rand(h, w) < 0.2)).astype(np.float32)
    s.gbuffer = ib.encode_gbuffer(n, rel_y, z, es, fb, dead)
    # lights
    lights = []
    for _ in range(int(rs.randint(1, 5))):
        l = ib.SphereLightSource(Position=(float(rs.uniform(-10, w + 10)), float(rs.uniform(-10, h + 10)), float(rs.uniform(-5, 100))),
                                 Radius=float(rs.uniform(1, 30)), RampLength=float(rs.uniform(5, 150)),
                                 RampMode=int(rs.choice([ib.LightSourceRampMode.Linear, ib.LightSourceRampMode.Exponential, ib.LightSourceRampMode.None_])),
                                 Color=(float(rs.rand()), float(rs.rand()), float(rs.rand()), float(rs.uniform(0.3, 1.5))),
                                 CastsShadows=bool(rs.rand() < 0.8))
        l.FalloffYFactor = float(rs.choice([1.0, 1.0, 0.5, 2.0]))
        if rs.rand() < 0.4:
            l.AmbientOcclusionRadius, l.AmbientOcclusionOpacity = float(rs.uniform(2, 20)), float(rs.uniform(0.2, 1.0))
        if rs.rand() < 0.3:
            l.SpecularColor, l.SpecularPower = (float(rs.rand()), float(rs.rand()), float(rs.rand())), float(rs.uniform(1, 8))
        if rs.rand() < 0.3:
            l.ShadowDistanceFalloff = float(rs.uniform(10, 60))
        l.ShadowFilter = int(rs.choice([ib.ShadowFilter.None_, ib.ShadowFilter.None_, ib.ShadowFilter.Shadowed, ib.ShadowFilter.Unshadowed]))
        lights.append(l)
    for _ in range(int(rs.randint(0, 3))):
        d = ib.DirectionalLightSource(Color=(float(rs.rand()), float(rs.rand()), float(rs.rand()), 1.0), CastsShadows=bool(rs.rand() < 0.8))
        d.Direction = (rs.uniform(-1, 1), rs.uniform(-1, 1), -rs.uniform(0.05, 1.0))
        d.ShadowTraceLength, d.ShadowSoftness, d.ShadowRampRate = float(rs.uniform(20, 300)), float(rs.uniform(1, 30)), float(rs.uniform(0.1, 1.0))
        if rs.rand() < 0.3:
            d.Bounds = ((float(rs.uniform(0, w / 2)), float(rs.uniform(0, h / 2))), (float(rs.uniform(w / 2, w)), float(rs.uniform(h / 2, h))))
        if rs.rand() < 0.3:
            d.AmbientOcclusionRadius, d.AmbientOcclusionOpacity = float(rs.uniform(2, 20)), float(rs.uniform(0.2, 1.0))
        lights.append(d)
    for _ in range(int(rs.randint(0, 3))):
        x0, y0, ang, length = rs.uniform(0, w), rs.uniform(0, h), rs.uniform(0, 2 * np.pi), rs.uniform(5, 120)
        c0 = (float(rs.rand()), float(rs.rand()), float(rs.rand()), float(rs.uniform(0.1, 0.6)))
        c1 = (float(rs.rand()), float(rs.rand()), float(rs.rand()), float(rs.uniform(0.1, 0.6)))
        lights.append(ib.LineLightSource(StartPosition=(float(x0), float(y0), float(rs.uniform(0, 60))),
                                         EndPosition=(float(x0 + np.cos(ang) * length), float(y0 + np.sin(ang) * length), float(rs.uniform(0, 60))),
                                         Radius=float(rs.uniform(1, 20)), StartColor=c0, EndColor=c1, CastsShadows=bool(rs.rand() < 0.8)))
    order = rs.permutation(len(lights))
    env.Lights = [lights[i] for i in order]     # types interleave in draw order: several batches, several passes
    return s, rs


import os

# ILB_FUZZ_SEEDS=n widens the sweep (the default keeps the suite short)
@pytest.mark.parametrize("seed", range(int(os.environ.get("ILB_FUZZ_SEEDS", "24"))))
def test_random_scene_matches_the_oracle(ctx, oracle, seed):
    s, rs = _random_scene(seed)
    no_field = rs.rand() < 0.15
    r, tex = make_renderer(ctx, s)
    if no_field:
        r.DistanceField = None
    if rs.rand() < 0.4:
        r.ViewportPosition = (float(rs.uniform(-8, 8)), float(rs.uniform(-8, 8)))
    if rs.rand() < 0.3:
        sc = float(rs.choice([0.75, 1.25, 1.5]))
        r.ViewportScale = (sc, sc)
    gpu = r.RenderLighting()
    ref = oracle_lightmap(oracle, r, tex, s)
    assert np.array_equal(np.isnan(gpu), np.isnan(ref))
    err = lighting_rel_err(np.nan_to_num(gpu), np.nan_to_num(ref))
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= LIGHTING_RTOL, f"seed {seed}: max rel err {err.max():.3e} at {worst}: gpu {gpu[worst]} ref {ref[worst]}"
    assert np.array_equal(gpu[..., 3], ref[..., 3]), f"seed {seed}: {int((gpu[..., 3] != ref[..., 3]).sum())} pixels with a different light count"
    # the same frame without the constant-bank light records: another instantiation of the kernel, whose smooth factors (AO,
    # specular: plain operators, contracted as the compiler sees fit) may differ in the last bits -- light counts may not
    from illuminant_b200 import _abi
    ctx.set_option(_abi.OPT_LIGHT_CONST_BANK, 0)
    try:
        other = r.RenderLighting()
    finally:
        ctx.set_option(_abi.OPT_LIGHT_CONST_BANK, 1)
    assert np.array_equal(other[..., 3], gpu[..., 3])
    d = lighting_rel_err(np.nan_to_num(other), np.nan_to_num(gpu))
    assert d.max() <= 2e-6, f"seed {seed}: constant bank on / off differ by {d.max():.3e}"
    # ... and as two row bands (the same instantiation): the same bits
    h = gpu.shape[0]
    cut = (h // 2 + 15) // 16 * 16
    halves = np.concatenate([r.RenderLighting(rows=(0, cut)), r.RenderLighting(rows=(cut, h))], axis=0)
    assert np.array_equal(halves, gpu, equal_nan=True)

"""Synthetic scenes of the shapes BASELINE.json names: deterministic functions of a seed
(numpy RandomState == MT19937, seed 0xB200 + config index).  Used by bench.py, smoke() and the parity tests; they
only build reference-API objects (environment, lights, obstructions, particle systems) -- no device work here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .distance_field import DistanceField, LightObstruction, LightObstructionType, RendererQualitySettings
from .lighting import (DirectionalLightSource, LightingEnvironment, LightProbe, LightSourceRampMode, LineLightSource,
                       RendererConfiguration, SphereLightSource, encode_gbuffer)
from .particles import (FMA, AreaType, Attractor, AttractorType, Formula, FormulaType, Gravity, Noise, ParticleCollision,
                        ParticleSystemConfiguration, Spawner, TransformArea)

SEED_BASE = 0xB200
F = np.float32


@dataclass
class LightingScene:
    width: int
    height: int
    environment: LightingEnvironment
    configuration: RendererConfiguration
    obstructions: List[LightObstruction]
    gbuffer: np.ndarray            # float32 [H, W, 4], encoded
    probes: List[LightProbe]
    df_slices: int = 8
    df_depth: float = 128.0
    df_resolution: float = 1.0


def make_obstructions(rs: np.random.RandomState, width: int, height: int, count: int, big: bool = True) -> List[LightObstruction]:
    """`W*H/2^15` random Box / Ellipsoid / Cylinder obstructions, size 16-96 px, z-size 16-100."""
    obs = []
    for _ in range(count):
        typ = [LightObstructionType.Box, LightObstructionType.Ellipsoid, LightObstructionType.Cylinder][rs.randint(0, 3)]
        lo, hi = (16.0, 96.0) if big else (4.0, 24.0)
        sx, sy = rs.uniform(lo, hi) / 2, rs.uniform(lo, hi) / 2
        sz = rs.uniform(16.0, 100.0) / 2
        cx, cy = rs.uniform(0, width), rs.uniform(0, height)
        obs.append(LightObstruction(typ, (float(F(cx)), float(F(cy)), float(F(sz))), (float(F(sx)), float(F(sy)), float(F(sz)))))
    return obs


def make_gbuffer(rs: np.random.RandomState, width: int, height: int, box_count: int) -> np.ndarray:
    """Ground plane (n = +z, z = 0, shadows on) plus axis-aligned raised boxes with top-face normals, z in [8, 64]."""
    z = np.zeros((height, width), dtype=F)
    normal = np.zeros((height, width, 3), dtype=F)
    normal[..., 2] = 1.0
    for _ in range(box_count):
        w, h = int(rs.randint(16, 96)), int(rs.randint(16, 96))
        x0, y0 = int(rs.randint(0, max(width - w, 1))), int(rs.randint(0, max(height - h, 1)))
        z[y0:y0 + h, x0:x0 + w] = F(rs.uniform(8.0, 64.0))
    return encode_gbuffer(normal, np.zeros_like(z), z, True, False, False)


def make_sphere_lights(rs, width, height, count, ramp_lo, ramp_hi) -> List[SphereLightSource]:
    lights = []
    for _ in range(count):
        lights.append(SphereLightSource(
            Position=(float(F(rs.uniform(0, width))), float(F(rs.uniform(0, height))), float(F(rs.uniform(8, 96)))),
            Radius=float(F(rs.uniform(8, 32))), RampLength=float(F(rs.uniform(ramp_lo, ramp_hi))),
            RampMode=LightSourceRampMode.Linear if rs.rand() < 0.5 else LightSourceRampMode.Exponential,
            Color=(float(F(rs.uniform(0.2, 1))), float(F(rs.uniform(0.2, 1))), float(F(rs.uniform(0.2, 1))), 1.0),
            CastsShadows=True, AmbientOcclusionRadius=0.0))
    return lights


def lighting_scene(config_index: int, width: int, height: int, n_sphere: int, n_directional: int = 0, n_line: int = 0,
                   n_probes: int = 0, ramp=(150.0, 450.0), quality: Optional[RendererQualitySettings] = None,
                   ao: bool = False, float4_lightmap: bool = False) -> LightingScene:
    rs = np.random.RandomState(SEED_BASE + config_index)
    env = LightingEnvironment(GroundZ=0.0, MaximumZ=128.0, ZToYMultiplier=1.0, Ambient=(0.05, 0.05, 0.08, 1.0))
    cfg = RendererConfiguration(MaximumRenderSize=(width, height), HighQuality=True, HighQualityGBuffer=True, TwoPointFiveD=False,
                                ScaleCompensation=True, LightOcclusion=0.0, RenderScale=(1.0, 1.0), Float4Lightmap=float4_lightmap,
                                MaximumLightProbeCount=max(256, n_probes))
    if quality is not None:
        cfg.DefaultQuality = quality
    obstructions = make_obstructions(rs, width, height, max(1, (width * height) >> 15), big=min(width, height) >= 512)
    gbuffer = make_gbuffer(rs, width, height, max(1, math.ceil(width * height / 65536)))
    lights = make_sphere_lights(rs, width, height, n_sphere, *ramp)
    if ao:
        for l in lights[::2]:
            l.AmbientOcclusionRadius, l.AmbientOcclusionOpacity = 12.0, 0.7
    for _ in range(n_directional):  # defaults LightSource.cs:136-144
        d = DirectionalLightSource(Color=(float(F(rs.uniform(0.05, 0.2))),) * 3 + (1.0,), CastsShadows=True)
        d.Direction = (rs.uniform(-1, 1), rs.uniform(-1, 1), -rs.uniform(0.3, 1.0))
        lights.append(d)
    for _ in range(n_line):
        x0, y0 = rs.uniform(0, width), rs.uniform(0, height)
        ang, length = rs.uniform(0, 2 * math.pi), rs.uniform(100, 600)
        zz = rs.uniform(8, 64)
        c = (float(F(rs.uniform(0.2, 1))), float(F(rs.uniform(0.2, 1))), float(F(rs.uniform(0.2, 1))), 0.35)
        lights.append(LineLightSource(StartPosition=(float(F(x0)), float(F(y0)), float(F(zz))),
                                      EndPosition=(float(F(x0 + math.cos(ang) * length)), float(F(y0 + math.sin(ang) * length)), float(F(zz))),
                                      Radius=float(F(rs.uniform(4, 20))), StartColor=c, EndColor=c, CastsShadows=True))
    env.Lights = lights
    env.Obstructions = obstructions
    probes = []
    if n_probes:
        side = int(round(math.sqrt(n_probes)))
        for j in range(side):
            for i in range(side):
                probes.append(LightProbe(Position=((i + 0.5) * width / side, (j + 0.5) * height / side, 16.0), Normal=(0.0, 0.0, 1.0)))
    return LightingScene(width, height, env, cfg, obstructions, gbuffer, probes)


# BASELINE.json configs (lighting side)
def config_c1() -> LightingScene:  # 1 SphereLightSource, 256x256
    s = lighting_scene(0, 256, 256, 0)
    s.environment.Lights = [SphereLightSource(Position=(128.0, 128.0, 32.0), Radius=16.0, RampLength=200.0, CastsShadows=True)]
    return s


def config_c2() -> LightingScene:  # 1920x1080, 32 sphere lights, 8-slice DF
    return lighting_scene(1, 1920, 1080, 32)


def config_c4(width: int = 3840, height: int = 2160) -> LightingScene:  # 4K, 128 mixed lights + probes
    return lighting_scene(3, width, height, 96, n_directional=8, n_line=24, n_probes=256, ramp=(100.0, 400.0))


def config_c5_lighting(width: int = 3840, height: int = 2160) -> LightingScene:  # 64-light 4K lighting of the combined loop
    return lighting_scene(4, width, height, 64, ramp=(100.0, 400.0))


def make_distance_field(ctx, scene: LightingScene, resolution: Optional[float] = None) -> DistanceField:
    return DistanceField(ctx, scene.width, scene.height, scene.df_depth, scene.df_slices,
                         scene.df_resolution if resolution is None else resolution, 128)


# ---------------------------------------------------------------------------------------------------- particles
@dataclass
class ParticleScene:
    count: int
    chunk_size: int
    positions: np.ndarray   # [count, 4]
    velocities: np.ndarray
    attributes: np.ndarray
    configuration: ParticleSystemConfiguration
    transforms: list
    dt: float = 1 / 60.0


def particle_scene(config_index: int, count: int, chunk_size: int, width: int = 1920, height: int = 1080, steps_hint: int = 1000,
                   collision_field: Optional[DistanceField] = None, spawn_rate: float = 60000.0) -> ParticleScene:
    """C3 / C5 chain: Spawner (Linear position, Spherical velocity; SimpleParticles.cs:140-163) -> Gravity (4 attractors,
    :164-180) -> Noise (defaults, Velocity scale 30) -> FMA (Velocity.Multiply 0.98, Ellipsoid area) ->
    UpdateWithDistanceField (bounce 0.95, escape 256, friction 0.1, LifeDecay 1.2)."""
    rs = np.random.RandomState(SEED_BASE + config_index)
    dt = 1 / 60.0
    P = np.zeros((count, 4), dtype=F)
    P[:, 0] = rs.uniform(0, width, count)
    P[:, 1] = rs.uniform(0, height, count)
    P[:, 2] = rs.uniform(0, 64, count)
    # life so that most particles stay alive for the whole run (decay 1.2/s)
    span = 1.2 * steps_hint * dt
    P[:, 3] = rs.uniform(200, 360, count) / 60.0 * span / 3.0 + span
    ang, speed = rs.uniform(0, 2 * math.pi, count), 40.0 * np.sqrt(rs.uniform(0, 1, count))
    V = np.zeros((count, 4), dtype=F)
    V[:, 0], V[:, 1] = np.cos(ang) * speed, np.sin(ang) * speed
    A = rs.uniform(0.2, 1.0, (count, 4)).astype(F)
    cfg = ParticleSystemConfiguration(Friction=0.1, MaximumVelocity=2048.0, LifeDecayPerSecond=1.2, RotationFromVelocity=True,
                                      RotationFromLife=30.0, RotationFromIndex=2.0)
    if collision_field is not None:
        cfg.Collision = ParticleCollision(DistanceField=collision_field, DistanceFieldMaximumZ=256.0, EscapeVelocity=256.0,
                                          BounceVelocityMultiplier=0.95, Distance=0.33, LifePenalty=0.0)
    spawner = Spawner(MinRate=spawn_rate, MaxRate=spawn_rate, Seed=SEED_BASE + config_index,
                      Position=Formula(Constant=(width / 2, height / 2, 0.0), RandomScale=(width * 0.9, height * 0.9, 0.0), Offset=(-0.5, -0.5, 0.0)),
                      Velocity=Formula(Constant=(0.0, 0.0, 0.0), RandomScale=(32.0, 32.0, 0.0), Offset=(8.0, 8.0, 0.0), Type=FormulaType.Spherical),
                      Life=(span * 2, 1.0, 0.0), ColorConstant=(1.0, 1.0, 1.0, 1.0), AlphaDiscardThreshold=1.0)
    gravity = Gravity(MaximumAcceleration=8.0, Attractors=[
        Attractor(Position=(width * 0.25, height * 0.25, 0.0), Radius=150.0, Strength=400.0, Type=AttractorType.Physical),
        Attractor(Position=(width * 0.75, height * 0.25, 0.0), Radius=350.0, Strength=1200.0, Type=AttractorType.Linear),
        Attractor(Position=(width * 0.25, height * 0.75, 0.0), Radius=550.0, Strength=900.0, Type=AttractorType.Exponential),
        Attractor(Position=(width * 0.75, height * 0.75, 0.0), Radius=250.0, Strength=700.0, Type=AttractorType.Linear)])
    noise = Noise(VelocityScale=(30.0, 30.0, 30.0), ReplaceOldVelocity=False, Seed=SEED_BASE + 100 + config_index)
    fma = FMA(VelocityMultiply=(0.98, 0.98, 0.98),
              Area=TransformArea(Type=AreaType.Ellipsoid, Center=(width / 2, height / 2, 0.0), Size=(width / 3, height / 3, 64.0), Falloff=64.0,
                                 Rotation=0.0))
    return ParticleScene(count, chunk_size, P, V, A, cfg, [spawner, gravity, noise, fma], dt)

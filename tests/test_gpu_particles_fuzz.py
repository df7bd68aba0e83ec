"""Randomised parity of the particle path: random transform chains (0-5 of Gravity with 1-6 attractors of every type, Noise in
both velocity modes with and without an area, FMA and MatrixMultiply over every area shape -- so the specialised
Gravity -> Noise -> FMA kernel, the empty chain and the generic op loop are all exercised), random system settings (friction,
maximum velocity, life decay, collision on / off with random bounce / escape / distance / life penalty, flat or full field
addressing), random spawners, and initial states that include dead particles, particles at rest, particles outside the field
and particles moving very fast.  Particle state must equal the oracle's EXACTLY after several steps; render outputs within 1e-5."""
import os

import numpy as np
import pytest

import illuminant_b200 as ib
from illuminant_b200 import scenes
from helpers import check_particles

pytestmark = pytest.mark.gpu
W, H = 320, 224


def _area(rs):
    if rs.rand() < 0.35:
        return None
    t = int(rs.choice([ib.AreaType.Ellipsoid, ib.AreaType.Box, ib.AreaType.Cylinder, ib.AreaType.Spheroid, ib.AreaType.Octagon]))
    return ib.TransformArea(Type=t, Center=(float(rs.uniform(0, W)), float(rs.uniform(0, H)), float(rs.uniform(0, 30))),
                            Size=(float(rs.uniform(20, 150)), float(rs.uniform(20, 150)), float(rs.uniform(10, 80))),
                            Falloff=float(rs.uniform(1, 60)), Rotation=float(rs.choice([0.0, rs.uniform(-2, 2)])))


def _transform(rs, kind):
    if kind == "gravity":
        return ib.Gravity(MaximumAcceleration=float(rs.uniform(1, 20)), Attractors=[
            ib.Attractor(Position=(float(rs.uniform(0, W)), float(rs.uniform(0, H)), float(rs.uniform(0, 40))), Radius=float(rs.uniform(5, 300)),
                         Strength=float(rs.uniform(-500, 1500)), Type=int(rs.choice([0, 1, 2]))) for _ in range(int(rs.randint(1, 7)))])
    if kind == "noise":
        return ib.Noise(VelocityScale=tuple(float(v) for v in rs.uniform(0, 40, 3)), ReplaceOldVelocity=bool(rs.rand() < 0.5),
                        PositionScale=tuple(float(v) for v in rs.uniform(0, 3, 4) * (rs.rand() < 0.5)), SpeedScale=float(rs.choice([0.0, rs.uniform(0, 6)])),
                        Interval=float(rs.choice([1000.0, 50.0, 333.0])), Seed=int(rs.randint(1, 1000)), Strength=float(rs.uniform(0.2, 1.0)),
                        CyclesPerSecond=float(rs.choice([10, 30])) if rs.rand() < 0.8 else None, Area=_area(rs))
    if kind == "fma":
        return ib.FMA(PositionAdd=tuple(float(v) for v in rs.uniform(-2, 2, 3)), PositionMultiply=tuple(float(v) for v in rs.uniform(0.98, 1.02, 3)),
                      VelocityAdd=tuple(float(v) for v in rs.uniform(-5, 5, 3)), VelocityMultiply=tuple(float(v) for v in rs.uniform(0.8, 1.1, 3)),
                      Strength=float(rs.uniform(0.2, 1.0)), CyclesPerSecond=float(rs.choice([10, 60])) if rs.rand() < 0.8 else None,
                      CategoryFilter=(0.0, 0.0) if rs.rand() < 0.2 else None, Area=_area(rs))
    a = rs.uniform(-0.3, 0.3)
    rot = (np.cos(a), np.sin(a), 0, 0, -np.sin(a), np.cos(a), 0, 0, 0, 0, 1, 0, rs.uniform(-2, 2), rs.uniform(-2, 2), 0, 1)
    return ib.MatrixMultiply(Position=tuple(float(v) for v in rot), Velocity=tuple(float(v) for v in rot[:12]) + (0.0, 0.0, 0.0, 1.0),
                             Strength=float(rs.uniform(0.2, 1.0)), Area=_area(rs))


def _random_case(ctx, seed):
    rs = np.random.RandomState(5000 + seed)
    chunk = int(rs.choice([64, 128]))
    count = int(rs.randint(chunk * chunk // 4, chunk * chunk * 2))
    fs = scenes.lighting_scene(300 + seed % 7, W, H, 0)
    collide = rs.rand() < 0.7
    df, tex = None, None
    if collide:
        df = scenes.make_distance_field(ctx, fs, resolution=float(rs.choice([0.25, 0.5, 1.0])))
        df.Rasterize(fs.obstructions)
        tex = df.Save()
    ps = scenes.particle_scene(400 + seed, count, chunk, W, H, steps_hint=20, collision_field=df, spawn_rate=float(rs.choice([0.0, 40000.0])))
    cfg = ps.configuration
    cfg.Friction = float(rs.choice([0.0, 0.1, rs.uniform(0, 2)]))
    cfg.MaximumVelocity = float(rs.choice([2048.0, 60.0, 9999.0]))
    cfg.LifeDecayPerSecond = float(rs.uniform(0.0, 20.0))
    if collide:
        c = cfg.Collision
        c.EscapeVelocity, c.BounceVelocityMultiplier = float(rs.uniform(0, 400)), float(rs.choice([0.0, 0.95, rs.uniform(0, 1.5)]))
        c.Distance, c.LifePenalty = float(rs.uniform(0.1, 3.0)), float(rs.choice([0.0, 0.3]))
        c.FullFieldAddressing = bool(rs.rand() < 0.3)
    P, V = ps.positions, ps.velocities
    n = P.shape[0]
    P[rs.rand(n) < 0.1, 3] = 0.0                                   # dead on entry
    V[rs.rand(n) < 0.1, :3] = 0.0                                  # at rest
    far = rs.rand(n) < 0.05
    P[far, :2] += rs.uniform(-400, 400, (int(far.sum()), 2)).astype(np.float32)   # outside the field
    V[rs.rand(n) < 0.05, :3] *= 60.0                               # fast: clamped by MaximumVelocity, long collision marches
    V[:, 3] = np.where(rs.rand(n) < 0.3, rs.randint(0, 4, n), 0).astype(np.float32)   # category / bounce delay
    P[:, 2] = rs.uniform(-5, 70, n).astype(np.float32)
    kinds = ["gravity", "noise", "fma", "matrix"]
    if rs.rand() < 0.3:
        chain = ["gravity", "noise", "fma"]                        # the specialised kernel
    else:
        chain = [kinds[i] for i in rs.randint(0, 4, int(rs.randint(0, 6)))]
    transforms = ([ps.transforms[0]] if ps.transforms[0].MinRate > 0 else []) + [_transform(rs, k) for k in chain]
    return ps, chunk, tex, transforms, rs


@pytest.mark.parametrize("seed", range(int(os.environ.get("ILB_FUZZ_SEEDS", "16"))))
def test_random_particle_system_matches_the_oracle(ctx, oracle, seed):
    from test_gpu_particles import _run_both
    ps, chunk, tex, transforms, rs = _random_case(ctx, seed)
    steps = int(rs.randint(2, 7))
    system, gpu, ref = _run_both(ctx, oracle, ps, chunk, steps, tex=tex, seed=11 + seed, max_chunks=6, transforms=transforms)
    check_particles(gpu, ref, f"seed {seed} ({[type(t).__name__ for t in transforms]}, {steps} steps)", allow_nan=True)
    assert system.LiveCount == int((ref[0][:, 3] > 0).sum())

"""N > 1 path on CPU: world_size-2 `gloo` processes run the same partitioning bench.py uses on GPUs -- row bands of the
lightmap reassembled by one all-gather, particle chunk ranges with no collective -- with the CPU oracle standing in for
the device kernel (test infrastructure may call the oracle).  Sharded results must equal the unsharded ones exactly."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, outdir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import illuminant_b200 as ib
    from illuminant_b200 import scenes, sharding
    from oracle import oracle

    # ---- lighting: row bands + one all-gather
    s = scenes.lighting_scene(90, 80, 45, 3, n_directional=1, ramp=(30.0, 90.0), float4_lightmap=True)    # 45 rows: uneven split
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, s.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField, r._gbuffer_shape = df, s.gbuffer.shape[:2]
    batches, nb, verts, nv = r.build_batches()
    h = sharding.band_height(s.height, world)
    r0, r1 = sharding.row_band(rank, world, s.height)
    band = np.zeros((h, s.width, 4), np.float32)
    band[:r1 - r0] = oracle.render_lighting(tex, s.gbuffer, r.build_frame(1.0, (r0, r1)), batches, nb, verts, nv, nthreads=2)
    full = torch.empty((h * world, s.width, 4))
    dist.all_gather_into_tensor(full, torch.from_numpy(band))
    whole = oracle.render_lighting(tex, s.gbuffer, r.build_frame(), batches, nb, verts, nv, nthreads=2)
    assert np.array_equal(full.numpy()[:s.height], whole)

    # ---- particles: chunk ranges, no collective (gather only to compare)
    ps = scenes.particle_scene(91, 3 * 32 * 32, 32, 80, 45, steps_hint=10)
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=32, RandomSeed=9))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=3)
    system.Transforms = ps.transforms[1:]
    ops, u = system.plan_ops(0.0), system.system_uniforms(ps.dt)
    per = 32 * 32
    c0, c1 = sharding.chunk_range(rank, world, 3)
    sl = slice(c0 * per, c1 * per)
    mine = oracle.particles_step(ps.positions[sl], ps.velocities[sl], ps.attributes[sl], 32, u, [], ops, engine.RandomnessTexture, None, 3, nthreads=2)
    allp = oracle.particles_step(ps.positions, ps.velocities, ps.attributes, 32, u, [], ops, engine.RandomnessTexture, None, 3, nthreads=2)
    for a, b in zip(mine, allp):
        assert np.array_equal(a, b[sl])
    (Path(outdir) / f"ok{rank}").write_text("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path, oracle):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

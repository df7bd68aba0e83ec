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


def _host_frame_worker(rank, world, name, outdir, depth=1):
    """The host-to-host leg of bench.py at N > 1 without a device: every rank writes its band of three consecutive frames into
    the one shared host frame (here the oracle stands in for ilb_render_lighting_frame), rank 0 consumes each complete frame."""
    sys.path.insert(0, str(ROOT))
    os.environ["OMP_NUM_THREADS"] = "2"
    import illuminant_b200 as ib
    from illuminant_b200 import scenes, sharding
    from oracle import oracle
    s = scenes.lighting_scene(92, 64, 45, 2, float4_lightmap=True)
    df = scenes.make_distance_field(None, s)
    tex = oracle.generate_distance_field(df, s.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, s.environment, s.configuration)
    r.DistanceField, r._gbuffer_shape = df, s.gbuffer.shape[:2]
    shared = sharding.SharedHostFrame(name, s.height, s.width, 4, np.float32, rank, world, depth=depth)
    r0, r1 = sharding.row_band(rank, world, s.height)
    last = 5
    for seq in range(1, last + 1):
        scale = float(seq)            # a different frame every time: a stale band would be noticed
        batches, nb, verts, nv = r.build_batches(scale)
        shared.begin(seq)             # with depth 2 the other rank may be one frame ahead of the consumer, never two
        shared.rows(r0, r1, seq)[...] = oracle.render_lighting(tex, s.gbuffer, r.build_frame(scale, (r0, r1)), batches, nb, verts, nv, nthreads=2)
        shared.publish(seq)
        if rank == 0:
            shared.wait_complete(seq)
            whole = oracle.render_lighting(tex, s.gbuffer, r.build_frame(scale), batches, nb, verts, nv, nthreads=2)
            assert np.array_equal(shared.frame_of(seq), whole), seq
            shared.release(seq)
    (Path(outdir) / f"hf{rank}").write_text("ok")
    if rank == 0:   # the others may still be mapping / unmapping; the name can go, the memory lives until the last unmap
        shared.wait_complete(last)
    shared.close()


@pytest.mark.parametrize("depth", [1, 2])
def test_shared_host_frame_is_reassembled_by_the_ranks_themselves(tmp_path, oracle, depth):
    name = f"ilb_test_frame_{os.getpid()}_{depth}"
    mp.spawn(_host_frame_worker, args=(2, name, str(tmp_path), depth), nprocs=2, join=True)
    assert (tmp_path / "hf0").exists() and (tmp_path / "hf1").exists()
    assert not os.path.exists(os.path.join("/dev/shm", name))


def test_shared_host_frame_single_rank_protocol():
    from illuminant_b200 import sharding
    name = f"ilb_test_frame1_{os.getpid()}"
    f = sharding.SharedHostFrame(name, 8, 4, 4, np.float16, 0, 1)
    assert f.frame.shape == (8, 4, 4) and f.frame.dtype == np.float16 and not f.frame.any()
    assert f.frame.ctypes.data % 4096 == 0          # page-aligned behind the header: page-lockable as one range
    f.begin(1)
    f.rows(2, 5)[...] = 1.0
    f.publish(1)
    f.wait_complete(1)
    assert float(f.frame.sum()) == 3 * 4 * 4
    with pytest.raises(TimeoutError):
        f.begin(3, timeout_s=0.05)                   # frame 2 was never released
    f.release(2)
    f.begin(3)
    f.close()
    g = sharding.SharedHostFrame(name, 8, 4, 4, np.float16, 0, 1, depth=2)      # two slots: frame s lives in slot s % 2
    assert g.frame_of(1) is g.frames[1] and g.frame_of(2) is g.frames[0] and g.frames[1].ctypes.data % 4096 == 0
    g.begin(1)
    g.begin(2)                                       # one frame ahead of the consumer is allowed ...
    with pytest.raises(TimeoutError):
        g.begin(3, timeout_s=0.05)                   # ... two are not
    g.rows(0, 8, 1)[...] = 2.0
    assert not g.frames[0].any() and float(g.frame_of(1).sum()) == 2.0 * 8 * 4 * 4
    g.close()
    assert not os.path.exists(os.path.join("/dev/shm", name))

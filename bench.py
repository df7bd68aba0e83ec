#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric on this framework (and, with --impl reference, on the CPU restatement
of the reference algorithm).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload both|lighting|particles]

Primary metric: lit Mpixels/s on config C4 (3840x2160, 128 mixed Sphere/Directional/Line lights + 256 light probes,
9-slice distance field); a "step" is one RenderLighting of the whole frame (+ the probe update).  The second hot path
is reported in the same JSON line under "particles": Mparticle-steps/s for 8M particles (32 chunks x 512^2) per GPU
through Spawner+Gravity+Noise+FMA+SDF collision; a step is one ParticleSystem.Update.

Timing: W >= 3 warm-up steps, then exactly K steps bracketed by barrier + synchronize, timed with CUDA events on the
library's stream, max over ranks.  Inputs are larger than L2 (C4: 265 MB field + 133 MB G-buffer; particles: 640 MB of
state), so every step streams from HBM -- no L2 flush needed (config.l2 says so).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LIGHT_BYTES_PER_PIXEL = 48       # G-buffer 16 + half4 lightmap 8 + 3 physical DF slices x 8
PARTICLE_BYTES_PER_STEP = 112    # read P,V,attr + write P,V,renderColor,renderData


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML in a thread every 5 ms (the timed region of
    the default run is ~0.1 s), `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop, self._thread, self._nvml = threading.Event(), None, None

    def _nvml_loop(self):
        n, h = self._nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for name, attr in names:
                    if mask & int(getattr(n, attr, 0)):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            # NVML enumerates all GPUs of the box; map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.device
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.device])
            self._nvml = (n, n.nvmlDeviceGetHandleByIndex(idx))
            self._thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_lighting_sample(oracle, scenes, ib, scene, target_seconds: float):
    """Times the oracle (reference algorithm, multi-pass, all host threads) on a band of rows of the same frame."""
    df = scenes.make_distance_field(None, scene)
    tex = oracle.generate_distance_field(df, scene.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, scene.environment, scene.configuration)
    r.DistanceField, r._gbuffer_shape = df, scene.gbuffer.shape[:2]
    batches, nb, verts, nv = r.build_batches()
    mid = scene.height // 2

    def run(rows):
        frame = r.build_frame(1.0, (mid - rows // 2, mid - rows // 2 + rows))
        t = time.perf_counter()
        oracle.render_lighting(tex, scene.gbuffer, frame, batches, nb, verts, nv)
        return time.perf_counter() - t
    rows = 4
    t = run(rows)
    rows = int(min(scene.height, max(4, rows * target_seconds / max(t, 1e-3))))
    t = run(rows)
    return rows * scene.width / t / 1e6, rows, t


def cpu_particle_sample(oracle, scenes, ib, tex, df_desc, target_seconds: float):
    ps = scenes.particle_scene(2, 1 << 18, 512, 3840, 2160, steps_hint=1000, collision_field=df_desc, spawn_rate=0.0)
    engine = ib.ParticleEngine(None, ib.ParticleEngineConfiguration(ChunkSize=512, RandomSeed=0xB200))
    system = ib.ParticleSystem(engine, ps.configuration, maxChunks=1)
    system.Transforms = ps.transforms
    ops, u = system.plan_ops(ps.dt), system.system_uniforms(ps.dt)
    P, V, A = ps.positions, ps.velocities, ps.attributes
    t = time.perf_counter()
    oracle.particles_step(P, V, A, 512, u, [], ops, engine.RandomnessTexture, tex, 1)
    t1 = time.perf_counter() - t
    steps = int(max(1, min(64, target_seconds / max(t1, 1e-3))))
    t = time.perf_counter()
    oracle.particles_step(P, V, A, 512, u, [], ops, engine.RandomnessTexture, tex, steps)
    t = time.perf_counter() - t
    return ps.count * steps / t / 1e6, steps, t


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import illuminant_b200 as ib
    from illuminant_b200 import scenes
    from oracle import oracle
    oracle.lib()
    cores = oracle.threads()
    scene = scenes.config_c4()
    df = scenes.make_distance_field(None, scene)
    tex = oracle.generate_distance_field(df, scene.obstructions)
    df.ValidSliceCount, df.handle = df.SliceCount, 1
    r = ib.LightingRenderer(None, scene.environment, scene.configuration)
    r.DistanceField, r._gbuffer_shape = df, scene.gbuffer.shape[:2]
    batches, nb, verts, nv = r.build_batches()
    rows = 8   # bounded sample per step: an 8-row band around the middle of the 4K frame, all 128 lights
    mid = scene.height // 2
    frame = r.build_frame(1.0, (mid - rows // 2, mid + rows // 2))
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        oracle.render_lighting(tex, scene.gbuffer, frame, batches, nb, verts, nv)
        if i >= args.warmup:
            times.append(time.perf_counter() - t)
    total = sum(times)
    value = rows * scene.width * args.steps / total / 1e6
    pvalue, psteps, pt = cpu_particle_sample(oracle, scenes, ib, tex, df, 6.0) if args.workload != "lighting" else (None, 0, 0)
    sample = f"{rows}-row band ({rows * scene.width} px) of the 3840x2160 / 128-light frame per step"
    line = {"impl": "reference", "metric": "lit Mpixels/s (4K, 128 lights)", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4 3840x2160, 96 sphere + 8 directional + 24 line lights, 9-slice DF", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "particles": None if pvalue is None else {"metric": "Mparticle-steps/s", "value": pvalue, "unit": "Mparticle-steps/s",
                                                      "cpu_baseline": {"value": pvalue, "unit": "Mparticle-steps/s", "cores": cores, "kind": "port",
                                                                       "sample": f"262144 particles x {psteps} steps of the C3/C5 chain"}},
            "note": "reference = CPU restatement (oracle port) of the HLSL path; the C#/HLSL reference cannot run on this box (no D3D/.NET)"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import illuminant_b200 as ib
    from illuminant_b200 import _abi, build, scenes
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    build.build()
    torch.cuda.set_device(local_rank)
    ctx = ib.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    peak, peak_src = measured_peaks()
    result = {}

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, steps, warmup):
        """Returns (total ms of `steps` steps, per-step ms list) measured with CUDA events on the library's stream."""
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_fn()
            barrier()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            evs[0].record(stream)
            for i in range(steps):
                step_fn()
                evs[i + 1].record(stream)
            barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return evs[0].elapsed_time(evs[steps]), per

    sampler = ClockSampler(local_rank)
    scene = scenes.config_c4()
    W, H = scene.width, scene.height

    # ------------------------------------------------------------------ lighting
    if args.workload in ("both", "lighting"):
        renderer = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
        df = scenes.make_distance_field(ctx, scene)
        df.Rasterize(scene.obstructions)      # N1 kernel: the field is produced on the GPU, replicated per rank
        renderer.DistanceField = df
        renderer.Probes = scene.probes
        gb_host = torch.from_numpy(scene.gbuffer).pin_memory()
        renderer.SetGBuffer(gb_host.numpy())
        from illuminant_b200 import sharding
        rows_per = sharding.band_height(H, world)
        r0, r1 = sharding.row_band(rank, world, H)
        # Reassembly of the lit buffer.  Preferred: the kernel itself stores every finished texel into
        # the full-frame buffer of EVERY rank through NVLink peer mappings (torch symmetric memory provides the mapped
        # pointers), followed by a device-side barrier -- compute and all-gather are one kernel.  Fallback: a plain NCCL
        # all-gather of the row bands.
        gather, hdl, peer_ptrs = "none", None, None
        if dist is not None and args.gather in ("auto", "peers"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                full = symm_mem.empty((rows_per * world, W, 4), dtype=torch.float16, device=torch.device("cuda", local_rank))
                hdl = symm_mem.rendezvous(full, dist.group.WORLD)
                peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
                gather = "peer-stores (in-kernel all-gather over NVLink) + symmetric-memory barrier"
            except Exception as e:   # noqa: BLE001
                if args.gather == "peers":
                    raise
                hdl, peer_ptrs = None, None
                print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
        if peer_ptrs is None:
            full = torch.empty((rows_per * world, W, 4), dtype=torch.float16, device="cuda")
            if dist is not None:
                gather = "nccl all_gather_into_tensor"
        band = full[rank * rows_per:(rank + 1) * rows_per]
        packed = renderer.build_batches()
        launches0 = ctx.launch_count

        def light_step():
            if peer_ptrs is not None:
                renderer.RenderLightingPeers(peer_ptrs, rows=(r0, r1), packed=packed)
                hdl.barrier(channel=0)
            else:
                renderer.RenderLightingDevice(band.data_ptr(), rows=(r0, r1), packed=packed)
                if dist is not None:   # one all-gather of row bands reassembles the lit buffer on every rank
                    dist.all_gather_into_tensor(full, band)

        sampler.start()
        total_ms, per = timed(light_step, args.steps, args.warmup)
        clocks = sampler.stop()
        launches = (ctx.launch_count - launches0) * args.steps // (args.steps + args.warmup)
        total_ms = max_over_ranks(total_ms)
        ms_step = total_ms / args.steps
        mpx = W * H / (ms_step * 1e-3) / 1e6
        # kernel-only duration (no collective) for the roofline: time the band kernel alone
        k_ms, _ = timed(lambda: renderer.RenderLightingDevice(band.data_ptr(), rows=(r0, r1), packed=packed), max(3, args.steps // 2), 1)
        k_ms = k_ms / max(3, args.steps // 2)
        achieved = LIGHT_BYTES_PER_PIXEL * W * (r1 - r0) / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("light_accumulate_kernel")
            except Exception:
                traffic = None

        # end to end through the public API with HOST buffers: per step the G-buffer (the per-frame input) goes
        # host->device from pinned memory and the lightmap band comes back into pinned host memory
        out_host = torch.empty((max(r1 - r0, 1), W, 4), dtype=torch.float16).pin_memory()
        frame = renderer.build_frame(1.0, (r0, r1))
        batches, nb, verts, nv = packed
        gb_ptr, out_ptr = C.c_void_p(gb_host.data_ptr()), C.c_void_p(out_host.data_ptr())

        def e2e_step():   # one C-ABI call per frame: G-buffer up, shade, lightmap down (pipelined over row bands inside)
            ctx.check(ctx.lib.ilb_render_lighting_frame(ctx.handle, df.handle, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                        C.cast(verts, C.c_void_p), nv, W, H, _abi.FORMAT_FLOAT4, gb_ptr, out_ptr))
        e_steps = max(3, args.steps // 2)
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e_steps)
        result.update({
            "metric": "lit Mpixels/s (4K, 128 lights)", "value": mpx, "unit": "Mpixels/s", "ms_per_step": ms_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "light_accumulate_kernel (line-light pass + sphere/directional pass, both launches of the step)",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_pixel": LIGHT_BYTES_PER_PIXEL, "peak_source": peak_src,
                         "note": "per-pixel work is O(lights x trace steps): issue- and latency-bound, not HBM-bound (see DESIGN.md); "
                                 "traffic = ncu DRAM bytes of both passes (the expanded distance-field planes trade traffic for instructions)"},
            "e2e": {"value": W * H / (e_ms * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": int((r1 - r0) * W * 16 + nv * 128),
                    "d2h_bytes_per_step": int(out_host.numel() * 2), "ms_per_step": e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "gather": gather,
        })
        if world > 1:   # checksum of the reassembled frame: every rank must hold the same lit buffer
            chk = torch.tensor([float(full[:H].float().sum().item())], device="cuda", dtype=torch.float64)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            result["gather_checksum_equal"] = bool(lo.item() == hi.item())
            whole = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")     # unsharded render of the same frame
            renderer.RenderLightingDevice(whole.data_ptr(), rows=(0, H), packed=packed)
            barrier()
            result["gather_matches_single_gpu"] = bool(torch.equal(whole.view(torch.int16), full[:H].view(torch.int16)))
        # probes (config 4's "GI probes"): timed separately, tiny
        t0 = time.perf_counter()
        renderer.UpdateLightProbes()
        result["probes_ms"] = (time.perf_counter() - t0) * 1e3

        # N3: resolve of the lit buffer -- HalfVector4 lightmap + Color albedo -> Color backbuffer, tone-mapped.
        # A streaming kernel: 16 algorithmic bytes per pixel.  Two buffer sets alternate (2 x 133 MB > L2) so every launch
        # streams from HBM.  Reported next to the headline, never part of it.
        try:
            from illuminant_b200 import hdr as hdr_mod
            hdr_cfg = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.2, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0))
            rp = hdr_mod.pack_resolve(W, H, _abi.FORMAT_HALF4, hdr_cfg, _abi.FORMAT_RGBA8, _abi.FORMAT_RGBA8)
            lms = [full[:H].contiguous(), full[:H].clone()]
            als = [torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
            outs = [torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
            flip = {"i": 0}

            R_INNER = 8   # launches per timed step: one 4K resolve is shorter than the Python launch path, so a step is 8 back to back

            def resolve_step():
                for _ in range(R_INNER):
                    i = flip["i"] = flip["i"] ^ 1
                    ctx.check(ctx.lib.ilb_resolve_lighting_device(ctx.handle, C.byref(rp), C.c_void_p(lms[i].data_ptr()),
                                                                  C.c_void_p(als[i].data_ptr()), C.c_void_p(outs[i].data_ptr())))
            r_steps = max(20, 2 * args.steps)
            r_total, r_per = timed(resolve_step, r_steps, 3)
            r_ms = float(np.median(r_per)) / R_INNER
            r_ach = 16 * W * H / (r_ms * 1e-3) / 1e9
            result["resolve"] = {"metric": "resolved Mpixels/s (4K, tone-mapped, with albedo)", "value": W * H / (r_ms * 1e-3) / 1e6,
                                 "unit": "Mpixels/s", "ms_per_step": r_ms, "steps": r_steps,
                                 "roofline": {"bound": "hbm", "achieved": r_ach, "peak": peak, "unit": "GB/s", "frac": r_ach / peak,
                                              "traffic": None, "kernel": "resolve_kernel<ToneMap, albedo, vec4>",
                                              "algorithmic_bytes_per_pixel": 16, "peak_source": peak_src}}
            del lms, als, outs
        except Exception as e:   # noqa: BLE001 -- the headline must not depend on the N3 side measurement
            result["resolve"] = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ particles (weak: 8M particles per GPU)
    if args.workload in ("both", "particles"):
        chunk, nchunks = 512, 32
        count = chunk * chunk * nchunks
        pscene_field = scenes.lighting_scene(1, 1920, 1080, 0)
        pdf = scenes.make_distance_field(ctx, pscene_field, resolution=0.25)   # quarter-res field like SimpleParticles.cs:216-219
        pdf.Rasterize(pscene_field.obstructions)
        ps = scenes.particle_scene(2 + rank, count, chunk, 1920, 1080, steps_hint=1000, collision_field=pdf, spawn_rate=0.0)
        engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=0xB200))
        system = ib.ParticleSystem(engine, ps.configuration, maxChunks=nchunks)
        system.Transforms = ps.transforms
        system.Spawn(ps.positions, ps.velocities, ps.attributes)
        state = {"now": 0.0}
        launches0 = ctx.launch_count

        def particle_step():
            state["now"] += ps.dt
            system.Update(state["now"], ps.dt)

        p_steps, p_warm = max(20 * args.steps, 200), max(args.warmup, 3)   # ~0.1 s timed region: enough clock samples
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        total_ms, per = timed(particle_step, p_steps, p_warm)
        pclocks = sampler2.stop()
        p_launches = (ctx.launch_count - launches0) * p_steps // (p_steps + p_warm)
        total_ms = max_over_ranks(total_ms)
        ms_step = total_ms / p_steps
        mps = count * world / (ms_step * 1e-3) / 1e6
        achieved = PARTICLE_BYTES_PER_STEP * count / (float(np.median(per)) * 1e-3) / 1e9
        live = C.c_int64(0)

        def p_e2e_step():
            particle_step()
            ctx.check(ctx.lib.ilb_particles_count_live(system.handle, C.byref(live)))   # the liveness readback (D2H)
        for _ in range(2):
            p_e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(p_steps):
            p_e2e_step()
        barrier()
        e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / p_steps)
        ptraffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                ptraffic = json.loads(tp.read_text()).get("particle_step_kernel")
            except Exception:
                ptraffic = None
        result["particles"] = {
            "metric": "Mparticle-steps/s", "value": mps, "unit": "Mparticle-steps/s", "ms_per_step": ms_step, "steps": p_steps, "scaling": "weak",
            "config": {"workload": f"{count} particles per GPU (32 chunks x 512^2), Gravity(4)+Noise+FMA+UpdateWithDistanceField, dt 1/60",
                       "live_fraction": system.LiveCount / count},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ptraffic,
                         "kernel": "particle_step_kernel<COLLIDE, Gravity, Noise, FMA>", "algorithmic_bytes_per_particle_step": PARTICLE_BYTES_PER_STEP, "peak_source": peak_src},
            "e2e": {"value": count * world / (e_ms * 1e-3) / 1e6, "unit": "Mparticle-steps/s",
                    "h2d_bytes_per_step": int(C.sizeof(_abi.PsysUniforms) + 3 * C.sizeof(_abi.Op)), "d2h_bytes_per_step": 8, "ms_per_step": e_ms},
            "gpu_launches": int(p_launches), "clocks": pclocks,
        }

        # N2: ParticleSystem.Render of the same 8M particles into a 4K half4 target, additive (the whole
        # render: vertex work, binning, stable sort, ordered per-tile shading).  Reported next to the headline, never part of it.
        try:
            rparams = system.render_params(W, H, "Additive", ib.ParticleRenderParameters(Scale=(2.0, 2.0)), clearColor=(0, 0, 0, 0),
                                           target_format=_abi.FORMAT_HALF4)
            rtarget = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")

            def render_step():
                ctx.check(ctx.lib.ilb_particles_render_device(system.handle, C.byref(rparams), None, C.c_void_p(rtarget.data_ptr())))
            rn = 10
            _, rper = timed(render_step, rn, 2)
            rms = float(np.median(rper))
            rbytes = 48 * count + 8 * W * H
            result["render"] = {"metric": "rasterised Mparticles/s (8M particles per GPU -> 4K half4 target, additive)",
                                "value": count / (rms * 1e-3) / 1e6, "unit": "Mparticles/s", "ms_per_step": rms, "steps": rn,
                                "lit_fraction": float((rtarget[..., 3] > 0).float().mean().item()),
                                "roofline": {"bound": "hbm", "achieved": rbytes / (rms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                             "frac": rbytes / (rms * 1e-3) / 1e9 / peak, "traffic": None,
                                             "kernel": "raster_count + scan + raster_emit + radix sort + raster_ranges + raster_shade (one render)",
                                             "algorithmic_bytes": "48 B per particle (P, RenderData, RenderColor) + 8 B per target pixel",
                                             "peak_source": peak_src}}
            del rtarget
        except Exception as e:   # noqa: BLE001 -- the headline must not depend on the N2 side measurement
            result["render"] = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ combined frame loop (config C5)
    if args.workload == "both":
        # ParticleLights.cs:333-378: System.Update (collision against the lighting field) -> RenderLighting -> reassemble.
        c5 = scenes.config_c5_lighting()
        c5r = ib.LightingRenderer(ctx, c5.environment, c5.configuration)
        c5r.DistanceField = df                      # particles and lights share the same 4K field
        c5r._gbuffer_shape = renderer._gbuffer_shape
        c5packed = c5r.build_batches()
        system.Configuration.Collision.DistanceField = df

        def frame_step():
            particle_step()
            if peer_ptrs is not None:
                c5r.RenderLightingPeers(peer_ptrs, rows=(r0, r1), packed=c5packed)
                hdl.barrier(channel=0)
            else:
                c5r.RenderLightingDevice(band.data_ptr(), rows=(r0, r1), packed=c5packed)
                if dist is not None:
                    dist.all_gather_into_tensor(full, band)
        f_steps = max(args.steps, 10)
        total_ms, _ = timed(frame_step, f_steps, 3)
        ms_frame = max_over_ranks(total_ms) / f_steps
        result["combined_c5"] = {"metric": "frames/s (8M particles per GPU + 64-light 4K lighting)", "value": 1e3 / ms_frame, "unit": "frames/s",
                                 "ms_per_frame": ms_frame, "lit_mpixels_per_s": W * H / (ms_frame * 1e-3) / 1e6,
                                 "mparticle_steps_per_s": count * world / (ms_frame * 1e-3) / 1e6, "steps": f_steps}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        cores = oracle.threads()
        if "metric" in result:
            v, rows, t = cpu_lighting_sample(oracle, scenes, ib, scene, 12.0)
            result["cpu_baseline"] = {"value": v, "unit": "Mpixels/s", "cores": cores, "kind": "port",
                                      "sample": f"{rows}-row band ({rows * W} px) of the same 4K / 128-light frame, {t:.1f} s"}
        if "particles" in result:
            cdf = scenes.make_distance_field(None, pscene_field, resolution=0.25)
            ctex = oracle.generate_distance_field(cdf, pscene_field.obstructions)
            cdf.ValidSliceCount, cdf.handle = cdf.SliceCount, 1
            v, steps, t = cpu_particle_sample(oracle, scenes, ib, ctex, cdf, 8.0)
            result["particles"]["cpu_baseline"] = {"value": v, "unit": "Mparticle-steps/s", "cores": cores, "kind": "port",
                                                   "sample": f"262144 particles x {steps} steps of the same chain, {t:.1f} s"}

    if rank == 0:
        primary = "metric" in result
        if not primary:   # particles-only run: promote the secondary block
            p = result.pop("particles")
            result.update(p)
        line = {"metric": result.get("metric"), "value": result.get("value"), "unit": result.get("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": result.get("ms_per_step"), "higher_is_better": True,
                "scaling": "strong" if primary else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C4: 3840x2160, 96 sphere + 8 directional + 24 line lights, 256 probes, 9-slice 3840x2160 distance field"
                           if primary else result.get("config", {}).get("workload"),
                           "parallelism": f"row bands x{world}, gather: {result.get('gather', 'none')}" if primary else f"chunk ranges x{world}, no collective",
                           "l2": "inputs larger than L2 (no flush)"}}
        for k in ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "gather_checksum_equal", "gather_matches_single_gpu", "particles", "combined_c5", "probes_ms", "resolve", "render"):
            if k in result:
                line[k] = result[k]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "lighting", "particles"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "peers", "nccl"], help="N>1 lit-buffer reassembly")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric on this framework (and, with --impl reference, on the CPU restatement
of the reference algorithm).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload both|lighting|particles]

Primary metric: lit Mpixels/s on config C4 (3840x2160, 128 mixed Sphere/Directional/Line lights + 256 light probes,
9-slice distance field); a "step" is one RenderLighting of the whole frame PLUS the UpdateLightProbes of the same frame,
both inside the timed region.  At N > 1 the frame is cut into row bands of equal MEASURED cost (calibrated before the
timed region, sharding.rebalance_rows), every rank stores its band into every rank's full-frame buffer from inside the
kernel (NVLink peer stores), and a symmetric-memory barrier ends the step.  The host-to-host number (`e2e`) goes through
ilb_render_lighting_frame_async / _wait with HOST buffers, two frames in flight (frame n is queued before frame n - 1 is
waited for; every frame uploads its G-buffer from pinned memory and downloads its lightmap and probes); at N > 1 it is one such
call per rank on its band, the lightmap rows of every rank landing in one page-locked shared-memory host frame (two slots) that
rank 0 consumes (sharding.SharedHostFrame: no collective on that leg).  The second hot path is reported in the same
JSON line under "particles": Mparticle-steps/s for 8M particles (32 chunks x 512^2) per GPU through
Spawner(60 000 / s)+Gravity+Noise+FMA+SDF collision; a step is one ParticleSystem.Update (spawn kernel, Noise table kernel,
step kernel).  The particles the Spawner adds during the run are updated too but NOT counted in the metric.

Timing: W >= 3 warm-up steps, then exactly K steps bracketed by barrier + synchronize, timed with CUDA events on the
library's stream, max over ranks.  Inputs are larger than L2 (C4: 265 MB field + 133 MB G-buffer; particles: 640 MB of
state), so every step streams from HBM -- no L2 flush needed (config.l2 says so).

roofline.traffic is measured in the run: rank 0 profiles the same launches under `ncu --metrics dram__bytes_*` in a child
process (byte counters only; no timing is ever taken under the profiler); null when ncu is unavailable.
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
ROW_STRIDE = 8   # the CPU arm shades every 8th row of the WHOLE frame per step (offset = step mod 8): 8 steps cover the frame once


class CpuLighting:
    """The oracle (reference algorithm, multi-pass, all host threads) on the C4 frame.  Lights are clustered, so a band of
    adjacent rows is not representative of the frame; a step here is a strided sample of rows spread over the whole frame."""

    def __init__(self, oracle, scenes, ib, scene):
        self.oracle, self.scene = oracle, scene
        df = scenes.make_distance_field(None, scene)
        self.tex = oracle.generate_distance_field(df, scene.obstructions)
        df.ValidSliceCount, df.handle = df.SliceCount, 1
        self.df = df
        r = ib.LightingRenderer(None, scene.environment, scene.configuration)
        r.DistanceField, r._gbuffer_shape = df, scene.gbuffer.shape[:2]
        self.r = r
        self.packed = r.build_batches()

    def strided_step(self, step: int, stride: int = ROW_STRIDE):
        """(rows shaded, seconds) for rows step % stride, + stride, ... of the whole frame."""
        batches, nb, verts, nv = self.packed
        frame = self.r.build_frame(1.0, (step % stride, self.scene.height))
        t = time.perf_counter()
        out = self.oracle.render_lighting(self.tex, self.scene.gbuffer, frame, batches, nb, verts, nv, row_stride=stride)
        return out.shape[0], time.perf_counter() - t


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
    cpu = CpuLighting(oracle, scenes, ib, scene)
    times, rows = [], 0
    for i in range(args.warmup + args.steps):
        n, t = cpu.strided_step(i)
        if i >= args.warmup:
            times.append(t)
            rows += n
    total = sum(times)
    value = rows * scene.width / total / 1e6
    pvalue, psteps, pt = cpu_particle_sample(oracle, scenes, ib, cpu.tex, cpu.df, 6.0) if args.workload != "lighting" else (None, 0, 0)
    sample = (f"every {ROW_STRIDE}th row of the whole 3840x2160 / 128-light frame per step ({scene.height // ROW_STRIDE} rows, "
              f"{scene.height // ROW_STRIDE * scene.width} px; the row offset advances each step, {ROW_STRIDE} steps cover the frame once)")
    line = {"impl": "reference", "metric": "lit Mpixels/s (4K, 128 lights)", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C4_WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "particles": None if pvalue is None else {"metric": "Mparticle-steps/s", "value": pvalue, "unit": "Mparticle-steps/s",
                                                      "cpu_baseline": {"value": pvalue, "unit": "Mparticle-steps/s", "cores": cores, "kind": "port",
                                                                       "sample": f"262144 particles x {psteps} steps of the C3/C5 chain"}},
            "note": "reference = CPU restatement (oracle port) of the HLSL path; the C#/HLSL reference cannot run on this box (no D3D/.NET)"}
    print(json.dumps(line), flush=True)


C4_WORKLOAD = "C4: 3840x2160, 96 sphere + 8 directional + 24 line lights, 256 probes, 9-slice 3840x2160 distance field"


# ---------------------------------------------------------------------------------------------------- DRAM traffic (ncu child)
def measure_traffic(what: str, pattern: str, skip: int, count: int, extra=()):
    """dram__bytes_read.sum + dram__bytes_write.sum of `count` launches matching `pattern` of profiles/microbench/profile_hot.py,
    counted by ncu in a child process.  Byte counters only -- nothing timed under the profiler is ever reported.  Returns
    (bytes, None) or (None, reason); the warp-instruction count of the same launches (smsp__inst_executed.sum) is left in
    measure_traffic.instructions (None when it was not reported)."""
    measure_traffic.instructions = None
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not Path(ncu).exists():
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum", "--clock-control", "none", "-k", f"regex:{pattern}", "-s", str(skip),
           "-c", str(count), "--csv", sys.executable, str(ROOT / "profiles" / "microbench" / "profile_hot.py"), what, *[str(e) for e in extra]]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    except Exception as e:   # noqa: BLE001
        return None, f"ncu failed: {type(e).__name__}"
    import csv, io
    total, seen = 0.0, 0
    rows = [r for r in csv.reader(io.StringIO(res.stdout)) if len(r) > 5]
    if not rows:
        return None, "ncu printed no counters (profiling not permitted on this box?)"
    hdr = rows[0]
    try:
        ni, ui, vi = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    except ValueError:
        return None, "unexpected ncu output"
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    insts, iseen = 0.0, 0
    for r in rows[1:]:
        if r[ni].startswith("dram__bytes_"):
            total += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
            seen += 1
        elif r[ni].startswith("smsp__inst_executed"):
            insts += float(r[vi].replace(",", ""))
            iseen += 1
    if seen < 2 * count:
        return None, f"ncu saw {seen // 2} of {count} launches"
    if iseen == count:
        measure_traffic.instructions = insts
    return total, None


def issue_roofline(warp_instructions, kernel_ms, clocks, device):
    """What bounds the two hot kernels is the instruction issue rate, not HBM: the warp-instructions of
    the launches (counted by ncu in the run) over the kernel time (CUDA events, no profiler), against the 4 issue slots per SM
    and cycle at the SM clock sampled during the timed region.  Reported NEXT TO the HBM roofline the contract asks for."""
    if not warp_instructions or not kernel_ms or not clocks or not clocks.get("sm_mhz"):
        return None
    import torch
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    peak = sms * 4 * float(clocks["sm_mhz"]) * 1e6 / 1e9
    achieved = warp_instructions / (kernel_ms * 1e-3) / 1e9
    return {"bound": "issue", "warp_instructions": warp_instructions, "achieved": achieved, "peak": peak, "unit": "G warp-instructions/s",
            "frac": achieved / peak, "peak_source": f"{sms} SMs x 4 schedulers x {clocks['sm_mhz']:.0f} MHz (sampled under load)"}


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import illuminant_b200 as ib
    from illuminant_b200 import _abi, build, scenes, sharding
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

    def reduce_ranks(v: float, op="max") -> float:
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return float(t.item())

    def gather_ranks(v: float):
        if dist is None:
            return [v]
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def timed(step_fn, steps, warmup, sync=True):
        """Returns (total ms of `steps` steps, per-step ms list) measured with CUDA events on the library's stream."""
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_fn()
            if sync:
                barrier()
            else:
                ctx.synchronize()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            evs[0].record(stream)
            for i in range(steps):
                step_fn()
                evs[i + 1].record(stream)
            if sync:
                barrier()
            else:
                ctx.synchronize()
                torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return evs[0].elapsed_time(evs[steps]), per

    scene = scenes.config_c4()
    W, H = scene.width, scene.height
    peers = {"ptrs": None, "hdl": None, "full": None, "gather": "none"}

    def lighting_buffers():
        """The full-frame lit buffer every rank ends a step with.  Preferred: symmetric memory, so that the kernel itself stores
        every finished texel into the buffer of EVERY rank through NVLink peer mappings; fallback: NCCL all-gather of bands."""
        if dist is not None and args.gather in ("auto", "peers"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                full = symm_mem.empty((H, W, 4), dtype=torch.float16, device=torch.device("cuda", local_rank))
                hdl = symm_mem.rendezvous(full, dist.group.WORLD)
                peers.update(ptrs=[int(p) for p in hdl.buffer_ptrs], hdl=hdl, full=full,
                             gather="peer-stores (in-kernel all-gather over NVLink) + symmetric-memory barrier")
                return
            except Exception as e:   # noqa: BLE001
                if args.gather == "peers":
                    raise
                print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
        rows_per = sharding.band_height(H, world)
        peers.update(full=torch.empty((rows_per * world, W, 4), dtype=torch.float16, device="cuda"),
                     gather="nccl all_gather_into_tensor" if dist is not None else "none")

    # ------------------------------------------------------------------ lighting
    if args.workload in ("both", "lighting"):
        renderer = ib.LightingRenderer(ctx, scene.environment, scene.configuration)
        df = scenes.make_distance_field(ctx, scene)
        df.Rasterize(scene.obstructions)      # N1 kernel: the field is produced on the GPU, replicated per rank
        renderer.DistanceField = df
        renderer.Probes = scene.probes
        gb_host = torch.from_numpy(scene.gbuffer).pin_memory()
        renderer.SetGBuffer(gb_host.numpy())
        lighting_buffers()
        full = peers["full"]
        packed = renderer.build_batches()
        probes_packed = renderer.pack_probes()
        d_probes = torch.zeros((max(probes_packed[2], 1), 4), dtype=torch.float16, device="cuda")
        bounds = [sharding.row_band(k, world, H)[0] for k in range(world)] + [H]
        scratch = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")   # kernel-only timing target (any band fits)

        def my_rows():
            return bounds[rank], bounds[rank + 1]

        def band_kernel_ms(reps=2, with_probes=False):
            r0, r1 = my_rows()

            def work():
                if r1 > r0:
                    renderer.RenderLightingDevice(scratch.data_ptr(), rows=(r0, r1), packed=packed)
                if with_probes and rank == 0:   # what rank 0 does on top of its band inside a step
                    renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
            ms, _ = timed(work, reps, 1, sync=False)
            return ms / reps

        # Row bands of equal measured cost (lights are clustered: equal-height bands finish at different times and the frame
        # waits for the slowest rank).  Needs the peer-store gather, which leaves the band edges free.
        calibration = None
        if dist is not None and peers["ptrs"] is not None and not args.equal_bands:
            # A band's cost is not uniform over its rows and a measurement carries a per cent of noise, so the plain quantile update
            # overshoots and oscillates: every edge moves only half of the way, and the bounds that gave the smallest slowest band
            # over the iterations are the ones the timed region uses.
            history, best = [], None
            for it in range(7):
                times = gather_ranks(band_kernel_ms(6, with_probes=True))
                history.append({"bounds": list(bounds), "band_ms": [round(t, 4) for t in times]})
                if best is None or max(times) < best[0]:
                    best = (max(times), list(bounds))
                if it == 6:
                    break
                target = sharding.rebalance_rows(bounds, times, H, quantum=1)
                moved = [bounds[0]] + [int(round((b + 0.5 * (t - b)) / 4.0)) * 4 for b, t in zip(bounds[1:-1], target[1:-1])] + [H]
                bounds = [moved[0]] + [min(max(e, moved[k]), H) for k, e in enumerate(moved[1:])]
            bounds = best[1]
            calibration = history
        r0, r1 = my_rows()
        launches0 = ctx.launch_count

        def light_step():
            if peers["ptrs"] is not None:
                renderer.RenderLightingPeers(peers["ptrs"], rows=(r0, r1), packed=packed)
            else:
                renderer.RenderLightingDevice(full[r0:].data_ptr(), rows=(r0, r1), packed=packed)
            if rank == 0:   # UpdateLightProbes of the same frame (256 probes x 128 lights): rank 0, asynchronous, inside the timed step
                renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
            if peers["ptrs"] is not None:
                peers["hdl"].barrier(channel=0)
            elif dist is not None:   # one all-gather of row bands reassembles the lit buffer on every rank
                dist.all_gather_into_tensor(full, full[r0:r0 + sharding.band_height(H, world)])

        sampler = ClockSampler(local_rank)
        sampler.start()
        total_ms, per = timed(light_step, args.steps, args.warmup)
        clocks = sampler.stop()
        launches = (ctx.launch_count - launches0) * args.steps // (args.steps + args.warmup)
        total_ms = reduce_ranks(total_ms)
        ms_step = total_ms / args.steps
        mpx = W * H / (ms_step * 1e-3) / 1e6
        # kernel-only duration of this rank's band (no probes, no collective) for the roofline; all ranks' values are reported
        k_ms = band_kernel_ms(max(3, args.steps // 2))
        rank_kernel_ms = gather_ranks(k_ms)
        t0 = time.perf_counter()
        renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
        ctx.synchronize()
        probes_host_ms = (time.perf_counter() - t0) * 1e3
        p_ms, _ = timed(lambda: renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed), 5, 1, sync=False)
        achieved = LIGHT_BYTES_PER_PIXEL * W * (r1 - r0) / (max(k_ms, 1e-6) * 1e-3) / 1e9

        # end to end through the C-ABI with HOST buffers: per step this frame's G-buffer goes host->device from pinned memory and
        # the lit buffer comes back into pinned host memory.  N = 1: one ilb_render_lighting_frame call (copies pipelined behind
        # the kernels over row bands).  N > 1: every rank uploads the G-buffer rows of its band, renders it with the in-kernel
        # peer-store gather, and RANK 0 downloads the REASSEMBLED full frame.
        frame = renderer.build_frame(1.0, (r0, r1))
        batches, nb, verts, nv = packed
        probes_out_host = torch.zeros((max(probes_packed[2], 1), 4), dtype=torch.float16).pin_memory()
        e2e_drain = None
        if dist is None:
            # Two frames in flight, as a renderer that double-buffers its lit frame has them: frame n is queued
            # (ilb_render_lighting_frame_async: G-buffer up, shade, lightmap down into buffer n % 2) before frame n - 1 is waited
            # for, so the fill and the drain of one frame's copy pipeline hide behind its neighbours.  Every step still uploads
            # its G-buffer and downloads its lightmap and probes; the timed region ends when the last frame is on the host.
            out_hosts = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(2)]
            probes_hosts = [torch.zeros((max(probes_packed[2], 1), 4), dtype=torch.float16).pin_memory() for _ in range(2)]
            gb_ptr, out_ptrs = C.c_void_p(gb_host.data_ptr()), [C.c_void_p(o.data_ptr()) for o in out_hosts]
            fl = {"n": 0, "pending": 0}

            def e2e_step():
                n = fl["n"]
                fl["n"] += 1
                # the probe update and its read-back are queued first (asynchronous, independent of the lightmap); the frame's
                # last download is behind them in stream order, so waiting for the frame covers them
                renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
                with torch.cuda.stream(stream):
                    probes_hosts[n % 2].copy_(d_probes, non_blocking=True)
                ticket = C.c_uint64(0)
                ctx.check(ctx.lib.ilb_render_lighting_frame_async(ctx.handle, df.handle, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                                  C.cast(verts, C.c_void_p), nv, W, H, _abi.FORMAT_FLOAT4, gb_ptr, out_ptrs[n % 2],
                                                                  C.byref(ticket)))
                if fl["pending"]:      # frame n - 1 is complete in host memory
                    ctx.check(ctx.lib.ilb_render_lighting_frame_wait(ctx.handle, C.c_uint64(fl["pending"])))
                fl["pending"] = int(ticket.value)

            def e2e_drain():
                if fl["pending"]:
                    ctx.check(ctx.lib.ilb_render_lighting_frame_wait(ctx.handle, C.c_uint64(fl["pending"])))
                    fl["pending"] = 0
                ctx.synchronize()
            h2d, d2h = H * W * 16 + nv * 128 + probes_packed[2] * 32, H * W * 8 + probes_packed[2] * 8
            e2e_note = ("one ilb_render_lighting_frame_async call per frame (+ probe update and read-back), two frames in flight: frame n is "
                        "queued before frame n - 1 is waited for; every frame uploads its G-buffer and downloads its lightmap")
        else:
            # N > 1, host to host: every rank makes ONE ilb_render_lighting_frame call for its band (its G-buffer rows go up from
            # pinned memory, its rows of the lightmap come down, both pipelined behind the kernels inside the call) whose
            # `lightmap_out` points into ONE host frame shared by the node's ranks (POSIX shared memory, page-locked on every rank
            # with ilb_host_register).  The frame is reassembled in RANK 0's host memory by the copy engines of all GPUs over
            # their own PCIe links: no collective and no GPU barrier on this leg.  Per frame, a sequence number per rank in the
            # same mapping tells rank 0 that the rank's band is in place; rank 0 takes the frame and releases it.
            port = os.environ.get("MASTER_PORT", "0")
            shared, shared_err = None, ""
            nbytes = 2 * H * W * 8 + sharding.SharedHostFrame.HEADER
            try:
                st = os.statvfs("/dev/shm")
                room = st.f_bavail * st.f_frsize >= nbytes + (16 << 20)
            except OSError:
                room = False
            if room and os.environ.get("ILB_BENCH_NO_SHARED_FRAME") != "1":
                for turn in (0, 1):      # rank 0 creates the object (replacing a stale one of a crashed run) before anyone else opens it
                    if (rank == 0) == (turn == 0):
                        try:
                            shared = sharding.SharedHostFrame(f"ilb_bench_frame_{port}", H, W, 4, "float16", rank, world, ctx=ctx, timeout_s=10.0, depth=2)
                        except Exception as e:   # noqa: BLE001  (no room after all, page-locking refused, ...)
                            shared_err = f"{type(e).__name__}: {e}"
                    barrier()
            else:
                shared_err = "no room in /dev/shm" if not room else "disabled by ILB_BENCH_NO_SHARED_FRAME"
            if reduce_ranks(1.0 if shared is not None else 0.0, op="min") < 0.5:   # all ranks or none
                if shared is not None:
                    shared.close()
                shared = None
            seq = {"n": 0}
            if shared is not None:
                # two frames in the mapping: the other ranks work on frame n + 1 while rank 0 still holds frame n, so no rank ever
                # waits for the hand-shake of the frame before (a consumer that double-buffers the lit frame, as a renderer does)
                gb_ptr = C.c_void_p(gb_host.data_ptr())
                out_ptrs = [C.c_void_p(shared.rows(r0, r1, k).ctypes.data if r1 > r0 else shared.frames[k].ctypes.data) for k in range(shared.depth)]

                def e2e_finish(m, ticket):   # frame m of this rank is in the shared frame: publish it; rank 0 takes the whole frame
                    if m <= 0:
                        return
                    if ticket:
                        ctx.check(ctx.lib.ilb_render_lighting_frame_wait(ctx.handle, C.c_uint64(ticket)))
                    shared.publish(m)
                    if rank == 0:
                        shared.wait_complete(m)      # the whole frame m is in rank 0's host memory here
                        shared.release(m)

                def e2e_step():
                    seq["n"] += 1
                    n = seq["n"]
                    shared.begin(n)                  # slot n % 2 is free: frame n - 2 has been taken
                    if rank == 0:   # queued ahead of the band (asynchronous; the frame's last download is behind them in stream order)
                        renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
                        with torch.cuda.stream(stream):
                            probes_out_host.copy_(d_probes, non_blocking=True)
                    ticket = C.c_uint64(0)
                    if r1 > r0:     # this rank's band of frame n is queued before its band of frame n - 1 is waited for
                        ctx.check(ctx.lib.ilb_render_lighting_frame_async(ctx.handle, df.handle, C.byref(frame), C.cast(batches, C.c_void_p), nb,
                                                                          C.cast(verts, C.c_void_p), nv, W, H, _abi.FORMAT_FLOAT4, gb_ptr,
                                                                          out_ptrs[n % shared.depth], C.byref(ticket)))
                    e2e_finish(n - 1, seq.get("ticket", 0))
                    seq["ticket"] = int(ticket.value)
                    seq["open"] = n

                def e2e_drain():
                    if seq.get("open"):
                        e2e_finish(seq["open"], seq.get("ticket", 0))
                        seq["open"], seq["ticket"] = 0, 0
                    ctx.synchronize()
                e2e_note = ("one ilb_render_lighting_frame_async call per rank on its row band, two frames in flight; every rank's lightmap rows "
                            "land directly in a page-locked shared-memory host frame owned by rank 0 (two frame slots, no collective, no GPU "
                            "barrier); rank 0 waits for every rank's band of frame n, then releases its slot")
            else:
                # Fallback (no shared host frame on this box: the reason is in the note): every rank uploads its G-buffer rows and
                # renders its band with the device-side gather, rank 0 downloads the reassembled frame from its own device alone.
                if rank == 0:
                    print(f"[bench] shared host frame unavailable ({shared_err}); rank 0 downloads the gathered frame", file=sys.stderr)
                out_host = torch.empty((H, W, 4), dtype=torch.float16).pin_memory() if rank == 0 else None
                use_peers = peers["ptrs"] is not None

                def e2e_step():
                    with torch.cuda.stream(stream):
                        if r1 > r0:
                            ctx.check(ctx.lib.ilb_gbuffer_upload_rows(ctx.handle, W, H, _abi.FORMAT_FLOAT4, r0, r1, C.c_void_p(gb_host[r0:r1].data_ptr())))
                            if use_peers:
                                renderer.RenderLightingPeers(peers["ptrs"], rows=(r0, r1), packed=packed)
                            else:
                                renderer.RenderLightingDevice(full[r0:].data_ptr(), rows=(r0, r1), packed=packed)
                        if rank == 0:
                            renderer.UpdateLightProbesDevice(d_probes.data_ptr(), packed=packed, probes=probes_packed)
                        if use_peers:
                            peers["hdl"].barrier(channel=0)
                        else:
                            dist.all_gather_into_tensor(full, full[r0:r0 + sharding.band_height(H, world)])
                        if rank == 0:
                            out_host.copy_(full[:H], non_blocking=True)
                            probes_out_host.copy_(d_probes, non_blocking=True)
                    ctx.synchronize()
                    torch.cuda.synchronize()
                e2e_note = (f"shared host frame unavailable ({shared_err}): per-rank G-buffer band upload and render with the device-side gather, "
                            "rank 0 downloads the reassembled full frame")
            h2d = H * W * 16 + world * nv * 128 + probes_packed[2] * 32     # all ranks' band uploads together = one G-buffer
            d2h = H * W * 8 + probes_packed[2] * 8
        e_steps = max(3, args.steps // 2)
        e2e_error = ""
        try:
            for _ in range(2):
                e2e_step()
            if e2e_drain is not None:
                e2e_drain()
        except Exception as e:   # noqa: BLE001  (a rank that lost the hand-shake must not take the device-timed line with it)
            e2e_error = f"{type(e).__name__}: {e}"
        barrier()
        t0 = time.perf_counter()
        try:
            if not e2e_error:
                for _ in range(e_steps):
                    e2e_step()
                if e2e_drain is not None:    # the last frame in flight is on the host before the clock stops
                    e2e_drain()
        except Exception as e:   # noqa: BLE001
            e2e_error = f"{type(e).__name__}: {e}"
        barrier()
        e_ms = reduce_ranks((time.perf_counter() - t0) * 1e3 / e_steps)
        if reduce_ranks(1.0 if e2e_error else 0.0) > 0.5:    # on any rank: no host-to-host number rather than a wrong one
            e_ms = float("nan")
            e2e_note = "FAILED (" + (e2e_error or "on another rank") + "): " + e2e_note
            print(f"[bench] rank {rank}: host-to-host leg failed: {e2e_error or 'on another rank'}", file=sys.stderr)
            try:
                ctx.synchronize()
            except Exception:   # noqa: BLE001
                pass
        if dist is None:   # both host frames of the two-deep queue hold the frame a plain device render gives, bit for bit
            whole = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
            renderer.RenderLightingDevice(whole.data_ptr(), rows=(0, H), packed=packed)
            ctx.synchronize()
            result["e2e_host_frame_matches_device_render"] = all(bool(torch.equal(whole.cpu().view(torch.int16), o.view(torch.int16))) for o in out_hosts)
            del whole
        if dist is not None:
            renderer.SetGBuffer(gb_host.numpy())    # the band uploads left the other rows as they were; restore for what follows
            if rank == 0:   # the frame that arrived on the host is the unsharded render, bit for bit
                whole = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
                renderer.RenderLightingDevice(whole.data_ptr(), rows=(0, H), packed=packed)
                ctx.synchronize()
                if shared is not None:    # every slot holds a complete frame (each was written at least once, zero-filled before)
                    out_host = [torch.from_numpy(f) for f in shared.frames]
                else:
                    out_host = [out_host]
                result["e2e_host_frame_matches_single_gpu"] = all(bool(torch.equal(whole.cpu().view(torch.int16), o.view(torch.int16))) for o in out_host)
                del whole
            out_host = None
            barrier()
            if shared is not None:
                shared.close()

        traffic, traffic_note = None, "not measured"
        if rank == 0 and not args.no_traffic:
            l0 = ctx.launch_count      # launches of one render of this rank's band: 2, or 4 when the band runs as two halves
            renderer.RenderLightingDevice(scratch.data_ptr(), rows=(r0, r1), packed=packed)
            ctx.synchronize()
            per_band = max(int(ctx.launch_count - l0), 1)
            traffic, why = measure_traffic("light", "light_accumulate", per_band, per_band, (2, r0, r1))
            traffic_note = why or (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of the {per_band} lighting launches of this rank's band, "
                                   "measured in this run")
            result["roofline_issue"] = issue_roofline(measure_traffic.instructions, k_ms, clocks, local_rank)
        result.update({
            "metric": "lit Mpixels/s (4K, 128 lights)", "value": mpx, "unit": "Mpixels/s", "ms_per_step": ms_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_note,
                         "kernel": "light_accumulate_kernel (line-light pass + sphere/directional pass, both launches of this rank's band)",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_pixel": LIGHT_BYTES_PER_PIXEL, "peak_source": peak_src,
                         "note": "per-pixel work is O(lights x trace steps): issue- and latency-bound, not HBM-bound (see DESIGN.md)"},
            "e2e": {"value": (W * H / (e_ms * 1e-3) / 1e6) if e_ms == e_ms else None, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e_ms if e_ms == e_ms else None, "what": e2e_note},
            "gpu_launches": int(launches), "clocks": clocks, "gather": peers["gather"],
            "probes": {"kernel_ms": p_ms / 5, "host_call_ms": probes_host_ms, "in_timed_step": True, "count": probes_packed[2]},
            "bands": {"bounds": list(bounds), "rank_kernel_ms": [round(t, 4) for t in rank_kernel_ms], "calibration": calibration},
        })
        if world > 1:   # checksum of the reassembled frame: every rank must hold the same lit buffer
            light_step()
            barrier()
            chk = torch.tensor([float(full[:H].float().sum().item())], device="cuda", dtype=torch.float64)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            result["gather_checksum_equal"] = bool(lo.item() == hi.item())
            whole = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")     # unsharded render of the same frame
            renderer.RenderLightingDevice(whole.data_ptr(), rows=(0, H), packed=packed)
            barrier()
            result["gather_matches_single_gpu"] = bool(torch.equal(whole.view(torch.int16), full[:H].view(torch.int16)))
            del whole

        # N3: resolve of the lit buffer -- HalfVector4 lightmap + Color albedo -> Color backbuffer, tone-mapped.
        # A streaming kernel: 16 algorithmic bytes per pixel.  Two buffer sets alternate (2 x 133 MB > L2) so every launch
        # streams from HBM.  Reported next to the headline, never part of it.
        try:
            from illuminant_b200 import hdr as hdr_mod
            hdr_cfg = ib.HDRConfiguration(Mode=ib.HDRMode.ToneMap, Exposure=1.2, ToneMapping=ib.ToneMappingConfiguration(WhitePoint=3.0))
            rp = hdr_mod.pack_resolve(W, H, _abi.FORMAT_HALF4, hdr_cfg, _abi.FORMAT_RGBA8, _abi.FORMAT_RGBA8)
            lms = [full[:H].contiguous().clone(), full[:H].clone()]
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
            r_total, r_per = timed(resolve_step, r_steps, 3, sync=False)
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

    # ------------------------------------------------------------------ particles
    def particle_system(count, chunk, nchunks, field, seed, headroom=2):
        ps = scenes.particle_scene(seed, count, chunk, 1920, 1080, steps_hint=1000, collision_field=field, spawn_rate=60000.0)
        engine = ib.ParticleEngine(ctx, ib.ParticleEngineConfiguration(ChunkSize=chunk, RandomSeed=0xB200))
        system = ib.ParticleSystem(engine, ps.configuration, maxChunks=nchunks + headroom)   # headroom: chunks for the Spawner to fill
        system.Transforms = ps.transforms
        system.Spawn(ps.positions, ps.velocities, ps.attributes)
        return ps, system

    if args.workload in ("both", "particles"):
        chunk, nchunks = 512, 32
        count = chunk * chunk * nchunks
        pscene_field = scenes.lighting_scene(1, 1920, 1080, 0)
        pdf = scenes.make_distance_field(ctx, pscene_field, resolution=0.25)   # quarter-res field like SimpleParticles.cs:216-219
        pdf.Rasterize(pscene_field.obstructions)
        ps, system = particle_system(count, chunk, nchunks, pdf, 2 + rank)      # weak: 8M particles per GPU
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
        total_ms = reduce_ranks(total_ms)
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
        e_ms = reduce_ranks((time.perf_counter() - t0) * 1e3 / p_steps)
        ptraffic, ptraffic_note = None, "not measured"
        if rank == 0 and not args.no_traffic:
            ptraffic, why = measure_traffic("particles", "particle_step_kernel", 4, 1, (5,))
            ptraffic_note = why or "ncu dram__bytes_read.sum + dram__bytes_write.sum of one particle_step_kernel launch, measured in this run"
            p_issue = issue_roofline(measure_traffic.instructions, ms_step, pclocks, local_rank)
        result["particles"] = {
            "metric": "Mparticle-steps/s", "value": mps, "unit": "Mparticle-steps/s", "ms_per_step": ms_step, "steps": p_steps, "scaling": "weak",
            "config": {"workload": f"{count} particles per GPU (32 chunks x 512^2 + spawn headroom), Spawner(60000/s)+Gravity(4)+Noise+FMA+UpdateWithDistanceField, dt 1/60",
                       "live_particles_at_end": int(live.value), "live_chunks_at_end": system.LiveChunkCount,
                       "counted_per_step": count, "note": "particles added by the Spawner during the run are updated but not counted"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ptraffic,
                         "traffic_source": ptraffic_note,
                         "kernel": "particle_step_kernel<COLLIDE, Gravity, Noise, FMA> (+ spawn and Noise-table launches of the step)",
                         "algorithmic_bytes_per_particle_step": PARTICLE_BYTES_PER_STEP, "peak_source": peak_src},
            "e2e": {"value": count * world / (e_ms * 1e-3) / 1e6, "unit": "Mparticle-steps/s",
                    "h2d_bytes_per_step": int(C.sizeof(_abi.PsysUniforms) + 3 * C.sizeof(_abi.Op) + C.sizeof(_abi.Spawn)), "d2h_bytes_per_step": 8, "ms_per_step": e_ms},
            "gpu_launches": int(p_launches), "clocks": pclocks,
        }
        if rank == 0 and not args.no_traffic and p_issue:
            result["particles"]["roofline_issue"] = p_issue

        # N2: ParticleSystem.Render of the same 8M particles into a 4K half4 target, additive (the whole
        # render: vertex work, binning, stable sort, ordered per-tile shading).  Reported next to the headline, never part of it.
        try:
            rparams = system.render_params(W, H, "Additive", ib.ParticleRenderParameters(Scale=(2.0, 2.0)), clearColor=(0, 0, 0, 0),
                                           target_format=_abi.FORMAT_HALF4)
            rtarget = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")

            def render_step():
                ctx.check(ctx.lib.ilb_particles_render_device(system.handle, C.byref(rparams), None, C.c_void_p(rtarget.data_ptr())))
            rn = 10
            _, rper = timed(render_step, rn, 2, sync=False)
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

        # N2 at N > 1: every rank renders its chunks (rank order = draw order) over a transparent float4 layer, then composites its
        # row band of ALL ranks' layers (P2P loads over NVLink) and stores the band into every rank's target (peer stores).
        if dist is not None and peers["ptrs"] is not None:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                from illuminant_b200.particles import composite_layers
                layer = symm_mem.empty((H, W, 4), dtype=torch.float32, device=torch.device("cuda", local_rank))
                lhdl = symm_mem.rendezvous(layer, dist.group.WORLD)
                layer_ptrs = [int(p) for p in lhdl.buffer_ptrs]
                b0, b1 = sharding.row_band(rank, world, H)

                def sharded_render_step():
                    system.RenderLayerDevice(layer.data_ptr(), W, H, "Additive", ib.ParticleRenderParameters(Scale=(2.0, 2.0)))
                    lhdl.barrier(channel=0)                      # every rank's layer is complete
                    composite_layers(ctx, layer_ptrs, W, H, (b0, b1), "Additive", _abi.FORMAT_HALF4, (0.0, 0.0, 0.0, 0.0), peers["ptrs"])
                    peers["hdl"].barrier(channel=0)              # every rank holds the whole image
                sn = 10
                s_total, _ = timed(sharded_render_step, sn, 2)
                s_ms = reduce_ranks(s_total) / sn
                result["render_sharded"] = {"metric": f"rasterised Mparticles/s ({count} particles per GPU, layers composited over NVLink into every rank's 4K half4 target)",
                                            "value": count * world / (s_ms * 1e-3) / 1e6, "unit": "Mparticles/s", "ms_per_step": s_ms, "steps": sn,
                                            "lit_fraction": float((peers["full"][:H, :, 3] > 0).float().mean().item())}
                del layer
            except Exception as e:   # noqa: BLE001 -- a side measurement
                result["render_sharded"] = {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ combined frame loop (config C5)
    if args.workload == "both":
        # ParticleLights.cs:333-378: System.Update (collision against the lighting field) -> RenderLighting -> reassemble.
        c5 = scenes.config_c5_lighting()
        c5r = ib.LightingRenderer(ctx, c5.environment, c5.configuration)
        c5r.DistanceField = df                      # particles and lights share the same 4K field
        c5r._gbuffer_shape = renderer._gbuffer_shape
        c5packed = c5r.build_batches()
        eq0, eq1 = (r0, r1) if dist is None else sharding.row_band(rank, world, H)   # C5's lights are spread differently: equal bands

        def c5_lighting():
            if peers["ptrs"] is not None:
                c5r.RenderLightingPeers(peers["ptrs"], rows=(eq0, eq1), packed=c5packed)
                peers["hdl"].barrier(channel=0)
            else:
                c5r.RenderLightingDevice(full[eq0:].data_ptr(), rows=(eq0, eq1), packed=c5packed)
                if dist is not None:
                    dist.all_gather_into_tensor(full, full[eq0:eq0 + sharding.band_height(H, world)])

        def c5_loop(update, n_particles, label):
            def frame_step():
                update()
                c5_lighting()
            f_steps = max(args.steps, 10)
            total_ms, _ = timed(frame_step, f_steps, 3)
            ms_frame = reduce_ranks(total_ms) / f_steps
            return {"metric": f"frames/s ({label} + 64-light 4K lighting)", "value": 1e3 / ms_frame, "unit": "frames/s",
                    "ms_per_frame": ms_frame, "lit_mpixels_per_s": W * H / (ms_frame * 1e-3) / 1e6,
                    "mparticle_steps_per_s": n_particles / (ms_frame * 1e-3) / 1e6, "steps": f_steps}

        system.Configuration.Collision.DistanceField = df
        result["combined_c5"] = dict(c5_loop(particle_step, count * world, "8M particles per GPU"), scaling="weak (particles) / strong (lighting)")
        # C5 as BASELINE.json words it: 8M particles IN TOTAL, chunk ranges sharded over the ranks (32 chunks -> 4 per GPU at N = 8)
        c0, c1 = sharding.chunk_range(rank, world, nchunks)
        sps, ssys = particle_system(chunk * chunk * (c1 - c0), chunk, c1 - c0, df, 40 + rank)
        sstate = {"now": 0.0}

        def strong_step():
            sstate["now"] += sps.dt
            ssys.Update(sstate["now"], sps.dt)
        result["combined_c5_strong"] = dict(c5_loop(strong_step, count, "8M particles in total, chunk ranges sharded"),
                                            scaling="strong", chunks_per_rank=c1 - c0)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        cores = oracle.threads()
        if "metric" in result:
            cpu = CpuLighting(oracle, scenes, ib, scene)
            t_all, rows_all = 0.0, 0
            for step in range(ROW_STRIDE):          # the whole frame once, in 8 strided passes; stops early past 20 s
                n, t = cpu.strided_step(step)
                t_all += t
                rows_all += n
                if t_all > 20.0:
                    break
            result["cpu_baseline"] = {"value": rows_all * W / t_all / 1e6, "unit": "Mpixels/s", "cores": cores, "kind": "port",
                                      "sample": f"{rows_all} rows ({rows_all * W} px) of the same 4K / 128-light frame, every {ROW_STRIDE}th row "
                                                f"per pass, {t_all:.1f} s"}
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
                "config": {"workload": C4_WORKLOAD if primary else result.get("config", {}).get("workload"),
                           "parallelism": f"row bands of equal measured cost x{world}, gather: {result.get('gather', 'none')}" if primary else f"chunk ranges x{world}, no collective",
                           "l2": "inputs larger than L2 (no flush)"}}
        for k in ("roofline", "roofline_issue", "cpu_baseline", "e2e", "gpu_launches", "clocks", "gather_checksum_equal", "gather_matches_single_gpu", "e2e_host_frame_matches_single_gpu", "e2e_host_frame_matches_device_render", "probes", "bands",
                  "particles", "combined_c5", "combined_c5_strong", "resolve", "render", "render_sharded"):
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
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child process that counts DRAM bytes")
    ap.add_argument("--equal-bands", action="store_true", help="N>1: equal-height row bands instead of bands of equal measured cost")
    ap.add_argument("--gather", default="auto", choices=["auto", "peers", "nccl"], help="N>1 lit-buffer reassembly")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- particle-substeps/sec of the solver hot path (collisions on), headless, on N B200s.

One "step" is one substep = particlesUpdate(dt) + particlesCollisionsCheck_Grid(W, H, dt) over the whole field
(reference src/main.c:92-98).  `value` times K substeps with the state resident in HBM (CUDA events on the library's
own stream); `e2e` runs the same substeps through the reference-facing C entry points and copies positions and
colours back to the host arrays every 8th substep ("only when a frame is drawn"), copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pile16m|uniform16m|uniform1m] [--impl b200|reference]

`--impl reference` times the reference's own CPU solver (oracle/_ref/libref.so, built from the unmodified sources)
on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RESOLVE_NAMES = {1: "sequential-equivalent wavefront (replays the reference's pair order: positions and velocities bit-exact)",
                 2: "jacobi",
                 3: "tile Gauss-Seidel in shared memory (deterministic, bounded passes; integer work and positions bit-exact, velocities = "
                    "the reference's sequential pass 2 in a tile order, checked bit for bit against its CPU model)"}
METRIC = "particle-substeps/sec (collisions on)"
UNIT = "particle-substeps/s"
ALGO_BYTES_PER_PARTICLE_SUBSTEP = 40.0   # read p,v,r,m (24 B) + write p,v (16 B)
SUBSTEPS_PER_FRAME = 8
DT = 1.0 / 144.0 / SUBSTEPS_PER_FRAME

WORKLOADS = {
    # name: (particles per GPU, description, generator kwargs)
    "pile16m": (1 << 24, "16M polydisperse particles (radius 1.5-6, 1:4), dense pile on the bottom wall (BASELINE configs[2])"),
    "uniform16m": (1 << 24, "16M equal-radius particles (r=2), uniform random, area fraction 0.30 (BASELINE configs[4] sweep point)"),
    "uniform1m": (1 << 20, "1M equal-radius particles (r=2), uniform random, area fraction 0.30, 8 substeps/frame (BASELINE configs[1])"),
}
# Multi-GPU (BASELINE configs[3]): 256M particles (r = 1, area fraction 0.40) over 8 slabs of cell rows = 2**25 per GPU; the box is
# 45,916 wide (3,827 cell columns) and 478 cell rows (5,736 px) tall PER GPU, so per-GPU work is fixed as N grows (weak scaling).
WORKLOADS["slab32m"] = (1 << 25, "256M equal-radius particles (r=1, area fraction 0.40) over 8 GPUs = 32M per GPU, slabs of 478 cell rows "
                                 "with halo exchange and migration over NCCL send/recv (BASELINE configs[3])")
SLAB_W, SLAB_ROWS_PER_GPU, SLAB_CELL, SLAB_RADIUS = 45916, 478, 12, 1.0
CPU_SAMPLE_PARTICLES = 1 << 20


def make_scene(workload: str, n: int, seed_offset: int = 0):
    from particlephysicsc_b200 import scenes
    if workload.startswith("slab"):   # one slab's worth of the config-4 field (also the CPU sample)
        rows = max(8, int(round(SLAB_ROWS_PER_GPU * n / float(1 << 25))))
        return scenes.uniform_slab(n, SLAB_RADIUS, SLAB_W, rows * SLAB_CELL, 0, rows * SLAB_CELL, seed=0x5EED0004 + seed_offset)
    if workload.startswith("pile"):
        return scenes.polydisperse_pile(n, seed=0x5EED0003 + seed_offset)
    return scenes.uniform_random(n, 2.0, 0.30, seed=0x5EED0002 + seed_offset, vmax=0.0)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def max_over_ranks(ms: float, dist=None, device="cpu") -> float:
    """Slowest rank's time: every multi-GPU number is the max over ranks of a device-side duration."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms)
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_value(particles_per_rank: int, world: int, steps: int, ms_total: float) -> float:
    """particle-substeps/s of the whole job: units processed by ALL ranks divided by the slowest rank's time."""
    return particles_per_rank * world * steps / (ms_total * 1e-3)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons, sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx, self.proc, self.path = device_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_reference_cpu(workload: str, steps: int, warmup: int, budget_s: float = 0.0):
    """The reference's own CPU solver on a bounded sample of the workload (single-threaded, as the reference is).

    The sample is a field of the same generator and density with 2**k particles.  With budget_s > 0 (the ``--impl
    reference`` arm, whose K and W come from the driver) k is chosen from a one-step calibration on 2**16 particles so
    that the whole run takes about budget_s seconds; otherwise 2**20 particles are used."""
    from oracle import oracle as om
    if om.reference_available():
        R, kind = om.Reference(), "reference"
        step = lambda s, W, H: R.substep(s, W, H, DT)
    else:   # the reference sources are not here and no prebuilt libref.so travelled: time our C restatement
        O, kind = om.Oracle(), "port"
        step = lambda s, W, H: O.substep(s, W, H, DT, 0, True)

    def field(n):
        sc = make_scene(workload, n)
        return sc, om.State(sc.px.copy(), sc.py.copy(), sc.vx.copy(), sc.vy.copy(), sc.mass.copy(), sc.radius.copy(), sc.rgba.copy())

    n = min(CPU_SAMPLE_PARTICLES, WORKLOADS[workload][0])
    if budget_s > 0:
        sc, s = field(1 << 16)
        t0 = time.perf_counter()
        step(s, sc.W, sc.H)
        per_particle = (time.perf_counter() - t0) / (1 << 16)
        want = budget_s / max(1, steps + warmup) / per_particle
        n = 1 << max(14, min(24, int(np.floor(np.log2(max(want, 1.0))))))
        n = min(n, WORKLOADS[workload][0])
    sc, s = field(n)
    for _ in range(warmup):
        step(s, sc.W, sc.H)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(s, sc.W, sc.H)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    sample = (f"{n} particles of the same generator and density ({sc.W}x{sc.H} box), {warmup} untimed + {steps} timed substeps, "
              f"1 thread of {os.cpu_count()} host cores (the reference is single-threaded)")
    return value, dt / steps * 1e3, {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "particles": n}


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank on the CPU cores next to its GPU (what `numactl` / the launcher's binding does on a production node), BEFORE
    the pinned host arrays are allocated: the frame copy-back of 8 ranks then writes 8 node-local buffers instead of crossing
    the socket interconnect.  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no CPU affinity reported for this GPU"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cores near GPU {local_rank}"
    except Exception as e:   # no NVML, a container without the call: run unbound
        return f"unbound ({type(e).__name__})"


def nccl_options():
    """NCCL's own stream at high priority: the neighbour exchange of the slab decomposition runs beside the interior integrate
    pass (slab.slab_step), and the few CTAs of its transfer kernels must not queue behind that pass's hundred thousand."""
    try:
        from torch.distributed import ProcessGroupNCCL
        o = ProcessGroupNCCL.Options()
        o.is_high_priority_stream = True
        return o
    except Exception:
        return None


def run_reference(args, rank: int):
    if rank != 0:
        return
    value, ms, base = time_reference_cpu(args.workload, args.steps, args.warmup, budget_s=170.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOADS[args.workload][1], "sample": base["sample"],
                                            "full_field": base["particles"] == WORKLOADS[args.workload][0],
                                            "note": "the reference's CPU solver, one thread (it has no threads); the field is the full workload when "
                                                    "(steps + warmup) substeps of it fit ~170 s (one 16M-particle substep takes ~16 s), else the largest "
                                                    "power-of-two sample of the same generator and density that does (the full 16M-particle substep of "
                                                    "tests/test_gpu_parity.py runs at ~1.0 M particle-substeps/s on the same host, the 2^20 sample at ~0.9 M)"},
            "cpu_baseline": base, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def one_slab(api, torch, local_rank: int, n_per: int, steps: int, settle: int, warmup: int, peak: float, resolve: int = 1, opts=None):
    """BASELINE configs[3]'s per-GPU workload as ONE slab on ONE GPU (no neighbours): the denominator of the weak-scaling
    efficiency.  Returns value (state resident) and e2e (owned positions + colours to the host every 8th substep)."""
    from particlephysicsc_b200 import scenes, slab
    rows_per = max(8, int(round(SLAB_ROWS_PER_GPU * n_per / float(1 << 25))))
    H = rows_per * SLAB_CELL
    sc = scenes.uniform_slab(n_per, SLAB_RADIUS, SLAB_W, H, 0, H, seed=0x5EED0004)
    o = dict(opts or {})
    o["resolve"] = resolve = int(o.get("resolve", resolve))
    if resolve == 3:   # the tile order is part of the result: every slab of a field is told the whole field's tile width
        o.update(slab.gs_options(n_per, SLAB_W, H, 6.0))
    me = slab.DeviceSlab(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba, np.arange(n_per, dtype=np.uint32), SLAB_W, H, 6.0, (0, rows_per), 4,
                         capacity=int(1.15 * n_per) + 65536, device=local_rank, options=o)

    def run(k, colour_every=0):
        for s in range(k):
            slab.local_step([me], DT, bool(colour_every) and (s % colour_every == colour_every - 1))

    run(settle + warmup)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(me.stream)
    run(steps)
    e1.record(me.stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    frames = max(1, steps // SUBSTEPS_PER_FRAME)
    e0.record(me.stream)
    for _ in range(frames):
        run(SUBSTEPS_PER_FRAME, colour_every=SUBSTEPS_PER_FRAME)
        me.sync_host_async()
    me.sync_host_wait()
    e1.record(me.stream)
    e1.synchronize()
    e2e_ms = e0.elapsed_time(e1) / (frames * SUBSTEPS_PER_FRAME)
    me.free()
    return {"workload": WORKLOADS["slab32m"][1] + " -- one slab on one GPU, no neighbours", "resolve_option": resolve, "particles": n_per,
            "value": n_per / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "e2e": {"value": n_per / (e2e_ms * 1e-3), "unit": UNIT},
            "roofline_substep_frac": ALGO_BYTES_PER_PARTICLE_SUBSTEP * n_per / (ms * 1e-3) / 1e9 / peak}


def state_digest(ids, *arrays) -> int:
    """Sum over particles of a 64-bit mix of (index, field bit patterns): independent of the order and of how particles are split."""
    with np.errstate(over="ignore"):
        h = ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        for a in arrays:
            h = (h ^ np.ascontiguousarray(a).view(np.uint32).astype(np.uint64)) * np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(29)
        return int(np.sum(h, dtype=np.uint64))


def slab_equivalence(api, torch, dist, rank: int, world: int, local_rank: int, particles: int = 1 << 21, steps: int = 8, halo: int = 8):
    """Before anything is timed: a small field cut into `world` slabs and stepped over NCCL must equal the same field stepped as a
    whole on one GPU -- an order-independent 64-bit digest of (global index, p_x, p_y, v_x, v_y bit patterns) over all particles."""
    from particlephysicsc_b200 import scenes, slab
    sc = scenes.uniform_random(particles, 2.0, 0.30, seed=0x5EED0006, vmax=200.0)
    me = slab.make_device_slabs(sc, world, halo, device_of=lambda k: local_rank, only_rank=rank)[0]
    tr = slab.DistTransport(dist, torch.device("cuda", local_rank))
    travelled = 0
    for _ in range(steps):
        slab.slab_step(me, tr, DT)
        travelled = max(travelled, int(me.info().taint_rows))
    d = me.download()
    full = state_digest(d["ids"], d["p_x"], d["p_y"], d["v_x"], d["v_y"])
    # digests are sums modulo 2^64; add them as two 32-bit halves so that the int64 all_reduce cannot overflow
    parts = torch.tensor([full & 0xFFFFFFFF, full >> 32, len(d["ids"]), int(d["halo_ok"])], dtype=torch.int64, device="cuda")
    dist.all_reduce(parts, op=dist.ReduceOp.SUM)   # set-up check, not the data path
    tmax = torch.tensor([travelled], dtype=torch.int64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    me.free()
    out = None
    if rank == 0:
        lo, hi, count, ok = (int(x) for x in parts.tolist())
        slabs = (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF
        p = api.Particles(sc.n, device=local_rank)
        p.set_sync_policy(api.PPC_SYNC_MANUAL)
        p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
        p.run_substeps(sc.W, sc.H, DT, steps)
        p.download(api.PPC_DL_ALL)
        whole = state_digest(np.arange(sc.n, dtype=np.uint32), p.array("p_x"), p.array("p_y"), p.array("v_x"), p.array("v_y"))
        p.Free()
        out = {"bit_identical": slabs == whole and count == sc.n and ok == world, "particles": sc.n, "substeps": steps, "slabs": world,
               "halo_rows": halo, "inexactness_travel_rows": int(tmax.item()), "all_ranks_report_exact": ok == world,
               "particles_owned_by_all_ranks": count, "digest_slabs": f"{slabs:016x}", "digest_single_gpu": f"{whole:016x}",
               "what": "r=2 uniform random field at area fraction 0.30, initial speeds up to 200 px/s (particles migrate between slabs), slabs over "
                       "NCCL send/recv vs the whole field on one GPU; digest over (global index, p_x, p_y, v_x, v_y) bit patterns"}
    dist.barrier()
    return out


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    from particlephysicsc_b200 import api

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=nccl_options())
    n_per_gpu, descr = WORKLOADS[args.workload]
    n = n_per_gpu
    sc = make_scene(args.workload, n, seed_offset=0x100 * rank)
    p = api.Particles(n, device=local_rank)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.set_option("resolve", args.resolve)
    for kv in args.opt:
        k, v = kv.split("=")
        p.set_option(k, int(v))
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    W, H = sc.W, sc.H
    torch.cuda.set_device(local_rank)
    ext = torch.cuda.ExternalStream(p.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(ms: float) -> float:
        return max_over_ranks(ms, dist, "cuda")

    # settle + warm-up (untimed).  State (28 B x n = %d MB at 16M) is several times the 126 MB L2, so consecutive
    # substeps cannot reuse each other's lines; no explicit flush is needed.
    p.run_substeps(W, H, DT, args.settle)
    p.run_substeps(W, H, DT, args.warmup)
    barrier()

    # ---- value: K substeps, state resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = api.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    p.run_substeps(W, H, DT, args.steps)
    e1.record(ext)
    e1.synchronize()
    barrier()
    launches = api.kernel_launches() - launches0
    ms_total = max_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = whole_job_value(n, world, args.steps, ms_total)

    # ---- e2e: the reference's entry points, copy-back of positions + colours every 8th substep ----
    frames = max(1, args.steps // SUBSTEPS_PER_FRAME)
    mask = api.PPC_DL_POS | api.PPC_DL_COLOR
    barrier()
    e0.record(ext)
    for _ in range(frames):
        for s in range(SUBSTEPS_PER_FRAME):
            if s == SUBSTEPS_PER_FRAME - 1:
                p.request_colors()
            p.Update(DT)
            p.CollisionsCheck_Grid(W, H, DT)
        p.download_async(mask)   # the frame's copy-back overlaps the substeps of the next frame
    p.download_wait()            # ... and the last one is waited for inside the timed region
    e1.record(ext)
    e1.synchronize()
    barrier()
    e2e_ms = max_ranks(e0.elapsed_time(e1))
    e2e_value = n * world * frames * SUBSTEPS_PER_FRAME / (e2e_ms * 1e-3)
    d2h_per_step = 12.0 * n / SUBSTEPS_PER_FRAME

    # ---- per-kernel durations (CUDA events around every stage, serialised) -> roofline of the dominant kernel ----
    prof_steps = min(args.steps, 16)
    p.profile_reset()
    p.profile(True)
    p.run_substeps(W, H, DT, prof_steps)
    p.synchronize()
    p.profile(False)
    stages = {k: (ms / prof_steps, cnt // prof_steps) for k, (ms, cnt) in p.profile_read().items() if cnt}
    peak, peak_src = measured_peak_gbs()
    dom = max(stages, key=lambda k: stages[k][0]) if stages else None
    # algorithmic bytes per particle of each kernel: what that kernel cannot avoid moving
    algo = {"integrate_walls_keys": 36.0, "collide": 40.0, "resolve_wavefront": 16.0, "canonical_gather": 60.0, "cell_sort": 8.0}.get(
        dom, ALGO_BYTES_PER_PARTICLE_SUBSTEP)
    traffic = substep_traffic = None
    try:   # DRAM bytes per launch (per substep for kernels launched several times) from the ncu --set full captures under profiles/
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            per_stage = json.load(f).get(f"{args.workload}_resolve{args.resolve}", {})
        traffic = per_stage.get(dom)
        substep_traffic = per_stage.get("whole_substep")
    except Exception:
        traffic = None
    roofline = None
    if dom:
        ach = algo * n / (stages[dom][0] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "ncu --set full capture under profiles/ (bytes per launch)" if traffic else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": algo,
                    "stage_ms_per_substep": {k: round(v[0], 4) for k, v in stages.items()}}
    substep_gbs = ALGO_BYTES_PER_PARTICLE_SUBSTEP * n * world * args.steps / (ms_total * 1e-3) / 1e9

    def quick_value(workload: str, steps: int = 16, resolve: int | None = None, scene=None, label: str | None = None):
        """particle-substeps/s of another single-GPU configuration (state resident, same timing rules), for context."""
        resolve = args.resolve if resolve is None else resolve
        sc2 = scene if scene is not None else make_scene(workload, WORKLOADS[workload][0])
        m = sc2.n
        q = api.Particles(m, device=local_rank)
        q.set_sync_policy(api.PPC_SYNC_MANUAL)
        q.set_option("resolve", resolve)
        q.fill(sc2.px, sc2.py, sc2.vx, sc2.vy, sc2.mass, sc2.radius, sc2.rgba)
        q.run_substeps(sc2.W, sc2.H, DT, args.settle + args.warmup)
        ext2 = torch.cuda.ExternalStream(q.stream(), device=torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(ext2)
        q.run_substeps(sc2.W, sc2.H, DT, steps)
        a1.record(ext2)
        a1.synchronize()
        ms = a0.elapsed_time(a1) / steps
        fell = q.gs_fallbacks() if resolve == 3 else None
        q.Free()
        return {"workload": label or WORKLOADS[workload][1], "resolve_option": resolve, "value": m / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "roofline_substep_frac": ALGO_BYTES_PER_PARTICLE_SUBSTEP * m / (ms * 1e-3) / 1e9 / peak,
                "gs_substeps_that_fell_back": fell,
                "l2": "state larger than L2" if 28.0 * m > 126e6 else "state fits the 126 MB L2 (no flush): not an HBM-roofline number"}

    def other_configs():
        """Context for the headline, measured in this same run: the same field under the other resolve, the other single-GPU
        configurations of BASELINE.json, three points of the configs[4] density sweep, and one slab of the configs[3] field
        (the weak-scaling denominator of the multi-GPU runs: value AND end-to-end)."""
        from particlephysicsc_b200 import scenes
        out = {}
        other = 1 if args.resolve == 3 else 3
        out[f"{args.workload}_resolve{other}"] = quick_value(args.workload, resolve=other)
        for w in ("uniform16m", "uniform1m"):
            if w != args.workload:
                out[w] = quick_value(w)
        out[f"uniform16m_resolve{other}"] = quick_value("uniform16m", resolve=other)
        out["density_sweep_16m"] = [
            dict(quick_value("uniform16m", steps=8, scene=scenes.uniform_random(1 << 24, 2.0, phi, seed=0x5EED0005, vmax=0.0),
                             label=f"16M equal-radius particles (r=2), uniform random, area fraction {phi} (BASELINE configs[4])"),
                 area_fraction=phi) for phi in (0.05, 0.40, 0.80)]
        out["slab32m_one_slab"] = one_slab(api, torch, local_rank, 1 << 25, args.steps, args.settle, args.warmup, peak)
        return out

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, _, cpu = time_reference_cpu(args.workload, 3, 1)
        others = None
        if world == 1 and not args.no_other_configs:
            others = other_configs()
        g = p.grid_info()
        wave = p.wave_stats()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": descr, "particles_per_gpu": n, "box": [W, H], "dt": DT, "grid": [g.cols, g.rows],
                           "substeps_per_frame": SUBSTEPS_PER_FRAME, "settle_substeps": args.settle,
                           "resolve": RESOLVE_NAMES.get(args.resolve, str(args.resolve)),
                           "resolve_option": args.resolve,
                           "gs_substeps_that_fell_back_to_the_exact_resolve": p.gs_fallbacks() if args.resolve == 3 else None,
                           "gs_fallback_reasons": p.gs_fallback_reasons() if args.resolve == 3 else None,
                           "l2": "state is %.0f MB per GPU, larger than the 126 MB L2; no flush between steps" % (28.0 * n / 1e6),
                           "parallelism": "independent replicas" if world > 1 else "single GPU"},
                "roofline": roofline,
                "roofline_substep": {"bound": "hbm", "achieved": substep_gbs / world, "peak": peak, "unit": "GB/s",
                                     "frac": substep_gbs / world / peak, "traffic": substep_traffic,
                                     "note": "40 B x particles x substeps / time, whole substep, per GPU; traffic = DRAM bytes of all kernels of "
                                             "one substep (ncu capture under profiles/)"},
                "cpu_baseline": cpu,
                "other_configs": others,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h_per_step,
                        "note": "particlesUpdate + particlesCollisionsCheck_Grid x8, then p_x, p_y, colors copied to the host arrays "
                                "(ppc_download_async: a frame's copy overlaps the next frame's substeps; the last copy is waited for inside the timed region)"},
                "resolve_wavefront": None if wave is None else {"rounds_last_substep": wave[0], "particle_visits_last_substep": wave[1]},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    p.Free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_slab(args, rank: int, world: int, local_rank: int):
    """Slab decomposition: one slab of cell rows per rank, neighbour exchange of migrants + ghost rows every substep."""
    import torch
    from particlephysicsc_b200 import api, scenes, slab

    dist = None
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else "not requested"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=nccl_options())
    dev = torch.device("cuda", local_rank)
    # weak scaling: BASELINE configs[3], every rank generates its own slab; strong scaling: one of the single-GPU fields
    # (e.g. the 16M pile), generated identically on every rank and cut into slabs of equal particle count
    split = args.workload != "slab32m"
    n_per, descr = WORKLOADS[args.workload]
    n_per = args.particles_per_gpu or n_per
    rows_per = max(8, int(round(SLAB_ROWS_PER_GPU * n_per / float(1 << 25))))
    n_total = n_per if split else n_per * world
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
    opts.setdefault("resolve", args.resolve)
    gs = opts["resolve"] == 3
    whole = make_scene(args.workload, n_per) if split else None
    box_w = whole.W if split else SLAB_W

    def make_rank(k: int, nranks: int, halo: int):
        if split:
            r = slab.make_device_slabs(whole, nranks, halo, device_of=lambda _: local_rank, only_rank=k, options=opts, by_load=True)[0]
            return r, whole.H
        H = rows_per * SLAB_CELL * nranks
        sc = scenes.uniform_slab(n_per, SLAB_RADIUS, SLAB_W, H, k * rows_per * SLAB_CELL, (k + 1) * rows_per * SLAB_CELL,
                                 seed=0x5EED0004 + 0x100 * k)
        ids = (np.arange(n_per, dtype=np.int64) + k * n_per).astype(np.uint32)
        ghosts = int(2.2 * halo * n_per / rows_per)
        cap = int(1.15 * (n_per + 3 * ghosts)) + 65536
        r = slab.DeviceSlab(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba, ids, SLAB_W, H, 6.0, (k * rows_per, (k + 1) * rows_per), halo,
                            capacity=cap, device=local_rank, options=opts)
        return r, H

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(me, tr, steps, colour_every=0):
        for s in range(steps):
            colour = bool(colour_every) and (s % colour_every == colour_every - 1)
            if tr is None:
                slab.local_step([me], DT, colour)
            else:
                slab.slab_step(me, tr, DT, colour)

    def settle_and_choose_halo(me, tr, group_max):
        """Relax the random field with a generous halo, then restart from the relaxed state with the halo this field needs:
        the rows the resolve's inexactness mark travels into the window (csrc/ppc_collide.cuh ST_TAINT), plus a margin."""
        run(me, tr, args.settle)
        travelled = rounds = 0
        for _ in range(6):
            run(me, tr, 1)
            i = me.info()
            travelled, rounds = max(travelled, int(i.taint_rows)), max(rounds, int(i.rounds))
        travelled, rounds = group_max(travelled), group_max(rounds)
        halo = slab.required_halo(travelled, args.halo_margin)
        me.reinit(halo)
        if tr is not None:
            tr.reset()
        return halo, rounds, travelled

    def timed(me, tr, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(me.stream)
        run(me, tr, steps)
        e1.record(me.stream)
        e1.synchronize()
        barrier()
        return e0.elapsed_time(e1)

    def group_max(v: int) -> int:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # set-up only; nothing on the data path is a collective
        return int(t.item())

    equivalence = None
    if dist is not None and not args.no_equivalence:
        equivalence = slab_equivalence(api, torch, dist, rank, world, local_rank)
    tr = slab.DistTransport(dist, dev) if dist is not None else None
    if gs:   # "resolve"=3: the ghost rows the tile order needs (exact by construction: nothing to measure, nothing to retry)
        if split:
            _, ranges, _, _ = slab.split_scene(whole, world, 1, True)
            _, _, rows_all = slab.grid_dims(whole.W, whole.H, float(max(6.0, float(np.max(whole.radius)))))
            opts.update(slab.gs_options(whole.n, whole.W, whole.H, float(max(6.0, float(np.max(whole.radius))))))
        else:
            ranges, rows_all = [(k * rows_per, (k + 1) * rows_per) for k in range(world)], rows_per * world
            opts.update(slab.gs_options(n_per * world, SLAB_W, rows_all * SLAB_CELL, 6.0))
        halo, rounds, travelled = (slab.gs_required_halo(ranges, rows_all) if world > 1 else 1), 0, 0
        me, H = make_rank(rank, world, halo)
        run(me, tr, args.settle)
    else:
        me, H = make_rank(rank, world, args.settle_halo)
        me.p.set_option("slab_strict", 0)   # a halo that turns out too thin is counted, then the phase is redone with a wider one
        halo, rounds, travelled = settle_and_choose_halo(me, tr, group_max)
    halo_retries = 0
    while True:
        run(me, tr, args.warmup)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = api.kernel_launches()
        ms_total = max_over_ranks(timed(me, tr, args.steps), dist, "cuda")
        launches = api.kernel_launches() - launches0
        clocks = sampler.stop() if rank == 0 else None
        if group_max(int(me.info().violations)) == 0 or halo_retries >= 3:
            break
        halo_retries += 1   # every rank takes this branch together: widen the halo, restart from the current state, measure again
        halo += 4
        me.reinit(halo)
        me.p.set_option("slab_strict", 0)
        if tr is not None:
            tr.reset()
    exact = group_max(int(me.info().violations)) == 0
    value = n_total * args.steps / (ms_total * 1e-3)

    # e2e: owned particles copied back to the host arrays every 8th substep (colours computed on that substep)
    frames = max(1, args.steps // SUBSTEPS_PER_FRAME)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(me.stream)
    n_down = 0
    for _ in range(frames):
        run(me, tr, SUBSTEPS_PER_FRAME, colour_every=SUBSTEPS_PER_FRAME)
        n_down = me.sync_host_async()   # overlaps the next frame's substeps
    me.sync_host_wait()
    e1.record(me.stream)
    e1.synchronize()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), dist, "cuda")
    e2e_value = n_total * frames * SUBSTEPS_PER_FRAME / (e2e_ms * 1e-3)

    # per-stage times of this rank (serialised) and exchange volume
    prof_steps = min(args.steps, 8)
    me.p.profile_reset()
    me.p.profile(True)
    run(me, tr, prof_steps)
    me.synchronize()
    me.p.profile(False)
    stages = {k: (ms / prof_steps, cnt // prof_steps) for k, (ms, cnt) in me.p.profile_read().items() if cnt}
    info = me.info()
    sent = me.counts
    phase_ms = None
    if tr is not None:   # where a distributed substep spends its wall time (every phase synchronised: diagnostics, not the metric)
        timers = {}
        for _ in range(prof_steps):
            slab.slab_step(me, tr, DT, False, timers)
        phase_ms = {k: round(v / prof_steps, 4) for k, v in timers.items()}
    peak, peak_src = measured_peak_gbs()

    # the same per-GPU workload on ONE GPU (a single slab, no neighbours), measured by rank 0 in this same job: the
    # denominator of the parallel efficiency
    ref_value = ref_e2e = None
    barrier()
    if rank == 0 and world > 1 and not args.no_scaling_reference:
        me.free()
        if split:
            solo, _ = make_rank(0, 1, 4)
            run(solo, None, args.settle + args.warmup)
            e0.record(solo.stream)
            run(solo, None, args.steps)
            e1.record(solo.stream)
            e1.synchronize()
            ref_value = n_total * args.steps / (e0.elapsed_time(e1) * 1e-3)
            solo.free()
        else:
            ref = one_slab(api, torch, local_rank, n_per, args.steps, args.settle, args.warmup, peak, opts=opts)
            ref_value, ref_e2e = ref["value"], ref["e2e"]["value"]
    if dist is not None:
        dist.barrier()

    if rank == 0:
        dom = max(stages, key=lambda k: stages[k][0]) if stages else None
        roofline = None
        if dom:
            algo = {"integrate_walls_keys": 36.0, "collide": 40.0, "resolve_wavefront": 16.0, "canonical_gather": 60.0, "cell_sort": 8.0}.get(
                dom, ALGO_BYTES_PER_PARTICLE_SUBSTEP)
            ach = algo * info.n_window / (stages[dom][0] * 1e-3) / 1e9
            traffic = None
            try:   # bytes per launch from the ncu --set full capture of one 2**25-particle slab (profiles/)
                with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                    traffic = json.load(f).get(f"slab32m_resolve{opts['resolve']}", {}).get(dom) if (not split and n_per == 1 << 25) else None
            except Exception:
                traffic = None
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                        "peak_source": peak_src, "algorithmic_bytes_per_particle": algo,
                        "stage_ms_per_substep": {k: round(v[0], 4) for k, v in stages.items()}}
        substep_gbs = ALGO_BYTES_PER_PARTICLE_SUBSTEP * (n_total / world) * args.steps / (ms_total * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if split else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": descr, "particles_per_gpu": n_total // world, "particles": n_total, "box": [box_w, H],
                           "grid": [info.cols, info.rows], "rows_per_gpu": rows_per, "halo_rows": halo, "dependency_rounds": rounds,
                           "inexactness_travel_rows": travelled, "halo_retries": halo_retries,
                           "owned_particles_exact_in_timed_region": exact, "dt": DT,
                           "substeps_per_frame": SUBSTEPS_PER_FRAME, "settle_substeps": args.settle,
                           "resolve": RESOLVE_NAMES.get(int(opts.get("resolve", args.resolve)), "?"), "resolve_option": int(opts.get("resolve", args.resolve)),
                           "l2": "state is %.0f MB per GPU, larger than the 126 MB L2; no flush between steps" % (28.0 * n_per / 1e6),
                           "parallelism": f"{world} slabs of cell rows, neighbour send/recv (NCCL) of migrants + {halo} ghost rows per substep"
                                          if world > 1 else "one slab (no neighbours)",
                           "records_sent_per_substep_rank0": list(sent), "bytes_sent_per_substep_rank0": 32 * sum(sent),
                           "window_particles_rank0": int(info.n_window), "phase_wall_ms_rank0_synchronised": phase_ms},
                "roofline": roofline,
                "roofline_substep": {"bound": "hbm", "achieved": substep_gbs, "peak": peak, "unit": "GB/s", "frac": substep_gbs / peak,
                                     "note": "40 B x owned particles x substeps / time, whole substep, per GPU"},
                "scaling_reference": None if ref_value is None else {
                    "n_gpus": 1, "value": ref_value, "unit": UNIT,
                    "note": ("the whole field as one slab on one GPU, measured by rank 0 in this job; parallel efficiency = value / (n_gpus x this)"
                             if split else "the same per-GPU workload as one slab on one GPU, measured by rank 0 in this job; "
                             "parallel efficiency = value / (n_gpus x this)"),
                    "parallel_efficiency": value / (world * ref_value),
                    "e2e": None if ref_e2e is None else {"value": ref_e2e, "unit": UNIT, "parallel_efficiency": e2e_value / (world * ref_e2e)}},
                "equivalence": equivalence, "host_binding_rank0": numa,
                "cpu_baseline": None,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 12.0 * n_down / SUBSTEPS_PER_FRAME,
                        "note": "8 slab substeps, then the owned particles' p_x, p_y, colors copied to the host arrays (asynchronously: a frame's "
                                "copy overlaps the next frame's substeps; the last copy is waited for inside the timed region)"},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            _, _, line["cpu_baseline"] = time_reference_cpu("slab32m", 2, 1)
        print(json.dumps(line), flush=True)
    if ref_value is None or rank != 0:
        try:
            me.free()
        except Exception:
            pass
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--settle", type=int, default=64, help="untimed substeps that relax the synthetic field first")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: pile16m on one GPU (the configuration the roofline target is stated on), slab32m on several")
    ap.add_argument("--particles-per-gpu", type=int, default=0, help="slab32m only: override 2**25")
    ap.add_argument("--settle-halo", type=int, default=40, help="slab32m: ghost rows while the random field relaxes")
    ap.add_argument("--halo-margin", type=int, default=3, help="slab32m: rows added to the halo the measured travel of the inexactness mark requires")
    ap.add_argument("--no-scaling-reference", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not bind each rank to the CPU cores next to its GPU")
    ap.add_argument("--no-equivalence", action="store_true", help="skip the N-slabs == one-GPU digest check that precedes the timing")
    ap.add_argument("--replicas", action="store_true", help="with --gpus N and a single-GPU workload: N independent copies instead of N slabs of one field")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (ppc_set_option), e.g. collide=3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the context measurements of the other single-GPU configurations")
    ap.add_argument("--resolve", type=int, default=0,
                    help="3 = tile Gauss-Seidel (deterministic bounded passes; default on one GPU), 1 = sequential-equivalent wavefront "
                         "(bit-exact replay of the reference's pair order; default for slabs), 2 = Jacobi")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "pile16m" if args.gpus == 1 else "slab32m"
    if args.resolve == 0:
        args.resolve = 1 if (args.workload == "slab32m" or (args.gpus > 1 and not args.replicas)) else 3
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "slab32m" or (world > 1 and not args.replicas):
        run_slab(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

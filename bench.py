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
CPU_SAMPLE_PARTICLES = 1 << 20


def make_scene(workload: str, n: int, seed_offset: int = 0):
    from particlephysicsc_b200 import scenes
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
        n = 1 << max(14, min(20, int(np.floor(np.log2(max(want, 1.0))))))
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
    return value, dt / steps * 1e3, {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample}


def run_reference(args, rank: int):
    if rank != 0:
        return
    value, ms, base = time_reference_cpu(args.workload, args.steps, args.warmup, budget_s=90.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOADS[args.workload][1], "sample": base["sample"]},
            "cpu_baseline": base, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    from particlephysicsc_b200 import api

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_per_gpu, descr = WORKLOADS[args.workload]
    n = n_per_gpu
    sc = make_scene(args.workload, n, seed_offset=0x100 * rank)
    p = api.Particles(n, device=local_rank)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.set_option("resolve", args.resolve)
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
        p.download(mask)
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
    algo = {"integrate_walls_keys": 36.0, "collide": 40.0}.get(dom, ALGO_BYTES_PER_PARTICLE_SUBSTEP)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get(dom)
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

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, _, cpu = time_reference_cpu(args.workload, 3, 1)
        g = p.grid_info()
        wave = p.wave_stats()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": descr, "particles_per_gpu": n, "box": [W, H], "dt": DT, "grid": [g.cols, g.rows],
                           "substeps_per_frame": SUBSTEPS_PER_FRAME, "settle_substeps": args.settle,
                           "resolve": {1: "sequential-equivalent wavefront (bit-exact)", 2: "jacobi"}.get(args.resolve, str(args.resolve)),
                           "l2": "state is %.0f MB per GPU, larger than the 126 MB L2; no flush between steps" % (28.0 * n / 1e6),
                           "parallelism": "independent replicas" if world > 1 else "single GPU"},
                "roofline": roofline,
                "roofline_substep": {"bound": "hbm", "achieved": substep_gbs / world, "peak": peak, "unit": "GB/s",
                                     "frac": substep_gbs / world / peak, "note": "40 B x particles x substeps / time, whole substep, per GPU"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h_per_step,
                        "note": "particlesUpdate + particlesCollisionsCheck_Grid x8, then p_x, p_y, colors copied to the host arrays"},
                "resolve_wavefront": None if wave is None else {"rounds_last_substep": wave[0], "particle_visits_last_substep": wave[1]},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    p.Free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--settle", type=int, default=64, help="untimed substeps that relax the synthetic field first")
    ap.add_argument("--workload", default="pile16m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resolve", type=int, default=1, help="1 = sequential-equivalent wavefront (default, bit-exact), 2 = Jacobi")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

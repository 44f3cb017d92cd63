#!/usr/bin/env python
"""BASELINE configs[4]: packing-density sweep at 16M particles (r = 2, uniform random centres, area fraction 5 % .. 80 %).

    python tools/density_sweep.py [--n 16777216] [--steps 16] > profiles/r01_density_sweep_16m.json

For every area fraction: settle, then time `steps` substeps with the state resident (CUDA events on the library's stream) and
read the per-stage times; prints one JSON document.  Low fractions show the streaming pipeline (few contacts), high
fractions the contact kernels (collision-bound regime)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from particlephysicsc_b200 import api, scenes  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--settle", type=int, default=32)
    ap.add_argument("--phis", default="0.05,0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8")
    args = ap.parse_args()
    peak, _ = bench.measured_peak_gbs()
    rows = []
    for phi in [float(x) for x in args.phis.split(",")]:
        sc = scenes.uniform_random(args.n, 2.0, phi, seed=0x5EED0005, vmax=0.0)
        p = api.Particles(args.n)
        p.set_sync_policy(api.PPC_SYNC_MANUAL)
        p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
        p.run_substeps(sc.W, sc.H, bench.DT, args.settle)
        ext = torch.cuda.ExternalStream(p.stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        p.run_substeps(sc.W, sc.H, bench.DT, args.steps)
        e1.record(ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        p.profile_reset()
        p.profile(True)
        p.run_substeps(sc.W, sc.H, bench.DT, 4)
        p.synchronize()
        p.profile(False)
        stages = {k: round(v / 4, 4) for k, (v, c) in p.profile_read().items() if c}
        pairs = int(p.lib.ppc_debug_pairs(p.s, None, 0))
        wave = p.wave_stats()
        g = p.grid_info()
        rows.append({"area_fraction": phi, "box": [sc.W, sc.H], "particles_per_cell": args.n / float(g.cells), "pairs_last_substep": pairs,
                     "contacts_per_particle": 2.0 * pairs / args.n, "ms_per_substep": ms, "particle_substeps_per_s": args.n / (ms * 1e-3),
                     "roofline_substep_frac": bench.ALGO_BYTES_PER_PARTICLE_SUBSTEP * args.n / (ms * 1e-3) / 1e9 / peak,
                     "wavefront_rounds": wave[0] if wave else None, "stage_ms": stages})
        print(f"phi {phi}: {ms:.3f} ms/substep, {pairs} pairs", file=sys.stderr)
        p.Free()
    print(json.dumps({"workload": f"{args.n} particles, r = 2, uniform random, area-fraction sweep (BASELINE configs[4])", "dt": bench.DT,
                      "settle_substeps": args.settle, "timed_substeps": args.steps, "peak_gbs": peak, "points": rows}, indent=1))


if __name__ == "__main__":
    main()

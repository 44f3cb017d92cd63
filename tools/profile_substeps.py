#!/usr/bin/env python
"""Minimal driver for profilers: build one synthetic field, settle it, run a few substeps through the C ABI.

    python tools/profile_substeps.py --workload pile16m --settle 16 --steps 4
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_substeps.py ...
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from particlephysicsc_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pile16m", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--settle", type=int, default=16)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--resolve", type=int, default=1)
    ap.add_argument("--collide", type=int, default=0)
    args = ap.parse_args()
    n = bench.WORKLOADS[args.workload][0]
    sc = bench.make_scene(args.workload, n)
    p = api.Particles(n)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.set_option("resolve", args.resolve)
    p.set_option("collide", args.collide)
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    p.run_substeps(sc.W, sc.H, bench.DT, args.settle)
    p.synchronize()
    p.run_substeps(sc.W, sc.H, bench.DT, args.steps)
    p.synchronize()
    print("done", args.workload, "rounds/visits of last resolve:", p.wave_stats())
    p.Free()


if __name__ == "__main__":
    main()

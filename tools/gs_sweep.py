#!/usr/bin/env python
"""Sweep "resolve"=3 tile parameters on one settled field, same store, CUDA-event stage times per substep.

    python tools/gs_sweep.py --workload pile16m --wc 0,16,24,33,49 --strip 0,700,850
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from particlephysicsc_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pile16m", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--settle", type=int, default=64)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--wc", default="0")
    ap.add_argument("--strip", default="0")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    n = bench.WORKLOADS[args.workload][0]
    sc = bench.make_scene(args.workload, n)
    p = api.Particles(n)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.set_option("resolve", 3)
    for kv in args.opt:
        k, v = kv.split("=")
        p.set_option(k, int(v))
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    p.run_substeps(sc.W, sc.H, bench.DT, args.settle)
    p.synchronize()
    for wc in [int(x) for x in args.wc.split(",")]:
        for strip in [int(x) for x in args.strip.split(",")]:
            p.set_option("gs_wc", wc)
            p.set_option("gs_strip", strip)
            p.run_substeps(sc.W, sc.H, bench.DT, 2)
            p.synchronize()
            p.profile_reset()
            p.profile(True)
            p.run_substeps(sc.W, sc.H, bench.DT, args.steps)
            p.synchronize()
            p.profile(False)
            st = {k: round(ms / args.steps, 4) for k, (ms, cnt) in p.profile_read().items() if cnt}
            print(json.dumps({"workload": args.workload, "wc": p.gs_params()["wc"], "strip": p.gs_params()["strip_limit"], "fallbacks": p.gs_fallbacks(),
                              "total": round(sum(st.values()), 4), "stages": st}), flush=True)
    p.Free()


if __name__ == "__main__":
    main()

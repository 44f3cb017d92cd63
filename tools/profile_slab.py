#!/usr/bin/env python
"""Minimal single-process driver of one slab of the configs[3] field for profilers (no neighbours, no NCCL):

    ncu --set full -k regex:k_collide_tile2 -s 8 -c 1 -o gpurun_out/prof python tools/profile_slab.py [particles] [substeps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from particlephysicsc_b200 import scenes, slab  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rows = max(8, int(round(bench.SLAB_ROWS_PER_GPU * n / float(1 << 25))))
    H = rows * bench.SLAB_CELL
    sc = scenes.uniform_slab(n, bench.SLAB_RADIUS, bench.SLAB_W, H, 0, H, seed=0x5EED0004)
    me = slab.DeviceSlab(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba, np.arange(n, dtype=np.uint32), bench.SLAB_W, H, 6.0, (0, rows), 4,
                         capacity=int(n * 1.05) + 65536)
    for _ in range(steps):
        slab.local_step([me], bench.DT)
    me.synchronize()
    print("done", n, "particles,", steps, "substeps; rounds of the last resolve:", me.info().rounds)
    me.free()


if __name__ == "__main__":
    main()

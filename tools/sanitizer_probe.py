#!/usr/bin/env python
"""Small run of every kernel on the default paths, for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_probe.py
Single store (pile, uniform, slot-0/add path), then two slabs in one process with migration."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlephysicsc_b200 import api, scenes, slab  # noqa: E402

DT = 1.0 / 144.0 / 8.0


def main():
    for sc in (scenes.polydisperse_pile(30000, seed=1, vmax=5.0), scenes.uniform_random(30000, 2.0, 0.5, seed=2, vmax=100.0),
               scenes.uniform_random(20000, 1.0, 0.4, seed=3, vmax=50.0)):
        for opts in ({}, {"wave_mode": 0}, {"collide": 3}, {"collide": 1}, {"resolve": 3}, {"resolve": 3, "gs_wc": 3}, {"resolve": 3, "gs_merge": 1},
                     {"resolve": 3, "scan_single": 0, "cell_ranks": 0, "defer_state": 0}):
            p = api.Particles(sc.n)
            p.set_sync_policy(api.PPC_SYNC_MANUAL)
            for k, v in opts.items():
                p.set_option(k, v)
            p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
            p.run_substeps(sc.W, sc.H, DT, 3, True)
            p.debug_pairs()
            p.download(api.PPC_DL_ALL)
            p.Free()
    q = api.Particles(64)
    for f in range(40):   # growth + dormant newest particle
        q.AddParticle(100.0 + f, 50.0, 5.0, 3.0, 10)
        q.Update(DT)
        q.CollisionsCheck_Grid(1200, 675, DT)
    q.RemoveParticle()
    q.Update(DT)
    q.CollisionsCheck_Grid(1200, 675, DT)
    q.Free()
    sc = scenes.uniform_random(40000, 2.0, 0.35, seed=4, vmax=300.0)
    ranks = slab.make_device_slabs(sc, 2, 6)
    for _ in range(4):
        slab.local_step(ranks, DT, True)
    got = sum(len(r.download()["ids"]) for r in ranks)
    assert got == sc.n
    for r in ranks:
        r.debug_pairs()
        r.free()
    print("sanitizer probe done")


if __name__ == "__main__":
    main()

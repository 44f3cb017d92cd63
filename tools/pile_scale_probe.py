#!/usr/bin/env python
"""Per-stage times of the settled polydisperse pile at 1M .. 16M particles: does the velocity wavefront get cheaper per pair
once its records fit the 126 MB L2?  (profiles/README.md quotes the result.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from particlephysicsc_b200 import api, scenes  # noqa: E402


def main():
    for n in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
        sc = scenes.polydisperse_pile(n, seed=0x5EED0003)
        p = api.Particles(n)
        p.set_sync_policy(api.PPC_SYNC_MANUAL)
        p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
        p.run_substeps(sc.W, sc.H, bench.DT, 40)
        p.profile_reset()
        p.profile(True)
        p.run_substeps(sc.W, sc.H, bench.DT, 8)
        p.synchronize()
        p.profile(False)
        st = {k: round(v / 8, 4) for k, (v, c) in p.profile_read().items() if c}
        pairs = int(p.lib.ppc_debug_pairs(p.s, None, 0))
        print(n, "pairs", pairs, "rounds", p.wave_stats()[0], "wavefront ns/pair", round(st["resolve_wavefront"] * 1e6 / pairs, 2),
              "narrow phase ns/particle", round(st["collide"] * 1e6 / n, 3), st, flush=True)
        p.Free()


if __name__ == "__main__":
    main()

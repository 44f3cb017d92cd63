#!/usr/bin/env python
"""Per-stage timing of the reference's debug scene (BASELINE configs[0]) through the drop-in entry points."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlephysicsc_b200 import api, scenes  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4500
    W, H, dt = 1200, 675, 1.0 / 144.0
    p = api.Particles(50000)
    t0 = time.perf_counter()
    for f, (x, y, m, r, cnt) in enumerate(scenes.debug_scene_frames(frames=frames, W=W, H=H)):
        p.AddParticle(x, y, m, r, cnt)
        if f == frames - 200:
            p.synchronize()
            t1 = time.perf_counter()
            p.profile_reset()
            p.profile(True)
        p.Update(dt)
        p.CollisionsCheck_Grid(W, H, dt)
    p.synchronize()
    t2 = time.perf_counter()
    p.profile(False)
    print(f"{frames} frames, {p.length} particles: {t1 - t0:.2f} s unprofiled for the first {frames - 200}, "
          f"{(t2 - t1) / 200 * 1e3:.3f} ms/frame over the last 200 (profiled, serialised)")
    for k, (ms, cnt) in p.profile_read().items():
        if cnt:
            print(f"  {k:22s} {ms / 200:8.4f} ms/frame  {cnt // 200} launches/frame")
    print("wavefront rounds/visits of the last frame:", p.wave_stats())
    p.Free()


if __name__ == "__main__":
    main()

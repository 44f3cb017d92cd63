import sys; sys.path.insert(0,'/root/repo')
import bench
from particlephysicsc_b200 import api, scenes
for n in (1<<20, 1<<21, 1<<22, 1<<23, 1<<24):
    sc = scenes.polydisperse_pile(n, seed=0x5EED0003)
    p = api.Particles(n); p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    p.run_substeps(sc.W, sc.H, bench.DT, 40)
    p.profile_reset(); p.profile(True); p.run_substeps(sc.W, sc.H, bench.DT, 8); p.synchronize(); p.profile(False)
    st = {k: round(v/8,4) for k,(v,c) in p.profile_read().items() if c}
    pairs = int(p.lib.ppc_debug_pairs(p.s, None, 0))
    print(n, "pairs", pairs, "rounds", p.wave_stats()[0], "wave ns/pair", round(st["resolve_wavefront"]*1e6/pairs,2), "collide ns/particle", round(st["collide"]*1e6/n,3), st, flush=True)
    p.Free()

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each file holds a seeded input state, the call parameters and what the reference produced: the state after
particlesUpdate, the state after particlesCollisionsCheck_Grid, the candidate-pair list captured at the
particlesResolveVelocity boundary (reference src/particles_solver.c:386), and the state after `steps` substeps.
The fixtures pin oracle/ppc_oracle.c (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py)
wherever /root/reference is not available.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, State, ParticlesT, Vector2  # noqa: E402
from particlephysicsc_b200 import scenes  # noqa: E402
import ctypes as C  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("px", "py", "vx", "vy", "mass", "radius", "rgba")


def pack(prefix, s):
    return {f"{prefix}_{f}": getattr(s, f).copy() for f in FIELDS}


def run_case(R, name, s0, W, H, dt, steps):
    out = dict(W=W, H=H, dt=np.float32(dt), steps=steps)
    out.update(pack("in", s0))
    s = s0.copy()
    R.update(s, dt)
    out.update(pack("upd", s))
    pairs = R.collisions_check_grid(s, W, H, dt, capture=True)
    out["pairs"] = pairs if pairs is not None else np.zeros((0, 2), np.int32)
    out.update(pack("col", s))
    for _ in range(steps - 1):
        R.substep(s, W, H, dt)
    out.update(pack("end", s))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: n={s0.n} pairs={len(out['pairs'])} steps={steps}")


def st(sc):
    return State(sc.px.copy(), sc.py.copy(), sc.vx.copy(), sc.vy.copy(), sc.mass.copy(), sc.radius.copy(), sc.rgba.copy())


def main():
    R = Reference()
    dt = 1.0 / 144.0
    # 1. equal radii, dense enough for multi-contact particles
    sc = scenes.uniform_random(300, 2.0, 0.55, seed=0x5EED1001, vmax=40.0)
    run_case(R, "uniform_300", st(sc), sc.W, sc.H, dt, 5)
    # 2. radii beyond the 6.0 floor: cell size follows the largest radius (reference :231-244)
    sc = scenes.uniform_random(257, 2.0, 0.45, seed=0x5EED1002, vmax=25.0)
    rng = scenes.Stream(0x5EED1003)
    sc.radius[:] = 1.5 + rng.unit(sc.n) * 7.5
    sc.px[:] = np.clip(sc.px, sc.radius, sc.W - sc.radius)
    sc.py[:] = np.clip(sc.py, sc.radius, sc.H - sc.radius)
    run_case(R, "polydisperse_257", st(sc), sc.W, sc.H, dt, 5)
    # 3. wall hits on all four sides, including particles outside the box and both x tests firing at once
    sc = scenes.uniform_random(64, 3.0, 0.2, seed=0x5EED1004, vmax=300.0)
    s = st(sc)
    s.px[:8] = [-5.0, 0.0, 3.0, sc.W - 3.0, sc.W + 10.0, sc.W, 2.9, 1.0]
    s.py[8:16] = [-1.0, 0.0, 3.0, sc.H - 3.0, sc.H + 7.0, sc.H, 2.5, 1.0]
    run_case(R, "walls_64", s, sc.W, sc.H, dt, 3)
    # 4. a box narrower than one particle: x <= r and x >= W - r both fire (stale locals, reference :208-219)
    s = State.empty(4)
    s.px[:] = [1.0, 2.0, 3.0, 1.5]
    s.py[:] = [10.0, 30.0, 50.0, 70.0]
    s.vx[:] = [5.0, -5.0, 2.0, 0.0]
    s.vy[:] = [1.0, 2.0, 3.0, 4.0]
    s.mass[:] = [5, 6, 7, 8]
    s.radius[:] = [2.0, 2.0, 2.5, 3.0]
    run_case(R, "narrow_box_4", s, 3, 100, dt, 2)
    # 5. tiny counts: N = 1 skips the whole collisions check (reference :204-205), N = 2 is the smallest real case
    for n in (1, 2):
        s = State.empty(n)
        s.px[:] = [10.0, 12.5][:n]
        s.py[:] = [-3.0, -2.0][:n]   # above the top wall: only N >= 2 gets clamped
        s.vx[:] = [3.0, -1.0][:n]
        s.vy[:] = [-2.0, 4.0][:n]
        s.mass[:] = [5, 9][:n]
        s.radius[:] = [2.0, 3.0][:n]
        run_case(R, f"tiny_{n}", s, 200, 100, dt, 2)
    # 6. coincident centres (particlesAddParticle writes `count` copies at one point, src/particles_arrays.c:265-277)
    s = State.empty(40)
    for c in range(4):
        s.px[c * 10:(c + 1) * 10] = 20.0 + 7.0 * c
        s.py[c * 10:(c + 1) * 10] = 15.0 + 3.0 * c
        s.mass[c * 10:(c + 1) * 10] = 5 + c
        s.radius[c * 10:(c + 1) * 10] = 3.0 + c
    run_case(R, "coincident_40", s, 120, 80, dt, 4)
    # 7. mass values that are not integers / exceed a byte: read through uint8_t in the resolve (reference :141-142)
    sc = scenes.uniform_random(128, 2.5, 0.6, seed=0x5EED1007, vmax=30.0)
    s = st(sc)
    s.mass[:] = (5.0 + scenes.Stream(0x5EED1008).unit(128) * 300.0)
    s.mass[s.mass.astype(np.int64) % 256 == 0] += 1.0   # a zero byte would divide by zero on both sides; keep it finite
    run_case(R, "odd_mass_128", s, sc.W, sc.H, dt, 3)
    # 8. the store: particlesCreate / particlesAddParticle / particlesRemoveParticle through the reference itself,
    #    interleaved with substeps like src/main.c:80-98 (slot 0 never written, newest particle dormant)
    lib = R.lib
    p = lib.particlesCreate(64)
    # The reference never initialises its arrays (aligned_alloc, src/particles_arrays.c:147-183) yet slot 0 is
    # processed without ever being written (src/particles_arrays.c:266-269).  A fresh 200 KB allocation is zero pages
    # in practice; the B200 store DEFINES it as zero.  Zero the reference's arrays here so the
    # fixture does not depend on heap history.
    for a in ("p_x", "p_y", "v_x", "v_y", "mass", "radius", "colors"):
        C.memset(getattr(p, a), 0, 4 * p.size)
    ops, snaps = [], []
    frames = list(scenes.debug_scene_frames(frames=12, per_frame=10, W=240, H=135, seed=0x5EED1009))
    for f, (x, y, m, r, cnt) in enumerate(frames):
        lib.particlesAddParticle(C.byref(p), Vector2(x, y), m, r, cnt)
        ops.append((x, y, m, r, cnt))
        if f == 7:
            lib.particlesRemoveParticle(C.byref(p))
        if p.length > 0:
            lib.particlesUpdate(C.byref(p), dt)
            lib.particlesCollisionsCheck_Grid(C.byref(p), 240, 135, dt)
        n = p.length
        snap = np.stack([np.ctypeslib.as_array(getattr(p, a), shape=(p.size,))[: n + 1].copy() for a in ("p_x", "p_y", "v_x", "v_y", "mass", "radius")])
        snaps.append(snap)
    np.savez_compressed(os.path.join(HERE, "store_frames_12.npz"), ops=np.array(ops, np.float32), remove_at=7, W=240, H=135,
                        dt=np.float32(dt), lengths=np.array([s.shape[1] - 1 for s in snaps]), size=p.size,
                        **{f"snap_{i}": s for i, s in enumerate(snaps)})
    print("store_frames_12: final length", p.length, "size", p.size)
    lib.particlesFree(C.byref(p))


if __name__ == "__main__":
    main()

"""GPU: the "resolve"=3 path (csrc/ppc_gs.cuh: narrow phase + position pushes + tile Gauss-Seidel velocity resolve in one
shared-memory kernel) through the C ABI.

Acceptance protocol:
  * integer work (keys, sorted order, cell-start table, candidate pairs) bit-exact against the reference's;
  * the pair set the HOT kernel itself resolved equals the reference's pair set;
  * positions bit-exact against the reference after a substep;
  * positions AND velocities bit-exact against the oracle's CPU model of the device's pair order
    (oracle/ppc_oracle.c:ppc_oracle_tilegs_order) over several substeps, for several tile widths incl. strips;
  * 1000 substeps from non-overlapping lattices (area fraction 0.30 and 0.60): kinetic energy and both momenta within 1 %
    of the compiled reference, with the reference's spread under a reversed / shuffled pair list printed beside it;
  * dense pile: deviation of the same order as the reference's own reordering spread, measured in the same test; finite
    for 1000 substeps at the 16M-particle headline size.
"""
import numpy as np
import pytest

from conftest import assert_bits_equal, state_from_scene
from oracle.oracle import State
from particlephysicsc_b200 import api, scenes
from test_gpu_parity import DT, SCENES, assert_state_bits, gpu_state, gpu_store

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def gs_store(s: State, wc: int = 0, dump: bool = False) -> api.Particles:
    p = gpu_store(s)
    p.set_option("resolve", 3)
    if wc:
        p.set_option("gs_wc", wc)
    if dump:
        p.set_option("gs_dump", 1)
    return p


def sorted_pairs(pairs):
    return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))] if len(pairs) else pairs


# ---- bit-exact against the CPU model of the device's order, integer work against the reference's ---------------------------
@pytest.mark.parametrize("wc", [0, 2, 3, 7, 30])
@pytest.mark.parametrize("name,make", SCENES)
def test_matches_the_cpu_model_of_its_pair_order(oracle, name, make, wc):
    sc = make()
    ref = state_from_scene(sc)
    p = gs_store(ref, wc, dump=True)
    for step in range(4):   # later steps start from cell-sorted storage and from velocities the previous resolve left
        p.Update(DT)
        p.CollisionsCheck_Grid(sc.W, sc.H, DT)
        par = p.gs_params()
        assert wc == 0 or par["wc"] == wc
        oracle.set_tilegs(par["wc"], par["tile_r"], par["strip_limit"])
        oracle.update(ref, DT, False)
        seq = ref.copy()
        res = oracle.collisions_check_grid(seq, sc.W, sc.H, DT, mode=0, want_grid=True)
        oracle.collisions_check_grid(ref, sc.W, sc.H, DT, mode=3)
        g = p.debug_grid()
        assert np.array_equal(g["keys"], res["keys"]), f"{name} step {step}: cell keys"
        assert np.array_equal(g["order"], res["order"]), f"{name} step {step}: sorted order"
        assert np.array_equal(g["cell_start"], res["cell_start"]), f"{name} step {step}: cell-start table"
        assert np.array_equal(p.debug_pairs(), res["pairs"]), f"{name} step {step}: candidate-pair list"
        assert np.array_equal(p.debug_hot_pairs(), sorted_pairs(res["pairs"])), f"{name} step {step}: pairs the hot kernel resolved"
        got = gpu_state(p)
        assert_bits_equal(got.px, seq.px, f"{name} step {step}: p_x vs the sequential reference")
        assert_bits_equal(got.py, seq.py, f"{name} step {step}: p_y vs the sequential reference")
        assert_state_bits(got, ref, f"{name} step {step} wc={par['wc']} vs the tile Gauss-Seidel model")
    p.Free()


def test_hot_pairs_of_the_exact_path_equal_the_reference_pair_set(oracle):
    """The same check for the default path: the contact lists k_collide_tile2 wrote hold exactly the reference's pairs,
    including an r = 1 field at coordinates above 60,000 px where one float ulp is 2**-8 px (the fine-bin margin's edge)."""
    far = scenes.uniform_random(30000, 1.0, 0.40, seed=411, vmax=50.0)
    off_x, off_y = 65535 - far.W, 65535 - far.H
    far.px += np.float32(off_x)
    far.py += np.float32(off_y)
    cases = [(n, m()) for n, m in SCENES] + [("far_corner_r1", far)]
    for name, sc in cases:
        W, H = (65535, 65535) if name == "far_corner_r1" else (sc.W, sc.H)
        for resolve in (1, 3):
            ref = state_from_scene(sc)
            p = gpu_store(ref)
            p.set_option("resolve", resolve)
            p.set_option("gs_dump", 1)
            for step in range(2):
                oracle.update(ref, DT, False)
                res = oracle.collisions_check_grid(ref, W, H, DT, mode=0, want_grid=True)
                p.Update(DT)
                p.CollisionsCheck_Grid(W, H, DT)
                assert np.array_equal(p.debug_hot_pairs(), sorted_pairs(res["pairs"])), f"{name} resolve={resolve} step {step}"
                if resolve == 3:   # keep the two sides on the same trajectory: continue from the device's state
                    got = gpu_state(p)
                    ref.vx[:], ref.vy[:] = got.vx, got.vy
            p.Free()


@pytest.mark.parametrize("phi", [0.05, 0.40, 0.80])
def test_density_sweep_points_match_their_models(oracle, phi):
    """BASELINE configs[4] (packing-density sweep 5 %-80 %), at test size: the exact resolve against the reference's sequential
    order and the tile Gauss-Seidel resolve against its CPU model, bit for bit, at both ends and the middle of the sweep (the same
    generator bench.py's other_configs.density_sweep_16m times at 16M particles)."""
    sc = scenes.uniform_random(30000, 2.0, phi, seed=0x5EED0005, vmax=0.0)
    for resolve in (1, 3):
        ref = state_from_scene(sc)
        p = gpu_store(ref)
        p.set_option("resolve", resolve)
        for step in range(4):
            before = p.gs_fallbacks() if resolve == 3 else 0
            p.Update(DT)
            p.CollisionsCheck_Grid(sc.W, sc.H, DT)
            fell_back = resolve == 3 and p.gs_fallbacks() != before   # (a substep that does not fit the tile kernel runs on the exact resolve)
            if resolve == 3:
                par = p.gs_params()
                oracle.set_tilegs(par["wc"], par["tile_r"], par["strip_limit"])
            oracle.substep(ref, sc.W, sc.H, DT, 0 if ( resolve == 1 or fell_back ) else 3, False)
            assert_state_bits(gpu_state(p), ref, f"area fraction {phi}, resolve={resolve}, substep {step}, fell back: {fell_back}")
        p.Free()


def test_tiny_and_degenerate_fields(oracle):
    for n in (2, 3, 17):
        sc = scenes.uniform_random(n, 5.0, 0.6, seed=420 + n, vmax=10.0)
        ref = state_from_scene(sc)
        p = gs_store(ref)
        for _ in range(3):
            p.Update(DT)
            p.CollisionsCheck_Grid(sc.W, sc.H, DT)
            par = p.gs_params()
            oracle.set_tilegs(par["wc"], par["tile_r"], par["strip_limit"])
            oracle.substep(ref, sc.W, sc.H, DT, 3, False)
        assert_state_bits(gpu_state(p), ref, f"n={n}")
        p.Free()


def test_an_overfull_tile_runs_that_substep_on_the_exact_resolve(oracle):
    """A region far denser than the tile kernel's shared memory takes (here: clumps of hundreds of overlapping particles) makes
    the WHOLE substep fall back to the exact resolve, so that substep equals the sequential reference bit for bit; the
    substeps that fit keep following the tile Gauss-Seidel model.  Never an abort, never a silent partial resolve."""
    from oracle.oracle import State
    n, W, H = 12000, 3000, 3000
    rng = scenes.Stream(121)
    s = State.empty(n)
    cx, cy = rng.unit(30) * (W - 400) + 200, rng.unit(30) * (H - 400) + 200      # 30 clumps of 400 particles within 90 px
    k = np.arange(n) % 30
    s.px[:] = cx[k] + (rng.unit(n) - 0.5) * 90
    s.py[:] = cy[k] + (rng.unit(n) - 0.5) * 90
    s.vx[:] = (rng.unit(n) - 0.5) * 100
    s.vy[:] = (rng.unit(n) - 0.5) * 100
    s.mass[:] = rng.ints(n, 5, 10)
    s.radius[:] = 1.5 + rng.unit(n) * 4.5
    ref = s.copy()
    p = gs_store(s)
    fell_back = []
    for step in range(6):
        before = p.gs_fallbacks()
        p.Update(DT)
        p.CollisionsCheck_Grid(W, H, DT)
        fell_back.append(p.gs_fallbacks() - before)
        par = p.gs_params()
        oracle.set_tilegs(par["wc"], par["tile_r"], par["strip_limit"])
        oracle.substep(ref, W, H, DT, 0 if fell_back[-1] else 3, False)
        assert_state_bits(gpu_state(p), ref, f"clumps, substep {step}, fell back: {fell_back[-1]}")
    print("clumps: substeps that fell back to the exact resolve:", fell_back)
    assert fell_back[0] == 1
    p.Free()


def test_performance_options_do_not_change_the_result():
    """The slice length of k_gs_resolve's record stream ("gs_slice"), one launch for the four tile colours or four ("gs_merge"), who applies the integrate + wall pass ("defer_state": the
    sorting gather, or the first pass as in round 1) and who hands out the positions inside a cell ("cell_ranks": the histogram
    pass, or the scatter's own atomics), one launch or three for the cell-start scan ("scan_single") are performance choices: same bits whatever they are set to."""
    sc = scenes.polydisperse_pile(60000, seed=131, vmax=5.0)
    want = None
    for opts in ({}, {"gs_slice": 64}, {"gs_slice": 512}, {"gs_merge": 1}, {"scan_single": 0}, {"defer_state": 0}, {"cell_ranks": 0}, {"defer_state": 0, "cell_ranks": 0}):
        p = gs_store(state_from_scene(sc))
        for k, v in opts.items():
            p.set_option(k, v)
        for _ in range(6):
            p.Update(DT)
            p.CollisionsCheck_Grid(sc.W, sc.H, DT)
        got = gpu_state(p)
        assert p.gs_fallbacks() == 0
        p.Free()
        if want is None:
            want = got
        else:
            assert_state_bits(got, want, f"options {opts}")


def test_frame_copy_back_follows_a_substep_that_fell_back(oracle):
    """The reference's own calling pattern (main.c: no harness calls, the default sync policy copies the frame back inside
    particlesCollisionsCheck_Grid) with "resolve"=3 on a field whose clumps do not fit the tile kernel: the verdict of the narrow
    phase is looked at before the copy-back, the substep runs on the exact resolve, and the host arrays hold its result."""
    n, W, H = 6000, 2000, 2000
    rng = scenes.Stream(122)
    s = State.empty(n)
    cx, cy = rng.unit(12) * (W - 400) + 200, rng.unit(12) * (H - 400) + 200      # 12 clumps of 500 particles within 90 px
    k = np.arange(n) % 12
    s.px[:] = cx[k] + (rng.unit(n) - 0.5) * 90
    s.py[:] = cy[k] + (rng.unit(n) - 0.5) * 90
    s.vx[:] = (rng.unit(n) - 0.5) * 100
    s.vy[:] = (rng.unit(n) - 0.5) * 100
    s.mass[:] = rng.ints(n, 5, 10)
    s.radius[:] = 1.5 + rng.unit(n) * 4.5
    ref = s.copy()
    p = gpu_store(s, policy=api.PPC_SYNC_FRAME)
    p.set_option("resolve", 3)
    fell = []
    for step in range(4):
        p.Update(DT)
        p.CollisionsCheck_Grid(W, H, DT)                 # ... and the frame is on the host when this returns
        got_x, got_y = p.array("p_x").copy(), p.array("p_y").copy()
        fell.append(p.gs_fallbacks())
        fell_back = fell[-1] != (fell[-2] if len(fell) > 1 else 0)
        par = p.gs_params()
        oracle.set_tilegs(par["wc"], par["tile_r"], par["strip_limit"])
        oracle.substep(ref, W, H, DT, 0 if fell_back else 3, False)
        assert_bits_equal(got_x, ref.px, f"substep {step}: p_x as copied back (fell back: {fell_back})")
        assert_bits_equal(got_y, ref.py, f"substep {step}: p_y as copied back (fell back: {fell_back})")
    assert fell[0] == 1, "the clumps were meant to overflow the tile kernel in the first substep"
    p.Free()


# ---- statistics against the compiled reference ----------------------------------------------------------------------
def _reference_runs(sc, steps, dt):
    from oracle import oracle as om
    if not om.reference_available():
        pytest.skip("oracle/_ref/libref.so not built")
    R = om.Reference()
    out = {}
    for label, mode in (("as built", 0), ("reversed", 1), ("shuffled", 2)):
        s = state_from_scene(sc)
        R.set_pair_order(mode, 7)
        try:
            for _ in range(steps):
                R.substep(s, sc.W, sc.H, dt)
        finally:
            R.set_pair_order(0)
        out[label] = s
    return out


def _report(oracle, sc, got, refs, title):
    base = oracle.stats(refs["as built"], sc.W, sc.H)
    keys = ("ke", "mom_x", "mom_y", "overlap", "contacts")
    dev = {k: rel(oracle.stats(got, sc.W, sc.H)[k], base[k]) for k in keys}
    spread = {lab: {k: rel(oracle.stats(refs[lab], sc.W, sc.H)[k], base[k]) for k in keys} for lab in ("reversed", "shuffled")}
    print(f"\n{title}: deviation from the reference as built, in %")
    print("  device, tile Gauss-Seidel :", {k: round(100 * v, 4) for k, v in dev.items()})
    for lab in spread:
        print(f"  reference, list {lab:8s}:", {k: round(100 * v, 4) for k, v in spread[lab].items()})
    return dev, spread


@pytest.mark.parametrize("phi,seed", [(0.30, 109), (0.60, 110)])
def test_thousand_substeps_within_one_percent_of_the_reference(oracle, phi, seed):
    n, steps, dt = 20000, 1000, 1.0 / 1152.0
    sc = scenes.jittered_lattice(n, 2.0, phi, seed=seed, vmax=20.0)   # non-overlapping start
    p = gs_store(state_from_scene(sc))
    p.run_substeps(sc.W, sc.H, dt, steps)
    got = gpu_state(p)
    p.Free()
    dev, _ = _report(oracle, sc, got, _reference_runs(sc, steps, dt), f"lattice, area fraction {phi}, {steps} substeps")
    assert dev["ke"] < 0.01 and dev["mom_x"] < 0.01 and dev["mom_y"] < 0.01


def test_thousand_substeps_on_the_dense_pile_within_the_references_own_spread(oracle):
    """A chaotic, energy-injecting state: the bar is the reference's own sensitivity to its pair order,
    measured in this same test (kinetic energy of the reference with its list reversed / shuffled against the reference as built:
    1-17 % on this pile, largest in the first few hundred substeps and itself erratic from checkpoint to checkpoint, so the bar at
    a checkpoint is the largest value seen up to it; tests/order_sensitivity.py, profiles/r02_order_sensitivity.txt)."""
    from oracle import oracle as om
    if not om.reference_available():
        pytest.skip("oracle/_ref/libref.so not built")
    n, dt, checkpoints = 20000, 1.0 / 1152.0, (100, 250, 500, 1000)
    sensitivity = 0.0   # largest deviation of the reordered reference from the reference as built seen so far in this run
    sc = scenes.polydisperse_pile(n, seed=104, vmax=5.0)
    p = gs_store(state_from_scene(sc))
    R = om.Reference()
    refs = {label: state_from_scene(sc) for label in ("as built", "reversed", "shuffled")}
    done = 0
    for upto in checkpoints:
        p.run_substeps(sc.W, sc.H, dt, upto - done)
        got = gpu_state(p)
        assert np.isfinite(got.px).all() and np.isfinite(got.py).all() and np.isfinite(got.vx).all() and np.isfinite(got.vy).all()
        for label, mode in (("as built", 0), ("reversed", 1), ("shuffled", 2)):
            R.set_pair_order(mode, 7)
            try:
                for _ in range(upto - done):
                    R.substep(refs[label], sc.W, sc.H, dt)
            finally:
                R.set_pair_order(0)
        done = upto
        dev, spread = _report(oracle, sc, got, refs, f"dense pile, {upto} substeps")
        sensitivity = max(sensitivity, spread["reversed"]["ke"], spread["shuffled"]["ke"])
        assert dev["ke"] <= sensitivity, (f"substep {upto}: kinetic energy off by {100 * dev['ke']:.2f}%, more than the reference moves under a "
                                          f"reordering of its own pair list anywhere in this run ({100 * sensitivity:.2f}%)")
    print(f"\ndense pile: {p.gs_fallbacks()} of {done} substeps fell back to the exact resolve (a sub-tile overflowed the tile kernel): {p.gs_fallback_reasons()}")
    assert p.gs_fallbacks() < done // 4
    p.Free()


def test_sixteen_million_pile_stays_finite_for_a_thousand_substeps():
    n = 1 << 24
    sc = scenes.polydisperse_pile(n, seed=0x5EED0003)
    p = gs_store(state_from_scene(sc))
    p.run_substeps(sc.W, sc.H, DT, 1000)
    got = gpu_state(p)
    print(f"\n16M pile: {p.gs_fallbacks()} of 1000 substeps fell back to the exact resolve")
    assert p.gs_fallbacks() == 0
    p.Free()
    for a in (got.px, got.py, got.vx, got.vy):
        assert np.isfinite(a).all()
    assert got.px.min() >= 0 and got.px.max() <= sc.W and got.py.min() >= 0 and got.py.max() <= sc.H

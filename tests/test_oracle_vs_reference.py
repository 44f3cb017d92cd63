"""CPU: the oracle restatement against the compiled, unmodified reference (oracle/_ref/libref.so), live, on seeded
fields of a few thousand particles.  Skipped where the reference sources (and a prebuilt libref.so) are absent."""
import numpy as np
import pytest

from conftest import assert_bits_equal, state_from_scene
from particlephysicsc_b200 import scenes

DT = 1.0 / 144.0 / 8.0

FIELDS = [
    ("uniform_phi30", lambda: scenes.uniform_random(6000, 2.0, 0.30, seed=11, vmax=20.0)),
    ("uniform_phi70", lambda: scenes.uniform_random(6000, 2.0, 0.70, seed=12, vmax=60.0)),
    ("lattice_phi50", lambda: scenes.jittered_lattice(6000, 2.0, 0.50, seed=13, vmax=30.0)),
    ("pile", lambda: scenes.polydisperse_pile(6000, seed=14, vmax=5.0)),
]


@pytest.mark.parametrize("name,make", FIELDS)
def test_substeps_bit_identical(oracle, reference, name, make):
    sc = make()
    a, b = state_from_scene(sc), state_from_scene(sc)
    for step in range(12):
        reference.update(a, DT)
        oracle.update(b, DT, True)
        for f in ("px", "py", "vx", "vy"):
            assert_bits_equal(getattr(a, f), getattr(b, f), f"{name} step {step} update {f}")
        assert np.array_equal(a.rgba, b.rgba)
        ref_pairs = reference.collisions_check_grid(a, sc.W, sc.H, DT, capture=True)
        res = oracle.collisions_check_grid(b, sc.W, sc.H, DT, mode=0)
        assert np.array_equal(ref_pairs, res["pairs"]), f"{name} step {step} pair list"
        for f in ("px", "py", "vx", "vy"):
            assert_bits_equal(getattr(a, f), getattr(b, f), f"{name} step {step} collide {f}")


def test_resolve_entry_point(oracle, reference):
    """particlesResolveVelocity itself (reference src/particles_solver.c:58) on an explicit pair list."""
    import ctypes as C
    sc = scenes.uniform_random(2000, 2.0, 0.6, seed=21, vmax=40.0)
    a, b = state_from_scene(sc), state_from_scene(sc)
    grid = oracle.grid_build(b, sc.W, sc.H)
    pairs = oracle.pairs(b, grid)
    assert len(pairs) > 100
    st = a.as_struct()
    reference.lib.particlesResolveVelocity(C.byref(st), pairs.ctypes.data_as(C.POINTER(C.c_int32)), len(pairs), DT)
    oracle.resolve(b, pairs, DT, jacobi=False)
    for f in ("px", "py", "vx", "vy"):
        assert_bits_equal(getattr(a, f), getattr(b, f), f)


def test_reference_noise_floor_is_reported(oracle, reference):
    """The reference's own sensitivity to pair order: reversing its list moves the kinetic energy
    by a small but non-zero amount from a non-overlapping start.  This is the floor below which a GPU-vs-reference
    difference carries no information."""
    sc = scenes.jittered_lattice(3000, 2.0, 0.5, seed=31, vmax=20.0)
    a, b = state_from_scene(sc), state_from_scene(sc)
    for _ in range(60):
        reference.substep(a, sc.W, sc.H, DT)
    reference.set_pair_order(1)
    try:
        for _ in range(60):
            reference.substep(b, sc.W, sc.H, DT)
    finally:
        reference.set_pair_order(0)
    ka, kb = oracle.stats(a, sc.W, sc.H)["ke"], oracle.stats(b, sc.W, sc.H)["ke"]
    assert abs(ka / kb - 1.0) < 0.01

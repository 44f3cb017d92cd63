"""CPU: the oracle restatement (oracle/ppc_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py generated from the unmodified reference."""
import glob
import os

import numpy as np
import pytest

from conftest import assert_bits_equal
from oracle.oracle import State

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "store_frames" not in p)
FIELDS = ("px", "py", "vx", "vy", "mass", "radius", "rgba")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def state(g, prefix):
    return State(*(np.ascontiguousarray(g[f"{prefix}_{f}"]).copy() for f in FIELDS))


def check_state(a, g, prefix, what):
    for f in ("px", "py", "vx", "vy"):
        assert_bits_equal(getattr(a, f), g[f"{prefix}_{f}"], f"{what}.{f}")
    assert np.array_equal(a.rgba, g[f"{prefix}_rgba"]), f"{what}.rgba"


def test_golden_files_present():
    assert len(CASES) >= 8


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(oracle, name):
    g = load(name)
    W, H, dt, steps = int(g["W"]), int(g["H"]), float(g["dt"]), int(g["steps"])
    s = state(g, "in")
    oracle.update(s, dt, True)
    check_state(s, g, "upd", name + " after particlesUpdate")
    res = oracle.collisions_check_grid(s, W, H, dt, mode=0)
    assert np.array_equal(res["pairs"], g["pairs"]), name + " candidate-pair list"
    check_state(s, g, "col", name + " after particlesCollisionsCheck_Grid")
    for _ in range(steps - 1):
        oracle.substep(s, W, H, dt, 0, True)
    check_state(s, g, "end", name + f" after {steps} substeps")


@pytest.mark.parametrize("name", CASES)
def test_jacobi_model_positions_equal_sequential_after_one_resolve(oracle, name):
    """Position pushes are sums of frozen terms: the Jacobi model of the device must reproduce the reference's
    positions bit for bit after ONE collisions check (velocities differ where a particle has several contacts)."""
    g = load(name)
    W, H, dt = int(g["W"]), int(g["H"]), float(g["dt"])
    s = state(g, "upd")
    oracle.collisions_check_grid(s, W, H, dt, mode=1)
    assert_bits_equal(s.px, g["col_px"], name + " px")
    assert_bits_equal(s.py, g["col_py"], name + " py")


def test_grid_restatement_is_consistent(oracle):
    """keys / order / cell_start describe the same partition, order is stable ascending (key, index)."""
    g = load("polydisperse_257")
    s = state(g, "upd")
    W, H = int(g["W"]), int(g["H"])
    oracle.walls(s, W, H)
    grid = oracle.grid_build(s, W, H)
    keys, order, start = grid["keys"], grid["order"], grid["cell_start"]
    assert grid["cell_size"] == pytest.approx(2 * max(6.0, s.radius.max()))
    assert start[0] == 0 and start[-1] == s.n and np.all(np.diff(start.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(order), np.arange(s.n))
    sk = keys[order]
    assert np.all(np.diff(sk.astype(np.int64)) >= 0)
    same = sk[1:] == sk[:-1]
    assert np.all(order[1:][same] > order[:-1][same])
    assert np.array_equal(np.bincount(keys, minlength=len(start) - 1), np.diff(start))
    assert np.array_equal(oracle.pairs(s, grid), g["pairs"])

"""Test configuration.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200 and the built CUDA library."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and the built libparticlephysicsc_b200.so")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled, unmodified reference (oracle/_ref/libref.so).  Built by __graft_entry__.build() where
    /root/reference exists; the prebuilt file travels to the GPU box."""
    from oracle import oracle as om
    if not om.reference_available():
        try:
            om.build(("ref",))
        except Exception:
            pass
    if not om.reference_available():
        pytest.skip("oracle/_ref/libref.so not built (reference sources absent)")
    return om.Reference()


def state_from_scene(sc):
    from oracle.oracle import State
    return State(sc.px.copy(), sc.py.copy(), sc.vx.copy(), sc.vy.copy(), sc.mass.copy(), sc.radius.copy(), sc.rgba.copy())


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bits_equal(a, b, what=""):
    """Bit-for-bit equality of float32 arrays (NaN payloads included)."""
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.nonzero(ba != bb)[0]
        raise AssertionError(f"{what}: {bad.size} of {ba.size} elements differ, first at {bad[:5]}: "
                             f"{np.asarray(a).ravel()[bad[:3]]} vs {np.asarray(b).ravel()[bad[:3]]}")

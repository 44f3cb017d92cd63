"""CPU: the synthetic field generators."""
import numpy as np

from particlephysicsc_b200 import scenes


def test_splitmix64_known_answers():
    # reference values of the published splitmix64 algorithm, seed 0 and seed 1234567
    assert [int(v) for v in scenes.splitmix64(0, 0, 3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    assert [int(v) for v in scenes.splitmix64(1234567, 0, 2)] == [6457827717110365317, 3203168211198807973]


def test_stream_is_position_independent():
    a = scenes.Stream(5)
    x = np.concatenate([a.u64(3), a.u64(5)])
    assert np.array_equal(x, scenes.splitmix64(5, 0, 8))


def test_uniform_field_geometry():
    sc = scenes.uniform_random(10000, 2.0, 0.30)
    assert abs(sc.area_fraction() - 0.30) < 0.01
    assert sc.px.min() >= 2.0 and sc.px.max() <= sc.W - 2.0
    assert sc.py.min() >= 2.0 and sc.py.max() <= sc.H - 2.0
    assert set(np.unique(sc.mass)) <= {5.0, 6.0, 7.0, 8.0, 9.0, 10.0}
    assert sc.px.dtype == np.float32 and sc.rgba.shape == (10000, 4)


def test_lattice_has_no_overlaps(oracle):
    from conftest import state_from_scene
    for phi in (0.3, 0.6):
        sc = scenes.jittered_lattice(5000, 2.0, phi)
        st = oracle.stats(state_from_scene(sc), sc.W, sc.H)
        assert st["contacts"] == 0 and st["overlap"] == 0.0


def test_pile_radii_and_fit():
    sc = scenes.polydisperse_pile(20000)
    assert sc.radius.min() >= 1.5 and sc.radius.max() <= 6.0
    assert sc.py.max() <= sc.H and sc.py.min() >= 0
    assert sc.px.max() <= sc.W and sc.px.min() >= 0


def test_box_limit():
    import pytest
    with pytest.raises(ValueError):
        scenes.uniform_random(1 << 28, 2.0, 0.05)


def test_debug_scene_matches_main_c_parameters():
    frames = list(scenes.debug_scene_frames(frames=50))
    assert len(frames) == 50
    for x, y, m, r, cnt in frames:
        assert 0 <= x <= 1200 and 0 <= y <= 675 and 5 <= m <= 10 and 2 <= r <= 8 and cnt == 10


def test_snapshot_file_round_trip(tmp_path):
    """The snapshot format (include/ppc_b200.h part 4) written and read by the host-side helpers."""
    import numpy as np
    from particlephysicsc_b200 import scenes
    sc = scenes.polydisperse_pile(5000, seed=9, vmax=3.0)
    path = str(tmp_path / "pile.ppcb")
    scenes.write_snapshot(path, sc)
    assert (tmp_path / "pile.ppcb").stat().st_size == 64 + 28 * sc.n
    back = scenes.read_snapshot(path)
    assert (back.W, back.H, back.n) == (sc.W, sc.H, sc.n)
    for f in ("px", "py", "vx", "vy", "mass", "radius", "rgba"):
        assert np.array_equal(getattr(back, f), getattr(sc, f))
    with open(path, "r+b") as f:
        f.truncate(64 + 10 * sc.n)
    import pytest
    with pytest.raises(ValueError):
        scenes.read_snapshot(path)


def test_uniform_slab_stays_inside_its_rows():
    """One rank's share of the configs[3] field: every particle lies in the rank's own cell rows and inside the walls."""
    import numpy as np
    from particlephysicsc_b200 import scenes
    W, H, cell = 45916, 478 * 12 * 4, 12.0
    for k in range(4):
        sc = scenes.uniform_slab(50000, 1.0, W, H, k * 478 * cell, (k + 1) * 478 * cell, seed=0x5EED0004 + 0x100 * k)
        rows = np.floor(sc.py / np.float32(cell)).astype(int)
        assert rows.min() >= k * 478 and rows.max() < (k + 1) * 478
        assert sc.px.min() >= 1.0 and sc.px.max() <= W - 1.0 and sc.py.min() >= 1.0 and sc.py.max() <= H - 1.0
        assert (sc.W, sc.H) == (W, H) and sc.n == 50000


def test_bench_workloads_build_small():
    """bench.make_scene for every workload name, at a small size (the generators behind the benchmark lines)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for name in bench.WORKLOADS:
        sc = bench.make_scene(name, 1 << 14)
        assert sc.n == 1 << 14 and sc.W <= 65535 and sc.H <= 65535 and sc.radius.min() > 0

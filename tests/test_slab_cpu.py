"""CPU: host logic of the slab decomposition (particlephysicsc_b200/slab.py) and the slab algorithm itself, with the
device replaced by a model built from the oracle's pieces (tests/slab_model.py).  Includes a world_size-2 gloo run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import assert_bits_equal, state_from_scene
from particlephysicsc_b200 import scenes, slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 1.0 / 144.0 / 8.0


def test_partition_rows_covers_every_row_once():
    for rows, world in ((10, 1), (10, 3), (3827, 8), (8, 8)):
        r = slab.partition_rows(rows, world)
        assert r[0][0] == 0 and r[-1][1] == rows and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(e > b for b, e in r) and max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    with pytest.raises(ValueError):
        slab.partition_rows(3, 4)


def test_partition_by_load_balances_particles():
    counts = np.zeros(100, np.int64)
    counts[60:] = 1000          # a pile: everything in the lower 40 rows
    r = slab.partition_rows_by_load(counts, 4)
    assert r[0][0] == 0 and r[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    loads = [counts[b:e].sum() for b, e in r]
    assert max(loads) <= 1.1 * counts.sum() / 4 + 1000


def test_next_capacity_is_a_pure_function_of_the_last_count():
    # both ends of a link compute it independently from the count the last message carried
    f = slab.DistTransport.next_capacity
    assert f(0) >= 4096 and f(1000000) >= 1000000 * 17 // 16 and f(12345) % 1024 == 0
    assert all(f(n) >= n + 4096 for n in (0, 1, 5000, 10 ** 6, 10 ** 8))


def test_cell_rows_follow_the_reference_division():
    py = np.array([0.0, 11.999999, 12.0, 35.9, 1e9, -5.0, np.nan], np.float32)
    assert slab.cell_rows(py, 12.0, 3).tolist() == [0, 0, 1, 2, 2, 0, 0]
    assert slab.owner_of_rows(np.array([0, 2, 3, 9]), [(0, 3), (3, 10)]).tolist() == [0, 0, 1, 1]


def model_slabs(oracle, sc, world, halo):
    from slab_model import ModelSlab
    max_radius, ranges, idx, _ = slab.split_scene(sc, world, halo)
    full = state_from_scene(sc)
    out = []
    for k in range(world):
        ii = idx[k]
        st = type(full)(*(np.ascontiguousarray(a[ii]) for a in (full.px, full.py, full.vx, full.vy, full.mass, full.radius, full.rgba)))
        out.append(ModelSlab(oracle, st, ii, sc.W, sc.H, max_radius, ranges[k], halo))
    return out


def gather(ranks, n):
    out = {k: np.zeros(n, np.float32) for k in ("px", "py", "vx", "vy")}
    seen = np.zeros(n, np.int32)
    for r in ranks:
        for k in out:
            out[k][r.gid] = getattr(r.s, k)
        seen[r.gid] += 1
    assert (seen == 1).all(), "every particle is owned by exactly one rank"
    return out


@pytest.mark.parametrize("name,make,world,steps", [
    ("lattice", lambda: scenes.jittered_lattice(24000, 2.0, 0.50, seed=301, vmax=900.0), 2, 10),
    ("uniform", lambda: scenes.uniform_random(24000, 2.0, 0.30, seed=306, vmax=400.0), 2, 6),
])
def test_slabs_equal_whole_field_bit_for_bit(oracle, name, make, world, steps):
    """K slabs whose halo the inexactness mark does not cross == the oracle on the whole field, positions and velocities.
    First with a generous halo (which also measures how far the mark travels), then with the tightest halo that measure allows."""
    sc = make()
    travelled = 0
    for attempt, halo in enumerate((13, None)):
        halo = halo if halo is not None else travelled + 1
        ref = state_from_scene(sc)
        ranks = model_slabs(oracle, sc, world, halo)
        moved = 0
        for step in range(steps):
            before = [set(r.gid.tolist()) for r in ranks]
            slab.local_step(ranks, DT)
            oracle.substep(ref, sc.W, sc.H, DT, 0, False)
            moved += sum(len(set(r.gid.tolist()) - b) for r, b in zip(ranks, before))
            assert not any(r.taint_reached_owned for r in ranks)
            travelled = max([travelled] + [r.taint_rows for r in ranks]) if attempt == 0 else travelled
            got = gather(ranks, sc.n)
            for k in ("px", "py", "vx", "vy"):
                assert_bits_equal(got[k], getattr(ref, k), f"{name} halo {halo} step {step} {k}")
        assert moved > 0, "the test field must make particles migrate between slabs"
    assert 0 < travelled < 12


def test_too_thin_a_halo_is_noticed(oracle):
    """With one ghost row the mark reaches owned particles in a dense field: the criterion must say so (it is conservative:
    marked velocities MAY differ from the whole-field run).  Positions only ever need one ghost row."""
    sc = scenes.uniform_random(6000, 2.0, 0.6, seed=305, vmax=100.0)
    ref = state_from_scene(sc)
    ranks = model_slabs(oracle, sc, 3, 1)
    slab.local_step(ranks, DT)
    oracle.substep(ref, sc.W, sc.H, DT, 0, False)
    assert any(r.taint_reached_owned for r in ranks)
    got = gather(ranks, sc.n)
    assert_bits_equal(got["px"], ref.px, "positions need one ghost row only")
    assert_bits_equal(got["py"], ref.py, "positions need one ghost row only")


def test_integer_outputs_of_slabs_concatenate_to_the_whole_field(oracle):
    sc = scenes.uniform_random(5000, 2.5, 0.5, seed=303, vmax=200.0)
    ref = state_from_scene(sc)
    ranks = model_slabs(oracle, sc, 3, 12)
    for _ in range(3):
        slab.local_step(ranks, DT)
        oracle.update(ref, DT, False)
        full = oracle.collisions_check_grid(ref, sc.W, sc.H, DT, 0, want_grid=True)
    keys = np.zeros(sc.n, np.uint32)
    for r in ranks:
        keys[r.last_keys["gid"]] = r.last_keys["key"]
    assert np.array_equal(keys, full["keys"])
    pairs = np.concatenate([r.last_pairs for r in ranks]).astype(np.int64)
    want = full["pairs"].astype(np.int64)
    assert len(pairs) == len(want)
    assert np.array_equal(pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])


# ---- world_size 2 over gloo ------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, fixed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    sc = scenes.jittered_lattice(12000, 2.0, 0.30, seed=304, vmax=600.0)
    ranks = model_slabs(Oracle(), sc, world, 12)
    me = ranks[rank]
    me.fixed_size_messages = fixed
    tr = slab.DistTransport(dist, torch.device("cpu"))
    for _ in range(6):
        slab.slab_step(me, tr, DT)
    out.put((rank, me.gid, me.s.px, me.s.py, me.s.vx, me.s.vy, me.taint_reached_owned))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fixed", [False, True], ids=["counts_then_records", "fixed_size_messages"])
def test_two_ranks_over_gloo_equal_whole_field(oracle, fixed):
    ctx = mp.get_context("spawn")
    out, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, fixed)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = scenes.jittered_lattice(12000, 2.0, 0.30, seed=304, vmax=600.0)
    ref = state_from_scene(sc)
    for _ in range(6):
        oracle.substep(ref, sc.W, sc.H, DT, 0, False)
    have = {k: np.zeros(sc.n, np.float32) for k in ("px", "py", "vx", "vy")}
    count = np.zeros(sc.n, np.int32)
    for rank, gid, px, py, vx, vy, rounds in got:
        assert not rounds
        count[gid] += 1
        for k, a in zip(("px", "py", "vx", "vy"), (px, py, vx, vy)):
            have[k][gid] = a
    assert (count == 1).all()
    for k in have:
        assert_bits_equal(have[k], getattr(ref, k), f"gloo world 2 {k}")


def _overflow_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    sc = scenes.jittered_lattice(12000, 2.0, 0.30, seed=304, vmax=600.0)
    me = model_slabs(Oracle(), sc, world, 12)[rank]
    me.fixed_size_messages = True
    tr = slab.DistTransport(dist, torch.device("cpu"))
    slab.slab_step(me, tr, DT)            # negotiates the message sizes
    if rank == 0:                         # rank 0's next message is "too large" for what was agreed ...
        tr.cap_send[1] = 8
    else:                                 # ... and rank 1 expects that agreed size
        tr.cap_recv[0] = 8
    try:
        slab.slab_step(me, tr, DT)
        out.put((rank, "no error"))
    except slab.MessageOverflow as e:
        out.put((rank, "MessageOverflow"))
    dist.barrier()
    dist.destroy_process_group()


def test_an_outgrown_fixed_size_message_raises_on_both_ranks_instead_of_hanging():
    """a rank whose halo message outgrew the agreed size used to raise before posting its send, leaving the
    neighbour in its receive for ever.  The cut message is now sent with the true count in its header; both ends raise."""
    ctx = mp.get_context("spawn")
    out, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_overflow_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: "MessageOverflow", 1: "MessageOverflow"}

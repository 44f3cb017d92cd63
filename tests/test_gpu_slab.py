"""GPU: the slab decomposition against the single-GPU run of the whole field, through the C ABI.

Several slabs are run in ONE process on one GPU (the exchange is a device copy instead of an NCCL send/recv; everything
else -- kernels, C entry points, host protocol -- is the multi-GPU code path).  Acceptance, as stated in section 8e:
integer outputs identical to the single-GPU run; here positions and velocities are required to be bit-identical too.
The two-process NCCL run of the same thing is test_two_ranks_over_nccl (needs two GPUs: `gpurun --gpus 2`).
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import assert_bits_equal
from particlephysicsc_b200 import api, scenes, slab

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 1.0 / 144.0 / 8.0


def single_gpu(sc, steps, options=None):
    p = api.Particles(sc.n)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    for k, v in (options or {}).items():
        p.set_option(k, v)
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    rounds = 0
    for _ in range(steps):
        p.run_substeps(sc.W, sc.H, DT, 1)
        rounds = max(rounds, (p.wave_stats() or (0, 0))[0])
    grid = p.debug_grid()
    pairs = p.debug_pairs()
    p.download(api.PPC_DL_ALL)
    out = {k: p.array(k).copy() for k in ("p_x", "p_y", "v_x", "v_y")}
    p.Free()
    return out, grid, pairs, rounds


def gather(ranks, n):
    out = {k: np.zeros(n, np.float32) for k in ("p_x", "p_y", "v_x", "v_y")}
    seen = np.zeros(n, np.int32)
    for r in ranks:
        d = r.download()
        assert d["halo_ok"]
        seen[d["ids"]] += 1
        for k in out:
            out[k][d["ids"]] = d[k]
    assert (seen == 1).all(), "every particle is owned by exactly one rank"
    return out


CASES = [
    ("uniform_phi30_w4", lambda: scenes.uniform_random(300000, 2.0, 0.30, seed=401, vmax=300.0), 4, 10, False, {}),
    ("lattice_phi50_w3", lambda: scenes.jittered_lattice(200000, 2.0, 0.50, seed=402, vmax=500.0), 3, 10, False, {}),
    ("pile_w2_by_load", lambda: scenes.polydisperse_pile(150000, seed=403, vmax=20.0), 2, 6, True, {"wave_mode": 0}),
    ("uniform_polling_w2", lambda: scenes.uniform_random(100000, 2.0, 0.30, seed=404, vmax=300.0), 2, 6, False, {"wave_mode": 0}),
    ("uniform_radix_sort_w3", lambda: scenes.uniform_random(100000, 2.0, 0.30, seed=408, vmax=300.0), 3, 5, False, {"sort": 1}),
    ("lattice_jacobi_w2", lambda: scenes.jittered_lattice(100000, 2.0, 0.40, seed=409, vmax=300.0), 2, 5, False, {"resolve": 2}),
]


@pytest.mark.parametrize("name,make,world,steps,by_load,options", CASES)
def test_slabs_on_one_gpu_equal_the_whole_field(name, make, world, steps, by_load, options):
    sc = make()
    want, grid, pairs, rounds = single_gpu(sc, steps, options)
    travelled = 0
    # first a generous halo (which measures how far the inexactness mark travels), then the halo the product would choose from that
    for attempt in range(2):
        halo = 2 * rounds + 4 if attempt == 0 else slab.required_halo(travelled, 2)   # the mark's travel varies a little with where it starts
        ranks = slab.make_device_slabs(sc, world, halo, options=options, by_load=by_load)
        owned0 = [set(r.download()["ids"].tolist()) for r in ranks]
        for _ in range(steps):
            slab.local_step(ranks, DT)
            if attempt == 0:
                travelled = max([travelled] + [int(r.info().taint_rows) for r in ranks])
        got = gather(ranks, sc.n)
        for k in want:
            assert_bits_equal(got[k], want[k], f"{name} halo {halo} {k}")
        # integer work: sorted order and cell-start table of the owned rows, concatenated over ranks, are the whole field's
        order, starts = [], []
        for r in ranks:
            w = r.debug_window()
            i = w["info"]
            order.append(w["order"][i.own_first:i.own_first + i.n_owned])
            c0, c1 = (i.own_row_begin - i.row_lo) * i.cols, (i.own_row_end - i.row_lo) * i.cols
            starts.append(w["cell_start"][c0:c1].astype(np.int64) - int(i.own_first) + sum(len(o) for o in order[:-1]))
        assert np.array_equal(np.concatenate(order), grid["order"])
        assert np.array_equal(np.concatenate(starts), grid["cell_start"][:-1].astype(np.int64))
        # candidate pairs: each listed by the rank that owns its i
        mine = np.concatenate([r.debug_pairs() for r in ranks]).astype(np.int64)
        assert len(mine) == len(pairs)
        p64 = pairs.astype(np.int64)
        assert np.array_equal(mine[np.lexsort((mine[:, 1], mine[:, 0]))], p64[np.lexsort((p64[:, 1], p64[:, 0]))])
        owned1 = [set(r.download()["ids"].tolist()) for r in ranks]
        if name.startswith("uniform_phi30"):
            assert any(a != b for a, b in zip(owned0, owned1)), "particles must have migrated between slabs in this case"
        for r in ranks:
            r.free()
    print(f"{name}: {rounds} dependency rounds, inexactness mark travelled {travelled} rows")
    assert travelled >= 1 or options.get("resolve") == 2   # Jacobi needs one ghost row and nothing travels


def test_empty_slabs_are_fine():
    """Slabs that own nothing (the field fills only the upper half of the box) take part in the exchange like any other."""
    sc = scenes.uniform_random(60000, 2.0, 0.35, seed=407, vmax=200.0)
    sc.H *= 2
    want, grid, pairs, rounds = single_gpu(sc, 4)
    ranks = slab.make_device_slabs(sc, 4, 6)
    assert ranks[3].info().n_owned == 0 and ranks[2].info().n_owned < 1000
    for _ in range(4):
        slab.local_step(ranks, DT)
    got = gather(ranks, sc.n)
    for k in want:
        assert_bits_equal(got[k], want[k], f"empty slabs {k}")
    for r in ranks:
        r.free()


def test_shallow_halo_is_reported():
    """A halo thinner than the dependency depth cannot guarantee exact velocities: the library says so instead of going on."""
    sc = scenes.uniform_random(100000, 2.0, 0.60, seed=405, vmax=100.0)
    ranks = slab.make_device_slabs(sc, 2, 1)
    slab.local_step(ranks, DT)
    with pytest.raises(slab.HaloTooShallow):
        slab.local_step(ranks, DT)
    for r in ranks:
        r.free()


def test_standard_entry_points_refuse_a_slab_store():
    """Misuse is loud, not silent: a slab store only holds part of the field, so the whole-field entry points abort with a
    message (checked in a child process: the library has no error return on the reference's void entry points)."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from particlephysicsc_b200 import scenes, slab\n"
        "sc = scenes.uniform_random(5000, 2.0, 0.30, seed=406)\n"
        "ranks = slab.make_device_slabs(sc, 2, 4)\n"
        "assert ranks[0].info().halo_rows == 4 and ranks[0].info().n_owned + ranks[1].info().n_owned == sc.n\n"
        "print('slabs ready', flush=True)\n"
        "ranks[0].p.Update(0.001)\n"
        "print('not reached', flush=True)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "slabs ready" in out.stdout and "not reached" not in out.stdout
    assert out.returncode != 0
    assert "one slab of a multi-GPU field" in out.stderr


def test_a_halo_deeper_than_a_neighbouring_slab_is_refused():
    """Ghost rows only come from the two direct neighbours, which send what they own: a window reaching past a neighbour's slab
    would stay partly empty without anything noticing, so the library refuses it at set-up."""
    sc = scenes.uniform_random(20000, 2.0, 0.30, seed=412)
    _, cols, rows = slab.grid_dims(sc.W, sc.H, 6.0)
    thin = rows // 8   # rows per slab when cut into eight
    with pytest.raises(ValueError, match="deeper than a neighbouring slab"):
        slab.make_device_slabs(sc, 8, thin + 1)
    for r in slab.make_device_slabs(sc, 8, thin):
        r.free()


# ---- "resolve"=3 (tile Gauss-Seidel) on slabs: exact by construction ----------------------------------------------------------
GS_CASES = [
    ("uniform_phi30_w4", lambda: scenes.uniform_random(300000, 2.0, 0.30, seed=421, vmax=300.0), 4, 8, False),
    ("lattice_phi50_w3", lambda: scenes.jittered_lattice(200000, 2.0, 0.50, seed=422, vmax=500.0), 3, 8, False),
    ("pile_w2_by_load", lambda: scenes.polydisperse_pile(150000, seed=423, vmax=20.0), 2, 6, True),
    ("uniform_r1_phi40_w5", lambda: scenes.uniform_random(400000, 1.0, 0.40, seed=424, vmax=100.0), 5, 4, False),
]


@pytest.mark.parametrize("name,make,world,steps,by_load", GS_CASES)
def test_tile_gauss_seidel_slabs_equal_the_whole_field(name, make, world, steps, by_load):
    """The tile order is anchored at the whole field, so with the ghost rows ppc_slab_gs_halo asks for every rank computes its
    owned particles exactly as the single-GPU run: positions, velocities, sorted order, cell-start table and pair set."""
    sc = make()
    want, grid, pairs, _ = single_gpu(sc, steps, {"resolve": 3})   # the single GPU chooses its tile width itself ...
    max_radius = float(max(6.0, float(np.max(sc.radius))))
    options = slab.gs_options(sc.n, sc.W, sc.H, max_radius)        # ... and the ranks are told the same one
    _, ranges, _, _ = slab.split_scene(sc, world, 1, by_load)
    _, cols, rows = slab.grid_dims(sc.W, sc.H, max_radius)
    halo = slab.gs_required_halo(ranges, rows)
    assert 1 <= halo <= 16
    ranks = slab.make_device_slabs(sc, world, halo, options=options, by_load=by_load)
    for _ in range(steps):
        slab.local_step(ranks, DT)
    got = gather(ranks, sc.n)
    for k in want:
        assert_bits_equal(got[k], want[k], f"{name} halo {halo} {k}")
    order, starts = [], []
    for r in ranks:
        w = r.debug_window()
        i = w["info"]
        order.append(w["order"][i.own_first:i.own_first + i.n_owned])
        c0, c1 = (i.own_row_begin - i.row_lo) * i.cols, (i.own_row_end - i.row_lo) * i.cols
        starts.append(w["cell_start"][c0:c1].astype(np.int64) - int(i.own_first) + sum(len(o) for o in order[:-1]))
    assert np.array_equal(np.concatenate(order), grid["order"])
    assert np.array_equal(np.concatenate(starts), grid["cell_start"][:-1].astype(np.int64))
    mine = np.concatenate([r.debug_pairs() for r in ranks]).astype(np.int64)
    p64 = pairs.astype(np.int64)
    assert np.array_equal(mine[np.lexsort((mine[:, 1], mine[:, 0]))], p64[np.lexsort((p64[:, 1], p64[:, 0]))])
    for r in ranks:
        r.free()
    # one ghost row less than asked for is refused, not computed inexactly
    if halo > 1:
        ranks = slab.make_device_slabs(sc, world, halo - 1, options=options, by_load=by_load)
        with pytest.raises(slab.HaloTooShallow):
            slab.local_step(ranks, DT)
        for r in ranks:
            r.free()


# ---- two processes, two GPUs, NCCL ---------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, halo, steps, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}", device_id=torch.device("cuda", rank))
    sc = scenes.uniform_random(300000, 2.0, 0.30, seed=401, vmax=300.0)
    me = slab.make_device_slabs(sc, world, halo, device_of=lambda k: k, only_rank=rank)[0]
    tr = slab.DistTransport(dist, torch.device("cuda", rank))
    for _ in range(steps):
        slab.slab_step(me, tr, DT)
    d = me.download()
    out.put((rank, d["ids"], d["p_x"], d["p_y"], d["v_x"], d["v_y"], d["halo_ok"]))
    dist.barrier()
    dist.destroy_process_group()
    me.free()


def test_two_ranks_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    sc = scenes.uniform_random(300000, 2.0, 0.30, seed=401, vmax=300.0)
    steps = 10
    want, _, _, rounds = single_gpu(sc, steps)
    halo = 2 * rounds + 4
    ctx = mp.get_context("spawn")
    out, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, halo, steps, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    have = {k: np.zeros(sc.n, np.float32) for k in ("p_x", "p_y", "v_x", "v_y")}
    count = np.zeros(sc.n, np.int32)
    for rank, ids, px, py, vx, vy, ok in got:
        assert ok
        count[ids] += 1
        for k, a in zip(("p_x", "p_y", "v_x", "v_y"), (px, py, vx, vy)):
            have[k][ids] = a
    assert (count == 1).all()
    for k in have:
        assert_bits_equal(have[k], want[k], f"nccl world 2 {k}")


def test_colours_travel_with_migrants():
    """A drawn substep: colours come back per particle exactly as from the single-GPU run, whichever rank owns the particle."""
    sc = scenes.uniform_random(80000, 2.0, 0.30, seed=410, vmax=400.0)
    p = api.Particles(sc.n)
    p.set_sync_policy(api.PPC_SYNC_MANUAL)
    p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
    p.run_substeps(sc.W, sc.H, DT, 8, True)
    p.download(api.PPC_DL_ALL)
    want = p.array("colors").copy()
    p.Free()
    ranks = slab.make_device_slabs(sc, 3, 8)
    for s in range(8):
        slab.local_step(ranks, DT, colour=(s == 7))
    got = np.zeros((sc.n, 4), np.uint8)
    for r in ranks:
        d = r.download()
        got[d["ids"]] = d["colors"]
        r.free()
    assert np.array_equal(got, want)
    assert len(np.unique(want, axis=0)) > 100   # the field is moving: colours are not all alike


def test_slabs_on_two_gpus_in_one_process():
    """Stores on different devices in one process: every entry point works on its own store's device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = scenes.uniform_random(200000, 2.0, 0.30, seed=411, vmax=300.0)
    want, _, _, rounds = single_gpu(sc, 6)
    ranks = slab.make_device_slabs(sc, 2, 2 * rounds + 4, device_of=lambda k: k)
    for _ in range(6):
        slab.local_step(ranks, DT)
    got = gather(ranks, sc.n)
    for k in want:
        assert_bits_equal(got[k], want[k], f"two devices, one process: {k}")
    for r in ranks:
        r.free()


def test_cpp_host_drives_the_slabs_through_the_c_abi():
    """examples/slab_multi_gpu.cu: no Python, no torch in the loop -- a C++ host cuts a field into slabs (spread over the GPUs
    there are), moves the exchange buffers with cudaMemcpyPeer, and compares every particle with the single-GPU run."""
    import subprocess
    exe = os.path.join(ROOT, "examples", "_bin", "slab_multi_gpu")
    if not os.path.exists(exe):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build_examples()
    out = subprocess.run([exe, "4", "300000", "6"], capture_output=True, text=True, timeout=600)
    print(out.stdout[-600:], out.stderr[-300:])
    assert out.returncode == 0, out.stdout[-400:] + out.stderr[-400:]
    assert "0 differ from the single-GPU run" in out.stdout

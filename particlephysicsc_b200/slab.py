"""Multi-GPU slab decomposition of the solver hot path: the host side.

The field is cut into slabs of whole cell rows (the cell index is row-major, reference src/particles_solver.c:291), one
rank -- one process, one GPU, one device store -- per slab.  One substep on every rank:

    begin    integrate + wall pass of the owned particles; count the records for each neighbour
    pack     owned particles within `halo` rows of (or beyond) a border -> one message per neighbour (migrants + ghosts)
    exchange neighbour send/recv (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)
    finish   received records join the store; grid over owned + ghost rows; narrow phase; resolve; owned rows stay

No collective is on the data path: only neighbour send/recv.  The device work is behind the C ABI (``ppc_slab_*`` in
include/ppc_b200.h, kernels in csrc/ppc_slab.cuh); this module holds what is independent of the device: the row
partition, the exchange protocol, and the in-process transport used to run several slabs on one GPU in tests.

A "rank backend" is any object with ``begin(dt, colour) -> (n_lo, n_hi)``, ``pack(buf_lo, buf_hi)``,
``finish(dt, buf_lo, n_lo, buf_hi, n_hi) -> n_owned`` and ``alloc(records) -> torch.uint8 tensor [records, 32]``;
``DeviceSlab`` below is the CUDA one.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api

RECORD_BYTES = 32


def grid_dims(W: int, H: int, max_radius: float):
    """(cell_size, cols, rows) of the whole field (reference src/particles_solver.c:231-249), from the library."""
    g = api.GridInfo()
    if api.load_library().ppc_grid_dims(W, H, float(max_radius), C.byref(g)) != 0:
        raise ValueError("empty grid")
    return float(g.cell_size), int(g.cols), int(g.rows)


def partition_rows(rows: int, world: int):
    """Contiguous, near-equal row ranges [begin, end) per rank; every rank gets at least one row."""
    if world < 1 or rows < world:
        raise ValueError(f"cannot cut {rows} cell rows into {world} slabs")
    return [(rows * k // world, rows * (k + 1) // world) for k in range(world)]


def partition_rows_by_load(row_counts: np.ndarray, world: int):
    """Row ranges with near-equal PARTICLE counts (a settled pile is not uniform in y)."""
    rows = len(row_counts)
    if world < 1 or rows < world:
        raise ValueError(f"cannot cut {rows} cell rows into {world} slabs")
    cum = np.concatenate([[0], np.cumsum(row_counts, dtype=np.int64)])
    cuts = [0]
    for k in range(1, world):
        r = int(np.searchsorted(cum, cum[-1] * k / world))
        r = min(max(r, cuts[-1] + 1), rows - (world - k))
        cuts.append(r)
    cuts.append(rows)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def cell_rows(py: np.ndarray, cell_size: float, rows: int) -> np.ndarray:
    """Cell row of each particle with the reference's float operations (:286-289): IEEE float32 division, truncation, clamp."""
    q = (py.astype(np.float32) / np.float32(cell_size)).astype(np.float32)
    cy = np.trunc(np.nan_to_num(q, nan=-1.0, posinf=-1.0, neginf=-1.0)).astype(np.int64)
    return np.clip(cy, 0, rows - 1)


def owner_of_rows(cy: np.ndarray, ranges) -> np.ndarray:
    ends = np.array([e for _, e in ranges], dtype=np.int64)
    return np.searchsorted(ends, cy, side="right")


def required_halo(taint_rows: int, margin: int = 2) -> int:
    """Ghost rows for a field whose inexactness mark travelled `taint_rows` rows into the window (ppc_slab_info): the mark
    starts in the window's outermost row and must stop short of the first owned row; `margin` rows absorb its fluctuation
    from substep to substep (too thin a halo is always reported, never silent)."""
    return int(taint_rows) + 1 + margin


def gs_required_halo(ranges, rows: int) -> int:
    """Ghost rows the "resolve"=3 path needs on a slab decomposition `ranges` of `rows` cell rows (the largest over the ranks):
    its tile order is anchored at the whole field, and a window must hold every tile an owned particle depends on."""
    lib = api.load_library()
    return max(int(lib.ppc_slab_gs_halo(b, e, rows)) for b, e in ranges)


def gs_options(n_total: int, W: int, H: int, max_radius: float) -> dict:
    """Options that make every rank run the "resolve"=3 tile order of the WHOLE field (the tile width is part of the result)."""
    _, cols, rows = grid_dims(W, H, max_radius)
    return {"resolve": 3, "gs_wc": int(api.load_library().ppc_gs_auto_tile_width(float(n_total), float(cols) * float(rows)))}


class MessageOverflow(RuntimeError):
    """A fixed-size halo message did not hold the records of this substep (raised on the sender AND on its receiver)."""


class HaloTooShallow(RuntimeError):
    """The dependency depth of a substep exceeded what the halo guarantees; owned velocities may be inexact."""


class DeviceSlab:
    """One slab on one GPU: a device store in slab mode plus its exchange buffers (torch tensors: plumbing only)."""

    def __init__(self, px, py, vx, vy, mass, radius, rgba, global_ids, W, H, max_radius, own_rows, halo, capacity=None, device=0,
                 options=None):
        import torch
        self.torch = torch
        n = len(px)
        capacity = int(capacity if capacity is not None else max(1024, int(n * 1.3) + 4096))
        self.p = api.Particles(capacity, device=device)
        self.p.set_sync_policy(api.PPC_SYNC_MANUAL)
        for k, v in (options or {}).items():
            self.p.set_option(k, v)
        self.p.s.length = n
        for name, src in (("p_x", px), ("p_y", py), ("v_x", vx), ("v_y", vy), ("mass", mass), ("radius", radius)):
            self.p.array(name)[:] = src
        self.p.array("colors")[:] = rgba
        self.geometry = (W, H, float(max_radius), own_rows)
        self.device = torch.device("cuda", device)
        self.bufs = {}
        self.counts = (0, 0)
        self.received = (0, 0)
        self.fixed_size_messages = True   # distributed steps: one message per neighbour, capacity agreed from the previous count
        self._init(np.ascontiguousarray(global_ids, dtype=np.uint32), halo)
        self.stream = torch.cuda.ExternalStream(self.p.stream(), device=self.device)
        self.side_stream = torch.cuda.Stream(device=self.device)   # the neighbour exchange runs here, beside the interior integrate pass

    def _init(self, ids, halo):
        W, H, max_radius, own_rows = self.geometry
        rc = self.p.lib.ppc_slab_init(C.byref(self.p.s), ids.ctypes.data_as(api.u32p), W, H, max_radius, own_rows[0], own_rows[1], halo)
        if rc == -5:
            raise ValueError(f"halo of {halo} rows is deeper than a neighbouring slab: ghost rows only come from the direct neighbours")
        if rc != 0:
            raise ValueError(f"ppc_slab_init failed ({rc})")
        self.halo = halo

    def reinit(self, halo: int) -> None:
        """Restart this slab from its current owned particles with another halo depth (mass and radius never leave the host
        arrays' order on download, so the state is read back and uploaded again)."""
        d = self.download(all_fields=True)
        n = len(d["ids"])
        for name in ("p_x", "p_y", "v_x", "v_y", "mass", "radius"):
            self.p.array(name)[:n] = d[name]
        self.p.array("colors")[:n] = d["colors"]
        self._init(d["ids"], halo)

    def alloc(self, records: int):
        return self.torch.empty((max(records, 1), RECORD_BYTES), dtype=self.torch.uint8, device=self.device)

    def buffer(self, name: str, records: int):
        b = self.bufs.get(name)
        if b is None or b.shape[0] < records:
            b = self.bufs[name] = self.alloc(int(records * 1.25) + 256)
        return b

    def begin(self, dt: float, colour: bool = False):
        c = (C.c_uint64 * 2)()
        rc = self.p.lib.ppc_slab_begin(C.byref(self.p.s), dt, int(colour), c)
        if rc == -2:
            raise HaloTooShallow(f"halo of {self.halo} rows is too thin: the last resolve's inexactness mark travelled {self.info().taint_rows} rows "
                                 "and reached an owned particle")
        if rc != 0:
            raise RuntimeError(f"ppc_slab_begin failed ({rc})")
        self.counts = (int(c[0]), int(c[1]))
        return self.counts

    def pack(self, buf_lo, buf_hi, header: bool = False) -> None:
        self.p.lib.ppc_slab_pack(C.byref(self.p.s), buf_lo.data_ptr(), buf_hi.data_ptr(), int(header))

    def interior(self) -> None:
        """Queue the part of the integrate pass that cannot produce records (ppc_slab_begin did the border slots only)."""
        self.p.lib.ppc_slab_interior(C.byref(self.p.s))

    def finish(self, dt: float, buf_lo, n_lo: int, buf_hi, n_hi: int, header: bool = False) -> int:
        """With header=True n_lo / n_hi are the capacities of fixed-size messages; self.received holds the counts they carried."""
        got = (C.c_uint64 * 2)()
        n = int(self.p.lib.ppc_slab_finish(C.byref(self.p.s), dt, buf_lo.data_ptr(), n_lo, buf_hi.data_ptr(), n_hi, int(header), got))
        if n == -4:
            raise HaloTooShallow(f'"resolve"=3 needs "gs_wc" set and at least {self.p.lib.ppc_slab_gs_halo(*self.geometry[3], 1 << 30)} ghost rows here; '
                                 f"this slab has {self.halo}")
        if n == -3:
            raise MessageOverflow("a neighbour's halo message outgrew the agreed size (its header counts more records than arrived)")
        if n < 0:
            raise RuntimeError(f"ppc_slab_finish failed ({n})")
        self.received = (int(got[0]), int(got[1]))
        return n

    def synchronize(self) -> None:
        self.p.synchronize()

    def info(self) -> api.SlabInfo:
        out = api.SlabInfo()
        self.p.lib.ppc_slab_info(C.byref(self.p.s), C.byref(out))
        return out

    def download(self, all_fields: bool = False):
        """Owned particles in storage order: dict of arrays plus their global indices."""
        n_max = self.p.size
        ids = np.zeros(n_max, np.uint32)
        mask = api.PPC_DL_ALL | api.PPC_DL_IDS | (api.PPC_DL_MASS_RADIUS if all_fields else 0)
        rc = self.p.lib.ppc_slab_download(C.byref(self.p.s), ids.ctypes.data_as(api.u32p), mask)
        if rc == -3:
            raise RuntimeError("host arrays too small for the owned particles")
        n = self.p.length
        out = {k: self.p.array(k).copy() for k in ("p_x", "p_y", "v_x", "v_y", "colors") + (("mass", "radius") if all_fields else ())}
        out["ids"] = ids[:n].copy()
        out["halo_ok"] = rc == 0
        return out

    def sync_host(self, mask: int = api.PPC_DL_POS | api.PPC_DL_COLOR) -> int:
        """Make host arrays of the store current (owned particles, storage order) without copying them again; with
        PPC_DL_IDS the global indices land in ``self.host_ids``.  Returns the number of owned particles.  The default mask is
        the "frame is drawn" copy-back: positions and colours."""
        if getattr(self, "host_ids", None) is None or len(self.host_ids) < self.p.size:
            self.host_ids = np.zeros(self.p.size, np.uint32)
        rc = self.p.lib.ppc_slab_download(C.byref(self.p.s), self.host_ids.ctypes.data_as(api.u32p), mask)
        if rc == -2:
            raise HaloTooShallow(f"halo of {self.halo} rows is too thin (the inexactness mark travelled {self.info().taint_rows} rows)")
        if rc != 0:
            raise RuntimeError(f"ppc_slab_download failed ({rc})")
        return self.p.length

    def sync_host_async(self, mask: int = api.PPC_DL_POS | api.PPC_DL_COLOR) -> int:
        """Start the copy-back of the owned particles and return at once (it overlaps the substeps that follow);
        sync_host_wait() completes it.  Returns the number of particles being copied."""
        rc = self.p.lib.ppc_slab_download_async(C.byref(self.p.s), None, mask & ~api.PPC_DL_IDS)
        if rc != 0:
            raise RuntimeError(f"ppc_slab_download_async failed ({rc})")
        return self.p.length

    def sync_host_wait(self) -> None:
        self.p.lib.ppc_download_wait(C.byref(self.p.s))

    def debug_pairs(self) -> np.ndarray:
        n = self.p.lib.ppc_slab_debug_pairs(C.byref(self.p.s), None, 0)
        if n < 0:
            raise RuntimeError("pair list unavailable")
        out = np.zeros((n, 2), np.int32)
        if n:
            self.p.lib.ppc_slab_debug_pairs(C.byref(self.p.s), out.ctypes.data_as(api.i32p), n)
        return out

    def debug_window(self):
        """Sorted global indices and cell-start table of the window (owned + ghost rows) of the last grid build."""
        i = self.info()
        order = np.zeros(i.n_window, np.uint32)
        start = np.zeros((i.row_hi - i.row_lo) * i.cols + 1, np.uint32)
        if i.n_window:
            assert self.p.lib.ppc_debug_sorted_order(C.byref(self.p.s), order.ctypes.data_as(api.u32p)) == 0
        assert self.p.lib.ppc_debug_cell_start(C.byref(self.p.s), start.ctypes.data_as(api.u32p)) == 0
        return dict(order=order, cell_start=start, info=i)

    def free(self) -> None:
        self.p.Free()


# ---- exchange ------------------------------------------------------------------------------------------------------
class DistTransport:
    """Neighbour exchange over torch.distributed point-to-point ops (NCCL send/recv on GPUs; gloo on CPU).

    Rank r holds slab r; its lower neighbour is r - 1, its upper r + 1.  Counts go first (two integers per neighbour),
    then the records.  No collective is involved."""

    def __init__(self, dist, device):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lo = self.rank - 1 if self.rank > 0 else None
        self.hi = self.rank + 1 if self.rank + 1 < self.world else None

    def _run(self, ops):
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_counts(self, n_lo: int, n_hi: int):
        t, d = self.torch, self.dist
        out = t.tensor([n_lo, n_hi], dtype=t.int64, device=self.device)
        got = t.zeros(2, dtype=t.int64, device=self.device)
        ops = []
        if self.lo is not None:
            ops += [d.P2POp(d.isend, out[0:1], self.lo), d.P2POp(d.irecv, got[0:1], self.lo)]
        if self.hi is not None:
            ops += [d.P2POp(d.isend, out[1:2], self.hi), d.P2POp(d.irecv, got[1:2], self.hi)]
        self._run(ops)
        r = got.tolist()
        return int(r[0]), int(r[1])

    # fixed-size messages: both ends derive the capacity of the next message from the count of the last one, so no count
    # has to cross before the records do (one send/recv per neighbour and substep instead of two)
    @staticmethod
    def next_capacity(count: int) -> int:
        return ((count + count // 16 + 4096 + 1023) // 1024) * 1024

    def negotiate(self, n_lo: int, n_hi: int) -> None:
        """Once, before the first fixed-size exchange: tell the neighbours the current counts."""
        r_lo, r_hi = self.exchange_counts(n_lo, n_hi)
        self.cap_send = [self.next_capacity(n_lo) if self.lo is not None else 0, self.next_capacity(n_hi) if self.hi is not None else 0]
        self.cap_recv = [self.next_capacity(r_lo) if self.lo is not None else 0, self.next_capacity(r_hi) if self.hi is not None else 0]

    def reset(self) -> None:
        """Forget the agreed capacities (the slabs were restarted with another halo): the next step negotiates again."""
        for a in ("cap_send", "cap_recv"):
            if hasattr(self, a):
                delattr(self, a)

    def advance(self, sent, received) -> None:
        for k, nb in enumerate((self.lo, self.hi)):
            if nb is not None:
                self.cap_send[k], self.cap_recv[k] = self.next_capacity(sent[k]), self.next_capacity(received[k])

    def exchange(self, send_lo, send_hi, recv_lo, recv_hi):
        """send_* / recv_* are [count, 32] uint8 tensors (count may be 0: nothing is posted for it)."""
        d, ops = self.dist, []
        if self.lo is not None:
            if send_lo.shape[0]:
                ops.append(d.P2POp(d.isend, send_lo, self.lo))
            if recv_lo.shape[0]:
                ops.append(d.P2POp(d.irecv, recv_lo, self.lo))
        if self.hi is not None:
            if send_hi.shape[0]:
                ops.append(d.P2POp(d.isend, send_hi, self.hi))
            if recv_hi.shape[0]:
                ops.append(d.P2POp(d.irecv, recv_hi, self.hi))
        self._run(ops)


def slab_step(rank, transport: DistTransport, dt: float, colour: bool = False, timers: dict | None = None) -> int:
    """One substep of one rank of a distributed job.  Device work and NCCL traffic are ordered on the store's stream.
    With `timers` (diagnostics only) every phase is followed by a device synchronise and its wall time accumulated."""
    import time
    ctx = rank.torch.cuda.stream(rank.stream) if getattr(rank, "stream", None) is not None else _null()

    def lap(name, t0):
        if timers is None:
            return t0
        rank.torch.cuda.synchronize()
        t1 = time.perf_counter()
        timers[name] = timers.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    with ctx:
        t = time.perf_counter()
        n_lo, n_hi = rank.begin(dt, colour)
        t = lap("begin (integrate, walls, count)", t)
        if not getattr(rank, "fixed_size_messages", False):   # counts first, then exactly that many records
            s_lo, s_hi = rank.buffer("send_lo", n_lo), rank.buffer("send_hi", n_hi)
            rank.pack(s_lo, s_hi)
            t = lap("pack", t)
            r_lo, r_hi = transport.exchange_counts(n_lo, n_hi)
            t = lap("exchange counts", t)
            b_lo, b_hi = rank.buffer("recv_lo", r_lo), rank.buffer("recv_hi", r_hi)
            transport.exchange(s_lo[:n_lo], s_hi[:n_hi], b_lo[:r_lo], b_hi[:r_hi])
            t = lap("exchange records", t)
            n = rank.finish(dt, b_lo, r_lo, b_hi, r_hi)
            lap("finish (grid, narrow phase, resolve)", t)
            return n
        if not hasattr(transport, "cap_send"):
            transport.negotiate(n_lo, n_hi)
        cs, cr = transport.cap_send, transport.cap_recv
        # A message that outgrew the agreed size is still sent, cut to that size: its header carries the true count, so the
        # receiver sees the overflow too and both ends raise in the same substep (an error on one rank only would leave the
        # neighbours waiting in their receive for ever).
        overflow = n_lo > cs[0] or n_hi > cs[1]
        s_lo, s_hi = rank.buffer("send_lo", max(cs[0], n_lo) + 1), rank.buffer("send_hi", max(cs[1], n_hi) + 1)
        rank.pack(s_lo, s_hi, header=True)
        t = lap("pack", t)
        b_lo, b_hi = rank.buffer("recv_lo", cr[0] + 1), rank.buffer("recv_hi", cr[1] + 1)
        side = getattr(rank, "side_stream", None) if getattr(rank, "overlap_exchange", True) else None
        if side is not None and hasattr(rank, "interior"):
            # The records travel on a second stream, ordered behind the pack; meanwhile the store's stream integrates the owned
            # particles that cannot be in any message (most of them), and waits for the exchange only before it takes the records in.
            # (The interior pass is queued first: issuing the NCCL group can block the host until the neighbour is ready.  Measured
            # the other way round: 4.64 instead of 4.57 ms per substep.  The process group should be created with a high-priority
            # stream, so that the few CTAs of the transfer kernels do not queue behind the integrate pass: bench.py does.)
            side.wait_stream(rank.stream)
            rank.interior()
            with rank.torch.cuda.stream(side):
                transport.exchange(s_lo[:cs[0] + 1 if cs[0] else 0], s_hi[:cs[1] + 1 if cs[1] else 0], b_lo[:cr[0] + 1 if cr[0] else 0],
                                   b_hi[:cr[1] + 1 if cr[1] else 0])
            rank.stream.wait_stream(side)
        else:
            transport.exchange(s_lo[:cs[0] + 1 if cs[0] else 0], s_hi[:cs[1] + 1 if cs[1] else 0], b_lo[:cr[0] + 1 if cr[0] else 0],
                               b_hi[:cr[1] + 1 if cr[1] else 0])
        t = lap("exchange records (+ interior integrate)", t)
        if overflow:
            raise MessageOverflow(f"halo message grew from a capacity of {cs} to {(n_lo, n_hi)} records within one substep")
        n = rank.finish(dt, b_lo, cr[0], b_hi, cr[1], header=True)
        transport.advance((n_lo, n_hi), rank.received)
        lap("finish (grid, narrow phase, resolve)", t)
        return n


def local_step(ranks, dt: float, colour: bool = False):
    """One substep of ALL slabs held by this process (several slabs on one GPU, or CPU model slabs): the same three phases
    with the exchange done by plain copies.  Slab k's lower neighbour is k - 1."""
    counts = [r.begin(dt, colour) for r in ranks]
    sends = []
    for r, (n_lo, n_hi) in zip(ranks, counts):
        s_lo, s_hi = r.buffer("send_lo", n_lo), r.buffer("send_hi", n_hi)
        r.pack(s_lo, s_hi)
        sends.append((s_lo, s_hi))
    for r in ranks:
        r.synchronize()
    owned = []
    for k, r in enumerate(ranks):
        r_lo = counts[k - 1][1] if k > 0 else 0                  # what the lower neighbour sent upwards
        r_hi = counts[k + 1][0] if k + 1 < len(ranks) else 0     # what the upper neighbour sent downwards
        b_lo, b_hi = r.buffer("recv_lo", r_lo), r.buffer("recv_hi", r_hi)
        ctx = r.torch.cuda.stream(r.stream) if getattr(r, "stream", None) is not None else _null()
        with ctx:
            if r_lo:
                b_lo[:r_lo].copy_(sends[k - 1][1][:r_lo])
            if r_hi:
                b_hi[:r_hi].copy_(sends[k + 1][0][:r_hi])
            owned.append(r.finish(dt, b_lo, r_lo, b_hi, r_hi))
    return owned


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def split_scene(sc, world: int, halo: int, by_load: bool = False):
    """Cut a generated scene (particlephysicsc_b200.scenes.Scene) into `world` slabs by the cell row of the initial positions.
    Returns (max_radius, ranges, [index arrays], [capacity each store needs for its window])."""
    max_radius = float(max(6.0, float(np.max(sc.radius)))) if sc.n else 6.0
    cell, cols, rows = grid_dims(sc.W, sc.H, max_radius)
    cy = cell_rows(sc.py, cell, rows)
    per_row = np.bincount(cy, minlength=rows)
    ranges = partition_rows_by_load(per_row, world) if by_load else partition_rows(rows, world)
    owner = owner_of_rows(cy, ranges)
    cum = np.concatenate([[0], np.cumsum(per_row, dtype=np.int64)])
    need = []
    for b, e in ranges:   # owned + received ghosts, stored behind the lower ghosts of the previous substep
        g_lo, g_hi = cum[b] - cum[max(b - halo, 0)], cum[min(e + halo, rows)] - cum[e]
        need.append(int(cum[e] - cum[b] + 2 * g_lo + g_hi))
    return max_radius, ranges, [np.nonzero(owner == k)[0] for k in range(world)], need


def make_device_slabs(sc, world: int, halo: int, device_of=lambda k: 0, only_rank=None, options=None, by_load: bool = False, slack: float = 1.25):
    """DeviceSlab objects for all (or one) of the `world` slabs of scene `sc`."""
    max_radius, ranges, idx, need = split_scene(sc, world, halo, by_load)
    out = []
    for k in range(world):
        if only_rank is not None and k != only_rank:
            continue
        ii = idx[k]
        cap = int(need[k] * slack) + 65536
        options = dict(options or {})
        # ghost rows come from the direct neighbours only: the library refuses a halo deeper than a neighbour's slab
        options["slab_nb_lo_rows"] = ranges[k - 1][1] - ranges[k - 1][0] if k > 0 else 0
        options["slab_nb_hi_rows"] = ranges[k + 1][1] - ranges[k + 1][0] if k + 1 < world else 0
        out.append(DeviceSlab(sc.px[ii], sc.py[ii], sc.vx[ii], sc.vy[ii], sc.mass[ii], sc.radius[ii], sc.rgba[ii], ii.astype(np.uint32), sc.W, sc.H,
                              max_radius, ranges[k], halo, capacity=cap, device=device_of(k), options=options))
    return out

"""CPU model of one slab rank, built from the oracle's pieces (TEST INFRASTRUCTURE).

Same interface as particlephysicsc_b200.slab.DeviceSlab, so the product's host logic (row partition, exchange protocol,
slab_step / local_step, DistTransport) runs unchanged on CPU -- over gloo with world_size 2 -- and the slab ALGORITHM
(one message per neighbour carrying migrants + ghosts; ownership by cell row; halo depth vs dependency depth) is
checked bit for bit against the oracle run on the whole field.
"""
import ctypes as C

import numpy as np
import torch

from oracle.oracle import Oracle, State, u32p
from particlephysicsc_b200 import slab

REC = np.dtype([("px", "<f4"), ("py", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("radius", "<f4"), ("mass", "<f4"), ("gid", "<u4"),
                ("rgba", "u1", (4,))])
assert REC.itemsize == slab.RECORD_BYTES


def dependency_depth(pairs: np.ndarray) -> int:
    """Depth of the reference's pair dependency graph: a pair waits for the previous pair of either particle."""
    last = {}
    depth = 0
    for i, j in pairs.tolist():
        lv = 1 + max(last.get(i, 0), last.get(j, 0))
        last[i] = last[j] = lv
        depth = max(depth, lv)
    return depth


def taint_travel(pairs: np.ndarray, tainted: np.ndarray):
    """The product's exactness criterion (csrc/ppc_collide.cuh, ST_TAINT): particles of the window's outermost rows start
    out marked; a pair with a marked end marks both ends, in the reference's pair order.  Returns the final marks."""
    t = tainted.copy()
    for i, j in pairs.tolist():
        if t[i] or t[j]:
            t[i] = t[j] = True
    return t


class ModelSlab:
    torch = torch
    stream = None

    def __init__(self, oracle: Oracle, st: State, gid, W, H, max_radius, own_rows, halo):
        self.o, self.W, self.H = oracle, W, H
        self.cell, self.cols, self.rows = 2.0 * max(6.0, float(max_radius)), None, None
        self.cell = float(np.float32(2.0) * np.float32(max(6.0, max_radius)))
        self.cols = int(np.ceil(np.float32(W) / np.float32(self.cell)))
        self.rows = int(np.ceil(np.float32(H) / np.float32(self.cell)))
        self.own, self.halo = own_rows, halo
        self.has_lo, self.has_hi = own_rows[0] > 0, own_rows[1] < self.rows
        self.s, self.gid = st, np.asarray(gid, np.uint32).copy()
        self.bufs, self.rounds, self.last_pairs = {}, 0, None
        self.fixed_size_messages, self.received = False, (0, 0)

    # -- plumbing ------------------------------------------------------------------------------------------------
    def alloc(self, records):
        return torch.zeros((max(records, 1), slab.RECORD_BYTES), dtype=torch.uint8)

    def buffer(self, name, records):
        b = self.bufs.get(name)
        if b is None or b.shape[0] < records:
            b = self.bufs[name] = self.alloc(records + 16)
        return b

    def synchronize(self):
        pass

    def keys_of(self, s: State):
        keys, order = np.zeros(s.n, np.uint32), np.zeros(s.n, np.uint32)
        start = np.zeros(self.cols * self.rows + 1, np.uint32)
        f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        if s.n:
            self.o.lib.ppc_oracle_grid_build(f(s.px), f(s.py), s.n, self.cell, self.cols, self.rows, keys.ctypes.data_as(u32p),
                                             order.ctypes.data_as(u32p), start.ctypes.data_as(u32p))
        return dict(cell_size=self.cell, cols=self.cols, rows=self.rows, keys=keys, order=order, cell_start=start)

    # -- the three phases ----------------------------------------------------------------------------------------
    def begin(self, dt, colour=False):
        s = self.s
        if s.n:
            self.o.update(s, dt, colour)
            self.o.walls(s, self.W, self.H)
        cy = (self.keys_of(s)["keys"] // self.cols).astype(np.int64)
        b, e, h = self.own[0], self.own[1], self.halo
        self.to_lo = np.nonzero(cy < b + h)[0] if self.has_lo else np.zeros(0, np.int64)
        self.to_hi = np.nonzero(cy >= e - h)[0] if self.has_hi else np.zeros(0, np.int64)
        lost = ((cy < b - h) & self.has_lo) | ((cy >= e + h) & self.has_hi)
        assert not lost.any(), "a particle moved past the halo"
        return len(self.to_lo), len(self.to_hi)

    def records(self, idx):
        r = np.zeros(len(idx), REC)
        s = self.s
        for name, a in (("px", s.px), ("py", s.py), ("vx", s.vx), ("vy", s.vy), ("radius", s.radius), ("mass", s.mass)):
            r[name] = a[idx]
        r["gid"], r["rgba"] = self.gid[idx], s.rgba[idx]
        return r

    def pack(self, buf_lo, buf_hi, header=False):
        for buf, idx in ((buf_lo, self.to_lo), (buf_hi, self.to_hi)):
            b = buf.numpy()
            if header:   # record 0 = number of records that follow
                b[0, :4] = np.array([len(idx)], "<u4").view(np.uint8)
                b = b[1:]
            if len(idx):
                b[:len(idx)] = self.records(idx).view(np.uint8).reshape(len(idx), slab.RECORD_BYTES)

    def finish(self, dt, buf_lo, n_lo, buf_hi, n_hi, header=False):
        parts = [self.records(np.arange(self.s.n))]
        got = []
        for buf, n in ((buf_lo, n_lo), (buf_hi, n_hi)):
            b = buf.numpy()
            if header and n:   # n is the capacity; the header says how many records are real
                cap, n = n, int(np.ascontiguousarray(b[0, :4]).view("<u4")[0])
                if n > cap:
                    raise slab.MessageOverflow("a neighbour's halo message outgrew the agreed size")
                b = b[1:]
            got.append(n)
            if n:
                parts.append(np.ascontiguousarray(b[:n]).view(REC).reshape(n))
        self.received = tuple(got)
        r = np.concatenate(parts)
        r = r[np.argsort(r["gid"], kind="stable")]   # local index order = global index order: the pair order carries over
        assert len(np.unique(r["gid"])) == len(r), "a particle arrived twice"
        c = np.ascontiguousarray
        u = State(c(r["px"]), c(r["py"]), c(r["vx"]), c(r["vy"]), c(r["mass"]), c(r["radius"]), c(r["rgba"]))
        grid = self.keys_of(u)
        cy = (grid["keys"] // self.cols).astype(np.int64)
        row_lo, row_hi = max(self.own[0] - self.halo, 0), min(self.own[1] + self.halo, self.rows)
        keep = (cy >= row_lo) & (cy < row_hi)
        if not keep.all():   # outside the window: dropped before the grid is built
            r, cy = r[keep], cy[keep]
            u = State(c(r["px"]), c(r["py"]), c(r["vx"]), c(r["vy"]), c(r["mass"]), c(r["radius"]), c(r["rgba"]))
            grid = self.keys_of(u)
        pairs = self.o.pairs(u, grid) if u.n >= 2 else np.zeros((0, 2), np.int32)
        self.rounds = dependency_depth(pairs)
        gids = r["gid"]
        owned = (cy >= self.own[0]) & (cy < self.own[1])
        edge = ((cy == row_lo) & (row_lo > 0)) | ((cy == row_hi - 1) & (row_hi < self.rows))
        marked = taint_travel(pairs, edge)
        self.taint_reached_owned = bool((marked & owned).any())
        depth = np.minimum(np.where(row_lo > 0, cy - row_lo, 1 << 30), np.where(row_hi < self.rows, row_hi - 1 - cy, 1 << 30))
        self.taint_rows = int(depth[marked].max()) if marked.any() else 0
        self.last_pairs = np.stack([gids[pairs[:, 0]], gids[pairs[:, 1]]], 1)[owned[pairs[:, 0]]] if len(pairs) else np.zeros((0, 2), np.uint32)
        self.last_keys = dict(gid=gids[owned], key=grid["keys"][owned])
        self.o.resolve(u, pairs, dt)
        idx = np.nonzero(owned)[0]
        self.s = State(*(c(a[idx]) for a in (u.px, u.py, u.vx, u.vy, u.mass, u.radius, u.rgba)))
        self.gid = c(gids[idx])
        return self.s.n

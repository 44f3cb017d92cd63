#!/usr/bin/env python
"""Multi-GPU acceptance check: the slab-decomposed run equals the single-GPU run of the whole field.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_equivalence.py [--particles 16777216] [--steps 8]

Every rank runs its slab over NCCL neighbour exchange; rank 0 also runs the whole field on its GPU through the ordinary entry
points.  Compared: an order-independent 64-bit digest of (global index, p_x, p_y, v_x, v_y bit patterns) over all particles,
the number of particles, and the digest of the cell keys.  Equal digests = bit-identical state on every particle."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlephysicsc_b200 import api, scenes, slab  # noqa: E402

DT = 1.0 / 144.0 / 8.0


def digest(ids, *arrays):
    """Sum over particles of a 64-bit mix of (index, fields): independent of the order and of how particles are split."""
    with np.errstate(over="ignore"):
        h = ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        for a in arrays:
            h = (h ^ np.ascontiguousarray(a).view(np.uint32).astype(np.uint64)) * np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(29)
        return int(np.sum(h, dtype=np.uint64))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", dest="n", type=int, default=1 << 24)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--phi", type=float, default=0.30)
    ap.add_argument("--halo", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.uniform_random(args.n, 2.0, args.phi, seed=0x5EED0006, vmax=200.0)
    me = slab.make_device_slabs(sc, world, args.halo, device_of=lambda k: local, only_rank=rank)[0]
    tr = slab.DistTransport(dist, torch.device("cuda", local))
    travelled = 0
    for _ in range(args.steps):
        slab.slab_step(me, tr, DT)
        travelled = max(travelled, int(me.info().taint_rows))
    d = me.download()
    # digests are sums modulo 2^64; add them as two 32-bit halves so that int64 all_reduce cannot overflow
    full = digest(d["ids"], d["p_x"], d["p_y"], d["v_x"], d["v_y"])
    parts = torch.tensor([full & 0xFFFFFFFF, full >> 32, len(d["ids"]), int(d["halo_ok"])], dtype=torch.int64, device="cuda")
    dist.all_reduce(parts, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([travelled], dtype=torch.int64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    me.free()
    if rank == 0:
        lo, hi, count, ok = (int(x) for x in parts.tolist())
        slabs_digest = (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF
        p = api.Particles(sc.n, device=local)
        p.set_sync_policy(api.PPC_SYNC_MANUAL)
        p.fill(sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius, sc.rgba)
        p.run_substeps(sc.W, sc.H, DT, args.steps)
        p.download(api.PPC_DL_ALL)
        whole = digest(np.arange(sc.n, dtype=np.uint32), p.array("p_x"), p.array("p_y"), p.array("v_x"), p.array("v_y"))
        p.Free()
        out = {"gpus": world, "particles": sc.n, "substeps": args.steps, "area_fraction": args.phi, "halo_rows": args.halo,
               "inexactness_travel_rows": int(tmax.item()), "all_ranks_report_exact": ok == world, "particles_owned_by_all_ranks": count,
               "digest_slabs": f"{slabs_digest:016x}", "digest_single_gpu": f"{whole:016x}", "bit_identical": slabs_digest == whole and count == sc.n}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU: the multi-rank arithmetic of bench.py (max over ranks, whole-job value) on a world_size-2 gloo group."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    slowest = bench.max_over_ranks(10.0 * (rank + 1), dist, "cpu")   # rank 1 is the slow one: 20 ms
    out.put((rank, slowest, bench.whole_job_value(1000, world, 4, slowest)))
    dist.barrier()
    dist.destroy_process_group()


def test_max_over_ranks_and_whole_job_value_world2():
    ctx = mp.get_context("spawn")
    out, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slowest, value in got:
        assert slowest == pytest.approx(20.0)
        assert value == pytest.approx(1000 * 2 * 4 / 0.020)   # units of ALL ranks over the slowest rank's time


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.max_over_ranks(3.5) == 3.5
    assert bench.whole_job_value(10, 1, 2, 1000.0) == 20.0


def test_reference_arm_line_shape(monkeypatch, capsys):
    """`--impl reference` prints one JSON line with the contract's keys (the CPU timing itself is stubbed out)."""
    sys.path.insert(0, ROOT)
    import json
    import bench
    base = {"value": 1.0e5, "unit": bench.UNIT, "cores": 1, "kind": "reference", "sample": "stub", "particles": 1 << 20}
    monkeypatch.setattr(bench, "time_reference_cpu", lambda *a, **k: (1.0e5, 10.0, base))

    class A:
        workload, steps, warmup, gpus = "pile16m", 2, 1, 1
    bench.run_reference(A, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["e2e"] == {"value": 1.0e5, "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["higher_is_better"] is True and line["config"]["full_field"] is False
    bench.run_reference(A, 1)   # other ranks print nothing
    assert capsys.readouterr().out == ""


def test_host_binding_and_nccl_options_degrade_gracefully():
    """Without a GPU / NVML the rank stays unbound and says so; the NCCL options object, where torch provides it, asks for a
    high-priority stream (the neighbour exchange runs beside the interior integrate pass)."""
    sys.path.insert(0, ROOT)
    import os
    import bench
    before = os.sched_getaffinity(0)
    note = bench.bind_to_gpu_numa_node(0)
    assert isinstance(note, str) and note
    assert os.sched_getaffinity(0) <= before and os.sched_getaffinity(0)   # never widened, never emptied
    os.sched_setaffinity(0, before)
    o = bench.nccl_options()
    assert o is None or o.is_high_priority_stream is True


"""N > 1 host logic on CPU: two gloo ranks shard the streams, detect nothing on a GPU (fake per-stream Signal
lists stand in for the engine output) and gather on rank 0 in deterministic stream order."""
import datetime
import os
import socket

import pytest

from pyradiotracking_b200 import messages, shard


def test_stream_range_is_a_partition():
    for n in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            parts = [shard.stream_range(n, w, r) for r in range(w)]
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.stream_range(4, 2, 2)


def _signals_for(stream: int):
    t0 = datetime.datetime(2026, 1, 1, tzinfo=datetime.timezone.utc)
    return [messages.Signal(str(stream), t0 + datetime.timedelta(milliseconds=10 * k + stream), 150e6 + 1000 * k, 0.02,
                            -60 - k, -62, 1.0, -94, 30) for k in range(stream % 3 + 1)]


def _worker(rank, world, port, n_streams, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.stream_range(n_streams, world, rank)
        local = [_signals_for(s) for s in mine]
        got = shard.gather_signals(local, n_streams)
        if rank == 0:
            q.put([[(s.device, s.ts.isoformat(), s.frequency, s.max) for s in lst] for lst in got])
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    n_streams, world = 7, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = [[(s.device, s.ts.isoformat(), s.frequency, s.max) for s in _signals_for(i)] for i in range(n_streams)]
    assert got == want


def test_gather_without_process_group_is_identity():
    local = [_signals_for(0), _signals_for(1)]
    assert shard.gather_signals(local, 2) == local

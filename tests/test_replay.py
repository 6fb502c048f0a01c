"""Recorded-capture ingest: the reader on the CPU, the replay loop on the GPU."""
import datetime

import numpy as np
import pytest

from pyradiotracking_b200 import synth
from pyradiotracking_b200.replay import CaptureReader, replay


def _write(tmp_path, arrays):
    paths = []
    for i, a in enumerate(arrays):
        p = tmp_path / f"chan{i}.bin"
        a.tofile(p)
        paths.append(str(p))
    return paths


def test_reader_yields_full_blocks_in_order_and_drops_the_tail(tmp_path):
    rng = np.random.default_rng(1)
    N = 1000
    a = rng.integers(0, 256, 2 * N * 5 + 123, dtype=np.uint8)       # 5 blocks + a ragged tail
    b = rng.integers(0, 256, 2 * N * 7, dtype=np.uint8)             # longer channel: cut at the shortest
    paths = _write(tmp_path, [a, b])
    r = CaptureReader(paths, N, n_buffers=3, pinned=False)
    assert r.n_blocks == 5 and r.n_streams == 2
    seen = []
    for k, blk in r:
        assert blk.shape == (2, 2 * N) and blk.dtype == np.uint8
        assert np.array_equal(blk[0], a[2 * N * k: 2 * N * (k + 1)]) and np.array_equal(blk[1], b[2 * N * k: 2 * N * (k + 1)])
        seen.append(k)
    assert seen == list(range(5))
    assert [k for k, _ in CaptureReader(paths, N, pinned=False, max_blocks=2)] == [0, 1]
    with pytest.raises(ValueError):
        CaptureReader([], N)


def test_reader_ring_keeps_the_last_views_valid(tmp_path):
    N = 64
    data = np.arange(2 * N * 10, dtype=np.uint32).astype(np.uint8)
    paths = _write(tmp_path, [data])
    held = []
    for k, blk in CaptureReader(paths, N, n_buffers=4, pinned=False):
        held.append((k, blk))
        for kk, bb in held[-3:]:                                     # the current view and the two before it
            assert np.array_equal(bb[0], data[2 * N * kk: 2 * N * (kk + 1)])


@pytest.mark.gpu
def test_replay_of_files_equals_feeding_the_blocks_directly(tmp_path):
    from oracle.cases import BY_NAME
    from pyradiotracking_b200.analyze import BatchAnalyzer
    from tests import parity

    w = synth.C1
    nb = 4
    caps = [synth.make_stream(w, 40 + i, nb) for i in range(3)]                       # [nb, 2N] each
    tails = [np.zeros(777, np.uint8), np.zeros(0, np.uint8), np.zeros(2 * w.block_samples - 2, np.uint8)]
    paths = _write(tmp_path, [np.concatenate([c.reshape(-1), t]) for c, t in zip(caps, tails)])
    kw = BY_NAME["c1_default_300k"].analyzer_kwargs()
    devs = ["0", "1", "2"]
    t0 = datetime.datetime(2026, 9, 9, 9, 9, 9)
    dt = datetime.timedelta(seconds=w.block_samples / w.sample_rate)
    a = BatchAnalyzer(**parity.batch_kwargs(kw, devices=devs, calibration=[0.0] * 3))
    b = BatchAnalyzer(**parity.batch_kwargs(kw, devices=devs, calibration=[0.0] * 3))
    try:
        got = {}
        n = replay(paths, a, t0, on_block=lambda k, res: got.__setitem__(k, res))
        assert n == nb and sorted(got) == list(range(nb))
        total = 0
        for k in range(nb):
            want = b.process_blocks(np.stack([c[k] for c in caps]), [t0 + k * dt] * 3)
            for s in range(3):
                assert [(x.ts, x.frequency, x.duration, x.max) for x in got[k][s][0]] == [(x.ts, x.frequency, x.duration, x.max) for x in want[s][0]]
                assert got[k][s][1] == len(want[s][1])
                total += len(want[s][0])
        assert total > 0
    finally:
        a.close()
        b.close()


def test_reader_groups_of_blocks_and_pads_the_last_group(tmp_path):
    """blocks_per_read = B (rt_config.blocks_per_launch): items are B consecutive blocks per stream; 7 blocks = 2 groups of 3 + 1
    block padded with zero bytes."""
    rng = np.random.default_rng(2)
    N = 500
    a = rng.integers(1, 256, 2 * N * 7 + 5, dtype=np.uint8)
    b = rng.integers(1, 256, 2 * N * 7, dtype=np.uint8)
    paths = _write(tmp_path, [a, b])
    r = CaptureReader(paths, N, pinned=False, blocks_per_read=3)
    assert r.n_blocks == 7
    got = [(k, blk.copy()) for k, blk in r]
    assert [k for k, _ in got] == [0, 3, 6]
    for k, blk in got:
        assert blk.shape == (2, 3 * 2 * N)
        have = min(3, 7 - k) * 2 * N
        assert np.array_equal(blk[0][:have], a[2 * N * k: 2 * N * k + have]) and np.array_equal(blk[1][:have], b[2 * N * k: 2 * N * k + have])
        assert not blk[:, have:].any()

"""BASELINE.json configs[2], configs[3] and configs[4] at their STATED sizes against the oracle, on the B200 (pytest -m gpu):

  configs[2]  one 20 MS/s stream, nperseg 1024 and 4096, two full 20 M-sample blocks (T = 19 531 / 4 882 columns: the 296-CTA
              round-robin of spectro_r16 and the stand-alone row-mean kernel above 48 partial rows), one burst straddling the
              block boundary (carry, analyze.py:383-398; block-end drop, :415-417)
  configs[4]  1000 pulses per 1-s block at 2.4 MS/s (near-threshold and loud), one stream and a 64-stream batch: candidate
              counts before and after the shadow filter (analyze.py:315-328)
  configs[3]  offline replay of 16 station channels x 60 blocks of 300 kS/s with carry: block-by-block, eight blocks per
              launch, and sharded over two ranks gathered with shard.gather_signals -- all identical, and equal to the oracle
"""
import datetime
import os
import socket
from dataclasses import replace

import numpy as np
import pytest

from oracle import restatement as R
from pyradiotracking_b200 import shard, synth
from pyradiotracking_b200.analyze import BatchAnalyzer
from pyradiotracking_b200.replay import replay
from tests import parity

pytestmark = pytest.mark.gpu

T0 = datetime.datetime(2026, 3, 3, 3, 3, 3)


def _kw(w, **over):
    kw = dict(device="0", calibration_db=0.0, sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=w.nperseg,
              fft_window="hamming", signal_min_duration_ms=w.signal_min_duration_ms, signal_max_duration_ms=w.signal_max_duration_ms,
              signal_threshold_dbw=w.signal_threshold_dbw, snr_threshold_db=w.snr_threshold_db, sdr_callback_length=w.block_samples)
    kw.update(over)
    return kw


def _sig_tuple(s):
    return (s.device, s.ts, s.frequency, s.duration, s.max, s.avg, s.std, s.noise, s.snr)


# ---------------------------------------------------------------------------------------------------------------------------
# configs[2]: wideband single stream at the full block length
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w", [synth.C3A, synth.C3B], ids=["n1024", "n4096"])
def test_wideband_20M_sample_blocks_match_the_oracle(w):
    nb = 2
    cap = synth.make_stream(w, 0, nb)                          # [2, 40 MB]; make_stream lays one burst across the block boundary
    kw = _kw(w)
    P = parity.oracle_params(kw)
    ora = R.OracleAnalyzer(P)
    ba = BatchAnalyzer(**parity.batch_kwargs(kw))
    try:
        assert ba.engine.T == w.block_samples // w.nperseg
        last = None
        tot = dict(oracle=0, gpu=0, near_threshold_mismatch=0)
        carried = 0
        for b in range(nb):
            ts0 = parity.block_ts(T0, b, w.block_samples, w.sample_rate)
            _, _, S, found, kept = ora.process_block(cap[b], ts0)
            filtered, sigs, keys = ba.process_blocks(cap[b][None, :], [ts0])[0]
            st = parity.compare_block(P, S, last, found, sigs, keys)
            parity.compare_spectrogram(P, S, ba.engine.read_spectrogram(0), ba.engine.read_row_means(0), tag=f"c3_full/{w.nperseg}/b{b}")
            for k in tot:
                tot[k] += st[k]
            if st["near_threshold_mismatch"] == 0:
                assert [(s.ts, s.frequency) for s in filtered] == [(d.ts, d.frequency) for d in kept]
            carried += sum(1 for k in keys if k[1] < 0)
            last = S
        assert tot["gpu"] >= 20 and tot["near_threshold_mismatch"] == 0
        assert carried >= 1                                    # the straddling burst was re-found from the second block
    finally:
        ba.close()


# ---------------------------------------------------------------------------------------------------------------------------
# configs[4]: dense pulses at 2.4 MS/s, one stream and a 64-stream batch
# ---------------------------------------------------------------------------------------------------------------------------
C5B_LOUD = replace(synth.C5B, name="c5-dense-loud-2.4M", amp_db_over_thr=(6.0, 14.0))


@pytest.fixture(scope="module", params=[synth.C5B, C5B_LOUD], ids=["near-threshold", "loud"])
def dense(request):
    w = request.param
    nb, n_distinct = 2, 4
    caps = [synth.make_stream(w, 200 + i, nb) for i in range(n_distinct)]
    P = parity.oracle_params(_kw(w))
    want = []
    for c in caps:
        ora = R.OracleAnalyzer(P)
        last, blocks = None, []
        for b in range(nb):
            _, _, S, found, kept = ora.process_block(c[b], parity.block_ts(T0, b, w.block_samples, w.sample_rate))
            blocks.append((S, last, found, kept))
            last = S
        want.append(blocks)
    return w, caps, P, want


def test_dense_pulses_single_stream_matches_the_oracle(dense):
    w, caps, P, want = dense
    ba = BatchAnalyzer(**parity.batch_kwargs(_kw(w)))
    try:
        n_pre = n_post = 0
        for b, (S, last, found, kept) in enumerate(want[0]):
            ts0 = parity.block_ts(T0, b, w.block_samples, w.sample_rate)
            filtered, sigs, keys = ba.process_blocks(caps[0][b][None, :], [ts0])[0]
            st = parity.compare_block(P, S, last, found, sigs, keys)
            assert st["near_threshold_mismatch"] == 0
            assert len(sigs) == len(found) and len(filtered) == len(kept)          # pre- and post-shadow counts
            assert [(s.ts, s.frequency) for s in filtered] == [(d.ts, d.frequency) for d in kept]
            n_pre += len(sigs)
            n_post += len(filtered)
        print(f"dense {w.name}: {n_pre} candidates -> {n_post} after the shadow filter (one stream, 2 blocks)")
        assert n_pre > 0
        if w is C5B_LOUD:
            assert n_pre >= 500 and n_post < n_pre // 4                            # ~900 -> ~25 per block
    finally:
        ba.close()


def test_dense_pulses_64_stream_batch_matches_the_oracle(dense):
    w, caps, P, want = dense
    n = 64
    cal = [0.0] * n
    ba = BatchAnalyzer(**parity.batch_kwargs(_kw(w), devices=[str(i) for i in range(n)], calibration=cal))
    try:
        pre = post = 0
        for b in range(2):
            ts0 = parity.block_ts(T0, b, w.block_samples, w.sample_rate)
            blk = np.stack([caps[s % len(caps)][b] for s in range(n)])
            res = ba.process_blocks(blk, [ts0] * n)
            for s in range(n):
                S, last, found, kept = want[s % len(caps)][b]
                filtered, sigs, keys = res[s]
                assert keys == [d.key() for d in found], (s, b)
                assert [(x.ts, x.frequency, x.duration) for x in filtered] == [(d.ts, d.frequency, d.duration) for d in kept]
                assert all(x.device == str(s) for x in sigs)
                if s < len(caps):
                    parity.compare_block(P, S, last, found, sigs, keys)
                pre += len(sigs)
                post += len(filtered)
        print(f"dense {w.name}: 64-stream batch, {pre} candidates -> {post} after the shadow filter (2 blocks)")
        assert pre > 0 and ba.engine.truncated == 0
    finally:
        ba.close()


# ---------------------------------------------------------------------------------------------------------------------------
# configs[3]: offline replay with carry, one process / several blocks per launch / two ranks
# ---------------------------------------------------------------------------------------------------------------------------
N_CH, N_BLK = 16, 60


@pytest.fixture(scope="module")
def recordings(tmp_path_factory):
    d = tmp_path_factory.mktemp("c4")
    w = synth.C4
    paths, caps = [], []
    for c in range(N_CH):
        cap = synth.make_stream(w, 300 + c, N_BLK)            # [60, 600 000]
        p = d / f"chan{c:02d}.bin"
        cap.tofile(p)
        paths.append(str(p))
        caps.append(cap)
    return w, paths, caps


def _replay_all(w, paths, devices, blocks_per_launch, cuda_device=0):
    kw = _kw(w)
    ba = BatchAnalyzer(**parity.batch_kwargs(kw, devices=devices, calibration=[0.0] * len(devices), cuda_device=cuda_device,
                                             blocks_per_launch=blocks_per_launch))
    per_stream = [[] for _ in devices]          # per stream: (block, n_before, [signal tuples])
    try:
        n = replay(paths, ba, T0, on_block=lambda b, res: [per_stream[s].append((b, res[s][1], [_sig_tuple(x) for x in res[s][0]]))
                                                           for s in range(len(devices))])
        assert n == N_BLK
    finally:
        ba.close()
    return per_stream


@pytest.fixture(scope="module")
def replay_single(recordings):
    w, paths, _ = recordings
    return _replay_all(w, paths, [str(c) for c in range(N_CH)], 1)


def test_replay_60_blocks_matches_the_oracle(recordings, replay_single):
    w, paths, caps = recordings
    P = parity.oracle_params(_kw(w))
    n_sig = n_carry = 0
    for c in (0, 5, 11, 15):
        ora = R.OracleAnalyzer(P._replace(device=str(c)))
        for b in range(N_BLK):
            ts0 = T0 + b * datetime.timedelta(seconds=w.block_samples / w.sample_rate)
            _, _, S, found, kept = ora.process_block(caps[c][b], ts0)
            blk, n_before, sigs = replay_single[c][b]
            assert blk == b and n_before == len(found)
            assert [(t[0], t[1], t[2], t[3]) for t in sigs] == [(str(c), d.ts, d.frequency, d.duration) for d in kept]
            for t, d in zip(sigs, kept):
                assert max(abs(t[4] - d.max), abs(t[5] - d.avg), abs(t[6] - d.std), abs(t[7] - d.noise), abs(t[8] - d.snr)) <= parity.DB_ATOL
            n_sig += len(kept)
            n_carry += sum(1 for d in kept if d.start < 0)
    assert n_sig > 200 and n_carry > 5          # the carry was exercised across block boundaries


def test_replay_eight_blocks_per_launch_is_identical(recordings, replay_single):
    """rt_config.blocks_per_launch: 60 blocks = 7 launches of 8 + one of 4 (padded); the carry between the blocks of a launch
    never leaves the engine.  Bit-identical Signals."""
    w, paths, _ = recordings
    got = _replay_all(w, paths, [str(c) for c in range(N_CH)], 8)
    assert got == replay_single


def _rank_worker(rank, world, port, paths, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.stream_range(len(paths), world, rank)
        local = _replay_all(synth.C4, [paths[i] for i in mine], [str(i) for i in mine], 4, cuda_device=rank % torch.cuda.device_count())
        got = shard.gather_signals(local, len(paths))
        if rank == 0:
            q.put(got)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_replay_sharded_over_two_ranks_gathers_the_single_process_result(recordings, replay_single):
    """One process per GPU (both ranks share cuda:0 on a one-GPU box), contiguous channel ranges, no data-path collective;
    rank 0 gathers the per-channel Signal lists with shard.gather_signals."""
    import torch.multiprocessing as mp

    w, paths, _ = recordings
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, paths, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == replay_single

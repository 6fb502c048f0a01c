"""BASELINE.json configs[1] at FULL size on the B200 (64 streams x 2.4 M samples per block): too big for the
float64 oracle to check cell by cell, so parity is held through size-independent properties:
  * streams are independent: every stream of the 64-batch == the same capture run alone;
  * determinism: the same blocks from a reset engine give byte-identical records;
  * the two FFT kernels (register 16x16 and shared-memory Stockham) agree on the detected runs;
  * Parseval: sum over bins of a spectrogram column == energy of the detrended, windowed segment (numpy, exact bytes);
  * the oracle itself on two full-size streams (it finishes in seconds per stream).
"""
import datetime

import numpy as np
import pytest

from oracle import restatement as R
from pyradiotracking_b200 import engine as E
from pyradiotracking_b200 import synth
from pyradiotracking_b200.analyze import BatchAnalyzer
from tests import parity

pytestmark = pytest.mark.gpu

W = synth.C2
N_STREAMS = 64
N_DISTINCT = 4
N_BLOCKS = 2


def _kwargs(n, **extra):
    return dict(devices=[str(i) for i in range(n)], calibration_db=[0.0] * n, sample_rate=W.sample_rate,
                center_freq=W.center_freq, fft_nperseg=W.nperseg, fft_window="hamming",
                signal_min_duration_ms=W.signal_min_duration_ms, signal_max_duration_ms=W.signal_max_duration_ms,
                signal_threshold_dbw=W.signal_threshold_dbw, snr_threshold_db=W.snr_threshold_db,
                sdr_callback_length=W.block_samples, **extra)


@pytest.fixture(scope="module")
def batch():
    distinct = [synth.make_stream(W, 100 + i, N_BLOCKS) for i in range(N_DISTINCT)]
    cap = np.empty((N_BLOCKS, N_STREAMS, W.block_bytes), dtype=np.uint8)
    for s in range(N_STREAMS):
        cap[:, s, :] = distinct[s % N_DISTINCT]
    return distinct, cap


@pytest.fixture(scope="module")
def full_records(batch):
    _, cap = batch
    ba = BatchAnalyzer(**_kwargs(N_STREAMS))
    try:
        recs = [ba.engine.process(cap[b]) for b in range(N_BLOCKS)]
        spec = ba.engine.read_spectrogram(5)
        rm = ba.engine.read_row_means(5)
    finally:
        ba.close()
    return recs, spec, rm


def _strip(rec, stream):
    r = rec[rec["stream"] == stream].copy()
    r["stream"] = 0
    return r


def test_batch_streams_equal_single_stream_runs(batch, full_records):
    distinct, cap = batch
    recs, _, _ = full_records
    for d in range(N_DISTINCT):
        one = BatchAnalyzer(**_kwargs(1))
        try:
            for b in range(N_BLOCKS):
                alone = one.engine.process(distinct[d][b][None, :])
                for s in range(d, N_STREAMS, N_DISTINCT):          # every copy of this capture inside the batch
                    assert _strip(recs[b], s).tobytes() == alone.tobytes()
                assert len(alone) > 0
        finally:
            one.close()


def test_full_batch_is_deterministic(batch, full_records):
    _, cap = batch
    recs, _, _ = full_records
    ba = BatchAnalyzer(**_kwargs(N_STREAMS))
    try:
        again = [ba.engine.process(cap[b]) for b in range(N_BLOCKS)]
    finally:
        ba.close()
    assert all(a.tobytes() == b.tobytes() for a, b in zip(recs, again))


def test_both_fft_kernels_find_the_same_runs(batch, full_records):
    distinct, _ = batch
    recs, _, _ = full_records
    gen = BatchAnalyzer(**_kwargs(N_DISTINCT, fft_impl=E.FFT_GENERIC))
    try:
        for b in range(N_BLOCKS):
            g = gen.engine.process(np.stack([d[b] for d in distinct]))
            for s in range(N_DISTINCT):
                a, c = _strip(recs[b], s), _strip(g, s)
                ka = {(int(x["fi"]), int(x["start"]), int(x["end"])) for x in a}
                kc = {(int(x["fi"]), int(x["start"]), int(x["end"])) for x in c}
                assert len(ka ^ kc) <= max(1, len(ka) // 100)      # fp32 rounding differs between the two FFTs
                common = sorted(ka & kc)
                ia = {(int(x["fi"]), int(x["start"]), int(x["end"])): x for x in a}
                ic = {(int(x["fi"]), int(x["start"]), int(x["end"])): x for x in c}
                for k in common:
                    assert abs(ia[k]["mean_lin"] / ic[k]["mean_lin"] - 1) < 1e-4
    finally:
        gen.close()


def test_tensor_core_kernel_finds_the_same_runs(batch, full_records):
    """RT_FFT_TC256 (tcgen05 stage 1, TMEM accumulators) against the register kernel at BASELINE configs[1] size."""
    distinct, _ = batch
    recs, spec, rm = full_records
    tc = BatchAnalyzer(**_kwargs(N_DISTINCT, fft_impl=E.FFT_TC256))
    try:
        for b in range(N_BLOCKS):
            g = tc.engine.process(np.stack([d[b] for d in distinct]))
            for s in range(N_DISTINCT):
                a, c = _strip(recs[b], s), _strip(g, s)
                ka = {(int(x["fi"]), int(x["start"]), int(x["end"])) for x in a}
                kc = {(int(x["fi"]), int(x["start"]), int(x["end"])) for x in c}
                assert len(ka ^ kc) <= max(1, len(ka) // 100)      # fp32 rounding differs between the two FFTs
                ia = {(int(x["fi"]), int(x["start"]), int(x["end"])): x for x in a}
                ic = {(int(x["fi"]), int(x["start"]), int(x["end"])): x for x in c}
                for k in sorted(ka & kc):
                    assert abs(ia[k]["mean_lin"] / ic[k]["mean_lin"] - 1) < 1e-4
                    assert abs(ia[k]["max_lin"] / ic[k]["max_lin"] - 1) < 1e-4
        # last block, stream 5 == distinct[1]: cells and row means against the register kernel
        got = tc.engine.read_spectrogram(5 % N_DISTINCT).astype(np.float64)
        ref = spec.astype(np.float64)
        big = ref >= 1e-3 * ref.max(axis=1, keepdims=True)
        assert np.max(np.abs(got[big] - ref[big]) / ref[big]) < 2e-4
        assert np.max(np.abs(tc.engine.read_row_means(5 % N_DISTINCT).astype(np.float64) / rm.astype(np.float64) - 1)) < 1e-5
    finally:
        tc.close()


def test_parseval_on_full_size_columns(batch, full_records):
    distinct, _ = batch
    _, spec, _ = full_records                       # stream 5 == distinct[1], last block
    u8 = distinct[5 % N_DISTINCT][N_BLOCKS - 1]
    win = R.resolve_window("hamming", 256)
    scale = 1.0 / (W.sample_rate * (win * win).sum()) / 127.5 ** 2
    T = spec.shape[0]
    for t in (0, 1, T // 3, T - 2, T - 1):
        b = u8[512 * t: 512 * t + 512].astype(np.float64)
        x = (b[0::2] - b[0::2].mean()) + 1j * (b[1::2] - b[1::2].mean())
        energy = 256 * np.sum(np.abs(x * win) ** 2) * scale         # Parseval: sum_k |X_k|^2 = n sum_n |x_n|^2
        assert abs(spec[t].astype(np.float64).sum() / energy - 1) < 2e-6


def test_oracle_on_two_full_size_streams(batch, full_records):
    distinct, _ = batch
    recs, spec, rm = full_records
    t0 = datetime.datetime(2026, 7, 7, 7, 7, 7)
    ba = BatchAnalyzer(**_kwargs(1))
    for d in (0, 1):
        P = R.Params.make(sample_rate=W.sample_rate, center_freq=W.center_freq)
        ora = R.OracleAnalyzer(P)
        last = None
        for b in range(N_BLOCKS):
            _, _, S, found, kept = ora.process_block(distinct[d][b], t0)
            per = ba.finalize(_strip(recs[b], d), [t0])[0]
            stats = parity.compare_block(P, S, last, found, per[0], per[1])
            assert stats["near_threshold_mismatch"] == 0 and stats["gpu"] > 0
            last = S
        if d == 1:                                   # stream 5 of the batch is distinct[1]: its last spectrogram
            parity.compare_spectrogram(P, S, spec, rm, tag="c2_full/stream5")
    ba.close()

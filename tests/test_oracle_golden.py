"""The numpy restatement (oracle/restatement.py) against fixtures produced by the
UNMODIFIED reference (tests/golden, see oracle/make_golden.py).  CPU only."""
import datetime

import numpy as np
import pytest

from oracle import restatement as R
from oracle.cases import CASES, sha256
from tests import golden_io


def _params(kw):
    win = kw["fft_window"]
    return R.Params.make(
        device=kw["device"], calibration_db=kw["calibration_db"], sample_rate=kw["sample_rate"],
        center_freq=kw["center_freq"], fft_nperseg=kw["fft_nperseg"], fft_window=tuple(win) if isinstance(win, list) else win,
        signal_min_duration_ms=kw["signal_min_duration_ms"], signal_max_duration_ms=kw["signal_max_duration_ms"],
        signal_threshold_dbw=kw["signal_threshold_dbw"], snr_threshold_db=kw["snr_threshold_db"])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_restatement_matches_reference_fixture(case):
    g = golden_io.load(case.name)
    cap = case.capture()
    assert sha256(cap) == g.meta["input_sha256"], "synthetic capture changed: regenerate tests/golden (python -m oracle.make_golden)"
    P = _params(g.meta["analyzer"])
    ora = R.OracleAnalyzer(P)
    block_len = datetime.timedelta(seconds=cap.shape[1] // 2 / P.sample_rate)
    n_total = 0
    for b, gb in enumerate(g.blocks):
        ts_start = g.t0 + (b - 1) * block_len                      # analyze.py:218-231 with a drift-free clock
        freqs, times, S, found, kept = ora.process_block(cap[b], ts_start)
        rows, cols = golden_io.digest_index(S.shape[0], S.shape[1])
        np.testing.assert_allclose(S[np.ix_(rows, cols)], gb.cells, rtol=1e-9, atol=0)
        np.testing.assert_allclose(S.mean(axis=1), gb.rowmean, rtol=1e-11)
        assert len(found) == len(gb.ts_us)
        assert [golden_io.us(d.ts - golden_io.EPOCH) for d in found] == gb.ts_us.tolist()
        assert [golden_io.us(d.duration) for d in found] == gb.dur_us.tolist()
        assert [d.frequency for d in found] == gb.freq.tolist()
        if len(found):
            got = np.array([[d.max, d.avg, d.std, d.noise, d.snr] for d in found])
            np.testing.assert_allclose(got, gb.stats, rtol=0, atol=1e-8)
        kept_ids = {id(d) for d in kept}
        assert [id(d) in kept_ids for d in found] == gb.kept.tolist()
        n_total += len(found)
    assert n_total > 0


@pytest.mark.parametrize("name", ["c1_default_300k", "int_stride_256k", "tiny_T64", "loud_floor"])
def test_sequential_and_run_formulations_agree(name):
    """extract_sequential (the reference's visiting order) == extract_runs (per maximal run)."""
    case = {c.name: c for c in CASES}[name]
    g = golden_io.load(name)
    cap = case.capture()
    P = _params(g.meta["analyzer"])
    a, b = R.OracleAnalyzer(P, sequential=True), R.OracleAnalyzer(P, sequential=False)
    t0 = g.t0
    for blk in cap[:2]:
        fa = a.process_block(blk, t0)[3]
        fb = b.process_block(blk, t0)[3]
        assert fa == fb

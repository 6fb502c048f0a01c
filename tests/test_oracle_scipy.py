"""Spectrogram restatement vs the installed scipy (the de-facto pin of the unvendored
dependency) and vs the reference itself when it is present."""
import datetime

import numpy as np
import pytest

from oracle import ref_harness
from oracle import restatement as R
from pyradiotracking_b200 import synth


@pytest.mark.parametrize("nperseg,window", [(256, "hamming"), (1024, "hann"), (64, ("kaiser", 6.0)), (100, "boxcar")])
def test_spectrogram_matches_scipy(nperseg, window):
    scipy_signal = pytest.importorskip("scipy.signal")
    u8 = synth.make_stream(synth.C1, 5, 1)[0][: 2 * 40_000]
    x = R.bytes_to_iq(u8)
    f0, t0, s0 = scipy_signal.spectrogram(x, fs=300000, window=window, nperseg=nperseg, noverlap=0, return_onesided=False)
    f1, t1, s1 = R.spectrogram(x, 300000, window, nperseg)
    assert np.array_equal(f0, f1) and np.array_equal(t0, t1)
    np.testing.assert_allclose(s1, s0, rtol=1e-9)


def test_window_is_scipy_periodic_window():
    sw = pytest.importorskip("scipy.signal")
    for name in ("hamming", "hann", "boxcar"):
        for n in (8, 256, 1000):
            np.testing.assert_array_equal(R.resolve_window(name, n), sw.get_window(name, n))


def test_empty_and_single_column():
    P = R.Params.make()
    f, t, S = R.spectrogram(np.zeros(100, complex), 300000, "hamming", 256)
    assert S.shape == (256, 0) and len(t) == 0
    assert R.extract_runs(P, f, t, S, None, datetime.datetime(2026, 1, 1)) == []
    f, t, S = R.spectrogram(R.bytes_to_iq(synth.make_stream(synth.C1, 0, 1)[0][:512]), 300000, "hamming", 256)
    with pytest.raises(IndexError):       # analyze.py:354 indexes times[1]
        R.extract_sequential(P, f, t, S, None, datetime.datetime(2026, 1, 1))


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_restatement_matches_live_reference():
    w = synth.C1
    cap = synth.make_stream(w, 17, 3)
    t0 = datetime.datetime(2026, 5, 5, 5, 5, 5)
    ref = ref_harness.ReferenceRunner(t0, sample_rate=w.sample_rate)
    ora = R.OracleAnalyzer(R.Params.make(sample_rate=w.sample_rate))
    bl = datetime.timedelta(seconds=1)
    for b in range(3):
        queued = ref.feed(cap[b])
        _, _, S, found, kept = ora.process_block(cap[b], t0 + (b - 1) * bl)
        np.testing.assert_allclose(S, ref.spectrogram_last, rtol=1e-9)
        assert [(s.ts, s.frequency, s.duration) for s in ref.pre_shadow[-1]] == [(d.ts, d.frequency, d.duration) for d in found]
        assert [(s.ts, s.frequency) for s in queued] == [(d.ts, d.frequency) for d in kept]

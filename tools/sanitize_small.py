#!/usr/bin/env python
"""Small runs of every product kernel for compute-sanitizer (memcheck / racecheck): the smoke case, an 8-stream batch with two
launches in flight (lean schedule forced), a 2-block nperseg-1024 / 4096 stream and a generic-size stream.
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import datetime
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from pyradiotracking_b200 import engine as E  # noqa: E402
from pyradiotracking_b200 import synth  # noqa: E402
from pyradiotracking_b200.analyze import BatchAnalyzer  # noqa: E402


def run(w, n_streams, nperseg, blocks=2, **kw):
    ba = BatchAnalyzer(devices=[str(i) for i in range(n_streams)], calibration_db=[0.0] * n_streams, sample_rate=w.sample_rate,
                       center_freq=w.center_freq, fft_nperseg=nperseg, fft_window="hamming",
                       signal_min_duration_ms=w.signal_min_duration_ms, signal_max_duration_ms=w.signal_max_duration_ms,
                       signal_threshold_dbw=w.signal_threshold_dbw, snr_threshold_db=w.snr_threshold_db,
                       sdr_callback_length=w.block_samples, cuda_device=0, **kw)
    caps = [synth.make_stream(w, i, blocks) for i in range(min(2, n_streams))]
    host = np.stack([np.stack([caps[s % len(caps)][b] for s in range(n_streams)]) for b in range(blocks)])
    ts = [datetime.datetime(2026, 1, 1)] * n_streams
    ba.submit(host[0])
    n = 0
    for b in range(blocks):
        if b + 1 < blocks:
            ba.submit(host[b + 1])
        n += sum(len(r[0]) for r in ba.collect(ts))
    ba.close()
    return n


if __name__ == "__main__":
    print("c1 single", run(synth.C1, 1, 256))
    print("c1 x8 lean", run(synth.C1, 8, 256, scan_schedule=E.SCAN_LEAN))
    print("c1 x4 serial tc256", run(synth.C1, 4, 256, scan_schedule=E.SCAN_SERIAL, fft_impl=E.FFT_TC256))
    import dataclasses
    small = dataclasses.replace(synth.C3A, block_samples=1 << 21) if dataclasses.is_dataclass(synth.C3A) else synth.C3A._replace(block_samples=1 << 21)
    print("nperseg 1024", run(small, 1, 1024))
    print("nperseg 4096", run(small, 1, 4096))
    print("nperseg 512 generic", run(synth.C1, 1, 512))

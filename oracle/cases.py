"""Golden parity cases: small instances of every BASELINE.json configuration plus the
edge cases of SURVEY.md §8a.  TEST INFRASTRUCTURE ONLY (see oracle/restatement.py).

Each case names a synthetic capture (regenerated from its seed, checked by sha256)
and the analyzer keys it is run with.  `oracle/make_golden.py` runs the unmodified
reference on them and writes `tests/golden/<name>.npz`.
"""
import hashlib
from dataclasses import dataclass, field, replace
from typing import Dict

import numpy as np

from pyradiotracking_b200 import synth


@dataclass(frozen=True)
class Case:
    name: str
    workload: synth.Workload
    stream: int = 0
    n_blocks: int = 3
    analyzer: Dict[str, object] = field(default_factory=dict)   # overrides of SignalAnalyzer keys

    def capture(self) -> np.ndarray:
        return synth.make_stream(self.workload, self.stream, self.n_blocks)

    def analyzer_kwargs(self) -> Dict[str, object]:
        w = self.workload
        kw = dict(
            device="0", calibration_db=0.0, sample_rate=w.sample_rate, center_freq=w.center_freq, gain=49.6,
            fft_nperseg=w.nperseg, fft_window="hamming", signal_min_duration_ms=w.signal_min_duration_ms,
            signal_max_duration_ms=w.signal_max_duration_ms, signal_threshold_dbw=w.signal_threshold_dbw,
            snr_threshold_db=w.snr_threshold_db, verbose=0, sdr_max_restart=3, sdr_timeout_s=2, state_update_s=300,
            sdr_callback_length=w.block_samples,
        )
        kw.update(self.analyzer)
        return kw


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


_W = synth.Workload
CASES = [
    # config 1: the reference's own default analyze path
    Case("c1_default_300k", synth.C1, n_blocks=4),
    # config 2: one stream of the 64 x 2.4 MS/s batch, incl. exact 75/375-column bursts (float64 duration boundary)
    Case("c2_stream_2400k", synth.C2, stream=3, n_blocks=3),
    # config 3: wideband, nperseg 1024 / 4096 (callback length cut to 2 M samples to keep the fixture small)
    Case("c3a_20M_n1024", replace(synth.C3A, block_samples=2_000_000), n_blocks=2),
    Case("c3b_20M_n4096", replace(synth.C3B, block_samples=2_000_000), n_blocks=2),
    # config 5: dense pulses near threshold -> O(S^2) shadow filter
    Case("c5_dense_300k", synth.C5, n_blocks=3),
    Case("c5_dense_loud_300k", synth.C5L, n_blocks=2),
    # integer probe stride (0.008 / (256/256000) == 8.0): hit-or-miss runs of stride-1 columns
    Case("int_stride_256k", _W(11, "int-stride", 256_000, 256_000, n_blocks=3, sigma_lsb=1.0,
                               pulses_per_block=(40, 40), pulse_ms=(6.5, 9.5)), n_blocks=3),
    # block length not a multiple of nperseg, calibration offset, other window / thresholds
    Case("calib_hann_ragged", _W(12, "calib-hann", 300_000, 250_123, n_blocks=3, sigma_lsb=1.5,
                                 pulses_per_block=(8, 8), pulse_ms=(10.0, 30.0)), n_blocks=3,
         analyzer=dict(calibration_db=3.5, fft_window="hann", signal_threshold_dbw=-88.0, snr_threshold_db=7.0,
                       device="7", center_freq=433_920_000)),
    Case("kaiser_n512", _W(13, "kaiser-512", 1_024_000, 512_000, nperseg=512, n_blocks=3, sigma_lsb=2.0,
                           pulses_per_block=(10, 10), pulse_ms=(8.0, 40.0)), n_blocks=3,
         analyzer=dict(fft_window=("kaiser", 8.0))),
    # tiny blocks: T = 64 columns, carry reaches almost a whole previous block
    Case("tiny_T64", _W(14, "tiny", 300_000, 16_384, n_blocks=6, sigma_lsb=1.0,
                        pulses_per_block=(1, 2), pulse_ms=(10.0, 30.0)), n_blocks=6),
    # noise floor ABOVE the absolute threshold: every cell passes the first test, SNR decides
    Case("loud_floor", _W(15, "loud-floor", 300_000, 300_000, n_blocks=2, sigma_lsb=6.0,
                          pulses_per_block=(6, 6), pulse_ms=(10.0, 30.0), amp_lsb=(40.0, 90.0)), n_blocks=2),
]
BY_NAME = {c.name: c for c in CASES}

"""Run the UNMODIFIED reference `SignalAnalyzer` offline (build container only).

TEST INFRASTRUCTURE ONLY.  `/root/reference` does not exist on the GPU box, so
nothing in `tests -m gpu`, `smoke()` or `bench.py` imports this module at run time;
it is used by `oracle/make_golden.py` to produce the fixtures in `tests/golden/`
and by `tests/test_oracle_vs_reference.py` (skipped when the reference is absent).

The reference imports `pytz` and `rtlsdr` at module top (analyze.py:10-11), neither
of which is installed; only `pytz.utc`/`pytz.UTC` are used on the hot path
(analyze.py:186,189,449) and `rtlsdr` is never touched when `device` is an integer
string (analyze.py:89-91).  Two stub modules make the import succeed; the analyzer
is then driven exactly like pyrtlsdr's async loop would drive it.
"""
import datetime
import os
import signal as _signal
import sys
import types
from typing import List

import numpy as np

REFERENCE_ROOT = os.environ.get("RT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "radiotracking", "analyze.py"))


def _install_stubs():
    if "pytz" not in sys.modules:
        try:
            import pytz  # noqa: F401
        except ImportError:
            m = types.ModuleType("pytz")
            m.utc = m.UTC = datetime.timezone.utc
            sys.modules["pytz"] = m
    if "rtlsdr" not in sys.modules:
        try:
            import rtlsdr  # noqa: F401
        except ImportError:
            m = types.ModuleType("rtlsdr")
            sub = types.ModuleType("rtlsdr.rtlsdr")

            class LibUSBError(Exception):
                pass

            class RtlSdr:  # never instantiated offline
                @staticmethod
                def get_device_index_by_serial(serial):
                    raise LibUSBError(serial)

            sub.LibUSBError = LibUSBError
            m.rtlsdr = sub
            m.RtlSdr = RtlSdr
            sys.modules["rtlsdr"] = m
            sys.modules["rtlsdr.rtlsdr"] = sub


def load_reference():
    """-> the reference's `radiotracking.analyze` module."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import radiotracking.analyze as ra  # type: ignore

    return ra


class _Clock:
    """Stand-in for the `datetime` module inside analyze.py: `datetime.datetime.now()`
    is scripted, everything else is the real thing (analyze.py:204,210,218-231)."""

    def __init__(self, t0: datetime.datetime):
        self._now = t0
        outer = self

        class _DT(datetime.datetime):
            @classmethod
            def now(cls, tz=None):
                return outer._now

        self.datetime = _DT
        self.timedelta = datetime.timedelta
        self.timezone = datetime.timezone

    def set(self, t):
        self._now = t


class _ListQueue:
    def __init__(self):
        self.items: List[object] = []

    def put(self, x):
        self.items.append(x)


class _FakeSdr:
    cancelled = False

    def cancel_read_async(self):
        self.cancelled = True


class ReferenceRunner:
    """One reference analyzer fed block by block with a deterministic clock."""

    def __init__(self, t0: datetime.datetime, **kwargs):
        import multiprocessing

        self.ra = load_reference()
        self.clock = _Clock(t0)
        self.ra.datetime = self.clock          # module-level name used by analyze.py
        self.t0 = t0
        self.queue = _ListQueue()
        cfg = dict(
            device="0", calibration_db=0.0, sample_rate=300000, center_freq=150150000, gain=49.6,
            fft_nperseg=256, fft_window="hamming", signal_min_duration_ms=8, signal_max_duration_ms=40,
            signal_threshold_dbw=-90.0, snr_threshold_db=5.0, verbose=0, sdr_max_restart=3,
            sdr_timeout_s=2, state_update_s=300, sdr_callback_length=None,
        )
        cfg.update(kwargs)
        self.an = self.ra.SignalAnalyzer(signal_queue=self.queue, last_data_ts=multiprocessing.Value("d", 0.0), **cfg)
        self.an.last_state = None              # normally set in run() (analyze.py:151)
        self.an.sdr = _FakeSdr()               # normally set in run() (analyze.py:148)
        self.n_blocks = 0
        self.pre_shadow: List[list] = []
        # record the pre-shadow list too (process_samples only queues the filtered one)
        inner = self.an.extract_signals

        def spy(freqs, times, spectrogram, ts_start):
            sigs = inner(freqs, times, spectrogram, ts_start)
            self.pre_shadow.append(list(sigs))
            return sigs

        self.an.extract_signals = spy

    def feed(self, u8_block: np.ndarray):
        """One callback (analyze.py:192-268) -> list of reference `Signal`s queued by it."""
        iq = np.ascontiguousarray(u8_block, dtype=np.uint8).astype(np.float64).view(np.complex128)
        iq /= 127.5
        iq -= 1 + 1j
        block_len = datetime.timedelta(seconds=len(iq) / self.an.sample_rate)
        # wall clock == virtual clock: no drift (analyze.py:217-229)
        self.clock.set(self.t0 + self.n_blocks * block_len)
        before = len(self.queue.items)
        old = _signal.signal(_signal.SIGALRM, _signal.SIG_IGN)
        try:
            self.an.process_samples(iq, None)
        finally:
            _signal.alarm(0)                   # analyze.py:208 arms SIGALRM on every callback
            _signal.signal(_signal.SIGALRM, old)
        self.n_blocks += 1
        Sig = sys.modules["radiotracking"].Signal
        return [m for m in self.queue.items[before:] if isinstance(m, Sig)]

    @property
    def spectrogram_last(self):
        return self.an._spectrogram_last

"""CPU restatement of the reference's cross-device matcher.

TEST INFRASTRUCTURE ONLY: nothing under `pyradiotracking_b200/` imports this module; it is the checker of the
native matcher (`include/rt_matcher.h`) in `tests/test_matcher.py` and the baseline of `tools/bench_matcher.py`.

Plain Python on the same objects the reference uses (datetime / timedelta / float), following
  radiotracking/match.py:33-82        SignalMatcher.__init__ / consume / add
  radiotracking/__init__.py:285-334   MatchingSignal: duration = max, ts = min, frequency = statistics.median, _avgs
  radiotracking/__init__.py:337-383   MatchingSignal.has_member
  radiotracking/__init__.py:385-406   MatchingSignal.add_member
Pinned against the unmodified reference by `oracle/make_matcher_golden.py` -> `tests/golden/matcher_*.json`.
"""
import datetime
import statistics
from typing import Dict, List, Optional


class OracleGroup:
    def __init__(self, devices: List[str]):
        self.devices = devices
        self._sigs: Dict[str, object] = {}

    # __init__.py:298-329
    @property
    def duration(self):
        return max([s.duration for s in self._sigs.values()])

    @property
    def ts(self):
        return min([s.ts for s in self._sigs.values()])

    @property
    def frequency(self):
        return statistics.median([s.frequency for s in self._sigs.values()])

    # __init__.py:337-383
    def has_member(self, sig, time_diff, bandwidth, duration_diff) -> bool:
        if sig.frequency - bandwidth / 2 > self.frequency:
            return False
        if sig.frequency + bandwidth / 2 < self.frequency:
            return False
        if sig.ts - time_diff > (self.ts + self.duration):
            return False
        if (sig.ts + sig.duration) + time_diff < self.ts:
            return False
        if duration_diff:
            if sig.duration - (duration_diff / 2) > self.duration:
                return False
            if sig.duration + (duration_diff / 2) < self.duration:
                return False
        return True

    # __init__.py:385-406
    def add_member(self, sig) -> None:
        if sig.device in self._sigs:
            if self._sigs[sig.device].avg < sig.avg:
                self._sigs[sig.device] = sig
        else:
            self._sigs[sig.device] = sig


class OracleMatcher:
    """`emitted` collects what the reference puts on its queue, `_matched` is its open list."""

    def __init__(self, device: List[str], matching_timeout_s: float, matching_time_diff_s: float, matching_bandwidth_hz: float,
                 matching_duration_diff_ms: Optional[float] = None):
        self.devices = device
        self.matching_timeout = datetime.timedelta(seconds=matching_timeout_s)          # match.py:41-44
        self.matching_time_diff = datetime.timedelta(seconds=matching_time_diff_s)
        self.matching_bandwidth_hz = float(matching_bandwidth_hz)
        self.matching_duration_diff = datetime.timedelta(milliseconds=matching_duration_diff_ms) if matching_duration_diff_ms else None
        self._matched: List[OracleGroup] = []
        self.emitted: List[OracleGroup] = []

    def add(self, signal) -> None:                                                      # match.py:54-82
        now = signal.ts
        for msig in list(self._matched):
            if msig.ts < now - self.matching_timeout:
                self.emitted.append(msig)
                self._matched.remove(msig)
                continue
            if msig.has_member(signal, bandwidth=self.matching_bandwidth_hz, time_diff=self.matching_time_diff,
                               duration_diff=self.matching_duration_diff):
                msig.add_member(signal)
                return
        msig = OracleGroup(self.devices)
        msig.add_member(signal)
        self._matched.append(msig)


# ---------------------------------------------------------------------------------------------
# seeded signal sequences shared by the fixture generator, the tests and the benchmark
# ---------------------------------------------------------------------------------------------
class Sig:
    """The attributes of radiotracking.Signal the matcher reads."""
    __slots__ = ("device", "ts", "frequency", "duration", "avg", "idx")

    def __init__(self, device, ts, frequency, duration, avg, idx):
        self.device, self.ts, self.frequency, self.duration, self.avg, self.idx = device, ts, frequency, duration, avg, idx


T0 = datetime.datetime(2026, 3, 1, 12, 0, 0, tzinfo=datetime.timezone.utc)

CASES = {
    # name: (n_devices, n_signals, seed, timeout_s, time_diff_s, bandwidth_hz, duration_diff_ms, tags, rate_hz)
    "default_keys":   (4, 600, 11, 2.0, 0.0, 0.0, None, 6, 40.0),        # the reference's CLI defaults (__main__.py:68-71)
    "with_margins":   (4, 600, 12, 2.0, 0.002, 4000.0, 3.0, 6, 40.0),
    "tight_timeout":  (3, 600, 13, 0.05, 0.001, 2000.0, 0.0015, 5, 200.0),  # odd number of microseconds / 2, frequent time-outs
    "dense_8dev":     (8, 1500, 14, 0.02, 0.004, 6000.0, 5.0, 40, 2000.0),
    "boundaries":     (4, 400, 15, 1.0, 0.001, 1000.0, 2.0, 3, 50.0),       # values placed exactly on the inequalities
}


def make_signals(name: str) -> List[Sig]:
    import numpy as np

    n_dev, n, seed, timeout_s, time_diff_s, bw, dd_ms, n_tags, rate = CASES[name]
    rng = np.random.default_rng(seed)
    tag_f = 150.0e6 + rng.integers(-60, 60, n_tags) * 2000.0
    out: List[Sig] = []
    t = 0.0
    k = 0
    while len(out) < n:
        t += float(rng.exponential(1.0 / rate))
        tag = int(rng.integers(n_tags))
        dur_us = int(rng.integers(8000, 40000))
        if name == "boundaries":
            dur_us = 10000 + 1000 * int(rng.integers(0, 4))          # multiples of the duration margin's half
        ts_us = int(round(t * 1e6))
        if name == "boundaries":
            ts_us -= ts_us % 1000
        heard = [d for d in range(n_dev) if rng.random() < 0.7] or [int(rng.integers(n_dev))]
        rng.shuffle(heard)
        for d in heard:
            jit_us = int(rng.integers(-1500, 1500)) if name != "boundaries" else 1000 * int(rng.integers(-2, 3))
            f = float(tag_f[tag] + (rng.integers(-2, 3) * 1000.0 if name != "boundaries" else rng.integers(-1, 2) * 500.0))
            dj = int(rng.integers(-1200, 1200)) if name != "boundaries" else 1000 * int(rng.integers(-1, 2))
            if name == "tight_timeout":
                dj = int(rng.integers(-2, 3))                     # around the 1.5 us / 2 duration margin (rounds half to even)
            dev = str(d) if rng.random() > 0.02 else "99"        # now and then a device the station does not list
            out.append(Sig(dev, T0 + datetime.timedelta(microseconds=ts_us + jit_us), f,
                           datetime.timedelta(microseconds=max(1000, dur_us + dj)), float(rng.normal(-60, 6)), k))
            k += 1
            if rng.random() < 0.05:                               # a second, louder or quieter detection on the same device
                out.append(Sig(dev, T0 + datetime.timedelta(microseconds=ts_us + jit_us + 200), f,
                               datetime.timedelta(microseconds=max(1000, dur_us + dj)), float(rng.normal(-60, 6)), k))
                k += 1
    return out[:n]


def matcher_kwargs(name: str) -> dict:
    n_dev, _, _, timeout_s, time_diff_s, bw, dd_ms, _, _ = CASES[name]
    return dict(device=[str(d) for d in range(n_dev)], matching_timeout_s=timeout_s, matching_time_diff_s=time_diff_s,
                matching_bandwidth_hz=bw, matching_duration_diff_ms=dd_ms)


def groups_as_ids(groups) -> List[List[int]]:
    return [[s.idx for s in g._sigs.values()] for g in groups]

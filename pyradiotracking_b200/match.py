"""Cross-device signal matching on the batched output of the detection engine.

Drop-in for `radiotracking.match.SignalMatcher` (radiotracking/match.py:12-82): same constructor keys
(`device`, `matching_timeout_s`, `matching_time_diff_s`, `matching_bandwidth_hz`, `signal_queue`,
`matching_duration_diff_ms`), same `add(signal)` entry, same MatchingSignal objects on the same queue in the same
order.  The first-fit walk over the open groups runs natively (include/rt_matcher.h -> librtb200.so) on integer
microseconds, for one Signal or -- `add_batch` -- for everything a `BatchAnalyzer.collect` returned at once.
"""
import ctypes
import datetime
from typing import Dict, Iterable, List, Optional

import numpy as np

from . import engine as _engine
from .messages import matching_signal_type, message_types

MATCH_SIGNAL_DTYPE = np.dtype([
    ("ts_us", "<i8"), ("duration_us", "<i8"), ("frequency", "<f8"), ("avg", "<f8"),
    ("device", "<i4"), ("reserved", "<i4"), ("id", "<i8"),
])
assert MATCH_SIGNAL_DTYPE.itemsize == 48

_US = datetime.timedelta(microseconds=1)
_EPOCH_AWARE = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
_EPOCH_NAIVE = datetime.datetime(1970, 1, 1)

_bound = False


def _lib() -> ctypes.CDLL:
    global _bound
    lib = _engine.load_library()
    if not _bound:
        p64 = ctypes.POINTER(ctypes.c_int64)
        lib.rt_matcher_create.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
        lib.rt_matcher_destroy.argtypes = [ctypes.c_void_p]
        lib.rt_matcher_destroy.restype = None
        lib.rt_matcher_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.rt_matcher_pending.argtypes = [ctypes.c_void_p, p64, p64]
        lib.rt_matcher_drain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.rt_matcher_open.argtypes = [ctypes.c_void_p, p64, p64]
        lib.rt_matcher_read_open.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _bound = True
    return lib


def _us(ts: datetime.datetime) -> int:
    return (ts - (_EPOCH_NAIVE if ts.tzinfo is None else _EPOCH_AWARE)) // _US


class SignalMatcher:
    """Consumes Signals of several devices and publishes MatchingSignals (reference: match.py:12-82)."""

    def __init__(self, device: List[str], matching_timeout_s: float, matching_time_diff_s: float, matching_bandwidth_hz: float,
                 signal_queue, matching_duration_diff_ms: Optional[float] = None, **kwargs):
        self.devices = device
        # the reference's own conversions (match.py:41-44): timedelta rounds to whole microseconds
        self.matching_timeout = datetime.timedelta(seconds=matching_timeout_s)
        self.matching_time_diff = datetime.timedelta(seconds=matching_time_diff_s)
        self.matching_bandwidth_hz = float(matching_bandwidth_hz)
        self.matching_duration_diff = datetime.timedelta(milliseconds=matching_duration_diff_ms) if matching_duration_diff_ms else None
        self.signal_queue = signal_queue
        self.Signal, _ = message_types()
        self.MatchingSignal = matching_signal_type()
        self._dev_index: Dict[str, int] = {d: i for i, d in enumerate(device)}
        self._live: Dict[int, object] = {}          # id -> Signal of every member of an open group
        self._next_id = 0
        self._h = ctypes.c_void_p()
        dd = self.matching_duration_diff // _US if self.matching_duration_diff else -1
        _engine._check(_lib().rt_matcher_create(self.matching_timeout // _US, self.matching_time_diff // _US,
                                                self.matching_bandwidth_hz, dd, ctypes.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib().rt_matcher_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's entry points -------------------------------------------------------
    def add(self, signal) -> None:
        """One message (reference: match.py:54-82); anything but a Signal is ignored (match.py:62-63)."""
        self.add_batch([signal])

    def add_batch(self, signals: Iterable) -> int:
        """Signals in arrival order, e.g. the flattened result of `BatchAnalyzer.collect`.  Returns how many
        MatchingSignals were published."""
        sigs = [s for s in signals if isinstance(s, self.Signal) or _looks_like_signal(s)]
        if not sigs:
            return 0
        rec = np.zeros(len(sigs), dtype=MATCH_SIGNAL_DTYPE)
        for i, s in enumerate(sigs):
            dev = self._dev_index.get(s.device)
            if dev is None:                          # a device the station does not list still forms groups (dict key)
                dev = self._dev_index[s.device] = len(self._dev_index)
            sid = self._next_id
            self._next_id += 1
            self._live[sid] = s
            rec[i] = (_us(s.ts), s.duration // _US, s.frequency, s.avg, dev, 0, sid)
        _engine._check(_lib().rt_matcher_add(self._h, rec.ctypes.data_as(ctypes.c_void_p), len(sigs)))
        return self._publish()

    # -- results ----------------------------------------------------------------------------
    def _groups(self, count_fn, read_fn):
        ng, nm = ctypes.c_int64(), ctypes.c_int64()
        _engine._check(count_fn(self._h, ctypes.byref(ng), ctypes.byref(nm)))
        sizes = np.empty(ng.value, dtype=np.int64)
        ids = np.empty(nm.value, dtype=np.int64)
        if ng.value:
            _engine._check(read_fn(self._h, sizes.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p)))
        return sizes, ids

    def _build(self, ids, forget: bool):
        msig = self.MatchingSignal(self.devices)
        for sid in ids:
            s = self._live.pop(int(sid)) if forget else self._live[int(sid)]
            msig._sigs[s.device] = s
        return msig

    def _publish(self) -> int:
        lib = _lib()
        sizes, ids = self._groups(lib.rt_matcher_pending, lib.rt_matcher_drain)
        k = 0
        for n in sizes:
            self.signal_queue.put(self._build(ids[k:k + n], forget=True))      # SignalMatcher.consume (match.py:50-52)
            k += int(n)
        # signals replaced by a louder one of the same device (__init__.py:397-404) are not in any group any more
        if len(self._live) > 4096:
            _, open_ids = self._groups(lib.rt_matcher_open, lib.rt_matcher_read_open)
            keep = set(int(i) for i in open_ids)
            self._live = {i: s for i, s in self._live.items() if i in keep}
        return len(sizes)

    @property
    def _matched(self) -> list:
        """The open groups, like the reference's `_matched` list (read-only view)."""
        lib = _lib()
        sizes, ids = self._groups(lib.rt_matcher_open, lib.rt_matcher_read_open)
        out, k = [], 0
        for n in sizes:
            out.append(self._build(ids[k:k + n], forget=False))
            k += int(n)
        return out


def _looks_like_signal(s) -> bool:
    return all(hasattr(s, a) for a in ("device", "ts", "frequency", "duration", "avg")) and not hasattr(s, "_sigs")

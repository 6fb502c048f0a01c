"""Throughput of the cross-device matcher: native loop vs the reference's Python walk.

    python tools/bench_matcher.py [n_signals=200000] [n_devices=8]

Rows: the oracle restatement of radiotracking/match.py (pure Python, = the reference's cost model; the reference
itself when /root/reference is present), `SignalMatcher.add_batch` on Signal objects (includes packing them into
records in Python), and the bare C call `rt_matcher_add` on a prepared record array.  Host code only: no GPU needed.
"""
import ctypes
import datetime
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import matcher as M  # noqa: E402
from pyradiotracking_b200 import engine as E  # noqa: E402
from pyradiotracking_b200 import match as native  # noqa: E402
from pyradiotracking_b200 import messages  # noqa: E402


def make(n, n_dev, seed=5):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(2000.0 / n_dev, n)).astype(np.int64)          # ~500 transmissions/s heard by every device
    dev = rng.integers(0, n_dev, n)
    f = 150e6 + rng.integers(-40, 41, n) * 2000.0
    dur = rng.integers(8000, 40000, n)
    avg = rng.normal(-60, 5, n)
    return [M.Sig(str(int(dev[i])), M.T0 + datetime.timedelta(microseconds=int(t[i])), float(f[i]),
                  datetime.timedelta(microseconds=int(dur[i])), float(avg[i]), i) for i in range(n)]


class _Q:
    def __init__(self):
        self.n = 0

    def put(self, x):
        self.n += 1


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    n_dev = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    kw = dict(device=[str(d) for d in range(n_dev)], matching_timeout_s=2.0, matching_time_diff_s=0.002,
              matching_bandwidth_hz=4000.0, matching_duration_diff_ms=3.0)
    sigs = make(n, n_dev)
    rows = []

    n_py = min(n, int(os.environ.get("RT_BENCH_MATCHER_PY", "4000")))      # the Python walks manage ~100 signals/s on this workload
    o = M.OracleMatcher(**kw)
    t0 = time.perf_counter()
    for s in sigs[:n_py]:
        o.add(s)
    dt = time.perf_counter() - t0
    rows.append(dict(impl="oracle restatement of match.py (Python)", signals=n_py, seconds=dt, signals_per_s=n_py / dt, emitted=len(o.emitted)))

    try:
        from oracle import ref_harness
        if ref_harness.available():
            from oracle.make_matcher_golden import load_reference_matcher
            radiotracking, rm = load_reference_matcher()
            q = _Q()
            ref = rm.SignalMatcher(signal_queue=q, **kw)
            rs = [radiotracking.Signal(s.device, s.ts, s.frequency, s.duration, s.avg + 3, s.avg, 1.0, -95.0, 10.0) for s in sigs[:n_py]]
            import logging
            logging.disable(logging.CRITICAL)
            t0 = time.perf_counter()
            for s in rs:
                ref.add(s)
            dt = time.perf_counter() - t0
            rows.append(dict(impl="unmodified reference SignalMatcher", signals=n_py, seconds=dt, signals_per_s=n_py / dt, emitted=q.n))
    except Exception as exc:  # pragma: no cover
        rows.append(dict(impl="unmodified reference SignalMatcher", error=str(exc)))

    ms = [messages.Signal(s.device, s.ts, s.frequency, s.duration, s.avg + 3, s.avg, 1.0, -95.0, 10.0) for s in sigs]
    q = _Q()
    m = native.SignalMatcher(signal_queue=q, **kw)
    t0 = time.perf_counter()
    for i in range(0, n, 4096):
        m.add_batch(ms[i:i + 4096])
    dt = time.perf_counter() - t0
    rows.append(dict(impl="native, SignalMatcher.add_batch on Signal objects (4096 per call)", signals=n, seconds=dt, signals_per_s=n / dt, emitted=q.n))
    m.close()

    rec = np.zeros(n, dtype=native.MATCH_SIGNAL_DTYPE)
    rec["ts_us"] = [native._us(s.ts) for s in sigs]
    rec["duration_us"] = [s.duration // datetime.timedelta(microseconds=1) for s in sigs]
    rec["frequency"] = [s.frequency for s in sigs]
    rec["avg"] = [s.avg for s in sigs]
    rec["device"] = [int(s.device) for s in sigs]
    rec["id"] = np.arange(n)
    lib = native._lib()
    h = ctypes.c_void_p()
    E._check(lib.rt_matcher_create(2000000, 2000, 4000.0, 3000, ctypes.byref(h)))
    t0 = time.perf_counter()
    E._check(lib.rt_matcher_add(h, rec.ctypes.data_as(ctypes.c_void_p), n))
    dt = time.perf_counter() - t0
    ng, nm = ctypes.c_int64(), ctypes.c_int64()
    lib.rt_matcher_pending(h, ctypes.byref(ng), ctypes.byref(nm))
    rows.append(dict(impl="native, rt_matcher_add on a record array (one call)", signals=n, seconds=dt, signals_per_s=n / dt, emitted=ng.value))
    lib.rt_matcher_destroy(h)

    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()

"""Generate tests/golden/matcher_<case>.json by running the UNMODIFIED reference matcher
(radiotracking.match.SignalMatcher + radiotracking.MatchingSignal) on the seeded sequences of oracle/matcher.py.

Build container only (needs /root/reference).  `radiotracking.consume` imports paho/cbor2 at module top, so the
two missing third-party modules are stubbed the same way oracle/ref_harness.py stubs pytz/rtlsdr; no reference
code is modified or copied.   python -m oracle.make_matcher_golden
"""
import json
import os
import sys
import types

from oracle import matcher as M
from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_matcher():
    ref_harness._install_stubs()
    for name in ("paho", "paho.mqtt", "paho.mqtt.client", "cbor2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if ref_harness.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    import radiotracking  # type: ignore
    import radiotracking.match as rm  # type: ignore

    return radiotracking, rm


class _Q:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def run_reference(name: str):
    radiotracking, rm = load_reference_matcher()
    sigs = M.make_signals(name)
    q = _Q()
    kw = M.matcher_kwargs(name)
    matcher = rm.SignalMatcher(signal_queue=q, **kw)
    ref_sigs = []
    for s in sigs:
        r = radiotracking.Signal(s.device, s.ts, s.frequency, s.duration, s.avg + 3.0, s.avg, 1.0, -95.0, 10.0)
        r.idx = s.idx
        ref_sigs.append(r)
        matcher.add(r)
    return sigs, M.groups_as_ids(q.items), M.groups_as_ids(matcher._matched), q.items


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in M.CASES:
        sigs, emitted, still_open, items = run_reference(name)
        doc = dict(case=name, kwargs=M.matcher_kwargs(name), n_signals=len(sigs),
                   emitted=emitted, open=still_open,
                   # derived views of the first emitted groups, as the reference computes them
                   views=[dict(ts=g.ts.isoformat(), frequency=g.frequency, duration_us=g.duration // M.datetime.timedelta(microseconds=1),
                               avgs=g._avgs) for g in items[:20]])
        path = os.path.join(ROOT, "tests", "golden", f"matcher_{name}.json")
        with open(path, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, len(sigs), "signals ->", len(emitted), "emitted,", len(still_open), "open;",
              sum(len(g) > 1 for g in emitted), "multi-device groups")


if __name__ == "__main__":
    main()

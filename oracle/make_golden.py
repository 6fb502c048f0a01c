"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden [case ...]

For every case of `oracle/cases.py` the reference `SignalAnalyzer.process_samples`
(/root/reference/radiotracking/analyze.py:192-268) is fed the synthetic capture block
by block through `oracle/ref_harness.py`; stored per block:

* every `Signal` returned by `extract_signals` (pre shadow filter) with all nine
  fields (ts/duration as integer microseconds), and which of them survived
  `filter_shadow_signals` and reached the queue;
* a digest of the spectrogram scipy produced: all row means, and a grid of cells.
"""
import datetime
import json
import os
import sys

import numpy as np

from oracle import ref_harness
from oracle.cases import CASES, BY_NAME, sha256

T0 = datetime.datetime(2026, 3, 1, 6, 30, 0)          # naive, like datetime.now() in the reference
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EPOCH = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)


def us(td: datetime.timedelta) -> int:
    return (td.days * 86400 + td.seconds) * 1_000_000 + td.microseconds


def digest_index(nperseg: int, T: int):
    rows = np.arange(0, nperseg, max(1, nperseg // 256))
    cols = np.unique(np.linspace(0, T - 1, 12).astype(np.int64))
    return rows, cols


def run_case(case) -> dict:
    import scipy

    cap = case.capture()
    kw = case.analyzer_kwargs()
    ref = ref_harness.ReferenceRunner(T0, **kw)
    out = {}
    for b in range(case.n_blocks):
        queued = ref.feed(cap[b])
        pre = ref.pre_shadow[-1]
        queued_ids = {id(s) for s in queued}
        S = ref.spectrogram_last
        rows, cols = digest_index(S.shape[0], S.shape[1])
        out[f"b{b}_ts_us"] = np.array([us(s.ts - EPOCH) for s in pre], dtype=np.int64)
        out[f"b{b}_dur_us"] = np.array([us(s.duration) for s in pre], dtype=np.int64)
        out[f"b{b}_freq"] = np.array([s.frequency for s in pre], dtype=np.float64)
        out[f"b{b}_stats"] = np.array([[s.max, s.avg, s.std, s.noise, s.snr] for s in pre], dtype=np.float64).reshape(-1, 5)
        out[f"b{b}_kept"] = np.array([id(s) in queued_ids for s in pre], dtype=bool)
        assert int(out[f"b{b}_kept"].sum()) == len(queued)
        out[f"b{b}_rowmean"] = np.array([np.mean(r) for r in S])
        out[f"b{b}_cells"] = np.ascontiguousarray(S[np.ix_(rows, cols)])
    meta = dict(
        name=case.name, n_blocks=case.n_blocks, stream=case.stream, t0=T0.isoformat(),
        analyzer={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
        input_sha256=sha256(cap), scipy=scipy.__version__, numpy=np.__version__,
        reference="Nature40/pyradiotracking radiotracking/analyze.py (unmodified, via oracle/ref_harness.py)",
    )
    out["meta"] = np.array(json.dumps(meta))
    return out


def main(argv):
    names = argv or [c.name for c in CASES]
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for n in names:
        case = BY_NAME[n]
        res = run_case(case)
        path = os.path.join(GOLDEN_DIR, f"{n}.npz")
        np.savez_compressed(path, **res)
        tot = sum(len(res[f"b{b}_kept"]) for b in range(case.n_blocks))
        kept = sum(int(res[f"b{b}_kept"].sum()) for b in range(case.n_blocks))
        print(f"{n}: {tot} signals pre-shadow, {kept} queued, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])

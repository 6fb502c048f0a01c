"""Count-based parity check of the CUDA path against the oracle (no asserts: `bench.py` reports the numbers in its JSON line,
BASELINE.md section 4 "parity gate reported with every number").  TEST INFRASTRUCTURE ONLY, like the rest of `oracle/`.

Classification follows SURVEY.md section 8(d): integer run keys `(fi, start, end)` must match exactly; a differing run counts
as "near threshold" when one of its cells or neighbours lies within 2e-4 of max(thr, snr * row mean); `ts` / `duration` /
`frequency` must be bit-identical; the dB fields within 5e-4 dB; power cells >= 1e-2 * threshold within 1e-4 relative (cells
more than 50 dB under the strongest cell of their own segment are the fp32 "deep tail" and are counted separately).
"""
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import restatement as R

POWER_RTOL, DB_ATOL, NEAR_THRESHOLD, DEEP_DB = 1e-4, 5e-4, 2e-4, 50.0


def _near(P: R.Params, S: np.ndarray, last, key) -> bool:
    fi, start, end = key
    row = S[fi]
    thr = max(P.signal_threshold, P.snr_threshold * np.mean(row))
    lo, hi = start - 1, min(len(row), end + 1)
    cells = [row[max(lo, 0):hi]]
    if lo < 0 and last is not None:
        cells.append(last[fi][lo:])
    c = np.concatenate(cells)
    return bool(np.min(np.abs(c / thr - 1.0)) <= NEAR_THRESHOLD)


def new_totals() -> Dict[str, float]:
    return dict(oracle_signals=0, gpu_signals=0, key_mismatches=0, near_threshold_mismatches=0, exact_field_mismatches=0,
                max_db_err=0.0, post_shadow_lists_identical=True, cells_checked=0, cells_max_rel=0.0, cells_over_1e4=0,
                deep_cells=0, deep_cells_max_rel=0.0, row_mean_max_rel=0.0)


def add_block(tot: Dict[str, float], P: R.Params, S: np.ndarray, last: Optional[np.ndarray], found: Sequence[R.Detection],
              kept: Sequence[R.Detection], sigs: List, keys: List, filtered: List, S32_T: Optional[np.ndarray] = None,
              rowmean32: Optional[np.ndarray] = None) -> None:
    """One block of one stream: the oracle's `found` / `kept` against the engine's `sigs` (+ `keys`) / `filtered`."""
    okeys = [d.key() for d in found]
    oset, gset = set(okeys), set(keys)
    odd = sorted(oset ^ gset)
    tot["oracle_signals"] += len(okeys)
    tot["gpu_signals"] += len(keys)
    tot["key_mismatches"] += len(odd)
    tot["near_threshold_mismatches"] += sum(1 for k in odd if _near(P, S, last, k))
    gidx = {k: i for i, k in enumerate(keys)}
    for d in found:
        i = gidx.get(d.key())
        if i is None:
            continue
        g = sigs[i]
        if not (g.ts == d.ts and g.duration == d.duration and g.frequency == d.frequency):
            tot["exact_field_mismatches"] += 1
        err = max(abs(getattr(g, n) - getattr(d, n)) for n in ("max", "avg", "std", "noise", "snr"))
        tot["max_db_err"] = max(tot["max_db_err"], float(err))
    if not odd and [(s.ts, s.frequency) for s in filtered] != [(d.ts, d.frequency) for d in kept]:
        tot["post_shadow_lists_identical"] = False
    if S32_T is not None:
        got = S32_T.T.astype(np.float64)
        rel = np.abs(got - S) / S
        big = S >= 1e-2 * P.signal_threshold
        deep = S < S.max(axis=0, keepdims=True) * 10 ** (-DEEP_DB / 10)
        main = big & ~deep
        tot["cells_checked"] += int(main.sum())
        if main.any():
            tot["cells_max_rel"] = max(tot["cells_max_rel"], float(rel[main].max()))
        tot["cells_over_1e4"] += int((rel[main] > POWER_RTOL).sum())
        tot["deep_cells"] += int((big & deep).sum())
        if (big & deep).any():
            tot["deep_cells_max_rel"] = max(tot["deep_cells_max_rel"], float(rel[big & deep].max()))
    if rowmean32 is not None:
        m = S.mean(axis=1)
        tot["row_mean_max_rel"] = max(tot["row_mean_max_rel"], float(np.max(np.abs(rowmean32.astype(np.float64) - m) / m)))


def verdict(tot: Dict[str, float]) -> bool:
    """The gate: every mismatch explained by a near-threshold cell, exact fields exact, dB and cell tolerances held."""
    return bool(tot["key_mismatches"] == tot["near_threshold_mismatches"] and tot["exact_field_mismatches"] == 0
                and tot["max_db_err"] <= DB_ATOL and tot["post_shadow_lists_identical"] and tot["cells_over_1e4"] == 0
                and tot["row_mean_max_rel"] <= 2e-5 and tot["gpu_signals"] > 0)

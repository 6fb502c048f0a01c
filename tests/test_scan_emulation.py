"""The device scan schedule (chained work items, member loop, block-wise walks, early exits) restated on the CPU
(tests/scan_emulation.py) must pick exactly the runs the reference's sequential probe loop picks (analyze.py:354-433)."""
import numpy as np
import pytest

from tests.scan_emulation import device_scan, reference_scan


def _pattern(rng, n, T, stride, density):
    """Random above-matrix: runs whose lengths cluster around the probe stride and the duration limits, plus speckle."""
    ab = rng.random((n, T)) < density * 0.2            # 1-3 cell noise runs
    for fi in range(n):
        t = int(rng.integers(0, stride))
        while t < T:
            L = int(rng.choice([1, 2, 3, 4, stride - 1, stride, stride + 1, 2 * stride, 5 * stride + 1, 6 * stride, 12 * stride]))
            if rng.random() < density:
                ab[fi, t:t + L] = True
            t += L + int(rng.integers(1, 3 * stride))
    return ab


@pytest.mark.parametrize("stride,min_cols,max_cols,T", [(9, 8, 49, 400), (74, 73, 377, 1500), (4, 3, 12, 97), (1, 0, 2, 40), (39, 38, 198, 900)])
@pytest.mark.parametrize("ppt,widen", [(32, False), (8, True), (8, False)])
def test_device_schedule_equals_reference_loop(stride, min_cols, max_cols, T, ppt, widen):
    rng = np.random.default_rng(1000 * stride + ppt)
    n_total = 0
    for trial in range(6):
        density = [0.05, 0.3, 0.6, 0.9, 0.3, 0.6][trial]
        prev = _pattern(rng, 12, T, stride, density)
        cur = _pattern(rng, 12, T, stride, density)
        if trial == 3:
            cur[0, :] = True                    # a carrier: one run as long as the block
            cur[1, : T - 1] = True              # ... ending one cell before the block end
            cur[2, 1:] = True
            prev[0, :] = True
        for ab_prev in (None, prev):
            want = reference_scan(cur, ab_prev, stride, min_cols, max_cols)
            got = device_scan(cur, ab_prev, stride, min_cols, max_cols, ppt=ppt, widen=widen)
            assert got == want
            n_total += len(want)
    assert n_total > 0


def test_chain_is_bounded_and_split_members_are_skipped():
    from tests.scan_emulation import PROBE_CHAIN, probe_items

    T, stride = 2000, 10
    ab = np.ones(T, dtype=bool)
    ab[-1] = False
    items = probe_items(ab, T, stride, 9, 32)
    assert max(m for _, m in items) <= PROBE_CHAIN
    assert sum(m for _, m in items) == (T + stride - 1) // stride          # every probe column is a member of exactly one item

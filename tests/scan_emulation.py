"""CPU emulation of the device scan algorithm (pyradiotracking_b200/csrc/rt_engine.cu: probe kernels + extract2_kernel).

Test infrastructure: it restates, in plain Python over a boolean "above" matrix, WHAT the CUDA kernels decide --
probe hits, the quick +-PROBE_QUICK resolution, chaining of consecutive surviving hits into one work item, the member
loop with `skip_to`, the block-wise walks down and up from the probe column with their early exits, the carry walk, the coarse
duration gate --
so that the schedule can be checked against the reference's sequential loop (oracle.restatement.extract_sequential,
analyze.py:354-433) on random patterns without a GPU.  The kernels' arithmetic (predicate, statistics) is not emulated.
"""
from typing import List, Optional, Set, Tuple

import numpy as np

PROBE_QUICK = 3
PROBE_CHAIN = 8


def probe_items(ab: np.ndarray, T: int, stride: int, min_cols: int, ppt: int) -> List[Tuple[int, int]]:
    """Work items (head probe column, members) of one bin: probe_kernel phases 2 and 3 / probe_lean_kernel."""
    n_probes = (T + stride - 1) // stride
    live = []
    for k in range(n_probes):
        ti = k * stride
        if not ab[ti]:
            live.append(False)
            continue
        lo = hi = -1
        drop = False
        for d in range(1, PROBE_QUICK + 1):
            t = ti - d
            if t < 0:
                break
            if not ab[t]:
                lo = t
                break
        for d in range(1, PROBE_QUICK + 1):
            t = ti + d
            if t >= T:
                drop = True
                break
            if not ab[t]:
                hi = t
                break
        live.append(not drop and not (lo >= 0 and hi >= 0 and hi - lo < min_cols))
    items = []
    k = 0
    while k < n_probes:
        if not live[k]:
            k += 1
            continue
        ln = 1
        # a chain ends at a multiple of PROBE_CHAIN within the thread's group of ppt probes (ppt is a multiple of it)
        while k + ln < n_probes and ((k + ln) % ppt) % PROBE_CHAIN != 0 and live[k + ln]:
            ln += 1
        items.append((k * stride, ln))
        k += ln
    return items


def extract_item(ab: np.ndarray, ab_prev: Optional[np.ndarray], T: int, stride: int, min_cols: int, max_cols: int,
                 ti0: int, members: int, widen: bool = False) -> List[Tuple[int, int]]:
    """(start, end) records of one work item: extract2_kernel's member loop -- aligned 32-column blocks, the probe's own block
    first, then one block down and `wf` blocks up per round trip while that side of the run is open, the span cap evaluated on
    what is known so far, the carry walk into the previous block."""
    out = []
    skip_to = 0
    span_cap = max_cols + 2
    for mem in range(members):
        ti = ti0 + mem * stride
        if ti < skip_to:
            continue
        lo_lim = max(ti - stride, 0)
        bh = ti >> 5

        def not_above(b, lo, hi):                        # not-above columns of block b inside [lo, hi)
            return [t for t in range(max(32 * b, lo), min(32 * b + 32, hi)) if not ab[t]]

        nb = end = -1
        own = not_above(bh, lo_lim, T)
        below, beyond = [t for t in own if t < ti], [t for t in own if t > ti]
        if below:
            nb = max(below)
        if beyond:
            end = min(beyond)
        kb, kf = bh - 1, bh + 1
        bopen = nb < 0 and 32 * bh > lo_lim
        fopen = end < 0 and 32 * kf < T
        too_long = False
        rounds = 0
        while bopen or fopen:
            if fopen and 32 * kf - (nb if nb >= 0 else 32 * (kb + 1)) > span_cap:
                too_long = True
                skip_to = 32 * kf
                break
            wf = 1 if (not widen or rounds < 3) else (2 if rounds < 5 else 4)
            rounds += 1
            if bopen:
                c = not_above(kb, lo_lim, T)
                if c:
                    nb = max(c)
                bopen = nb < 0 and 32 * kb > lo_lim
                kb -= 1
            if fopen:
                for w in range(wf):
                    if end < 0:
                        c = not_above(kf + w, 0, T)
                        if c:
                            end = min(c)
                kf += wf
                fopen = end < 0 and 32 * kf < T
        if nb >= 0:
            start = nb
        elif ti - stride >= 0:
            continue
        elif ab_prev is None:
            start = 0
        elif too_long:
            continue
        else:
            jmax = T - 2
            jcap = min(jmax, max_cols + 2)
            jf = 0
            for jj in range(1, jcap + 1):
                if not ab_prev[T - jj]:
                    jf = jj
                    break
            if jf > 0:
                start = -jf
            elif jcap == jmax:
                start = -(T - 1)
            else:
                continue
        if too_long or end < 0:
            if not too_long:
                skip_to = T
            continue
        skip_to = end
        cols = end - start + (1 if start < 0 else 0)
        if cols < min_cols or cols > max_cols:
            continue
        out.append((start, end))
    return out


def device_scan(ab: np.ndarray, ab_prev: Optional[np.ndarray], stride: int, min_cols: int, max_cols: int,
                ppt: int = 32, widen: bool = False) -> Set[Tuple[int, int, int]]:
    """All (bin, start, end) the device emits for a block; `ab` is [bins][T] booleans (predicate already applied)."""
    out = set()
    n, T = ab.shape
    for fi in range(n):
        prev = None if ab_prev is None else ab_prev[fi]
        for ti0, members in probe_items(ab[fi], T, stride, min_cols, ppt):
            for start, end in extract_item(ab[fi], prev, T, stride, min_cols, max_cols, ti0, members, widen):
                assert (fi, start, end) not in out, "a run was emitted twice"
                out.add((fi, start, end))
    return out


def reference_scan(ab: np.ndarray, ab_prev: Optional[np.ndarray], stride: int, min_cols: int, max_cols: int) -> Set[Tuple[int, int, int]]:
    """The reference's sequential probe loop (analyze.py:357-417) on the same booleans, with the coarse column gate the
    device applies (the exact float64 duration test happens on the host for both)."""
    out = set()
    n, T = ab.shape
    reach = 0 if ab_prev is None else -T + 1
    for fi in range(n):
        skip = 0
        for ti in range(0, T, stride):
            if ti < skip or not ab[fi, ti]:
                continue
            start = ti
            while start > reach and (ab_prev[fi, start] if start < 0 else ab[fi, start]):
                start -= 1
            end = ti
            while end < T and ab[fi, end]:
                end += 1
            if end == T:
                continue
            skip = end
            cols = end - start + (1 if start < 0 else 0)
            if min_cols <= cols <= max_cols:
                out.add((fi, start, end))
    return out

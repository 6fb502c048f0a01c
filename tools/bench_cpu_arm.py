#!/usr/bin/env python
"""Is the CPU arm honest?  Times the oracle port (oracle/restatement.py, what `bench.py --impl reference` and
`cpu_baseline` run on the GPU box) against the UNMODIFIED reference (`oracle/ref_harness.ReferenceRunner.feed`, i.e.
`SignalAnalyzer.process_samples` incl. the uint8 -> complex128 conversion pyrtlsdr does) on the same bytes, one core.

Build container only (`/root/reference` does not travel).  The port must not be slower than what it stands in for:

    python tools/bench_cpu_arm.py > profiles/r02_cpu_arm_port_vs_reference.txt
"""
import datetime
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import scipy  # noqa: E402

from oracle import ref_harness as H  # noqa: E402
from oracle import restatement as R  # noqa: E402
from pyradiotracking_b200 import synth  # noqa: E402


def main():
    if not H.available():
        raise SystemExit("the reference is not present: run this in the build container")
    reps, n_blocks = 5, 4
    t0 = datetime.datetime(2026, 1, 1)
    print(f"# python {platform.python_version()} numpy {np.__version__} scipy {scipy.__version__}, 1 core, {n_blocks} blocks per run, best of {reps} (runs interleaved)")
    print("# workload                      port Msamples/s   unmodified reference Msamples/s   port/reference   signals (both)")
    for w in (synth.C2, synth.C5B, synth.C1, synth.C5, synth.C5L):
        cap = synth.make_stream(w, 0, n_blocks)
        P = R.Params.make(sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=w.nperseg)
        best_p = best_r = 1e9
        n_p = n_r = 0
        for _ in range(reps):
            ora = R.OracleAnalyzer(P)
            t = time.perf_counter()
            n_p = sum(len(ora.process_block(cap[b], t0)[4]) for b in range(n_blocks))
            best_p = min(best_p, time.perf_counter() - t)
            rr = H.ReferenceRunner(t0, sample_rate=w.sample_rate, sdr_callback_length=w.block_samples)
            t = time.perf_counter()
            n_r = sum(len(rr.feed(cap[b])) for b in range(n_blocks))
            best_r = min(best_r, time.perf_counter() - t)
        assert n_p == n_r, (w.name, n_p, n_r)
        ms = n_blocks * w.block_samples / 1e6
        print(f"{w.name:30s} {ms / best_p:12.2f} {ms / best_r:28.2f} {best_r / best_p:22.2f} {n_p:14d}")


if __name__ == "__main__":
    main()

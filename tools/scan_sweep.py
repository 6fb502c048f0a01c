#!/usr/bin/env python
"""configs[1] step time under the engine's schedule knobs (rt_config.scan_schedule / launch_streams / chunk_segs).

The kernel variants that round 1 swept through environment variables live in tools/ now (spectro256_lab.cuh); what is left
to sweep in the product is what rt_config exposes.

Run on the GPU box:  python tools/scan_sweep.py [--steps 20]   (one JSON line per variant)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SERIAL, OVERLAP, LEAN = 1, 2, 3
VARIANTS = [
    ("default: lean scan, two launch streams, chunks of 192 segments", {}),
    ("one launch stream", {"launch_streams": 1}),
    ("full-size scan kernels on the scan stream", {"scan_schedule": OVERLAP}),
    ("serial (everything on the launch stream)", {"scan_schedule": SERIAL}),
    ("chunks of 128 segments", {"chunk_segs": 128}),
    ("chunks of 256 segments", {"chunk_segs": 256}),
    ("per-kernel events on every launch", {"kernel_timing": 1}),
    ("no per-kernel events", {"kernel_timing": 0}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    args = ap.parse_args()
    import contextlib
    import io

    import torch

    from pyradiotracking_b200 import synth
    from pyradiotracking_b200.analyze import BatchAnalyzer
    from tools.bench_configs import run

    for name, kw in VARIANTS:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            run(name, synth.C2, 64, args.steps, args.warmup, torch, synth, BatchAnalyzer, **kw)
        print(buf.getvalue().strip().splitlines()[-1], flush=True)


if __name__ == "__main__":
    main()

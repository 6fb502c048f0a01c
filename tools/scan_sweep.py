#!/usr/bin/env python
"""configs[1] step time under the engine's schedule / scan knobs (environment variables read by rt_engine_create).

Run on the GPU box:  python tools/scan_sweep.py [--steps 20]   (one JSON line per variant)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VARIANTS = [
    ("default (pinned addresses + running TMA pointer)", {}),
    ("serial, default", {"RT_SCAN_OVERLAP": "0"}),
    ("unpinned v7", {"RT_V7_MAXR": "0"}),
    ("serial, unpinned v7", {"RT_SCAN_OVERLAP": "0", "RT_V7_MAXR": "0"}),
    ("default (again)", {}),
    ("serial, default (again)", {"RT_SCAN_OVERLAP": "0"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--extra", action="append", default=[], help="NAME=VALUE applied to every variant")
    args = ap.parse_args()
    import contextlib
    import io

    import torch

    from pyradiotracking_b200 import synth
    from pyradiotracking_b200.analyze import BatchAnalyzer
    from tools.bench_configs import run

    keys = {"RT_PROBE_PLANE", "RT_SCAN_OVERLAP", "RT_SCAN_LEAN", "RT_V7_MAXR", "RT_LEAN_NO_CARVEOUT", "RT_CHUNK_SEGS", "RT_SCAN_EXPERIMENT_L2", "RT_S_LAYOUT", "RT_LEAN_EX", "RT_BENCH_NO_KERNEL_TIMING", "RT_TIMING_PERIOD", "RT_LAUNCH_STREAMS"} | {kv.split("=")[0] for kv in args.extra}
    for name, env in VARIANTS:
        for k in keys:
            os.environ.pop(k, None)
        for kv in args.extra:
            k, v = kv.split("=", 1)
            os.environ[k] = v
        os.environ.update(env)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            run(name, synth.C2, 64, args.steps, args.warmup, torch, synth, BatchAnalyzer)
        row = json.loads(buf.getvalue().strip().splitlines()[-1])
        row["env"] = env
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

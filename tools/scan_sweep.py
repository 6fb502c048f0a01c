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
    ("default: v7n, lean scan 8 CTAs/SM <1,0>, two launch streams", {}),
    ("one launch stream", {"RT_LAUNCH_STREAMS": "1"}),
    ("lean extraction windows <2,2>", {"RT_LEAN_EX": "22"}),
    ("lean scan 1 CTA/SM", {"RT_SCAN_LEAN": "1"}),
    ("full-size scan kernels", {"RT_SCAN_LEAN": "0"}),
    ("unpinned v7", {"RT_V7_MAXR": "0"}),
    ("v7 112 registers", {"RT_V7_MAXR": "112"}),
    ("v7n + ALU byte sums + packed row sums", {"RT_V7_MAXR": "-4"}),
    ("probe plane", {"RT_PROBE_PLANE": "1"}),
    ("S time-blocked 8", {"RT_S_LAYOUT": "8"}),
    ("S time-blocked 32", {"RT_S_LAYOUT": "32"}),
    ("chunk 256", {"RT_CHUNK_SEGS": "256"}),
    ("scan reads an L2-resident S (timing experiment, wrong results)", {"RT_SCAN_EXPERIMENT_L2": "1"}),
    ("per-kernel events on every launch", {"RT_TIMING_PERIOD": "1"}),
    ("serial", {"RT_SCAN_OVERLAP": "0"}),
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

#!/usr/bin/env python
"""What each scan kernel costs the configs[1] step: the lab build of the engine (tools/build_lab_lib.sh, -DRT_LAB) leaves
kernels out on request (wrong results, timing only).  Run on the GPU box:  python tools/scan_skip.py [--steps 40]"""
import argparse
import contextlib
import ctypes
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VARIANTS = [("all kernels", 0), ("no row-mean kernel (stale means)", 1), ("no extraction", 4), ("no probe, no extraction", 6), ("spectrogram only", 7), ("no probe (extraction finds an empty list)", 2)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--lib", default="librtb200_lab.so", help="lab build under tools/ (RT_LAB_DEFS=-DRT_PERM_TG=4 sh tools/build_lab_lib.sh tg4 -> librtb200_lab_tg4.so)")
    ap.add_argument("--extract-mode", type=int, nargs="*", default=[0], help="bit 0: no statistics, bit 1: walks read an L2-resident S, bit 2: first block only")
    ap.add_argument("--extract-per-sm", type=int, nargs="*", default=[0])
    ap.add_argument("--lean-per-sm", type=int, nargs="*", default=[0], help="lean scan CTAs per SM to sweep (0 = the engine's default)")
    ap.add_argument("--nowait", type=int, default=0, help="1: launch i does not wait for the scan of launch i - 2 (racy, timing only)")
    args = ap.parse_args()
    import torch

    from pyradiotracking_b200 import build, engine, synth
    from pyradiotracking_b200.analyze import BatchAnalyzer
    from tools.bench_configs import run

    build.LIB = os.path.join(ROOT, "tools", args.lib)
    lib = engine.load_library()
    skip = ctypes.c_int.in_dll(lib, "rt_lab_skip")
    lean = ctypes.c_int.in_dll(lib, "rt_lab_lean_per_sm")
    exps = ctypes.c_int.in_dll(lib, "rt_lab_extract_per_sm")
    xmode = ctypes.c_int.in_dll(lib, "rt_lab_extract_mode")
    ctypes.c_int.in_dll(lib, "rt_lab_nowait").value = args.nowait
    for per_sm, ex_sm, xm in [(p, x, m) for p in args.lean_per_sm for x in args.extract_per_sm for m in args.extract_mode]:
      lean.value = per_sm
      xmode.value = xm
      exps.value = ex_sm
      for name, mask in (VARIANTS if (per_sm, ex_sm, xm) == (args.lean_per_sm[0], args.extract_per_sm[0], args.extract_mode[0]) and len(args.extract_mode) == 1 else VARIANTS[:1]):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            # the e2e part of run() needs records: skip only inside the device-timed region is not possible, so e2e numbers of masked runs are meaningless
            skip.value = mask
            run(name, getattr(synth, args.workload), args.streams, args.steps, args.warmup, torch, synth, BatchAnalyzer, kernel_timing=0)
        d = json.loads(buf.getvalue().strip().splitlines()[-1])
        print(json.dumps({"variant": name, "skip_mask": mask, "lean_ctas_per_sm": per_sm, "extract_ctas_per_sm": ex_sm, "extract_mode": xm, "ms_per_step": d["ms_per_step"]}), flush=True)


if __name__ == "__main__":
    main()

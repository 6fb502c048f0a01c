#!/usr/bin/env python
"""bench.py on the lab build of the engine (tools/build_lab_lib.sh) with its timing-only switches set -- WRONG RESULTS, timing only.
  python tools/bench_lab.py --extract-mode 16 -- --workload c4 --no-parity --steps 40"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    argv = sys.argv[1:]
    cut = argv.index("--") if "--" in argv else len(argv)
    mine, rest = argv[:cut], argv[cut + 1:]
    knobs = {"--extract-mode": "rt_lab_extract_mode", "--skip": "rt_lab_skip", "--nowait": "rt_lab_nowait", "--lean-per-sm": "rt_lab_lean_per_sm"}
    from pyradiotracking_b200 import build, engine

    build.LIB = os.path.join(ROOT, "tools", "librtb200_lab.so")
    lib = engine.load_library()
    for i in range(0, len(mine), 2):
        ctypes.c_int.in_dll(lib, knobs[mine[i]]).value = int(mine[i + 1])
    import bench

    sys.argv = ["bench.py"] + rest
    bench.main()


if __name__ == "__main__":
    main()

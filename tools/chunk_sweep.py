import contextlib, io, json, os, sys
sys.path.insert(0, "/root/repo")
import torch
from pyradiotracking_b200 import synth
from pyradiotracking_b200.analyze import BatchAnalyzer
from tools.bench_configs import run
for rep in range(3):
    for ch in (128, 144, 160, 176, 192):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            run(f"chunk {ch}", synth.C2, 64, 40, 4, torch, synth, BatchAnalyzer, kernel_timing=0, chunk_segs=ch)
        d = json.loads(buf.getvalue().strip().splitlines()[-1])
        print(ch, d["ms_per_step"], flush=True)

#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json configuration through the C ABI (one GPU).

bench.py measures configs[1] (the headline).  This tool times the other configurations the same way
(K launches on HBM-resident input, CUDA events, per-kernel times from the engine) so that DESIGN.md can
quote a measured row for each.  Run on the GPU box:  python tools/bench_configs.py [--steps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def run(name, w, n_streams, steps, warmup, torch, synth, BatchAnalyzer, fft_impl=0, kernel_timing=4, **engine_kw):
    distinct = [synth.make_stream(w, i, 2) for i in range(min(4, n_streams))]
    host = np.empty((2, n_streams, w.block_bytes), dtype=np.uint8)
    for s in range(n_streams):
        host[:, s, :] = distinct[s % len(distinct)]
    dev = torch.from_numpy(host).cuda()
    ba = BatchAnalyzer(devices=[str(i) for i in range(n_streams)], calibration_db=[0.0] * n_streams,
                       sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=w.nperseg, fft_window="hamming",
                       signal_min_duration_ms=w.signal_min_duration_ms, signal_max_duration_ms=w.signal_max_duration_ms,
                       signal_threshold_dbw=w.signal_threshold_dbw, snr_threshold_db=w.snr_threshold_db,
                       sdr_callback_length=w.block_samples, cuda_device=0, fft_impl=fft_impl, **engine_kw)
    eng = ba.engine
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    for i in range(warmup):
        eng.launch(dev[i % 2])
    n_rec = len(eng.fetch())
    timing = kernel_timing > 0      # per-kernel events cost a few us of launch gaps per step: every `kernel_timing`-th launch only
    eng.enable_timing(kernel_timing)
    eng.timing(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        eng.launch(dev[(warmup + i) % 2])
    eng.join()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tim = eng.timing(reset=True)
    eng.enable_timing(0)
    if not timing:
        tim = {"spectrogram_ms": 0.0, "probe_ms": 0.0, "extract_ms": 0.0, "launches": 1}
    n_rec = len(eng.fetch())
    work, _ = eng.last_counts()
    samples = n_streams * w.block_samples
    # end to end through BatchAnalyzer.submit / collect on pinned host buffers (two launches in flight), like bench.py's e2e
    import datetime
    import time
    pin = torch.empty(host.shape, dtype=torch.uint8, pin_memory=True)
    pin.numpy()[...] = host
    hp = pin.numpy()
    ts = [datetime.datetime(2026, 1, 1)] * n_streams
    for i in range(2):
        ba.process_blocks(hp[i % 2], ts)
    for k in ba.timings:
        ba.timings[k] = 0
    n_e2e = max(4, min(steps, 12))
    kept = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ba.submit(hp[0])
    for i in range(n_e2e):
        if i + 1 < n_e2e:
            ba.submit(hp[(i + 1) % 2])
        kept += sum(len(r[0]) for r in ba.collect(ts))
    e2e_s = (time.perf_counter() - t0) / n_e2e
    out = {"config": name, "streams": n_streams, "nperseg": w.nperseg, "block_samples": w.block_samples,
           "ms_per_step": round(ms, 4), "msamples_per_s": round(samples / (ms * 1e-3) / 1e6),
           "algorithmic_gb_per_s": round(2 * samples / (ms * 1e-3) / 1e9, 1),
           "spectrogram_ms": round(tim["spectrogram_ms"] / tim["launches"], 4),
           "probe_ms": round(tim["probe_ms"] / tim["launches"], 4), "extract_ms": round(tim["extract_ms"] / tim["launches"], 4),
           "records_per_step": n_rec, "work_items": work, "fft_impl": fft_impl, "engine": engine_kw,
           "e2e_msamples_per_s": round(samples / e2e_s / 1e6), "e2e_ms_per_step": round(1e3 * e2e_s, 3), "e2e_signals_kept_per_step": round(kept / n_e2e, 1),
           "e2e_host_ms_per_step": {k: round(1e3 * v / n_e2e, 3) for k, v in ba.timings.items() if k.endswith("_s")}}
    ba.close()
    del dev
    torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="", help="run only the configurations whose name contains this")
    args = ap.parse_args()
    import torch

    from pyradiotracking_b200 import engine as E
    from pyradiotracking_b200 import synth
    from pyradiotracking_b200.analyze import BatchAnalyzer

    a = (args.steps, args.warmup, torch, synth, BatchAnalyzer)
    global run
    _run = run

    def run(name, *rest, **kw):
        if args.only in name:
            _run(name, *rest, **kw)

    run("configs[0] single 300 kS/s stream", synth.C1, 1, *a)
    run("one 2.4 MS/s stream (a single live SDR at the RTL-SDR maximum)", synth.C2, 1, *a)
    run("configs[1] 64 x 2.4 MS/s (register kernel)", synth.C2, 64, *a)
    run("configs[1] 64 x 2.4 MS/s (tensor-core kernel)", synth.C2, 64, *a, fft_impl=E.FFT_TC256)
    run("configs[2] 20 MS/s nperseg 1024", synth.C3A, 1, *a)
    run("configs[2] 20 MS/s nperseg 4096", synth.C3B, 1, *a)
    run("configs[3] replay: 64 of 512 channels x 300 kS/s per GPU", synth.C4, 64, *a)
    run("configs[4] dense pulses 300 kS/s", synth.C5, 1, *a)
    run("configs[4] dense pulses 64 x 2.4 MS/s", synth.C5B, 64, *a)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: IQ Msamples/s from uint8 IQ to Signal records (BASELINE.json `metric`).

Workload = BASELINE.json configs[1]: a batch of 64 concurrent 2.4 MS/s streams per GPU,
nperseg 256 Hamming, -90 dBW / 5 dB SNR, 8..40 ms.  One *step* = one callback block (1 s of
samples) of every stream of the batch.  With N GPUs each rank owns 64 streams (weak scaling, no
collective on the data path; streams are independent analyzers).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

`value`   device-resident: the batch is in HBM before the timed region; K launches timed with CUDA
          events on the launching stream, max over ranks.
`e2e`     the public API (`BatchAnalyzer.process_blocks`) on pinned HOST buffers: H2D copy, kernels,
          D2H of the records, float64 finalisation into Signal objects and the shadow filter.
`roofline` the spectrogram kernel against the measured HBM copy bandwidth (algorithmic 2 B/sample).
`cpu_baseline` the oracle port of the reference's scipy path on this box's host cores (bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "iq_msamples_per_s"
UNIT = "Msamples/s"
N_DISTINCT = 8          # distinct seeded streams per rank, tiled to the batch of 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="streams per GPU")
    ap.add_argument("--cpu-sample", type=int, default=8, help="stream-blocks timed for cpu_baseline")
    ap.add_argument("--profile", action="store_true", help="device-resident region only (for runs under ncu)")
    ap.add_argument("--fft-impl", default="auto", choices=["auto", "reg256", "tc256", "generic"],
                    help="spectrogram kernel: auto = reg256 (registers, packed fp32x2); tc256 = tensor-core stage 1 (tcgen05)")
    return ap.parse_args()


def workload():
    from pyradiotracking_b200 import synth

    return synth.C2


def analyzer_kwargs(w, n_streams, rank):
    return dict(
        devices=[str(rank * n_streams + i) for i in range(n_streams)], calibration_db=[0.0] * n_streams,
        sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=w.nperseg, fft_window="hamming",
        signal_min_duration_ms=w.signal_min_duration_ms, signal_max_duration_ms=w.signal_max_duration_ms,
        signal_threshold_dbw=w.signal_threshold_dbw, snr_threshold_db=w.snr_threshold_db,
        sdr_callback_length=w.block_samples)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def recorded_traffic():
    """dram bytes per launch of the spectrogram kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("spectrogram_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (the reference is pure Python: no oracle/_ref binary)
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One analyzer process (the reference runs one process per SDR, __main__.py:94-130)."""
    import datetime

    from oracle import restatement as R
    from pyradiotracking_b200 import synth

    stream, n_blocks = args
    w = synth.C2
    cap = synth.make_stream(w, stream, n_blocks)
    P = R.Params.make(sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=w.nperseg)
    ora = R.OracleAnalyzer(P)
    t0 = datetime.datetime(2026, 1, 1)
    t = time.perf_counter()
    n_sig = 0
    for b in range(n_blocks):
        n_sig += len(ora.process_block(cap[b], t0)[4])
    return time.perf_counter() - t, n_blocks * w.block_samples, n_sig


def cpu_port_single(n_stream_blocks):
    """1 core: `n_stream_blocks` callback blocks of one 2.4 MS/s stream through the oracle port."""
    dt, samples, _ = _cpu_worker((0, n_stream_blocks))
    return samples / dt / 1e6


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), one analyzer
    process per host core, each step = one callback block per process."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload()
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    ctx = mp.get_context("fork")
    per_step = []
    with ctx.Pool(procs) as pool:
        for i in range(args.warmup + args.steps):
            t = time.perf_counter()
            res = pool.map(_cpu_worker, [(s, 1) for s in range(procs)])
            wall = time.perf_counter() - t
            gen = 0.0   # generation of the synthetic block happens inside the worker but outside its timer
            busy = max(r[0] for r in res)
            if i >= args.warmup:
                per_step.append((busy, sum(r[1] for r in res)))
            del wall, gen
    tot_t = sum(p[0] for p in per_step)
    tot_s = sum(p[1] for p in per_step)
    val = tot_s / tot_t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(per_step), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 64x2.4MS/s nperseg256 hamming -90dBW/5dB 8-40ms", "streams_per_step": procs,
                   "block_samples": w.block_samples},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{procs} analyzer processes x 1 callback block (2.4 M samples) per step, oracle port of scipy path"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from pyradiotracking_b200 import synth
    from pyradiotracking_b200.analyze import BatchAnalyzer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = workload()
    S = args.streams
    n_blk = 2
    # synthetic batch: N_DISTINCT seeded streams x 2 consecutive blocks, tiled to S streams
    distinct = [synth.make_stream(w, rank * S + i, n_blk) for i in range(min(N_DISTINCT, S))]
    host = torch.empty((n_blk, S, w.block_bytes), dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    for s in range(S):
        hnp[:, s, :] = distinct[s % len(distinct)]
    dev = host.cuda()
    from pyradiotracking_b200 import engine as _E
    impl = {"auto": _E.FFT_AUTO, "reg256": _E.FFT_REG256, "tc256": _E.FFT_TC256, "generic": _E.FFT_GENERIC}[args.fft_impl]
    ba = BatchAnalyzer(**analyzer_kwargs(w, S, rank), cuda_device=local, fft_impl=impl)
    eng = ba.engine
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    samples_per_step = S * w.block_samples

    # ---- device-resident timed region ------------------------------------------------------------
    clocks = ClockSampler(local)      # nvidia-smi needs ~0.2 s before its first line: start it ahead of the warm-up
    clocks.start()
    for i in range(args.warmup):
        eng.launch(dev[i % n_blk])
    n_rec = len(eng.fetch())
    eng.enable_timing(True)
    eng.timing(reset=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        eng.launch(dev[(args.warmup + i) % n_blk])
    eng.join()          # the scan kernels of the last launch run on the engine's scan stream: wait for them too
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    tim = eng.timing(reset=True)
    eng.enable_timing(False)
    n_rec = len(eng.fetch())

    # ---- end to end through the public API with host buffers -----------------------------------------
    import datetime

    t0 = datetime.datetime(2026, 1, 1)
    ts = [t0] * S
    # Two blocks in flight (submit i+1 before collect i): the H2D copy and the kernels of the next block
    # overlap the float64 finalisation of the current one, like a live multi-SDR ingest loop would run.
    e2e_steps = 0 if args.profile else max(3, min(args.steps, 40))      # 6 ms each (PCIe-bound): the pipeline fill and drain amortise
    for i in range(0 if args.profile else 2):
        ba.process_blocks(hnp[i % n_blk], ts)
    barrier()
    t_e2e = time.perf_counter()
    d2h = 0
    n_sig = 0
    if e2e_steps:
        ba.submit(hnp[0])
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            ba.submit(hnp[(i + 1) % n_blk])
        res = ba.collect(ts)
        n_sig += sum(len(r[0]) for r in res)
        d2h += 8 + 40 * ba.last_record_count
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t_e2e)
    # The timed region of a default run lasts tens of milliseconds -- shorter than nvidia-smi's start-up.  If it yielded fewer
    # than 5 clock samples, keep the GPU under the very same load (untimed launches of the same batch) until it has.
    clock_extra_s = 0.0
    t_ex = time.perf_counter()
    while len(clocks.lines) < 5 and time.perf_counter() - t_ex < 3.0 and clocks.proc is not None:
        for i in range(50):
            eng.launch(dev[i % n_blk])
        eng.join()
        torch.cuda.synchronize()
        clock_extra_s = time.perf_counter() - t_ex
    if clock_extra_s > 0:
        eng.fetch()
    clk = clocks.stop()
    clk["sampled_over"] = "timed region" if clock_extra_s == 0 else f"timed region + {clock_extra_s:.2f} s of identical untimed launches"

    if rank == 0:
        value = world * samples_per_step * args.steps / (ms * 1e-3) / 1e6
        peak, peak_kind = measured_peak()
        k_ms = tim["spectrogram_ms"] / max(1, tim["launches"])
        # consecutive spectrogram kernels run on alternating streams and overlap each other's tails: a kernel's own
        # event-bracketed duration then exceeds the time the GPU spends per launch, which is at most the step time
        k_eff = min(k_ms, ms / args.steps)
        achieved = 2.0 * samples_per_step / (k_eff * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 64x2.4MS/s nperseg256 hamming -90dBW/5dB 8-40ms", "streams_per_gpu": S,
                       "block_samples": w.block_samples, "distinct_streams": min(N_DISTINCT, S),
                       "l2": "inputs larger than L2 (307 MB per step, 2 alternating blocks)",
                       "records_per_step": n_rec, "extract_work_items_per_step": eng.last_counts()[0], "fft_impl": args.fft_impl,
                       "scan_overlap": os.environ.get("RT_SCAN_OVERLAP", "1") != "0"},
            "clocks": clk,
            "e2e": None if args.profile else {
                "value": world * samples_per_step * e2e_steps / e2e_s / 1e6, "unit": UNIT,
                "h2d_bytes_per_step": S * w.block_bytes, "d2h_bytes_per_step": d2h // e2e_steps,
                "signals_per_step": n_sig / e2e_steps},
            "gpu_launches": int(tim["kernels"]),
            "roofline": {"bound": "hbm", "kernel": "spectro_tc256" if args.fft_impl == "tc256" else "spectro_reg256", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic(), "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": 2 * samples_per_step, "kernel_ms": k_eff,
                         "kernel_ms_event_bracketed": k_ms,
                         "kernel_share_of_step": k_eff / (ms / args.steps),
                         "other_kernels_ms": {k: tim[k] / max(1, tim["launches"]) for k in ("rowmean_ms", "probe_ms", "extract_ms")}},
        }
        if world == 1 and not args.profile:
            v = cpu_port_single(args.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{args.cpu_sample} callback blocks (2.4 M samples each) of one stream, oracle port (numpy pocketfft + run extraction + shadow filter)"}
        print(json.dumps(line), flush=True)
    ba.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

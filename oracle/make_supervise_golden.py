"""Generate tests/golden/supervise_states.json: life-cycle messages of three UNMODIFIED reference analyzers driven with
scripted receive times (on time / jittered then one block 2.5 s late -> clock-drift stop / short state_update_s).
Build container only.   python -m oracle.make_supervise_golden
"""
import datetime
import json
import os
import signal as _signal
import sys

import numpy as np

from oracle import ref_harness
from pyradiotracking_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T0 = datetime.datetime(2026, 8, 1, 6, 0, 0)
N_BLOCKS = 6
# seconds after T0 at which block k of each device is received
RECV = {
    "0": [0.0, 1.0, 2.0, 3.0, 4.0, 5.0],
    "1": [0.1, 1.4, 2.2, 3.3, 6.9, 7.9],          # block 4 arrives 2.5 s late: drift > 2 blocks -> STOPPED + cancel
    "2": [0.2, 1.2, 2.2, 3.2, 4.2, 5.2],
}
STATE_UPDATE_S = {"0": 300, "1": 300, "2": 2}     # device 2 repeats RUNNING every > 2 s


def run_device(dev: str):
    w = synth.C1
    cap = synth.make_stream(w, 70 + int(dev), N_BLOCKS)
    r = ref_harness.ReferenceRunner(T0, device=dev, sample_rate=w.sample_rate, state_update_s=STATE_UPDATE_S[dev])
    rows = []
    Sig = None
    for k in range(N_BLOCKS):
        if r.an.sdr.cancelled:
            break                                   # the read loop was cancelled: no further callbacks
        iq = np.ascontiguousarray(cap[k], dtype=np.uint8).astype(np.float64).view(np.complex128)
        iq /= 127.5
        iq -= 1 + 1j
        r.clock.set(T0 + datetime.timedelta(seconds=RECV[dev][k]))
        before = len(r.queue.items)
        old = _signal.signal(_signal.SIGALRM, _signal.SIG_IGN)
        try:
            r.an.process_samples(iq, None)
        finally:
            _signal.alarm(0)
            _signal.signal(_signal.SIGALRM, old)
        Sig = sys.modules["radiotracking"].Signal
        msgs = r.queue.items[before:]
        rows.append(dict(
            block=k, last_data_ts=r.an.last_data_ts.value, cancelled=bool(r.an.sdr.cancelled),
            states=[[m.device, m.ts.isoformat(), m.state.name] for m in msgs if not isinstance(m, Sig)],
            signals=[[s.ts.isoformat(), s.frequency, s.duration.total_seconds()] for s in msgs if isinstance(s, Sig)]))
    return rows


def main():
    doc = dict(t0=T0.isoformat(), recv=RECV, state_update_s=STATE_UPDATE_S, devices={d: run_device(d) for d in RECV})
    path = os.path.join(ROOT, "tests", "golden", "supervise_states.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    for d, rows in doc["devices"].items():
        print(d, [(r["block"], [s[2] for s in r["states"]], len(r["signals"]), r["cancelled"]) for r in rows])


if __name__ == "__main__":
    main()

"""Supervision of a multi-stream engine against the life-cycle messages of three
UNMODIFIED reference analyzers (oracle/make_supervise_golden.py)."""
import datetime
import json
import os

import numpy as np
import pytest

from pyradiotracking_b200 import messages, supervise, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "supervise_states.json")


class _Q:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


class _FakeBatch:
    """The state logic does not need a GPU: an engine that finds nothing."""

    def __init__(self, n):
        self.n, self.resets = n, []

    def process_blocks(self, buffers, ts_start):
        self.ts_start = list(ts_start)
        return [([], [], []) for _ in range(self.n)]

    def reset_stream(self, s):
        self.resets.append(s)

    def close(self):
        pass


def _make(monkeypatch, q, fake=True, **over):
    g = json.load(open(GOLDEN))
    devs = sorted(g["devices"])
    w = synth.C1
    if fake:
        monkeypatch.setattr(supervise, "BatchAnalyzer", lambda **kw: _FakeBatch(len(kw["devices"])))
    kw = dict(device=devs, calibration_db=[0.0] * len(devs), sample_rate=w.sample_rate, center_freq=w.center_freq, fft_nperseg=256,
              fft_window="hamming", signal_min_duration_ms=8, signal_max_duration_ms=40, signal_threshold_dbw=-90.0,
              snr_threshold_db=5.0, sdr_max_restart=3, sdr_timeout_s=2, state_update_s=300, sdr_callback_length=None, signal_queue=q)
    kw.update(over)
    return g, devs, supervise.MultiStreamAnalyzer(**kw)


def _drive(g, devs, m, q, caps=None):
    """Feed the blocks with the fixture's receive times; per device state_update_s as in the fixture."""
    t0 = datetime.datetime.fromisoformat(g["t0"])
    w = synth.C1
    per_dev = {d: [] for d in devs}
    n_blocks = max(len(g["devices"][d]) for d in devs)
    for k in range(n_blocks):
        recv = [t0 + datetime.timedelta(seconds=g["recv"][d][k]) for d in devs]
        before = len(q.items)
        buf = np.zeros((len(devs), 2 * w.block_samples), np.uint8) if caps is None else np.stack([caps[d][k] for d in devs])
        m.process_bytes(buf, recv)
        for msg in q.items[before:]:
            per_dev[msg.device].append((k, msg))
    return per_dev


def test_state_messages_and_heartbeats_match_three_reference_analyzers(monkeypatch):
    q = _Q()
    g, devs, m = _make(monkeypatch, q)
    # the fixture uses a different state_update_s per device: the reference key is per analyzer process
    upd = g["state_update_s"]
    orig = m.update_state

    def update_state(s, ts, state):
        m.state_update_s = upd[devs[s]]
        orig(s, ts, state)

    m.update_state = update_state
    per_dev = _drive(g, devs, m, q)
    for si, d in enumerate(devs):
        want = [(r["block"], st) for r in g["devices"][d] for st in r["states"]]
        got = [(k, [msg.device, msg.ts.isoformat(), msg.state.name]) for k, msg in per_dev[d] if isinstance(msg, messages.StateMessage) or hasattr(msg, "state")]
        assert got == want, d
        last = g["devices"][d][-1]
        assert m.stopped[si] == last["cancelled"]
        assert m.last_data_ts[si].value == last["last_data_ts"]
    # the stopped stream is skipped afterwards and its carry dropped
    assert m.batch.resets == [devs.index("1")]


def test_runner_check_restarts_within_budget(monkeypatch):
    q = _Q()
    g, devs, m = _make(monkeypatch, q, sdr_max_restart=1)
    t0 = datetime.datetime(2026, 8, 1, 6, 0, 0)
    w = synth.C1
    buf = np.zeros((len(devs), 2 * w.block_samples), np.uint8)
    assert m.check_streams(t0)                                   # nothing started yet: nothing to do (__main__.py:161-162)
    m.process_bytes(buf, [t0] * 3)
    assert m.check_streams(t0 + datetime.timedelta(seconds=1)) and not any(m.stopped)
    n0 = len(q.items)
    assert m.check_streams(t0 + datetime.timedelta(seconds=5))   # all three heartbeats are older than sdr_timeout_s
    stops = q.items[n0:]
    assert [(x.device, x.state.name) for x in stops] == [(d, "STOPPED") for d in devs]
    assert all(x.ts == t0.astimezone(datetime.timezone.utc) for x in stops)   # stamped with the last heartbeat (__main__.py:171)
    assert m.restarts_left == [0, 0, 0] and all(v.value == 0.0 for v in m.last_data_ts) and m._ts == [None] * 3
    m.process_bytes(buf, [t0 + datetime.timedelta(seconds=6)] * 3)            # restarted analyzers announce themselves again
    assert [x.state.name for x in q.items[-3:]] == ["STARTED"] * 3
    assert not m.check_streams(t0 + datetime.timedelta(seconds=20))            # dead beyond the restart count (__main__.py:178-181)


@pytest.mark.gpu
def test_signals_and_states_on_the_gpu_match_the_reference_analyzers(monkeypatch):
    q = _Q()
    g, devs, m = _make(monkeypatch, q, fake=False)
    w = synth.C1
    caps = {d: synth.make_stream(w, 70 + int(d), 6) for d in devs}
    upd = g["state_update_s"]
    orig = m.update_state

    def update_state(s, ts, state):
        m.state_update_s = upd[devs[s]]
        orig(s, ts, state)

    m.update_state = update_state
    try:
        per_dev = _drive(g, devs, m, q, caps)
        for d in devs:
            for r in g["devices"][d]:
                got = [[x.ts.isoformat(), x.frequency, x.duration.total_seconds()] for k, x in per_dev[d] if k == r["block"] and hasattr(x, "snr")]
                assert got == r["signals"], (d, r["block"])
            n_ref = len(g["devices"][d])
            assert not [x for k, x in per_dev[d] if k >= n_ref]     # nothing after the read loop was cancelled
    finally:
        m.close()

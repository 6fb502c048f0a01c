"""Bulk serialisation against payloads produced by the UNMODIFIED reference consumers
(oracle/make_publish_golden.py: CSVConsumer and MQTTConsumer.add with a recording paho stub)."""
import base64
import datetime
import io
import json
import os

import pytest

from pyradiotracking_b200 import messages, publish

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "publish_rows.json")


def _messages():
    """The same values as oracle/make_publish_golden.py::messages, on the mirror classes."""
    utc = datetime.timezone.utc
    t0 = datetime.datetime(2026, 5, 17, 23, 59, 59, 999999, tzinfo=utc)
    vals = [(150.123456e6, 0.010666666666666666, -43.21987654321, -47.5, 1.25e-3, -93.00000000000001, 45.5),
            (149999999.99999997, 0.04, -1e-7, -60.0, 0.0, -94.5, 34.5),
            (150.2e6, 0.008, -55.55555555555556, -58.125, 12.0, -90.0, 31.875),
            (1.5015e8, 0.0213333, -70.0, -71.0, 3.3333333333333335, -95.25, 24.25)]
    sigs = [messages.Signal(str(i % 3), t0 + datetime.timedelta(milliseconds=137 * i, microseconds=i), f,
                            datetime.timedelta(seconds=d), mx, av, sd, nz, snr) for i, (f, d, mx, av, sd, nz, snr) in enumerate(vals)]
    ms = messages.MatchingSignal(["0", "1", "2", "3"])
    for s in sigs[:3]:
        ms._sigs[s.device] = s
    states = [messages.StateMessage("0", t0, messages.StateMessage.State.STARTED),
              messages.StateMessage("1", t0 + datetime.timedelta(seconds=1), messages.StateMessage.State.RUNNING)]
    return sigs, ms, states


def test_csv_files_are_byte_identical_to_the_reference_consumer():
    g = json.load(open(GOLDEN))
    sigs, ms, states = _messages()
    out = io.StringIO(newline="")
    c = publish.BulkCSVConsumer(out, messages.Signal, messages.Signal.header)
    assert c.add_batch(sigs + [ms] + states) == len(sigs)          # other message types are skipped (consume.py:194)
    assert out.getvalue() == g["signal_csv"]
    one = io.StringIO(newline="")
    c1 = publish.BulkCSVConsumer(one, messages.Signal, messages.Signal.header)
    for m in sigs + [ms] + states:
        c1.add(m)
    assert one.getvalue() == g["signal_csv"]
    mout = io.StringIO(newline="")
    publish.BulkCSVConsumer(mout, messages.MatchingSignal, ms.header).add_batch([ms])
    assert mout.getvalue() == g["matched_csv"]


def test_mqtt_payloads_are_identical_to_the_reference_publisher():
    g = json.load(open(GOLDEN))
    sigs, ms, states = _messages()
    msgs = sigs + [ms] + states
    want = {}
    for topic, payload in g["published"]:
        want.setdefault(topic, []).append(payload)
    got = {}
    js, cs = publish.json_payloads(msgs), publish.csv_payloads(msgs)
    for m, j, c in zip(msgs, js, cs):
        stem = publish.mqtt_topic("/radiotracking", m)
        got.setdefault(stem + "/json", []).append(j)
        got.setdefault(stem + "/csv", []).append(c)
    for topic in got:
        assert got[topic] == want[topic], topic
    assert {t for t in want if not t.endswith("/cbor")} == set(got)
    pytest.importorskip("cbor2")
    for m, p in zip(msgs, publish.cbor_payloads(msgs)):
        stem = publish.mqtt_topic("/radiotracking", m)
        assert base64.b64encode(p).decode() in want[stem + "/cbor"]


def test_bulk_rows_equal_per_message_rows_for_many_signals():
    import numpy as np

    rng = np.random.default_rng(3)
    t0 = datetime.datetime(2026, 1, 1, tzinfo=datetime.timezone.utc)
    sigs = [messages.Signal(str(int(rng.integers(8))), t0 + datetime.timedelta(microseconds=int(rng.integers(10 ** 9))), float(rng.normal(150e6, 1e5)),
                            datetime.timedelta(microseconds=int(rng.integers(8000, 40000))), *[float(x) for x in rng.normal(-60, 20, 5)]) for _ in range(2000)]
    bulk = publish.csv_rows(sigs)
    assert bulk == "".join(publish.csv_rows([s]) for s in sigs)
    assert publish.csv_payloads(sigs) == [r for r in bulk.split("\r\n") if r]
    assert [json.loads(p)["Frequency"] for p in publish.json_payloads(sigs)] == [s.frequency for s in sigs]

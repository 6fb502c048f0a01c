"""Generate tests/golden/publish_rows.json by running the UNMODIFIED reference consumers
(radiotracking.consume.CSVConsumer and MQTTConsumer.add) on seeded messages.

Build container only.  paho-mqtt is not installed: a recording stub stands in for `paho.mqtt.client.Client`, so that
MQTTConsumer.add (consume.py:127-162) runs as written and its publish() calls are captured.
    python -m oracle.make_publish_golden
"""
import base64
import datetime
import io
import json
import os
import sys
import types

from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Client:
    published = []

    def __init__(self, *a, **k):
        pass

    def connect(self, *a, **k):
        pass

    def loop_start(self):
        pass

    def loop_stop(self):
        pass

    def publish(self, topic, payload, qos=0):
        _Client.published.append((topic, payload))


def load():
    ref_harness._install_stubs()
    paho = types.ModuleType("paho")
    mqtt = types.ModuleType("paho.mqtt")
    client = types.ModuleType("paho.mqtt.client")
    client.Client = _Client
    paho.mqtt = mqtt
    mqtt.client = client
    sys.modules.update({"paho": paho, "paho.mqtt": mqtt, "paho.mqtt.client": client})
    if ref_harness.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    import radiotracking  # type: ignore
    import radiotracking.consume as rc  # type: ignore

    return radiotracking, rc


def messages(radiotracking):
    """Signals, a MatchingSignal and StateMessages with awkward values (exponents, negative zero-ish, long fractions)."""
    utc = datetime.timezone.utc
    t0 = datetime.datetime(2026, 5, 17, 23, 59, 59, 999999, tzinfo=utc)
    sigs = []
    vals = [(150.123456e6, 0.010666666666666666, -43.21987654321, -47.5, 1.25e-3, -93.00000000000001, 45.5),
            (149999999.99999997, 0.04, -1e-7, -60.0, 0.0, -94.5, 34.5),
            (150.2e6, 0.008, -55.55555555555556, -58.125, 12.0, -90.0, 31.875),
            (1.5015e8, 0.0213333, -70.0, -71.0, 3.3333333333333335, -95.25, 24.25)]
    for i, (f, d, mx, av, sd, nz, snr) in enumerate(vals):
        sigs.append(radiotracking.Signal(str(i % 3), t0 + datetime.timedelta(milliseconds=137 * i, microseconds=i), f,
                                         datetime.timedelta(seconds=d), mx, av, sd, nz, snr))
    ms = radiotracking.MatchingSignal(["0", "1", "2", "3"])
    for s in sigs[:3]:
        ms.add_member(s)
    states = [radiotracking.StateMessage("0", t0, radiotracking.StateMessage.State.STARTED),
              radiotracking.StateMessage("1", t0 + datetime.timedelta(seconds=1), radiotracking.StateMessage.State.RUNNING)]
    return sigs, ms, states


def main():
    radiotracking, rc = load()
    sigs, ms, states = messages(radiotracking)
    out = io.StringIO(newline="")
    cons = rc.CSVConsumer(out, radiotracking.Signal, radiotracking.Signal.header)
    for m in sigs + [ms] + states:
        cons.add(m)
    mout = io.StringIO(newline="")
    mcons = rc.CSVConsumer(mout, radiotracking.MatchingSignal, ms.header)
    mcons.add(ms)
    _Client.published = []
    mq = rc.MQTTConsumer("localhost", 1883, 1, 60, 0, prefix="/radiotracking")
    for m in sigs + [ms] + states:
        mq.add(m)
    pub = [(t, base64.b64encode(p).decode() if isinstance(p, bytes) else p) for t, p in _Client.published]
    doc = dict(signal_csv=out.getvalue(), matched_csv=mout.getvalue(), published=pub)
    path = os.path.join(ROOT, "tests", "golden", "publish_rows.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    print(path, len(out.getvalue()), "bytes of CSV,", len(pub), "MQTT payloads")


if __name__ == "__main__":
    main()

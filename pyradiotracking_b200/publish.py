"""Bulk serialisation of the messages that leave the detection path.

The reference formats one message at a time: `CSVConsumer.add` writes a row and flushes the file per Signal
(radiotracking/consume.py:192-200), `MQTTConsumer.add` builds a JSON, a CSV and a CBOR payload per message through a
fresh `StringIO` / `json.dumps` / `cbor.dumps` (consume.py:127-162).  That is fine at five signals per second and the
dominant cost behind an engine that emits 10^5.  These helpers produce the SAME BYTES for a whole batch at once:

    csv_rows(messages)        -> the text CSVConsumer would have appended (`;`-separated, excel dialect, CRLF rows)
    json_payloads(messages)   -> list of the MQTT JSON payloads (consume.py:141-145)
    csv_payloads(messages)    -> list of the MQTT CSV payloads (consume.py:147-151)
    cbor_payloads(messages)   -> list of the MQTT CBOR payloads (consume.py:153-160; needs cbor2, like the reference)
    mqtt_topic(prefix, msg)   -> the topic stem the reference publishes a message under (consume.py:129-139)
    BulkCSVConsumer           -> CSVConsumer's interface (`add`) plus `add_batch`, one flush per batch

Formats are the reference's: `timedelta` -> seconds as float (consume.py:24-58), datetimes as `str()` in CSV and ISO 8601
in JSON, CBOR tag 1337 for durations and timestamps as epoch numbers.
"""
import csv
import datetime
import io
import json
from typing import Any, Iterable, List, Optional, Type


def _csvify(o: Any) -> Any:
    """consume.py:50-55"""
    if isinstance(o, datetime.timedelta):
        return o.total_seconds()
    return o


def _jsonify(o: Any) -> Any:
    """consume.py:24-34"""
    if isinstance(o, datetime.datetime):
        return o.isoformat()
    if isinstance(o, datetime.timedelta):
        return o.total_seconds()
    raise TypeError(f"Object of type {type(o)} is not JSON serializable")


def csv_rows(messages: Iterable, header: Optional[List[str]] = None) -> str:
    """All rows `CSVConsumer` (consume.py:165-200) would write for `messages`, as one string."""
    buf = io.StringIO()
    w = csv.writer(buf, dialect="excel", delimiter=";")
    if header:
        w.writerow(header)
    w.writerows([_csvify(v) for v in m.as_list] for m in messages)
    return buf.getvalue()


def csv_payloads(messages: Iterable) -> List[str]:
    """MQTT CSV payloads (consume.py:147-151): one row each, without the line terminator."""
    text = csv_rows(messages)
    # a row never contains a line break (no field of the message types does), so the terminator splits rows
    return text.split("\r\n")[:-1] if text else []


_ENCODER = json.JSONEncoder(default=_jsonify)          # json.dumps' defaults, built once


def json_payloads(messages: Iterable) -> List[str]:
    """MQTT JSON payloads (consume.py:141-145): `json.dumps(msg.as_dict, default=jsonify)`."""
    enc = _ENCODER.encode
    return [enc(m.as_dict) for m in messages]


def cbor_payloads(messages: Iterable) -> List[bytes]:
    """MQTT CBOR payloads (consume.py:153-160).  cbor2 is a dependency of the reference (requirements.txt); without it
    this raises ImportError instead of inventing an encoding."""
    import cbor2

    def cborify(encoder, o):                            # consume.py:37-41
        if isinstance(o, datetime.timedelta):
            encoder.encode(cbor2.CBORTag(1337, o.total_seconds()))

    return [cbor2.dumps(m.as_list, timezone=datetime.timezone.utc, datetime_as_timestamp=True, default=cborify) for m in messages]


def mqtt_topic(prefix: str, message) -> Optional[str]:
    """Topic stem of consume.py:129-139 (`/json`, `/csv`, `/cbor` are appended by the publisher); None for unknown types."""
    if hasattr(message, "_sigs"):
        return f"{prefix}/matched"
    if hasattr(message, "state"):
        return f"{prefix}/state"
    if hasattr(message, "snr") and hasattr(message, "device"):
        return f"{prefix}/device/{message.device}"
    return None


class BulkCSVConsumer:
    """`CSVConsumer` (consume.py:165-200) with a batch entry: same file content, one write + flush per batch."""

    def __init__(self, out, cls: Type, header: Optional[List[str]] = None):
        self.out = out
        self.cls = cls
        if header:
            self.out.write(csv_rows([], header=header))
        self.out.flush()

    def add(self, signal) -> None:
        self.add_batch([signal])

    def add_batch(self, signals: Iterable) -> int:
        rows = [s for s in signals if isinstance(s, self.cls)]
        if rows:
            self.out.write(csv_rows(rows))
            self.out.flush()
        return len(rows)

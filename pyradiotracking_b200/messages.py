"""Message types that leave the detection hot path.

These mirror the public surface of the reference's message classes
(`radiotracking/__init__.py:13-22` dB helpers, `:61-93` StateMessage,
`:110-202` Signal) so that the reference's consumers (`consume.py`,
`match.py`, `present.py`) can take them unchanged: same attribute names, same
`header` / `as_list` / `as_dict` views, same constructor coercions.

When the reference package itself is importable (a live deployment), use
`message_types()` to obtain the reference's own classes instead, so that
`isinstance` checks in its consumers (`match.py:62`, `consume.py:130-160`)
keep working.
"""
import datetime
import enum
import math
from typing import Any, Dict, List, Tuple, Union


def dB(val):
    """Power ratio -> decibel (reference: radiotracking/__init__.py:13-17)."""
    import numpy as np

    return 10 * np.log10(val)


def from_dB(dB_val: float) -> float:
    """Decibel -> power ratio (reference: radiotracking/__init__.py:20-22)."""
    return 10 ** (dB_val / 10)


class _Message:
    header: List[str] = []

    @property
    def as_list(self) -> List[Any]:
        raise NotImplementedError

    @property
    def as_dict(self) -> Dict[str, Any]:
        return dict(zip(self.header, self.as_list))


class StateMessage(_Message):
    """Analyzer life-cycle heartbeat (reference: radiotracking/__init__.py:61-93)."""

    class State(enum.Enum):
        STOPPED = 0
        RUNNING = 1
        STARTED = 2

    header = ["Device", "Time", "State"]

    def __init__(self, device: str, ts: datetime.datetime, state: Union["StateMessage.State", int, str]):
        self.device = device
        self.ts = ts
        self.state = state if isinstance(state, StateMessage.State) else StateMessage.State(int(state))

    @property
    def as_list(self) -> List[Any]:
        return [self.device, self.ts, self.state.value]

    def __repr__(self) -> str:
        return f"StateMessage({self.device}, {self.ts}, {self.state})"


class Signal(_Message):
    """One detection on one device (reference: radiotracking/__init__.py:110-202).

    Fields: device, ts (tz-aware datetime), frequency [Hz], duration (timedelta),
    max / avg [dBW], std [dB], noise [dBW], snr [dB].
    """

    header = ["Device", "Time", "Frequency", "Duration", "max (dBW)", "avg (dBW)", "std (dB)", "noise (dBW)", "snr (dB)"]

    def __init__(self, device, ts, frequency, duration, max_dBW, avg_dBW, std_dB, noise_dBW, snr_dB):
        self.device = device
        self.ts = ts if isinstance(ts, datetime.datetime) else datetime.datetime.fromisoformat(ts)
        self.frequency = float(frequency)
        self.duration = duration if isinstance(duration, datetime.timedelta) else datetime.timedelta(seconds=float(duration))
        self.max = float(max_dBW)
        self.avg = float(avg_dBW)
        self.std = float(std_dB)
        self.noise = float(noise_dBW)
        self.snr = float(snr_dB)

    @property
    def as_list(self) -> List[Any]:
        return [self.device, self.ts, self.frequency, self.duration, self.max, self.avg, self.std, self.noise, self.snr]

    def __repr__(self) -> str:
        return "Signal(" + ", ".join(str(v) for v in self.as_list) + ")"

    def __str__(self) -> str:
        return f"Signal<SDR {self.device}, {self.frequency / 1e6:.3f} MHz, {self.duration.total_seconds() * 1e3:.2f} ms, {self.max:.1f} dBW>"


def message_types() -> Tuple[type, type]:
    """(Signal, StateMessage) classes to emit: the reference's own when the
    `radiotracking` package is installed next to us, otherwise the mirrors."""
    try:
        import radiotracking  # type: ignore

        return radiotracking.Signal, radiotracking.StateMessage
    except Exception:
        return Signal, StateMessage


def _isfinite(x: float) -> bool:
    return not (math.isnan(x) or math.isinf(x))


class MatchingSignal(_Message):
    """A group of Signals of different devices that belong to one transmission
    (reference: radiotracking/__init__.py:205-334, MatchedSignal + MatchingSignal).

    `_sigs` maps device name -> Signal in first-insertion order; ts / duration / frequency / `_avgs` are the
    reference's derived views (earliest start, longest duration, median frequency, per-device average power).
    """

    def __init__(self, devices: List[str]):
        self.devices = devices
        self._sigs: Dict[str, Signal] = {}

    @property
    def duration(self) -> datetime.timedelta:
        return max(sig.duration for sig in self._sigs.values())

    @property
    def ts(self) -> datetime.datetime:
        return min(sig.ts for sig in self._sigs.values())

    @property
    def frequency(self) -> float:
        import statistics

        return statistics.median(sig.frequency for sig in self._sigs.values())

    @property
    def _avgs(self) -> List[Any]:
        return [self._sigs[d].avg if d in self._sigs else None for d in self.devices]

    @property
    def header(self) -> List[str]:  # type: ignore[override]
        return ["Time", "Frequency", "Duration", *self.devices]

    @property
    def as_list(self) -> List[Any]:
        return [self.ts, self.frequency, self.duration, *self._avgs]

    def __repr__(self) -> str:
        return f"MatchedSignal({self.devices}, {self.ts}, {self.frequency}, {self.duration}, " + ", ".join(repr(a) for a in self._avgs) + ")"


def matching_signal_type() -> type:
    """The reference's own MatchingSignal when `radiotracking` is importable (so that `isinstance` checks in
    consume.py keep working), otherwise the mirror."""
    try:
        import radiotracking  # type: ignore

        return radiotracking.MatchingSignal
    except Exception:
        return MatchingSignal

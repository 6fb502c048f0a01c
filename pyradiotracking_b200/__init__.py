"""B200-native detection hot path of Nature40/pyradiotracking.

`SignalAnalyzer` is a drop-in for `radiotracking.analyze.SignalAnalyzer`
(/root/reference/radiotracking/analyze.py:20): same constructor keys, same
`process_samples(buffer, context)` callback, same `Signal` / `StateMessage` traffic on
the queue.  The spectrogram, predicate, run extraction and per-signal statistics run in
hand-written sm_100a CUDA kernels behind the C ABI of `include/rt_engine.h`.
"""
from .messages import Signal, StateMessage, dB, from_dB  # noqa: F401

__all__ = ["Signal", "StateMessage", "dB", "from_dB", "SignalAnalyzer", "BatchAnalyzer", "Engine"]


def __getattr__(name):
    # lazy: importing the package must not need numpy-heavy or CUDA pieces
    if name in ("SignalAnalyzer", "BatchAnalyzer", "DetectionPlan"):
        from . import analyze

        return getattr(analyze, name)
    if name == "Engine":
        from .engine import Engine

        return Engine
    raise AttributeError(name)

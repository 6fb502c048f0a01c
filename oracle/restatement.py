"""CPU restatement (numpy, float64) of the reference detection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `pyradiotracking_b200/` may import this
module; it is the checker for `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.

What it restates (all paths relative to /root/reference):

* `radiotracking/analyze.py:113-117`   threshold / duration preparation
* `radiotracking/analyze.py:231-241`   the `scipy.signal.spectrogram` call
* `radiotracking/analyze.py:330-452`   `extract_signals`
* `radiotracking/analyze.py:282-328`   `is_shadow_of` / `filter_shadow_signals`
* `radiotracking/__init__.py:13-22`    `dB`, `from_dB`

Third-party arithmetic that is not vendored in the reference and is restated
here from its published algorithm:

* scipy (requirements.txt:2, unpinned; 1.18.1 in this image):
  `scipy/signal/_spectral_py.py` `spectrogram -> _spectral_helper -> _fft_helper`
  with `detrend='constant'`, `scaling='density'`, `mode='psd'`,
  `return_onesided=False`, `noverlap=0`.
* pyrtlsdr (requirements.txt:1, unpinned, absent): `packed_bytes_to_iq`,
  `x/127.5 - (1+1j)`.

Parity pin: the reference has no tests and no golden vectors, so
this restatement is pinned by running the *unmodified reference* in the build
container (`oracle/ref_harness.py`) on the synthetic captures of
`pyradiotracking_b200/synth.py` and committing its outputs as fixtures under
`tests/golden/` (`oracle/make_golden.py`).  `tests/test_oracle_golden.py` holds the
restatement to those fixtures; `tests/test_oracle_scipy.py` holds the spectrogram
restatement to the installed scipy.
"""
import datetime
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

UTC = datetime.timezone.utc


def dB(x):
    return 10 * np.log10(x)


def from_dB(x):
    return 10 ** (x / 10)


def bytes_to_iq(u8: np.ndarray) -> np.ndarray:
    """pyrtlsdr `packed_bytes_to_iq` (see module docstring)."""
    iq = np.ascontiguousarray(u8, dtype=np.uint8).astype(np.float64).view(np.complex128)
    iq /= 127.5
    iq -= 1 + 1j
    return iq


def resolve_window(window, nperseg: int) -> np.ndarray:
    """scipy `_triage_segments`: str/tuple -> `get_window(window, nperseg)` (DFT-even,
    i.e. periodic), array-like -> used verbatim (must have length nperseg)."""
    if isinstance(window, (str, tuple)):
        if window == "boxcar":
            return np.ones(nperseg)
        if window == "hamming":
            return _general_cosine_periodic(nperseg, (0.54, 1.0 - 0.54))   # general_hamming(alpha): [alpha, 1 - alpha]
        if window == "hann":
            return _general_cosine_periodic(nperseg, (0.5, 0.5))
        from scipy.signal import get_window  # other named windows: defer to scipy

        return get_window(window, nperseg)
    win = np.asarray(window, dtype=np.float64)
    if win.ndim != 1 or win.shape[0] != nperseg:
        raise ValueError("window must be 1-D with length nperseg")
    return win


def _general_cosine_periodic(n: int, a: Sequence[float]) -> np.ndarray:
    """scipy.signal.windows.general_cosine(n, a, sym=False): build n+1 symmetric
    points on linspace(-pi, pi) and drop the last one."""
    fac = np.linspace(-np.pi, np.pi, n + 1)
    w = np.zeros(n + 1)
    for k, ak in enumerate(a):
        w += ak * np.cos(k * fac)
    return w[:-1]


def spectrogram(x: np.ndarray, fs: float, window, nperseg: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """`scipy.signal.spectrogram(x, fs, window, nperseg, noverlap=0, return_onesided=False)`
    for complex `x` (call site analyze.py:234-241).  Returns `(freqs, times, S)` with
    `S.shape == (nperseg, T)`, T = len(x)//nperseg, frequencies in FFT order (no fftshift)."""
    x = np.asarray(x)
    N = x.shape[-1]
    win = resolve_window(window, nperseg)
    T = N // nperseg if N >= nperseg else 0
    seg = x[: T * nperseg].reshape(T, nperseg)
    # scipy keeps the transforms time-major (its Sxx is a transposed view whose per-bin rows are 16*nperseg-byte gathers);
    # the values are the same whichever way they are stored, so they are written bin-major here (pocketfft copies every
    # 1-D transform in and out anyway) and extract_* walk contiguous rows.  Blocks of segments / of bins keep the
    # intermediate arrays cache-resident; every segment sees exactly scipy's operations in scipy's order.
    XT = np.empty((nperseg, T), dtype=np.complex128)
    for a in range(0, T, 512):
        blk = seg[a:a + 512]
        d = blk - np.mean(blk, axis=-1, keepdims=True)     # detrend='constant'
        d *= win                                            # window (real x complex, same products as scipy's win * d)
        np.fft.fft(d, n=nperseg, axis=-1, out=XT[:, a:a + 512].T)   # two-sided
    scale = 1.0 / (fs * (win * win).sum())                  # density scaling
    S = np.empty((nperseg, T))
    for r in range(0, nperseg, 8):
        Y = XT[r:r + 8]
        np.multiply(np.conjugate(Y), Y, out=Y)              # scipy: conj(result) * result, then *= scale, then .real
        Y *= scale
        S[r:r + 8] = Y.real
    freqs = np.fft.fftfreq(nperseg, 1 / fs)
    times = np.arange(nperseg / 2, N - nperseg / 2 + 1, nperseg) / float(fs)
    return freqs, times, S


class Detection(NamedTuple):
    fi: int
    start: int          # first column of the statistics window; negative = reaches into the previous block
    end: int            # exclusive
    ts: datetime.datetime
    frequency: float
    duration: datetime.timedelta
    max: float
    avg: float
    std: float
    noise: float
    snr: float

    def key(self):
        return (self.fi, self.start, self.end)


class Params(NamedTuple):
    device: str
    calibration_db: float
    sample_rate: int
    center_freq: int
    fft_nperseg: int
    fft_window: object
    signal_min_duration: float      # seconds  (analyze.py:113)
    signal_max_duration: float      # seconds  (analyze.py:114)
    signal_threshold: float         # linear   (analyze.py:115)
    snr_threshold: float            # linear   (analyze.py:116)

    @classmethod
    def make(cls, device="0", calibration_db=0.0, sample_rate=300000, center_freq=150150000, fft_nperseg=256,
             fft_window="hamming", signal_min_duration_ms=8, signal_max_duration_ms=40,
             signal_threshold_dbw=-90.0, snr_threshold_db=5.0):
        return cls(device, calibration_db, sample_rate, center_freq, fft_nperseg, fft_window,
                   signal_min_duration_ms / 1000, signal_max_duration_ms / 1000,
                   from_dB(signal_threshold_dbw + calibration_db), from_dB(snr_threshold_db))


def _above(p, avg, P: Params) -> bool:
    """analyze.py:370-379 / 391-396 / 403-410: a cell belongs to a run unless it
    undershoots the absolute threshold or the SNR-vs-row-mean threshold."""
    if p < P.signal_threshold:
        return False
    if p / avg < P.snr_threshold:
        return False
    return True


def _finish(P: Params, freqs, times, S, last, fi, start, end, avg, ts_start, row=None) -> Optional[Detection]:
    """Duration test and per-signal statistics (analyze.py:419-450)."""
    start_dt = -times[-start] if start < 0 else times[start]
    duration_s = times[end] - start_dt
    if duration_s < P.signal_min_duration or duration_s > P.signal_max_duration:
        return None
    if row is None:
        row = S[fi]
    data = np.concatenate((last[fi][start:], row[:end])) if start < 0 else row[start:end]
    mean = np.mean(data)
    ts = (ts_start + datetime.timedelta(seconds=start_dt)).astimezone(UTC)
    return Detection(
        fi, start, end, ts, float(freqs[fi] + P.center_freq), datetime.timedelta(seconds=duration_s),
        float(dB(np.max(data)) - P.calibration_db), float(dB(mean) - P.calibration_db),
        float(np.std(dB(data))), float(dB(avg)), float(dB(mean / avg)),
    )


def extract_sequential(P: Params, freqs, times, S, last, ts_start) -> List[Detection]:
    """`extract_signals` restated cell by cell in the reference's own visiting order
    (strided probe, backward scan, forward scan, `ti_skip`).  Slow: for small cases and
    to validate `extract_runs`."""
    out: List[Detection] = []
    T = len(times)
    if T == 0:
        return out
    stride = max(1, int(P.signal_min_duration / (times[1] - times[0])))     # IndexError for T == 1, like :354
    reach = 0 if last is None else -len(last[0]) + 1
    for fi in range(S.shape[0]):
        row = S[fi]
        avg = None
        skip = 0
        for ti in range(0, T, stride):
            if ti < skip or row[ti] < P.signal_threshold:
                continue
            if avg is None:
                avg = np.mean(row)
            if row[ti] / avg < P.snr_threshold:
                continue
            start = ti
            while start > reach and _above(last[fi, start] if start < 0 else row[start], avg, P):
                start -= 1
            end = ti
            while end < T and _above(row[end], avg, P):
                end += 1
            if end == T:
                continue            # touches the block end: dropped, re-found from the next block
            skip = end
            det = _finish(P, freqs, times, S, last, fi, start, end, avg, ts_start)
            if det is not None:
                out.append(det)
    return out


def extract_runs(P: Params, freqs, times, S, last, ts_start) -> List[Detection]:
    """Same result as `extract_sequential`, computed per maximal run (SURVEY.md §8a,
    "parallel formulation"): every maximal run of above-cells is evaluated at most once,
    iff a probe column k*stride lies inside it and it does not touch the block end.

    Rows with a probe hit are handled as one bin-major block; the row mean is taken lazily, only for rows where a probed cell reaches
    the absolute threshold, like analyze.py:370-375."""
    out: List[Detection] = []
    T = len(times)
    if T == 0:
        return out
    stride = max(1, int(P.signal_min_duration / (times[1] - times[0])))
    reach = 0 if last is None else -len(last[0]) + 1
    rows = np.nonzero((S[:, ::stride] >= P.signal_threshold).any(axis=1))[0]
    if len(rows) == 0:
        return out
    Sc = S if (len(rows) == S.shape[0] and S.flags.c_contiguous) else np.ascontiguousarray(S[rows])   # bin-major candidate rows
    avgs = Sc.mean(axis=1)                                  # bit-identical to np.mean(S[fi]) (pairwise sums of the same values)
    with np.errstate(divide="ignore", invalid="ignore"):
        ab = ~(Sc < P.signal_threshold) & ~(Sc / avgs[:, None] < P.snr_threshold)
    pad = np.zeros((len(rows), T + 2), dtype=np.int8)
    pad[:, 1:-1] = ab
    edge = pad[:, 1:] != pad[:, :-1]                        # first cell of a run / one past its last, alternating along a row
    er, ec = np.nonzero(edge)                               # sorted by (row, then time)
    rs, run_s, run_e = er[0::2], ec[0::2], ec[1::2]
    seen = (-(-run_s // stride) * stride < run_e) & (run_e != T)
    for r, s, e in zip(rs[seen].tolist(), run_s[seen].tolist(), run_e[seen].tolist()):
        fi = int(rows[r])
        avg = avgs[r]
        if s > 0:
            start = s - 1               # the not-above cell before the run is part of the window
        else:
            start = 0                   # analyze.py:382-398 from column 0 (which is above)
            while start > reach and _above(last[fi, start] if start < 0 else Sc[r, 0], avg, P):
                start -= 1
        det = _finish(P, freqs, times, S, last, fi, start, e, avg, ts_start, row=Sc[r])
        if det is not None:
            out.append(det)
    return out


def is_shadow_of(sig: Detection, signals: Sequence[Detection]) -> Optional[int]:
    """analyze.py:283-313: index of the first time-overlapping, strictly louder signal."""
    for i, other in enumerate(signals):
        if sig.ts > other.ts + other.duration:
            continue
        if sig.ts + sig.duration < other.ts:
            continue
        if other.max > sig.max:
            return i
    return None


def filter_shadow(signals: Sequence[Detection]) -> List[Detection]:
    """analyze.py:315-328."""
    return [s for s in signals if is_shadow_of(s, signals) is None]


class OracleAnalyzer:
    """Block-by-block driver with the reference's carry (`_spectrogram_last`, analyze.py:268)."""

    def __init__(self, params: Params, sequential: bool = False):
        self.P = params
        self.last = None
        self.sequential = sequential

    def process_block(self, u8_block: np.ndarray, ts_start: datetime.datetime):
        """-> (freqs, times, S, detections_before_shadow_filter, detections_after)."""
        P = self.P
        freqs, times, S = spectrogram(bytes_to_iq(u8_block), P.sample_rate, P.fft_window, P.fft_nperseg)
        fn = extract_sequential if self.sequential else extract_runs
        found = fn(P, freqs, times, S, self.last, ts_start)
        kept = filter_shadow(found)
        self.last = S
        return freqs, times, S, found, kept

"""Recorded-capture ingest for the batched engine (BASELINE configs[3]).

The reference ingests only live (`sdr.read_samples_async`, radiotracking/analyze.py:143-157) and has no file reader;
its wire format is what librtlsdr delivers and `rtl_sdr -f ... file.bin` records: interleaved uint8 I,Q bytes.  This
module feeds such recordings -- one file per station channel -- to a `BatchAnalyzer`:

    CaptureReader     blocks `[n_streams, 2*block_samples]` uint8 out of a ring of (pinned) host buffers, filled by a
                      reader thread so that disk reads overlap the GPU
    replay()          submit / collect with two blocks in flight (the engine's staging ring copies block i+1 while the
                      kernels of block i run); timestamps follow the reference's drift-free virtual
                      clock: block b starts at t0 + b * block_samples / sample_rate (analyze.py:218-231)

A recording ends like an SDR that is unplugged mid-callback: the incomplete last block is dropped, for every channel
at the length of the shortest one.
"""
import datetime
import os
import queue
import threading
from typing import Callable, Iterator, List, Optional, Sequence

import numpy as np


def _alloc(shape, pinned: bool) -> np.ndarray:
    """Host buffer, page-locked when torch sees a CUDA device (plumbing only: the H2D copies of the engine then run
    at full PCIe speed and asynchronously)."""
    if pinned:
        try:
            import torch

            if torch.cuda.is_available():
                t = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
                arr = t.numpy()
                _alloc.keep.append(t)          # the tensor owns the memory
                return arr
        except ImportError:
            pass
    return np.empty(shape, dtype=np.uint8)


_alloc.keep = []


class CaptureReader:
    """Iterate over the blocks of `paths` (one raw uint8 IQ file per stream).

    Yields `(block_index, array)`; `array` is a view of one of `n_buffers` ring buffers and stays valid until
    `n_buffers - 1` further blocks have been taken (with the default 4: two blocks in flight in the engine plus the one
    being filled)."""

    def __init__(self, paths: Sequence[str], block_samples: int, n_buffers: int = 4, pinned: bool = True,
                 max_blocks: Optional[int] = None):
        if not paths:
            raise ValueError("no capture files")
        if n_buffers < 2:
            raise ValueError("need at least two ring buffers")
        self.paths = list(paths)
        self.block_bytes = 2 * int(block_samples)
        sizes = [os.path.getsize(p) for p in self.paths]
        self.n_blocks = min(sizes) // self.block_bytes              # tail dropped
        if max_blocks is not None:
            self.n_blocks = min(self.n_blocks, max_blocks)
        self.n_streams = len(self.paths)
        self._bufs = [_alloc((self.n_streams, self.block_bytes), pinned) for _ in range(n_buffers)]
        self._free: "queue.Queue[int]" = queue.Queue()
        self._full: "queue.Queue" = queue.Queue()
        for i in range(n_buffers):
            self._free.put(i)
        self._thread: Optional[threading.Thread] = None
        self._stop = False

    def _fill(self):
        try:
            files = [open(p, "rb", buffering=0) for p in self.paths]
            try:
                for b in range(self.n_blocks):
                    i = self._free.get()
                    if self._stop:
                        return
                    buf = self._bufs[i]
                    for s, f in enumerate(files):
                        view = memoryview(buf[s])
                        got = 0
                        while got < self.block_bytes:            # raw files may return short reads
                            n = f.readinto(view[got:])
                            if not n:
                                raise IOError(f"{self.paths[s]}: unexpected end of file in block {b}")
                            got += n
                    self._full.put((b, i))
            finally:
                for f in files:
                    f.close()
            self._full.put(None)
        except Exception as exc:        # hand the error to the consumer thread
            self._full.put(exc)

    def __iter__(self) -> Iterator:
        if self._thread is not None:
            raise RuntimeError("a CaptureReader can be iterated once")
        self._thread = threading.Thread(target=self._fill, daemon=True)
        self._thread.start()
        held: List[int] = []
        while True:
            item = self._full.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            b, i = item
            held.append(i)
            if len(held) == len(self._bufs):                     # the oldest view is no longer needed by contract
                self._free.put(held.pop(0))
            yield b, self._bufs[i]
        self._thread.join()

    def close(self):
        self._stop = True
        self._free.put(0)


def replay(paths: Sequence[str], analyzer, t0: datetime.datetime, on_block: Optional[Callable] = None,
           max_blocks: Optional[int] = None, pinned: bool = True) -> int:
    """Run every full block of the recordings through `analyzer` (a `BatchAnalyzer` with one stream per file).

    `on_block(block_index, per_stream)` receives what `BatchAnalyzer.collect` returns for that block
    (`per_stream[s] = (shadow-filtered Signals, candidates before the filter)`), in block order.  Returns the number of
    blocks processed."""
    if analyzer.n_streams != len(paths):
        raise ValueError("one capture file per analyzer stream")
    reader = CaptureReader(paths, analyzer.block_samples, pinned=pinned, max_blocks=max_blocks)
    dt = datetime.timedelta(seconds=analyzer.block_samples / analyzer.sample_rate)
    pending: List[int] = []

    def drain():
        b = pending.pop(0)
        res = analyzer.collect([t0 + b * dt] * analyzer.n_streams)
        if on_block is not None:
            on_block(b, res)

    done = 0
    try:
        for b, block in reader:
            analyzer.submit(block)
            pending.append(b)
            if len(pending) == 2:                                # two blocks in flight
                drain()
                done += 1
        while pending:
            drain()
            done += 1
    finally:
        reader.close()
    return done

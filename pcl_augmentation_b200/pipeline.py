"""Streaming driver: a sequence of staged batches through several engines whose copies and kernels overlap.

The reference augments one frame after the other and its only scale-out is "start the script several times"
(object_detection/README.md:36).  Here one process per GPU keeps ``depth`` engines, each with its own CUDA stream,
its own device-resident batch and its own host thread (ctypes releases the GIL inside the C ABI calls), so that at
steady state one engine uploads the next batch of scans over PCIe (H2D), one runs the placement / occlusion rounds
and one downloads the augmented clouds (D2H): copy engines and SMs are busy at the same time, and the latency-bound
placement kernels of two engines fill each other's idle SMs.  Results are handed to ``on_result`` in submission
order together with the engine that produced them.
"""
from __future__ import annotations

import queue
import threading
import time

from .engine import Real3DEngine


class ScanPipeline:
    def __init__(self, task, config, db, *, depth=4, exclusive_run=True, tail_fraction=8, **engine_kwargs):
        assert depth >= 1
        self.tail_fraction = tail_fraction
        self._run_lock = threading.Lock() if exclusive_run else None
        engine_kwargs.setdefault('sub_batches', 2)       # measured best when several engines share the GPU
        self.engines = [Real3DEngine(task, config, db, **engine_kwargs) for _ in range(depth)]
        self.depth = depth
        self._buffers = [None] * depth

    def warmup(self, staged):
        """Run one batch through every engine (allocates the pinned result buffers)."""
        for i, eng in enumerate(self.engines):
            eng.load(staged)
            eng.run()
            self._buffers[i] = eng.fetch_raw(self._buffers[i])

    def process(self, staged_batches, on_result=None, trace=None):
        """Stream ``staged_batches`` (iterable of ``Real3DEngine.stage`` results; they may repeat) through the
        engines.  ``on_result(index, engine, buffers)`` is called from the worker thread, in any order, while the
        engine's pinned result buffers are still valid; the return value of the callback (or, without a callback,
        the D2H byte count) is collected per batch and returned as a list in submission order.  ``trace`` (a list)
        switches on per-phase host timestamps (worker, index, load start, load end, run end, fetch end) and makes the
        phases synchronous, for diagnosis only."""
        work = queue.Queue()
        items = list(staged_batches)
        for i, st in enumerate(items):
            work.put((i, st))
        out = [None] * len(items)
        errors = []

        def worker(w):
            eng = self.engines[w]
            while not errors:
                try:
                    i, st = work.get_nowait()
                except queue.Empty:
                    return
                try:
                    t0 = time.perf_counter()
                    eng.load(st)
                    if trace is not None:
                        eng.sync()
                    t1 = time.perf_counter()
                    if self._run_lock is not None:
                        eng.sync()                        # the upload must not hold up the engine that computes
                        # only the busy head of a run is exclusive: once a quarter of the scans is left the next
                        # engine may start, its first rounds fill the SMs this engine's thinning rounds leave idle
                        with self._run_lock:
                            more = eng.run_until(max(1, st['n'] // self.tail_fraction))
                        if more:
                            eng.run_until(0)
                    else:
                        eng.run()
                    if trace is not None:
                        eng.sync()
                    t2 = time.perf_counter()
                    buf = self._buffers[w] = eng.fetch_raw(self._buffers[w])
                    if trace is not None:
                        trace.append((w, i, t0, t1, t2, time.perf_counter()))
                    out[i] = on_result(i, eng, buf) if on_result is not None else buf['out_bytes']
                except BaseException as exc:          # surfaced on the caller's thread
                    errors.append(exc)
                    return

        threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(min(self.depth, len(items)))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

    def augment_stream(self, scan_batches):
        """Lists of ScanInput in, lists of ScanResult out (same order).  Staging (the dataset reader's job) happens on
        the caller's thread while the engines work on earlier batches."""
        staged = [self.engines[0].stage(b) for b in scan_batches]
        return self.process(staged, on_result=lambda i, eng, buf: eng.unpack(buf))

    def close(self):
        for eng in self.engines:
            eng.close()
        self.engines = []

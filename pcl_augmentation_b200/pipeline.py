"""Streaming driver: a sequence of staged batches through several engines whose copies and kernels overlap.

The reference augments one frame after the other and its only scale-out is "start the script several times"
(object_detection/README.md:36).  Here one process per GPU keeps ``depth`` engines, each with its own CUDA stream,
its own device-resident batch and its own host thread (ctypes releases the GIL inside the C ABI calls), so that at
steady state one engine uploads the next batch of scans over PCIe (H2D), one runs the per-scan walker and one
downloads the augmented clouds (D2H): copy engines and SMs are busy at the same time.  An engine's ``load`` and ``run``
only queue work on its stream (the walker needs no host polling); the thread blocks in ``fetch_raw`` alone.
Batches are pulled LAZILY from the iterable the caller hands in (a bounded look-ahead of ``depth`` batches), so a
dataset-sized stream never has more than ``2 * depth`` staged batches page-locked at a time.
"""
from __future__ import annotations

import queue
import threading
import time

from .engine import Real3DEngine


class ScanPipeline:
    def __init__(self, task, config, db, *, depth=4, exclusive_run=None, tail_fraction=8, **engine_kwargs):
        assert depth >= 1
        self.tail_fraction = tail_fraction
        self.engines = [Real3DEngine(task, config, db, **engine_kwargs) for _ in range(depth)]
        staged_rounds = self.engines[0].staged_rounds
        if exclusive_run is None:           # only the staged round kernels poll the device from the host
            exclusive_run = staged_rounds
        self._run_lock = threading.Lock() if exclusive_run else None
        self.depth = depth
        self._buffers = [None] * depth

    def warmup(self, staged):
        """Run one batch through every engine (allocates the pinned result buffers)."""
        for i, eng in enumerate(self.engines):
            eng.load(staged)
            eng.run()
            self._buffers[i] = eng.fetch_raw(self._buffers[i])

    def process(self, staged_batches, on_result=None, trace=None):
        """Stream ``staged_batches`` (any iterable of ``Real3DEngine.stage`` results, consumed lazily; items may
        repeat) through the engines.  ``on_result(index, engine, buffers)`` is called from the worker thread, in any
        order, while the engine's pinned result buffers are still valid; the return value of the callback (or,
        without a callback, the D2H byte count) is collected per batch and returned as a list in submission order.
        ``trace`` (a list) switches on per-phase host timestamps (worker, index, load start, load end, run end, fetch
        end) and makes the phases synchronous, for diagnosis only."""
        work = queue.Queue(maxsize=self.depth)
        out = {}
        errors = []
        done = object()

        def producer():
            try:
                for i, st in enumerate(staged_batches):
                    while not errors:
                        try:
                            work.put((i, st), timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if errors:
                        break
            except BaseException as exc:              # a failing reader / stager surfaces on the caller's thread
                errors.append(exc)
            finally:
                for _ in range(self.depth):           # the workers drain the queue until each has seen its sentinel
                    work.put(done)

        def worker(w):
            eng = self.engines[w]
            while True:
                item = work.get()
                if item is done:
                    return
                if errors:
                    continue                          # drain the queue so the producer can finish
                i, st = item
                try:
                    t0 = time.perf_counter()
                    eng.load(st)
                    if trace is not None:
                        eng.sync()
                    t1 = time.perf_counter()
                    if self._run_lock is not None:
                        eng.sync()                        # the upload must not hold up the engine that computes
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

        threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(self.depth)]
        prod = threading.Thread(target=producer, daemon=True)
        for t in threads:
            t.start()
        prod.start()
        prod.join()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return [out[i] for i in range(len(out))]

    def augment_stream(self, scan_batches):
        """Lists of ScanInput in (any iterable), lists of ScanResult out (same order).  Staging (the dataset reader's
        job: packing into pinned host buffers) runs on the producer thread while the engines work on earlier batches."""
        stage = self.engines[0].stage
        return self.process((stage(b) for b in scan_batches), on_result=lambda i, eng, buf: eng.unpack(buf))

    def close(self):
        for eng in self.engines:
            eng.close()
        self.engines = []

"""ORACLE (test infrastructure) — timed CPU baseline for bench.py.

Runs the oracle detection pipeline (cv2 letterbox restatement-equivalent + cv2.dnn forward of the
reference's .tflite + restated decode / weighted NMS, oracle/pipeline.py) over a bounded sample of
frames on the host cores: one worker process per core, cv2 pinned to one thread per worker.  cv2.dnn
fp32 is a stand-in for TFLite/XNNPACK, which cannot be installed in this image.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_P = None
_FRAMES = None


def _init(det_bytes, model, frame_fn_module, frame_fn_name):
    """Worker start-up (spawned, never forked: OpenCV's thread pool does not survive fork): builds its
    own pipeline and regenerates the seeded synthetic frames locally."""
    global _P, _FRAMES
    import importlib
    import cv2
    cv2.setNumThreads(1)
    from oracle.pipeline import OraclePipeline
    _P = OraclePipeline(det_bytes, model, None, "cv2dnn")
    _FRAMES = getattr(importlib.import_module(frame_fn_module), frame_fn_name)()


def _ready(_):
    return _P is not None


def _work(span):
    lo, hi = span
    n = 0
    for k in range(lo, hi):
        n += len(_P.detect(_FRAMES[k % len(_FRAMES)]))
    return n


class CpuPipeline:
    def __init__(self, det_bytes: bytes, model: str, frame_fn_module: str, frame_fn_name: str, workers: int = 0):
        self.workers = workers or len(os.sched_getaffinity(0)) or 1
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.workers, initializer=_init, initargs=(det_bytes, model, frame_fn_module, frame_fn_name))
        self.pool.map(_ready, range(self.workers * 2))

    def run(self, count: int):
        """Processes `count` frames (cycling over the sample); returns (seconds, faces)."""
        per = max(1, (count + self.workers * 4 - 1) // (self.workers * 4))
        spans = [(i, min(count, i + per)) for i in range(0, count, per)]
        t = time.perf_counter()
        faces = sum(self.pool.map(_work, spans))
        return time.perf_counter() - t, faces

    def close(self):
        self.pool.close()
        self.pool.join()

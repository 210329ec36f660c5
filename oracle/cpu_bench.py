"""ORACLE (test infrastructure) — timed CPU baseline for bench.py.

The reference's CPU pipeline as BASELINE.md section 3 states it, with the real OpenCV calls the reference makes
(lib/src/util/helpers.dart:303-421, :583-625) instead of their numpy restatements:
    cv2.resize(INTER_LINEAR) -> cv2.copyMakeBorder(BORDER_CONSTANT) -> BGR2RGB + convertTo(1/127.5, -1)
    (cv2.dnn.blobFromImage does the last two in one C++ pass) -> cv2.dnn forward of the reference's .tflite
    -> vectorised `raw >= 0` candidate scan -> restated decode / weighted NMS on the handful of candidates
    [-> cv2.getRotationMatrix2D + cv2.warpAffine 192x192 -> cv2.dnn forward of face_landmark -> unpack]
over a bounded sample of frames on the host cores: one worker process per core, cv2 pinned to one thread per
worker.  cv2.dnn fp32 is a stand-in for TFLite/XNNPACK, which cannot be installed in this image.  tests/
check that this fast path returns the same detections as the restated pipeline (oracle/pipeline.py).
"""
from __future__ import annotations

import math
import multiprocessing as mp
import os
import time

import numpy as np

_P = None
_FRAMES = None


class FastCpuPipeline:
    """One worker's pipeline.  mode: "fast" (detector only) or "standard" (+ warp + mesh)."""

    def __init__(self, det_bytes: bytes, model: str, mesh_bytes=None):
        import cv2
        from . import cv_ops, detect_post as dp, tflite_reader as tr
        self.cv2, self.dp, self.cv_ops = cv2, dp, cv_ops
        m = tr.read_tflite(det_bytes)
        self.S = m.tensors[m.inputs[0]].shape[1]
        self.names = [m.tensors[i].name for i in m.outputs]
        self.net = cv2.dnn.readNetFromTFLite(np.frombuffer(det_bytes, np.uint8))
        self.anchors = dp.generate_anchors(dp.ssd_options_for(model))
        self.mesh = None
        if mesh_bytes is not None:
            mm = tr.read_tflite(mesh_bytes)
            self.mesh_names = [mm.tensors[i].name for i in mm.outputs]
            self.mesh_sizes = [int(np.prod(mm.tensors[i].shape)) for i in mm.outputs]
            self.mesh = cv2.dnn.readNetFromTFLite(np.frombuffer(mesh_bytes, np.uint8))

    def detect(self, frame):
        cv2, dp, S = self.cv2, self.dp, self.S
        h, w = frame.shape[:2]
        lp = self.cv_ops.compute_letterbox_params(w, h, S, S)
        img = frame if (lp.new_w == w and lp.new_h == h) else cv2.resize(frame, (lp.new_w, lp.new_h), interpolation=cv2.INTER_LINEAR)
        if lp.pad_top or lp.pad_bottom or lp.pad_left or lp.pad_right:
            img = cv2.copyMakeBorder(img, lp.pad_top, lp.pad_bottom, lp.pad_left, lp.pad_right, cv2.BORDER_CONSTANT, value=(0, 0, 0))
        blob = cv2.dnn.blobFromImage(img, 1.0 / 127.5, (S, S), (127.5, 127.5, 127.5), swapRB=True, crop=False)
        self.net.setInput(blob)
        boxes, scores = self.net.forward(self.names)
        scores = np.asarray(scores, np.float32).reshape(-1)
        boxes = np.asarray(boxes, np.float32).reshape(-1, 16)
        idx = np.nonzero(scores >= 0.0)[0]                       # _collectCandidateScores, vectorised
        if idx.size == 0:
            return []
        idx = [int(i) for i in idx]
        sc = [dp.sigmoid_clipped(float(scores[i])) for i in idx]
        dets = dp.to_detections_filtered(dp.decode_boxes(boxes, self.anchors, idx, S), sc, idx)
        pad = (lp.pad_top / S, lp.pad_bottom / S, lp.pad_left / S, lp.pad_right / S)
        return dp.letterbox_removal(dp.weighted_nms(dets), pad)

    def detect_standard(self, frame, min_presence: float = 0.5):
        """detect + extractAlignedSquare (real cv2.getRotationMatrix2D / cv2.warpAffine) + face_landmark + unpack."""
        from . import geometry as geo
        cv2 = self.cv2
        h, w = frame.shape[:2]
        out = []
        for d in self.detect(frame):
            theta, cx, cy, size = geo.compute_face_alignment(d.kp, float(w), float(h))
            si = self.cv_ops.dart_round(size)
            if si <= 0:
                continue
            sc = 192.0 / si
            M = cv2.getRotationMatrix2D((cx, cy), theta * 180.0 / math.pi, sc)     # extractAlignedSquare(..., -theta): angle = +theta
            M[0, 2] += 96.0 + 0.5 * (sc - 1.0) - cx
            M[1, 2] += 96.0 + 0.5 * (sc - 1.0) - cy
            crop = cv2.warpAffine(frame, M, (192, 192), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
            blob = cv2.dnn.blobFromImage(crop, 1.0 / 127.5, (192, 192), (127.5, 127.5, 127.5), swapRB=True, crop=False)
            self.mesh.setInput(blob)
            o = self.mesh.forward(self.mesh_names)
            li = max((i for i in range(len(o)) if self.mesh_sizes[i] % 3 == 0), key=lambda i: self.mesh_sizes[i])
            fi = next(i for i in range(len(o)) if self.mesh_sizes[i] == 1)
            score = self.dp.sigmoid_clipped(float(np.asarray(o[fi]).reshape(-1)[0]))
            if min_presence > 0 and score < min_presence:
                continue
            raw = np.asarray(o[li], np.float64).reshape(-1, 3) / 192.0            # _unpackLandmarks (pad 0), vectorised
            raw[:, :2] = np.clip(raw[:, :2], 0.0, 1.0)
            out.append((d, geo.transform_mesh_to_absolute(raw, cx, cy, size, theta), score))
        return out


def _init(det_bytes, model, mesh_bytes, frame_fn_module, frame_fn_name, frame_fn_arg):
    """Worker start-up (spawned, never forked: OpenCV's thread pool does not survive fork): builds its
    own pipeline and regenerates the seeded synthetic frames locally."""
    global _P, _FRAMES
    import importlib
    import cv2
    cv2.setNumThreads(1)
    _P = FastCpuPipeline(det_bytes, model, mesh_bytes)
    _FRAMES = getattr(importlib.import_module(frame_fn_module), frame_fn_name)(frame_fn_arg)


def _ready(_):
    return _P is not None


def _work(span):
    lo, hi, standard = span
    n = 0
    for k in range(lo, hi):
        f = _FRAMES[k % len(_FRAMES)]
        n += len(_P.detect_standard(f) if standard else _P.detect(f))
    return n


class CpuPipeline:
    def __init__(self, det_bytes: bytes, model: str, frame_fn_module: str, frame_fn_name: str, frame_fn_arg=None,
                 mesh_bytes=None, workers: int = 0):
        self.workers = workers or len(os.sched_getaffinity(0)) or 1
        self.standard = mesh_bytes is not None
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.workers, initializer=_init,
                             initargs=(det_bytes, model, mesh_bytes, frame_fn_module, frame_fn_name, frame_fn_arg))
        self.pool.map(_ready, range(self.workers * 2))

    def run(self, count: int):
        """Processes `count` frames (cycling over the sample); returns (seconds, faces)."""
        per = max(1, (count + self.workers * 4 - 1) // (self.workers * 4))
        spans = [(i, min(count, i + per), self.standard) for i in range(0, count, per)]
        t = time.perf_counter()
        faces = sum(self.pool.map(_work, spans))
        return time.perf_counter() - t, faces

    def close(self):
        self.pool.close()
        self.pool.join()

"""ORACLE (test infrastructure, not product code).

End-to-end CPU restatement of the reference pipeline for one frame:
_FaceDetectorCore.detectFacesDirect (lib/src/isolate/face_detector_core.dart:215-394) in
`fast` (detector only) and `standard` (+ aligned crop + 468-point mesh) modes; iris /
blendshapes / embeddings are out of scope.

Inference back ends (both restate flutter_litert's Interpreter.invoke, un-vendored):
  * "f64" / "f32": oracle.graph_exec (torch CPU) — the parity reference;
  * "cv2dnn": cv2.dnn.readNetFromTFLite forward — an independent fp32 implementation, used as
    the timed CPU baseline in bench.py (a stand-in for TFLite/XNNPACK, which is not installable).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import cv_ops, detect_post as dp, geometry as geo, graph_exec, tflite_reader as tr

MESH_INPUT = 192
K_MIN_FACE_PRESENCE = 0.5   # face_model_config.dart:62


@dataclass
class FaceResult:
    det: dp.Detection
    mesh_px: Optional[np.ndarray] = None       # [468,3] absolute pixels (x,y,z)
    mesh_score: Optional[float] = None
    align: Optional[tuple] = None              # (theta,cx,cy,size)
    crop: Optional[np.ndarray] = None          # u8 [192,192,3] BGR
    mesh_raw: Optional[np.ndarray] = None      # f32 [1404] model output


class Net:
    """One TFLite model with a selectable CPU back end."""

    def __init__(self, tflite_bytes: bytes, backend: str = "f64"):
        import torch
        self.model = tr.read_tflite(tflite_bytes)
        self.backend = backend
        self.in_shape = self.model.tensors[self.model.inputs[0]].shape
        if backend == "cv2dnn":
            import cv2
            self.net = cv2.dnn.readNetFromTFLiteFromBuffer(bytes(tflite_bytes)) \
                if hasattr(cv2.dnn, "readNetFromTFLiteFromBuffer") else None
            if self.net is None:
                self.net = cv2.dnn.readNetFromTFLite(np.frombuffer(tflite_bytes, np.uint8))
            self.names = [self.model.tensors[i].name for i in self.model.outputs]
        else:
            self.exe = graph_exec.GraphExecutor(
                self.model, {"f64": torch.float64, "f32": torch.float32}[backend])

    def run(self, x_nhwc: np.ndarray) -> List[np.ndarray]:
        """x [B,H,W,3] f32 -> list of outputs in graph output order, each [B, ...] f32."""
        if self.backend == "cv2dnn":
            outs = [[] for _ in self.names]
            for b in range(x_nhwc.shape[0]):
                self.net.setInput(np.ascontiguousarray(x_nhwc[b:b + 1].transpose(0, 3, 1, 2)))
                r = self.net.forward(self.names)
                for k, o in enumerate(r):
                    outs[k].append(np.asarray(o, np.float32).reshape(-1))
            return [np.stack(o) for o in outs]
        res = self.exe.run(x_nhwc)
        return [res[i].astype(np.float32).reshape(x_nhwc.shape[0], -1) for i in self.model.outputs]


class OraclePipeline:
    def __init__(self, det_bytes: bytes, model: str = "shortRange",
                 mesh_bytes: Optional[bytes] = None, backend: str = "f64",
                 min_score: float = 0.0, min_face_size: float = 0.0,
                 min_face_presence: float = K_MIN_FACE_PRESENCE):
        self.det = Net(det_bytes, backend)
        self.opts = dp.ssd_options_for(model)
        self.anchors = dp.generate_anchors(self.opts)
        self.in_h, self.in_w = self.det.in_shape[1], self.det.in_shape[2]
        self.mesh = Net(mesh_bytes, backend) if mesh_bytes is not None else None
        self.min_score, self.min_face_size, self.min_presence = min_score, min_face_size, min_face_presence

    # -- stages -------------------------------------------------------------------------------
    def preprocess(self, frame_bgr: np.ndarray):
        return cv_ops.convert_image_to_tensor(frame_bgr, self.in_w, self.in_h)

    def raw_heads(self, tensor: np.ndarray):
        """-> (boxes [N,16] f32, scores [N] f32); output 0 = boxes, 1 = scores
        (face_detection_model.dart:19)."""
        o = self.det.run(tensor[None])
        return o[0][0].reshape(-1, 16), o[1][0].reshape(-1)

    def detect(self, frame_bgr: np.ndarray) -> List[dp.Detection]:
        tensor, pad, _ = self.preprocess(frame_bgr)
        boxes, scores = self.raw_heads(tensor)
        dets = dp.postprocess(boxes, scores, self.anchors, self.in_h, pad)
        return dp.apply_detection_gates(dets, self.min_score, self.min_face_size,
                                        float(frame_bgr.shape[1]))

    def detect_faces(self, frame_bgr: np.ndarray, mode: str = "fast") -> List[FaceResult]:
        h, w = frame_bgr.shape[:2]
        out = []
        for d in self.detect(frame_bgr):
            theta, cx, cy, size = geo.compute_face_alignment(d.kp, float(w), float(h))
            if mode == "fast":
                if cv_ops.dart_round(size) > 0:           # face_detector_core.dart:255-265
                    out.append(FaceResult(d, align=(theta, cx, cy, size)))
                continue
            crop = cv_ops.extract_aligned_square(frame_bgr, cx, cy, size, -theta, MESH_INPUT)
            if crop is None:                               # face dropped (face_detector_core.dart:266-268)
                continue
            t = cv_ops.normalize_bgr_u8(crop)              # no resize / pad: crop is already 192^2
            o = self.mesh.run(t[None])
            sizes = [x.shape[1] for x in o]
            # face_landmark.dart:154-166: landmarks = largest output divisible by 3, score = first 1-element output
            li = max((i for i in range(len(o)) if sizes[i] % 3 == 0), key=lambda i: sizes[i])
            si = next((i for i in range(len(o)) if sizes[i] == 1), -1)
            lm = geo.unpack_landmarks(o[li][0], MESH_INPUT, MESH_INPUT, (0.0, 0.0, 0.0, 0.0),
                                      clamp=True, normalize_z=True)
            score = dp.sigmoid_clipped(float(o[si][0][0])) if si >= 0 else None
            if score is not None and self.min_presence > 0 and score < self.min_presence:
                continue                                   # presence gate (face_detector_core.dart:353)
            out.append(FaceResult(d, geo.transform_mesh_to_absolute(lm, cx, cy, size, theta), score,
                                  (theta, cx, cy, size), crop, o[li][0].copy()))
        return out

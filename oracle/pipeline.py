"""ORACLE (test infrastructure, not product code).

End-to-end CPU restatement of the reference pipeline for one frame:
_FaceDetectorCore.detectFacesDirect (lib/src/isolate/face_detector_core.dart:215-394) in
`fast` (detector only), `standard` (+ aligned crop + 468-point mesh) and `full` (+ two eye crops,
iris_landmark, iris-refined eye keypoints; _irisFromMesh :532-596, :356-373) modes; the blendshape
classifier is out of scope.

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
IRIS_INPUT = 64
K_MIN_FACE_PRESENCE = 0.5   # face_model_config.dart:62


@dataclass
class FaceResult:
    det: dp.Detection
    mesh_px: Optional[np.ndarray] = None       # [468,3] absolute pixels (x,y,z)
    mesh_score: Optional[float] = None
    align: Optional[tuple] = None              # (theta,cx,cy,size)
    crop: Optional[np.ndarray] = None          # u8 [192,192,3] BGR
    mesh_raw: Optional[np.ndarray] = None      # f32 [1404] model output
    iris_px: Optional[np.ndarray] = None       # [152,3] absolute pixels x,y + raw z (76 left then 76 right), full mode
    eye_rois: Optional[list] = None            # [(cx,cy,size,theta)] x 2
    eye_crops: Optional[list] = None           # u8 [64,64,3] BGR x 2 (right one mirrored)
    iris_raw: Optional[list] = None            # per eye: (contours f32[213], iris f32[15])


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
                 min_face_presence: float = K_MIN_FACE_PRESENCE, iris_bytes: Optional[bytes] = None):
        self.det = Net(det_bytes, backend)
        self.opts = dp.ssd_options_for(model)
        self.anchors = dp.generate_anchors(self.opts)
        self.in_h, self.in_w = self.det.in_shape[1], self.det.in_shape[2]
        self.mesh = Net(mesh_bytes, backend) if mesh_bytes is not None else None
        self.iris = Net(iris_bytes, backend) if iris_bytes is not None else None
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
            res = FaceResult(d, geo.transform_mesh_to_absolute(lm, cx, cy, size, theta), score,
                             (theta, cx, cy, size), crop, o[li][0].copy())
            if mode == "full" and self.iris is not None:
                self._iris_from_mesh(frame_bgr, res)
            out.append(res)
        return out

    def _iris_from_mesh(self, frame_bgr: np.ndarray, res: FaceResult) -> None:
        """_irisFromMesh (face_detector_core.dart:532-596) + the eye-keypoint refinement (:356-373)."""
        import dataclasses
        h, w = frame_bgr.shape[:2]
        rois = geo.eye_rois_from_mesh(res.mesh_px)
        crops = [cv_ops.extract_aligned_square(frame_bgr, r[0], r[1], r[2], r[3], IRIS_INPUT) for r in rois]
        if crops[0] is None or crops[1] is None:
            return
        crops[1] = np.ascontiguousarray(crops[1][:, ::-1])            # cv.flip(rightCropRaw, 1)
        pts, raws = [], []
        for e in range(2):
            o = self.iris.run(cv_ops.normalize_bgr_u8(crops[e])[None])
            lm = []
            for out in o:                                              # every output in order, clamp: false, z untouched
                lm += list(geo.unpack_landmarks(out[0], IRIS_INPUT, IRIS_INPUT, (0.0, 0.0, 0.0, 0.0), clamp=False))
            raws.append(tuple(x[0].copy() for x in o))
            pts.append(geo.transform_iris_norm_to_absolute(lm, rois[e], e == 1))
        res.iris_px = np.concatenate(pts)
        res.eye_rois, res.eye_crops, res.iris_raw = rois, crops, raws
        kp = list(res.det.kp)
        lc = geo.iris_center_from_points(res.iris_px[71:76])
        rc = geo.iris_center_from_points(res.iris_px[147:152])
        kp[0], kp[1] = lc[0] / w, lc[1] / h
        kp[2], kp[3] = rc[0] / w, rc[1] / h
        res.det = dataclasses.replace(res.det, kp=kp)

"""ORACLE (test infrastructure, not product code).

Restatement of the detector post-processing with the reference's mixed f32/f64 arithmetic:
  * SSD anchors           flutter_litert generateAnchors (un-vendored) with the option sets of
                          lib/src/shared/face_model_config.dart:80-125
  * candidate collection  lib/src/models/face_detection_model.dart:473-492
  * box/keypoint decode   lib/src/models/face_detection_model.dart:431-467
  * degenerate filter     lib/src/models/face_detection_model.dart:498-516
  * weighted NMS          lib/src/util/helpers.dart:183-221 -> flutter_litert weightedNms (un-vendored)
  * letterbox removal     lib/src/util/helpers.dart:101-136
  * detection gates       lib/src/shared/face_gates.dart:115-146
weightedNms follows MediaPipe's WeightedNonMaxSuppression (strict IoU > thr against the top
box, score-weighted box average, score and keypoints of the top detection); its exact source
is not in the reference tree -> PARITY UNPINNED beyond the reference's own NMS tests
(test/helpers_coverage_test.dart:220-291), which tests/test_oracle_postprocess.py reproduces.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

K_MIN_SCORE = 0.5                 # face_model_config.dart:53
K_MIN_SUPPRESSION = 0.3           # face_model_config.dart:77
K_RAW_SCORE_LIMIT = 100.0         # face_model_config.dart:49
K_MAX_DETECTIONS = 100            # helpers.dart:187


@dataclass(frozen=True)
class SSDAnchorOptions:
    num_layers: int
    input_h: int
    input_w: int
    strides: Sequence[int]
    interpolated_scale_aspect_ratio: float
    aspect_ratios: Sequence[float] = (1.0,)
    offset_x: float = 0.5
    offset_y: float = 0.5


SSD_FRONT = SSDAnchorOptions(4, 128, 128, (8, 16, 16, 16), 1.0)     # face_model_config.dart:80-93
SSD_BACK = SSDAnchorOptions(4, 256, 256, (16, 32, 32, 32), 1.0)     # :96-109
SSD_FULL = SSDAnchorOptions(1, 192, 192, (4,), 0.0)                 # :112-125


def ssd_options_for(model: str) -> SSDAnchorOptions:
    """ssdOptionsFor (face_model_config.dart:128-134)."""
    return {"frontCamera": SSD_FRONT, "shortRange": SSD_FRONT, "backCamera": SSD_BACK,
            "full": SSD_FULL, "fullSparse": SSD_FULL}[model]


def generate_anchors(o: SSDAnchorOptions) -> np.ndarray:
    """[N,2] float64 anchor centres (cx, cy); layers of equal stride are merged, each
    contributes len(aspect_ratios) (+1 if interpolated_scale_aspect_ratio > 0) anchors per cell."""
    out = []
    layer = 0
    while layer < o.num_layers:
        last = layer
        repeats = 0
        while last < o.num_layers and o.strides[last] == o.strides[layer]:
            repeats += len(o.aspect_ratios) + (1 if o.interpolated_scale_aspect_ratio > 0 else 0)
            last += 1
        stride = o.strides[layer]
        fh = -(-o.input_h // stride)
        fw = -(-o.input_w // stride)
        for y in range(fh):
            cy = (y + o.offset_y) / fh
            for x in range(fw):
                cx = (x + o.offset_x) / fw
                out += [(cx, cy)] * repeats
        layer = last
    return np.array(out, np.float64)


def sigmoid_clipped(x: float, limit: float = K_RAW_SCORE_LIMIT) -> float:
    """flutter_litert sigmoidClipped: clip to +-limit then 1/(1+exp(-x)) in f64."""
    x = min(max(float(x), -limit), limit)
    return 1.0 / (1.0 + math.exp(-x))


@dataclass
class Detection:
    xmin: float
    ymin: float
    xmax: float
    ymax: float
    score: float
    kp: List[float] = field(default_factory=list)   # 12 values: x0,y0,...,x5,y5
    anchor: int = -1

    def as_row(self):
        return [self.xmin, self.ymin, self.xmax, self.ymax, self.score] + list(self.kp)


def raw_score_threshold(min_score: float = K_MIN_SCORE) -> float:
    return math.log(min_score / (1.0 - min_score))   # face_detection_model.dart:473-475


def collect_candidates(raw_scores: np.ndarray):
    """Ascending anchor indices with raw >= logit(0.5) = 0.0 (NaN rejected) and their scores."""
    thr = raw_score_threshold()
    raw = np.asarray(raw_scores, np.float32).reshape(-1)
    idx = [int(i) for i in range(raw.shape[0]) if float(raw[i]) >= thr]
    return idx, [sigmoid_clipped(float(raw[i])) for i in idx]


def decode_boxes(raw_boxes: np.ndarray, anchors: np.ndarray, indices, input_h: int):
    """f32 storage of every intermediate (`tmp` is a Float32List), f64 arithmetic per operation."""
    raw = np.asarray(raw_boxes, np.float32).reshape(-1, 16)
    scale = float(input_h)
    out = []
    for i in indices:
        tmp = np.empty(16, np.float32)
        for j in range(16):
            tmp[j] = np.float32(float(raw[i, j]) / scale)
        ax, ay = float(anchors[i, 0]), float(anchors[i, 1])
        tmp[0] = np.float32(float(tmp[0]) + ax)
        tmp[1] = np.float32(float(tmp[1]) + ay)
        for j in range(4, 16, 2):
            tmp[j] = np.float32(float(tmp[j]) + ax)
            tmp[j + 1] = np.float32(float(tmp[j + 1]) + ay)
        xc, yc, w, h = (float(tmp[0]), float(tmp[1]), float(tmp[2]), float(tmp[3]))
        out.append((xc - w * 0.5, yc - h * 0.5, xc + w * 0.5, yc + h * 0.5,
                    [float(v) for v in tmp[4:16]]))
    return out


def to_detections_filtered(boxes, scores, indices) -> List[Detection]:
    res = []
    for b, s, i in zip(boxes, scores, indices):
        if b[2] <= b[0] or b[3] <= b[1]:
            continue
        res.append(Detection(b[0], b[1], b[2], b[3], s, list(b[4]), i))
    return res


def iou(a, b) -> float:
    iw = min(a[2], b[2]) - max(a[0], b[0])
    ih = min(a[3], b[3]) - max(a[1], b[1])
    if iw <= 0 or ih <= 0:
        return 0.0
    inter = iw * ih
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union > 0 else 0.0


def weighted_nms(dets: List[Detection], iou_thresh: float = K_MIN_SUPPRESSION,
                 score_thresh: float = K_MIN_SCORE, max_det: int = K_MAX_DETECTIONS) -> List[Detection]:
    # helpers.dart:189-191 (Dart's sort is unstable; ties are broken here by input order,
    # i.e. ascending anchor index, and the CUDA path does the same)
    f = [d for d in dets if d.score >= score_thresh]
    f.sort(key=lambda d: -d.score)
    out = []
    remaining = list(range(len(f)))
    while remaining and len(out) < max_det:
        top = f[remaining[0]]
        tb = (top.xmin, top.ymin, top.xmax, top.ymax)
        cluster, rest = [], []
        for j in remaining:
            d = f[j]
            (cluster if iou((d.xmin, d.ymin, d.xmax, d.ymax), tb) > iou_thresh else rest).append(j)
        if not cluster:          # cannot happen for a non-degenerate top box (IoU with itself is 1)
            cluster, rest = [remaining[0]], remaining[1:]
        tot = 0.0
        acc = [0.0, 0.0, 0.0, 0.0]
        for j in cluster:
            d = f[j]
            tot += d.score
            acc[0] += d.xmin * d.score
            acc[1] += d.ymin * d.score
            acc[2] += d.xmax * d.score
            acc[3] += d.ymax * d.score
        out.append(Detection(acc[0] / tot, acc[1] / tot, acc[2] / tot, acc[3] / tot,
                             top.score, list(top.kp), top.anchor))
        remaining = rest
    return out


def letterbox_removal(dets: List[Detection], padding) -> List[Detection]:
    pt, pb, pl, pr = padding
    sx = 1.0 - (pl + pr)
    sy = 1.0 - (pt + pb)
    out = []
    for d in dets:
        kp = list(d.kp)
        for i in range(0, len(kp), 2):
            kp[i] = (kp[i] - pl) / sx
            kp[i + 1] = (kp[i + 1] - pt) / sy
        out.append(Detection((d.xmin - pl) / sx, (d.ymin - pt) / sy, (d.xmax - pl) / sx,
                             (d.ymax - pt) / sy, d.score, kp, d.anchor))
    return out


def box_visible_width_fraction(d: Detection, image_width: float) -> float:
    if image_width <= 0:
        return 0.0
    left = d.xmin * image_width
    right = d.xmax * image_width
    vis = min(right, image_width) - max(left, 0.0)
    return vis / image_width if vis > 0 else 0.0


def apply_detection_gates(dets, min_score: float, min_face_size: float, image_width: float):
    if min_score <= 0.0 and min_face_size <= 0.0:
        return dets
    return [d for d in dets if d.score >= min_score and
            (min_face_size <= 0.0 or box_visible_width_fraction(d, image_width) >= min_face_size)]


def postprocess(raw_boxes, raw_scores, anchors, input_h, padding) -> List[Detection]:
    """FaceDetection._postprocess (face_detection_model.dart:401-425)."""
    idx, sc = collect_candidates(raw_scores)
    boxes = decode_boxes(raw_boxes, anchors, idx, input_h)
    dets = to_detections_filtered(boxes, sc, idx)
    return letterbox_removal(weighted_nms(dets), padding)

"""Result gates, mirroring lib/src/shared/face_gates.dart: validateFaceGates :31-59, applyFaceGates :84-104,
boxVisibleWidthFraction :115-121, applyDetectionGates :130-146.  The detector applies the detection gates on the
device (k_decode_nms) and the presence gate in the C-ABI layer; these host functions are the reference's own
late-gate entry points for callers that re-filter a result list."""
from __future__ import annotations

import math
from typing import List


def validateFaceGates(*, minScore: float, minFaceSize: float, minFacePresenceConfidence: float = 0.0) -> None:
    for name, v in (("minScore", minScore), ("minFaceSize", minFaceSize), ("minFacePresenceConfidence", minFacePresenceConfidence)):
        if math.isnan(v) or v < 0.0 or v > 1.0:
            raise ValueError("%s must be in the inclusive range [0.0, 1.0]" % name)     # Dart ArgumentError.value


def boxVisibleWidthFraction(box, imageWidth: float) -> float:
    if imageWidth <= 0:
        return 0.0
    left, right = box.xmin * imageWidth, box.xmax * imageWidth
    visible = min(right, imageWidth) - max(left, 0.0)
    return visible / imageWidth if visible > 0 else 0.0


def applyFaceGates(faces: List, *, minScore: float, minFaceSize: float, minFacePresenceConfidence: float = 0.0) -> List:
    if minScore <= 0.0 and minFaceSize <= 0.0 and minFacePresenceConfidence <= 0.0:
        return faces
    return [f for f in faces
            if f.score >= minScore and f.widthFraction >= minFaceSize and
            (minFacePresenceConfidence <= 0.0 or
             (f.meshScore if f.meshScore is not None else float("inf")) >= minFacePresenceConfidence)]


def applyDetectionGates(detections: List, *, minScore: float, minFaceSize: float, imageWidth: float) -> List:
    if minScore <= 0.0 and minFaceSize <= 0.0:
        return detections
    return [d for d in detections
            if d.score >= minScore and (minFaceSize <= 0.0 or boxVisibleWidthFraction(d.boundingBox, imageWidth) >= minFaceSize)]

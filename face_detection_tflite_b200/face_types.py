"""Result and configuration types of the drop-in API, mirroring the reference's Dart types
(lib/src/shared/face_types.dart): Face :1070, Detection :1483, RectF :1439, FaceMesh :749,
FaceLandmarkType :19-37, FaceDetectionModel :100-115, FaceDetectionMode :118-127, and
flutter_litert's Point / BoundingBox.  Only what the detection hot path produces is mirrored."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


class FaceLandmarkType(enum.IntEnum):
    leftEye = 0
    rightEye = 1
    noseTip = 2
    mouth = 3
    leftEyeTragion = 4
    rightEyeTragion = 5


class FaceDetectionModel(enum.IntEnum):
    frontCamera = 0
    backCamera = 1
    shortRange = 2
    full = 3
    fullSparse = 4


class FaceDetectionMode(enum.IntEnum):
    fast = 0
    standard = 1
    full = 2


# faceDetectionModelFile (lib/src/shared/face_model_config.dart:137-143); front == short range
# (byte-identical files)
MODEL_FILES = {
    FaceDetectionModel.frontCamera: "face_detection_short_range.tflite",
    FaceDetectionModel.backCamera: "face_detection_back.tflite",
    FaceDetectionModel.shortRange: "face_detection_short_range.tflite",
    FaceDetectionModel.full: "face_detection_full_range.tflite",
    FaceDetectionModel.fullSparse: "face_detection_full_range_sparse.tflite",
}
MESH_MODEL_FILE = "face_landmark.tflite"


@dataclass(frozen=True)
class Point:
    x: float
    y: float
    z: Optional[float] = None


@dataclass(frozen=True)
class Size:
    width: float
    height: float


@dataclass(frozen=True)
class RectF:
    xmin: float
    ymin: float
    xmax: float
    ymax: float

    @property
    def w(self) -> float:
        return self.xmax - self.xmin

    @property
    def h(self) -> float:
        return self.ymax - self.ymin


@dataclass(frozen=True)
class BoundingBox:
    topLeft: Point
    topRight: Point
    bottomRight: Point
    bottomLeft: Point

    @property
    def width(self) -> float:
        return self.topRight.x - self.topLeft.x

    @property
    def height(self) -> float:
        return self.bottomLeft.y - self.topLeft.y

    @property
    def center(self) -> Point:
        return Point((self.topLeft.x + self.bottomRight.x) / 2.0, (self.topLeft.y + self.bottomRight.y) / 2.0)

    @property
    def corners(self) -> List[Point]:
        return [self.topLeft, self.topRight, self.bottomRight, self.bottomLeft]


@dataclass
class Detection:
    boundingBox: RectF
    score: float
    keypointsXY: List[float]
    imageSize: Optional[Size] = None

    def __getitem__(self, i: int) -> float:
        return self.keypointsXY[i]

    @property
    def landmarks(self) -> Dict[FaceLandmarkType, Point]:
        if self.imageSize is None:
            raise RuntimeError("Detection.imageSize is null; cannot produce pixel landmarks.")
        w, h = float(self.imageSize.width), float(self.imageSize.height)
        return {t: Point(self.keypointsXY[t * 2] * w, self.keypointsXY[t * 2 + 1] * h) for t in FaceLandmarkType}


@dataclass
class FaceMesh:
    """468 points in absolute pixels, packed float32 [468,3] like FaceMesh.packed (face_types.dart:773-809)."""
    packed: np.ndarray
    score: Optional[float] = None

    @property
    def points(self) -> List[Point]:
        return [Point(float(p[0]), float(p[1]), float(p[2])) for p in self.packed]

    def __len__(self) -> int:
        return int(self.packed.shape[0])


@dataclass
class Face:
    detectionData: Detection
    mesh: Optional[FaceMesh]
    originalSize: Size
    irisPoints: List[Point] = field(default_factory=list)
    trackingId: Optional[int] = None
    anchorIndex: int = -1   # parity aid: SSD anchor of the cluster's top detection

    @property
    def boundingBox(self) -> BoundingBox:
        r, w, h = self.detectionData.boundingBox, float(self.originalSize.width), float(self.originalSize.height)
        return BoundingBox(Point(r.xmin * w, r.ymin * h), Point(r.xmax * w, r.ymin * h),
                           Point(r.xmax * w, r.ymax * h), Point(r.xmin * w, r.ymax * h))

    @property
    def landmarks(self) -> Dict[FaceLandmarkType, Point]:
        return self.detectionData.landmarks

    @property
    def score(self) -> float:
        return self.detectionData.score

    @property
    def meshScore(self) -> Optional[float]:
        return self.mesh.score if self.mesh is not None else None

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
IRIS_MODEL_FILE = "iris_landmark.tflite"


@dataclass(frozen=True)
class Point:
    x: float
    y: float
    z: Optional[float] = None

    def toMap(self) -> dict:
        return {"x": self.x, "y": self.y, "z": self.z} if self.z is not None else {"x": self.x, "y": self.y}

    @staticmethod
    def fromMap(m: dict) -> "Point":
        z = m.get("z")
        return Point(float(m["x"]), float(m["y"]), None if z is None else float(z))


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

    def toMap(self) -> dict:          # face_types.dart:1465-1470
        return {"xmin": self.xmin, "ymin": self.ymin, "xmax": self.xmax, "ymax": self.ymax}

    @staticmethod
    def fromMap(m: dict) -> "RectF":
        return RectF(float(m["xmin"]), float(m["ymin"]), float(m["xmax"]), float(m["ymax"]))


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

    def toMap(self) -> dict:          # face_types.dart:1525-1531
        m = {"boundingBox": self.boundingBox.toMap(), "score": self.score, "keypointsXY": list(self.keypointsXY)}
        if self.imageSize is not None:
            m["imageSize"] = {"width": self.imageSize.width, "height": self.imageSize.height}
        return m

    @staticmethod
    def fromMap(m: dict) -> "Detection":   # :1534-1541
        sz = m.get("imageSize")
        return Detection(RectF.fromMap(m["boundingBox"]), float(m["score"]), [float(v) for v in m["keypointsXY"]],
                         Size(sz["width"], sz["height"]) if sz is not None else None)


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

    def __getitem__(self, i: int) -> Point:
        p = self.packed[i]
        return Point(float(p[0]), float(p[1]), float(p[2]))

    def toMap(self) -> dict:          # face_types.dart:821-824
        m = {"points": [p.toMap() for p in self.points]}
        if self.score is not None:
            m["score"] = self.score
        return m

    @staticmethod
    def fromMap(m: dict) -> "FaceMesh":    # :827-830
        pts = [Point.fromMap(p) for p in m["points"]]
        sc = m.get("score")
        return FaceMesh(np.array([[p.x, p.y, 0.0 if p.z is None else p.z] for p in pts], np.float32).reshape(-1, 3),
                        None if sc is None else float(sc))


def irisCenterFromPoints(pts: List[Point]) -> Point:
    """The iris point closest to the centroid (face_types.dart:976-997)."""
    if not pts:
        return Point(0, 0, 0)
    if len(pts) == 1:
        return pts[0]
    cx = cy = 0.0
    for p in pts:
        cx += p.x
        cy += p.y
    cx /= len(pts)
    cy /= len(pts)
    best, best_d = 0, float("inf")
    for i, p in enumerate(pts):
        dx, dy = p.x - cx, p.y - cy
        d = dx * dx + dy * dy
        if d < best_d:
            best_d, best = d, i
    return pts[best]


@dataclass
class Eye:
    """Iris centre + 4 iris contour points + 71-point eye mesh (face_types.dart:852-895)."""
    irisCenter: Point
    irisContour: List[Point]
    mesh: List[Point] = field(default_factory=list)


@dataclass
class EyePair:
    leftEye: Optional[Eye]
    rightEye: Optional[Eye]


@dataclass
class Face:
    detectionData: Detection
    mesh: Optional[FaceMesh]
    originalSize: Size
    irisPoints: List[Point] = field(default_factory=list)
    trackingId: Optional[int] = None
    anchorIndex: int = -1   # parity aid: SSD anchor of the cluster's top detection
    irisPacked: Optional[np.ndarray] = None   # f32 [152,3] as it crosses the C ABI (wire form 'iris', face_detector.dart:1174)

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

    @property
    def widthFraction(self) -> float:
        """Face.widthFraction (face_types.dart:1189-1190) -> boxVisibleWidthFraction."""
        from .face_gates import boxVisibleWidthFraction
        return boxVisibleWidthFraction(self.detectionData.boundingBox, float(self.originalSize.width))

    @staticmethod
    def _parseIris(points: List[Point]) -> Optional[Eye]:
        """Face._parseIris (face_types.dart:1146-1173): 76 points per eye = 71 eye-mesh + 5 iris."""
        if len(points) < 5:
            return None
        if len(points) == 76:
            eye_mesh, iris = points[:71], points[71:76]
        elif len(points) > 5:
            eye_mesh, iris = points[:-5], points[-5:]
        else:
            eye_mesh, iris = [], points
        c = irisCenterFromPoints(iris)
        return Eye(c, [p for p in iris if p is not c], eye_mesh)

    @property
    def eyes(self) -> Optional[EyePair]:
        """Face.eyes (face_types.dart:1193, _computeEyes :1289-1310): None without iris data."""
        n = len(self.irisPoints)
        if n == 0:
            return None
        left = right = None
        if n == 152:
            left, right = self._parseIris(self.irisPoints[:76]), self._parseIris(self.irisPoints[76:152])
        elif n == 76:
            left = self._parseIris(self.irisPoints)
        elif n == 10:
            left, right = self._parseIris(self.irisPoints[:5]), self._parseIris(self.irisPoints[5:10])
        elif n > 10 and n % 2 == 0:
            left, right = self._parseIris(self.irisPoints[:n // 2]), self._parseIris(self.irisPoints[n // 2:])
        elif n >= 5:
            left = self._parseIris(self.irisPoints)
        if left is None and right is None:
            return None
        return EyePair(left, right)

    def toMap(self) -> dict:
        """Face.toMap (face_types.dart:1350-1361)."""
        m = {"detection": self.detectionData.toMap()}
        if self.trackingId is not None:
            m["trackingId"] = self.trackingId
        if self.mesh is not None:
            m["mesh"] = self.mesh.toMap()
        m["irisPoints"] = [p.toMap() for p in self.irisPoints]
        m["originalSize"] = {"width": self.originalSize.width, "height": self.originalSize.height}
        return m

    @staticmethod
    def fromMap(m: dict) -> "Face":
        """Face.fromMap (face_types.dart:1363-1378)."""
        return Face(Detection.fromMap(m["detection"]), FaceMesh.fromMap(m["mesh"]) if m.get("mesh") is not None else None,
                    Size(m["originalSize"]["width"], m["originalSize"]["height"]),
                    irisPoints=[Point.fromMap(p) for p in m["irisPoints"]], trackingId=m.get("trackingId"))

"""FaceDetector — host-side mirror of the reference's public API for the detection hot path
(lib/src/face_detector.dart): create / initialize / detectFacesFromMatBytes / dispose keep their
names, argument meaning and error behaviour; detectFacesBatch is the new batched entry point.

All compute happens in libfdt_cuda.so behind the C ABI (include/fdt_api.h).  The isolate RPC of
the reference (face_detector.dart:1132-1584) is replaced by one FFI call."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .face_types import (IRIS_MODEL_FILE, MESH_MODEL_FILE, MODEL_FILES, Detection, Face, FaceDetectionMode,
                         FaceDetectionModel, FaceMesh, Point, RectF, Size)

ASSETS = Path(__file__).resolve().parent.parent / "assets" / "models"


class StateError(RuntimeError):
    """Dart StateError (not initialised / disposed / double initialise)."""


class FormatException(ValueError):
    """Dart FormatException: the image bytes cannot be decoded (face_detector.dart:476)."""


def _raise(lib, handle, code: int):
    msg = lib.fdt_last_error(handle)
    msg = msg.decode() if msg else "error %d" % code
    if code == _ffi.FDT_ERR_NOT_READY:
        raise StateError(msg)
    if code in (_ffi.FDT_ERR_BAD_ARG, _ffi.FDT_ERR_SIZE_MISMATCH):
        raise ValueError(msg)          # Dart ArgumentError
    if code == _ffi.FDT_ERR_FORMAT:
        raise FormatException(msg)
    if code == _ffi.FDT_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


class FaceDetector:
    modelVersion = "1.1.1"   # lib/src/face_detector.dart:54-64

    def __init__(self):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        self._ready = False
        self._max_faces = _ffi.FDT_MAX_FACES

    # -- lifecycle ------------------------------------------------------------------------------
    @classmethod
    def create(cls, model: FaceDetectionModel = FaceDetectionModel.backCamera, *, minScore: float = 0.0,
               minFaceSize: float = 0.0, minFacePresenceConfidence: float = 0.5, device: int = 0,
               maxBatch: int = 0, maxFaces: int = 0, fuseLevel: int = -1, withMesh: bool = True, withIris: bool = True,
               detectorBytes: Optional[bytes] = None, meshBytes: Optional[bytes] = None,
               irisBytes: Optional[bytes] = None, devices: Optional[Sequence[int]] = None) -> "FaceDetector":
        """FaceDetector.create (face_detector.dart:84-119).  `devices` (new): a list of CUDA device ordinals; every
        detect call then splits its batch across them inside the library (fdt_create_ex)."""
        d = cls()
        d.initialize(model, minScore=minScore, minFaceSize=minFaceSize,
                     minFacePresenceConfidence=minFacePresenceConfidence, device=device, maxBatch=maxBatch,
                     maxFaces=maxFaces, fuseLevel=fuseLevel, withMesh=withMesh, withIris=withIris,
                     detectorBytes=detectorBytes, meshBytes=meshBytes, irisBytes=irisBytes, devices=devices)
        return d

    def initialize(self, model: FaceDetectionModel = FaceDetectionModel.backCamera, *, minScore: float = 0.0,
                   minFaceSize: float = 0.0, minFacePresenceConfidence: float = 0.5, device: int = 0,
                   maxBatch: int = 0, maxFaces: int = 0, fuseLevel: int = -1, withMesh: bool = True, withIris: bool = True,
                   detectorBytes: Optional[bytes] = None, meshBytes: Optional[bytes] = None,
                   irisBytes: Optional[bytes] = None, devices: Optional[Sequence[int]] = None) -> None:
        """FaceDetector.initialize (face_detector.dart:297-415)."""
        if self._ready:
            raise StateError("FaceDetector already initialized")          # :315-317
        model = FaceDetectionModel(model)
        if detectorBytes is None:
            detectorBytes = (ASSETS / MODEL_FILES[model]).read_bytes()   # rootBundle.load (:353-372)
        if meshBytes is None and withMesh:
            meshBytes = (ASSETS / MESH_MODEL_FILE).read_bytes()
        if irisBytes is None and withMesh and withIris:
            irisBytes = (ASSETS / IRIS_MODEL_FILE).read_bytes()
        if not withMesh:
            meshBytes = irisBytes = None
        cfg = _ffi.FdtConfig()
        self._lib.fdt_default_config(C.byref(cfg))
        cfg.model, cfg.device, cfg.max_batch, cfg.max_faces, cfg.fuse_level = int(model), device, maxBatch, maxFaces, fuseLevel
        cfg.min_score, cfg.min_face_size, cfg.min_face_presence = minScore, minFaceSize, minFacePresenceConfidence
        h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices) if devices else None
        rc = self._lib.fdt_create_ex(C.byref(cfg), detectorBytes, len(detectorBytes), meshBytes,
                                     len(meshBytes) if meshBytes else 0, irisBytes, len(irisBytes) if irisBytes else 0,
                                     devs, len(devices) if devices else 0, C.byref(h))
        if rc != _ffi.FDT_OK:
            _raise(self._lib, None, rc)
        self._h = h
        self._ready = True
        self.model = model
        iw, ih, na, mf, mb = (C.c_int32() for _ in range(5))
        self._lib.fdt_get_info(h, C.byref(iw), C.byref(ih), C.byref(na), C.byref(mf), C.byref(mb))
        self.inputWidth, self.inputHeight, self.numAnchors = iw.value, ih.value, na.value
        self._max_faces, self.maxBatch = mf.value, mb.value

    @property
    def isReady(self) -> bool:
        return self._ready

    def dispose(self) -> None:
        """FaceDetector.dispose (face_detector.dart:1061-1081)."""
        if self._ready:
            self._lib.fdt_destroy(self._h)
        self._ready = False
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    def _check(self):
        if not self._ready:
            raise StateError("FaceDetector not initialized. Call initialize() first.")   # :1083-1089

    # -- detection ------------------------------------------------------------------------------
    def detectFacesFromMatBytes(self, data, *, width: int, height: int, matType: int = 16,
                                mode: FaceDetectionMode = FaceDetectionMode.full) -> List[Face]:
        """detectFacesFromMatBytes (face_detector.dart:588-609); the default mode is `full` like the reference's
        (detector + mesh + iris; the blendshape classifier is outside this path)."""
        self._check()
        buf = np.ascontiguousarray(np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1))
        faces = (_ffi.FdtFace * self._max_faces)()
        count = C.c_int32(0)
        mode = FaceDetectionMode(mode)
        want_mesh = mode != FaceDetectionMode.fast
        want_iris = mode == FaceDetectionMode.full
        mesh = np.empty((self._max_faces, 468, 3), np.float32) if want_mesh else None
        iris = np.empty((self._max_faces, 152, 3), np.float32) if want_iris else None
        rc = self._lib.fdt_detect_one(self._h, buf.ctypes.data, buf.size, width, height, matType, int(mode), faces,
                                      C.byref(count), mesh.ctypes.data_as(_ffi.f32p) if want_mesh else None,
                                      iris.ctypes.data_as(_ffi.f32p) if want_iris else None)
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)
        return [self._to_face(faces[i], mesh[i] if want_mesh else None, width, height, iris[i] if want_iris else None)
                for i in range(count.value)]

    def detectFacesFromBytes(self, imageBytes, *, mode: FaceDetectionMode = FaceDetectionMode.full) -> List[Face]:
        """detectFacesFromBytes (face_detector.dart:477-485): encoded image bytes (JPEG) -> faces.  The reference decodes with
        cv.imdecode; here the host runs the Huffman decoder and the device does the rest (bit-exact with cv.imdecode, EXIF
        orientation included).  Raises FormatException when the bytes cannot be decoded, like the reference."""
        self._check()
        buf = np.frombuffer(bytes(imageBytes), np.uint8)
        faces = (_ffi.FdtFace * self._max_faces)()
        count = C.c_int32(0)
        mode = FaceDetectionMode(mode)
        want_mesh = mode != FaceDetectionMode.fast
        want_iris = mode == FaceDetectionMode.full
        mesh = np.empty((self._max_faces, 468, 3), np.float32) if want_mesh else None
        iris = np.empty((self._max_faces, 152, 3), np.float32) if want_iris else None
        wh = (C.c_int32 * 2)()
        rc = self._lib.fdt_detect_jpeg(self._h, buf.ctypes.data, buf.size, int(mode), faces, C.byref(count),
                                       mesh.ctypes.data_as(_ffi.f32p) if want_mesh else None,
                                       iris.ctypes.data_as(_ffi.f32p) if want_iris else None, wh)
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)
        return [self._to_face(faces[i], mesh[i] if want_mesh else None, wh[0], wh[1], iris[i] if want_iris else None)
                for i in range(count.value)]

    def decodeImage(self, imageBytes) -> np.ndarray:
        """The decoder alone (cv.imdecode(bytes, IMREAD_COLOR)): HxWx3 BGR uint8."""
        self._check()
        buf = np.frombuffer(bytes(imageBytes), np.uint8)
        wh = (C.c_int32 * 2)()
        rc = self._lib.fdt_decode_jpeg(self._h, buf.ctypes.data, buf.size, None, 0, wh)
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)
        out = np.empty((wh[1], wh[0], 3), np.uint8)
        rc = self._lib.fdt_get_decoded_frame(self._h, out.ctypes.data, out.size)
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)
        return out

    def detectFacesFromMat(self, mat: np.ndarray, *, mode: FaceDetectionMode = FaceDetectionMode.full) -> List[Face]:
        """detectFacesFromMat (face_detector.dart:559-572): `mat` is an HxW[xC] uint8 array (cv.Mat)."""
        ch = 1 if mat.ndim == 2 else mat.shape[2]
        mt = {1: _ffi.FDT_MAT_8UC1, 3: _ffi.FDT_MAT_8UC3, 4: _ffi.FDT_MAT_8UC4}.get(ch)
        if mt is None or mat.dtype != np.uint8:
            raise ValueError("unsupported Mat type")
        return self.detectFacesFromMatBytes(np.ascontiguousarray(mat), width=mat.shape[1], height=mat.shape[0],
                                            matType=mt, mode=mode)

    def detectFacesBatch(self, frames, *, count: int, width: int, height: int, matType: int = 16,
                         mode: FaceDetectionMode = FaceDetectionMode.fast) -> List[List[Face]]:
        """New batched entry point: `frames` holds `count` packed frames (bytes / uint8 array)."""
        self._check()
        faces, counts, mesh, iris = self.detectBatchRaw(frames, count=count, width=width, height=height,
                                                        matType=matType, mode=mode, withIris=True)
        out = []
        for b in range(count):
            out.append([self._to_face(faces[b * self._max_faces + j], mesh[b, j] if mesh is not None else None, width, height,
                                      iris[b, j] if iris is not None else None)
                        for j in range(int(counts[b]))])
        return out

    def detectBatchRaw(self, frames, *, count: int, width: int, height: int, matType: int = 16,
                       mode: FaceDetectionMode = FaceDetectionMode.fast, memKind: int = _ffi.FDT_MEM_HOST,
                       rowStride: Optional[int] = None, withIris: bool = False):
        """fdt_detect_batch with array outputs: (FdtFace array, counts int32[count], mesh f32 or None) and, with
        withIris=True, a fourth element iris f32 [count, max_faces, 152, 3] (None unless mode is full).
        `frames` may be a numpy array / bytes (host) or an integer device pointer with memKind=FDT_MEM_DEVICE."""
        self._check()
        ch = {_ffi.FDT_MAT_8UC1: 1, _ffi.FDT_MAT_8UC3: 3, _ffi.FDT_MAT_8UC4: 4}.get(matType, 0)
        stride = rowStride if rowStride is not None else width * ch
        if isinstance(frames, int):
            ptr = frames
        else:
            buf = np.ascontiguousarray(np.frombuffer(frames, np.uint8) if not isinstance(frames, np.ndarray) else frames.reshape(-1))
            if buf.size != count * height * stride:
                raise ValueError("frames length does not equal count * height * rowStride")   # helpers.dart:440-447
            ptr = buf.ctypes.data
        faces = (_ffi.FdtFace * max(1, count * self._max_faces))()
        counts = np.zeros(max(1, count), np.int32)
        mode = FaceDetectionMode(mode)
        want_mesh = mode != FaceDetectionMode.fast
        want_iris = mode == FaceDetectionMode.full
        mesh = np.zeros((count, self._max_faces, 468, 3), np.float32) if want_mesh else None
        iris = np.zeros((count, self._max_faces, 152, 3), np.float32) if want_iris else None
        rc = self._lib.fdt_detect_batch(self._h, ptr, count, width, height, stride, matType, int(mode), memKind, faces,
                                        counts.ctypes.data_as(_ffi.i32p),
                                        mesh.ctypes.data_as(_ffi.f32p) if want_mesh else None,
                                        iris.ctypes.data_as(_ffi.f32p) if want_iris else None)
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)
        return (faces, counts[:count], mesh, iris) if withIris else (faces, counts[:count], mesh)

    @staticmethod
    def _to_face(f, mesh, width, height, iris=None) -> Face:
        size = Size(float(width), float(height))
        det = Detection(RectF(f.xmin, f.ymin, f.xmax, f.ymax), f.score, [f.keypoints[k] for k in range(12)], size)
        fm = FaceMesh(np.array(mesh, np.float32), f.mesh_score) if (mesh is not None and f.has_mesh) else None
        pts, packed = [], None
        if iris is not None and f.has_iris:
            packed = np.array(iris, np.float32)
            pts = [Point(float(p[0]), float(p[1]), float(p[2])) for p in packed]
        return Face(det, fm, size, irisPoints=pts, anchorIndex=f.anchor_index, irisPacked=packed)

    # -- embedding alignment (lib/src/models/face_embedding.dart:362-384, face_detector_core.dart:419-452) ----------
    def extractAlignedSquares(self, mat: np.ndarray, rois, outSize: int):
        """extractAlignedSquare (helpers.dart:583-625) on the device for ROIs (cx, cy, size, theta) of one frame.
        Returns (crops u8 [n, outSize, outSize, 3], ok bool[n])."""
        self._check()
        mat = np.ascontiguousarray(mat)
        ch = 1 if mat.ndim == 2 else mat.shape[2]
        mt = {1: _ffi.FDT_MAT_8UC1, 3: _ffi.FDT_MAT_8UC3, 4: _ffi.FDT_MAT_8UC4}[ch]
        r = np.ascontiguousarray(np.asarray(rois, np.float64).reshape(-1, 4))
        out = np.empty((r.shape[0], outSize, outSize, 3), np.uint8)
        ok = np.zeros(r.shape[0], np.int32)
        self._rc(self._lib.fdt_extract_aligned_squares(self._h, mat.ctypes.data, mat.shape[1], mat.shape[0], mat.shape[1] * ch, mt,
                                                       r.ctypes.data, r.shape[0], outSize, out.ctypes.data, ok.ctypes.data_as(_ffi.i32p)))
        return out, ok.astype(bool)

    def embeddingCrop(self, face: Face, mat: np.ndarray, outSize: int = 112) -> np.ndarray:
        """The aligned 112x112 crop getFaceEmbeddingFromEyesDirect feeds MobileFaceNet (face_detector_core.dart:419-452):
        computeEmbeddingAlignment on the (iris-refined) eye keypoints, then extractAlignedSquare(..., -theta).  The
        embedding model itself is not shipped by the reference."""
        lm = face.landmarks
        le, re = lm[0], lm[1]
        out4 = (C.c_double * 4)()
        self._lib.fdt_host_embedding_roi((C.c_double * 2)(le.x, le.y), (C.c_double * 2)(re.x, re.y), out4)
        theta, cx, cy, size = out4
        crops, ok = self.extractAlignedSquares(mat, [[cx, cy, size, -theta]], outSize)
        if not ok[0]:
            raise StateError("Failed to extract aligned face crop for embedding")
        return crops[0]

    # -- parity taps (tests only) -----------------------------------------------------------------
    def debugLetterboxed(self, n: int) -> np.ndarray:
        out = np.empty((n, self.inputHeight, self.inputWidth, 3), np.uint8)
        self._rc(self._lib.fdt_debug_get_letterboxed(self._h, n, out.ctypes.data))
        return out

    def debugInputTensor(self, n: int) -> np.ndarray:
        out = np.empty((n, self.inputHeight, self.inputWidth, 3), np.float32)
        self._rc(self._lib.fdt_debug_get_input_tensor(self._h, n, out.ctypes.data))
        return out

    def debugRawHeads(self, n: int):
        boxes = np.empty((n, self.numAnchors, 16), np.float32)
        scores = np.empty((n, self.numAnchors), np.float32)
        self._rc(self._lib.fdt_debug_get_raw_heads(self._h, n, boxes.ctypes.data, scores.ctypes.data))
        return boxes, scores

    def debugCandidates(self, image: int) -> np.ndarray:
        idx = np.empty(self.numAnchors, np.int32)
        n = C.c_int32(0)
        self._rc(self._lib.fdt_debug_get_candidates(self._h, image, idx.ctypes.data, idx.size, C.byref(n)))
        return idx[:n.value].copy()

    def debugTensor(self, which: int, tensor: int, n: int) -> np.ndarray:
        dims = (C.c_int32 * 4)()
        self._rc(self._lib.fdt_debug_get_tensor(self._h, which, tensor, n, None, 0, dims))
        out = np.empty(tuple(dims), np.float32)
        self._rc(self._lib.fdt_debug_get_tensor(self._h, which, tensor, n, out.ctypes.data, out.size, dims))
        return out

    def debugMeshStage(self, n: int):
        crops = np.empty((n, 192, 192, 3), np.uint8)
        raw = np.empty((n, 1404), np.float32)
        flag = np.empty((n,), np.float32)
        got = C.c_int32(0)
        self._rc(self._lib.fdt_debug_get_mesh_stage(self._h, n, crops.ctypes.data, raw.ctypes.data, flag.ctypes.data, C.byref(got)))
        m = min(n, got.value)
        return crops[:m], raw[:m], flag[:m]

    def debugIrisStage(self, n: int):
        """(eye crops u8 [2m,64,64,3], rois f64 [2m,4] = cx,cy,size,theta, contours f32 [2m,213], iris f32 [2m,15])."""
        crops = np.empty((2 * n, 64, 64, 3), np.uint8)
        rois = np.empty((2 * n, 4), np.float64)
        cont = np.empty((2 * n, 213), np.float32)
        ir = np.empty((2 * n, 15), np.float32)
        got = C.c_int32(0)
        self._rc(self._lib.fdt_debug_get_iris_stage(self._h, n, crops.ctypes.data, rois.ctypes.data, cont.ctypes.data, ir.ctypes.data, C.byref(got)))
        m = 2 * min(n, got.value)
        return crops[:m], rois[:m], cont[:m], ir[:m]

    def debugDecode(self, rawBoxes, rawScores, *, anchors=None, scale: float, scoreThresh: float = 0.5, iouThresh: float = 0.3,
                    padding=None):
        """k_decode_nms on caller-supplied raw heads: boxes [B,N,16], scores [B,N] -> (list of list of FdtFace-like
        dicts, decoded candidates [B][n_cand, 18])."""
        boxes = np.ascontiguousarray(np.asarray(rawBoxes, np.float32))
        scores = np.ascontiguousarray(np.asarray(rawScores, np.float32))
        if boxes.ndim == 2:
            boxes, scores = boxes[None], scores[None]
        B, N = scores.shape
        anc = np.ascontiguousarray(np.asarray(anchors, np.float64).reshape(N, 2)) if anchors is not None else None
        pad = (C.c_double * 4)(*padding) if padding is not None else None
        faces = (_ffi.FdtFace * (B * _ffi.FDT_MAX_FACES))()
        counts = np.zeros(B, np.int32)
        dec = np.zeros((B, N, 18), np.float64)
        ndec = np.zeros(B, np.int32)
        self._rc(self._lib.fdt_debug_decode(self._h, boxes.ctypes.data, scores.ctypes.data, anc.ctypes.data if anc is not None else None,
                                            B, N, float(scale), float(scoreThresh), float(iouThresh), pad, faces,
                                            counts.ctypes.data_as(_ffi.i32p), dec.ctypes.data, ndec.ctypes.data_as(_ffi.i32p)))
        out = [[self._face_row(faces[b * _ffi.FDT_MAX_FACES + j]) for j in range(int(counts[b]))] for b in range(B)]
        return out, [dec[b, :int(ndec[b])] for b in range(B)]

    def debugNms(self, dets17, *, scoreThresh: float, iouThresh: float, padding=None):
        """weightedNms + letterbox removal on the device for detections [n,17] = xmin,ymin,xmax,ymax,score,kp12."""
        d = np.ascontiguousarray(np.asarray(dets17, np.float64).reshape(-1, 17))
        pad = (C.c_double * 4)(*padding) if padding is not None else None
        faces = (_ffi.FdtFace * _ffi.FDT_MAX_FACES)()
        count = C.c_int32(0)
        self._rc(self._lib.fdt_debug_nms(self._h, d.ctypes.data if d.shape[0] else None, d.shape[0], float(scoreThresh), float(iouThresh), pad, faces, C.byref(count)))
        return [self._face_row(faces[j]) for j in range(count.value)]

    @staticmethod
    def _face_row(f) -> dict:
        return {"box": (f.xmin, f.ymin, f.xmax, f.ymax), "score": f.score, "kp": [f.keypoints[k] for k in range(12)], "index": f.anchor_index}

    def numDevices(self) -> int:
        return int(self._lib.fdt_num_devices(self._h))

    def anchors(self) -> np.ndarray:
        a = np.empty((self.numAnchors, 2), np.float64)
        self._rc(self._lib.fdt_get_anchors(self._h, a.ctypes.data_as(_ffi.f64p)))
        return a

    def lastLaunchCount(self) -> int:
        return int(self._lib.fdt_last_launch_count(self._h))

    def _rc(self, rc):
        if rc != _ffi.FDT_OK:
            _raise(self._lib, self._h, rc)

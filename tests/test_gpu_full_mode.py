"""GPU parity of the widened path: FaceDetectionMode.full (eye ROIs -> 64x64 warps -> iris_landmark ->
iris points + refined eye keypoints), the embedding alignment crop, the in-library multi-device split, and the full
benchmark frame sets / batch sizes of configs C2, C3 and C4.  Everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle import cv_ops as co, detect_post as dp, geometry as geo
from oracle.pipeline import OraclePipeline

pytestmark = pytest.mark.gpu

HEAD_REL_TOL = 1e-4
COORD_TOL = 1e-3


@pytest.fixture(scope="module")
def fdt(lib):
    import face_detection_tflite_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def iris_bytes():
    from conftest import ASSETS
    return (ASSETS / "models" / "iris_landmark.tflite").read_bytes()


@pytest.fixture(scope="module")
def images():
    import cv2
    from conftest import ASSETS
    return {n: cv2.imread(str(ASSETS / "samples" / n)) for n in ("iris-detection-ex1.jpg", "iris-detection-ex2.jpg", "landmark-ex1.jpg")}


_dets, _oracles = {}, {}


def detector(fdt, model, **kw):
    key = (model, tuple(sorted(kw.items())))
    if key not in _dets:
        _dets[key] = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], **kw)
    return _dets[key]


def oracle(model_bytes, iris_bytes, model, backend="f64"):
    if (model, backend) not in _oracles:
        _oracles[(model, backend)] = OraclePipeline(model_bytes[model], model, model_bytes["mesh"], backend, iris_bytes=iris_bytes)
    return _oracles[(model, backend)]


@pytest.mark.parametrize("model", ["backCamera", "shortRange"])
def test_full_mode_matches_oracle(fdt, model_bytes, iris_bytes, images, model):
    """detectFacesFromMatBytes with the reference's DEFAULT mode (full, face_detector.dart:588-594)."""
    d = detector(fdt, model)
    o = oracle(model_bytes, iris_bytes, model)
    for name, img in images.items():
        h, w = img.shape[:2]
        faces = d.detectFacesFromMatBytes(img.tobytes(), width=w, height=h)           # mode defaults to full
        want = o.detect_faces(img, "full")
        assert len(faces) == len(want) == 1                                             # all_model_variants_test.dart:28-32, :193-206
        g, r = faces[0], want[0]
        assert g.anchorIndex == r.det.anchor and len(g.irisPoints) == 152 and len(g.mesh) == 468
        crops, rois, cont, ir = d.debugIrisStage(1)
        assert crops.shape[0] == 2
        for e in range(2):
            # like for like: the oracle's warp of the SAME roi is bit-exact (the right eye arrives mirrored)
            wc = co.extract_aligned_square(img, rois[e][0], rois[e][1], rois[e][2], rois[e][3], 64)
            if e == 1:
                wc = wc[:, ::-1]
            assert np.array_equal(crops[e], wc)
            ref = o.iris.run(co.normalize_bgr_u8(crops[e])[None])
            assert np.abs(cont[e] - ref[0][0]).max() <= HEAD_REL_TOL * np.abs(ref[0][0]).max()
            assert np.abs(ir[e] - ref[1][0]).max() <= HEAD_REL_TOL * np.abs(ref[1][0]).max()
            # the ROI itself follows the mesh: <= 1e-3 of its own size from the all-oracle ROI
            assert np.abs(np.array(rois[e][:3]) - np.array(r.eye_rois[e][:3])).max() <= COORD_TOL * r.eye_rois[e][2]
            assert abs(rois[e][3] - r.eye_rois[e][3]) <= 2e-3
        # end to end: iris points within 1e-3 of the eye ROI side, refined eye keypoints within 1e-3 (normalised)
        for e in range(2):
            sl = slice(76 * e, 76 * e + 76)
            assert np.abs(g.irisPacked[sl, :2] - r.iris_px[sl, :2]).max() <= 2 * COORD_TOL * r.eye_rois[e][2]
            # z is the model's raw depth output (not geometric, face_geometry.dart:123): it moves with the crop's sub-pixel
            # placement, i.e. with the fp32 mesh; like-for-like (same crop) it is within 1e-4 above
            assert np.abs(g.irisPacked[sl, 2] - r.iris_px[sl, 2]).max() <= 2e-2 * max(1.0, np.abs(r.iris_px[sl, 2]).max())
        assert np.abs(np.array(g.detectionData.keypointsXY) - np.array(r.det.kp)).max() <= COORD_TOL
        # the refined keypoints ARE iris points (closest to the centroid), the others are the detector's
        lc = geo.iris_center_from_points(g.irisPacked[71:76].astype(np.float64))
        assert abs(g.detectionData.keypointsXY[0] - lc[0] / w) <= 1e-6 and abs(g.detectionData.keypointsXY[1] - lc[1] / h) <= 1e-6
        std = d.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.standard)[0]
        assert std.detectionData.keypointsXY[4:] == g.detectionData.keypointsXY[4:] and std.irisPoints == []
        assert std.detectionData.keypointsXY[:4] != g.detectionData.keypointsXY[:4]
        assert np.array_equal(std.mesh.packed, g.mesh.packed)
        eyes = g.eyes                                                                   # face_detection_integration_test.dart:403-421
        assert len(eyes.leftEye.mesh) == 71 and len(eyes.leftEye.irisContour) == 4 and len(eyes.rightEye.mesh) == 71


def test_full_mode_batch_and_fast_unchanged(fdt, images):
    d = detector(fdt, "backCamera")
    img = images["iris-detection-ex2.jpg"]
    h, w = img.shape[:2]
    frames = np.stack([img, img[:, ::-1].copy(), np.zeros_like(img), img])
    batch = d.detectFacesBatch(frames, count=4, width=w, height=h, mode=fdt.FaceDetectionMode.full)
    one = d.detectFacesFromMat(img)
    assert [len(x) for x in batch] == [1, 1, 0, 1]
    for k in (0, 3):
        assert batch[k][0].detectionData == one[0].detectionData
        assert np.array_equal(batch[k][0].irisPacked, one[0].irisPacked) and np.array_equal(batch[k][0].mesh.packed, one[0].mesh.packed)
    fast = d.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    assert fast[0].mesh is None and fast[0].irisPoints == [] and fast[0].detectionData.boundingBox == one[0].detectionData.boundingBox
    nomesh = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withMesh=False)
    with pytest.raises(fdt.StateError):
        nomesh.detectFacesFromMat(img)                                                  # full needs mesh + iris models
    noiris = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withIris=False)
    with pytest.raises(fdt.StateError):
        noiris.detectFacesFromMat(img)
    assert len(noiris.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.standard)) == 1
    nomesh.dispose(); noiris.dispose()


def test_iris_layers(fdt, model_bytes, iris_bytes, images):
    """Every materialised tensor of the iris_landmark graph (fuse levels 0 and 2) vs the fp64 oracle."""
    img = images["iris-detection-ex1.jpg"]
    o = oracle(model_bytes, iris_bytes, "backCamera")
    for fuse in (0, 2):
        d = detector(fdt, "backCamera", fuseLevel=fuse, maxBatch=4)
        assert len(d.detectFacesFromMat(img)) == 1
        crops, _, _, _ = d.debugIrisStage(1)
        ref = o.iris.exe.run(np.stack([co.normalize_bgr_u8(c) for c in crops]), taps="all")
        n = 0
        for tf_idx, want in ref.items():
            try:
                got = d.debugTensor(2, tf_idx, 2)
            except ValueError:
                continue
            want = want.reshape(got.shape)
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
            assert err <= HEAD_REL_TOL, "iris tensor %d: %.3e (fuse %d)" % (tf_idx, err, fuse)
            n += 1
        assert n >= (20 if fuse else 60)


def test_extract_aligned_squares_and_embedding_crop(fdt, images):
    d = detector(fdt, "backCamera")
    img = images["landmark-ex1.jpg"]
    h, w = img.shape[:2]
    rng = np.random.default_rng(5)
    rois = [[rng.uniform(-50, w + 50), rng.uniform(-50, h + 50), rng.uniform(1, 600), rng.uniform(-3.2, 3.2)] for _ in range(24)]
    rois += [[100.0, 100.0, 0.4, 0.3], [w / 2, h / 2, 112.0, 0.0]]           # round(size) == 0 -> null ; identity scale
    for out in (112, 64, 192):
        crops, ok = d.extractAlignedSquares(img, rois, out)
        for i, r in enumerate(rois):
            want = co.extract_aligned_square(img, r[0], r[1], r[2], r[3], out)
            assert ok[i] == (want is not None)
            if want is not None:
                assert np.array_equal(crops[i], want), (out, i)
    # embedding path (face_detector_core.dart:419-452): computeEmbeddingAlignment on the iris-refined eyes, -theta, 112x112
    face = d.detectFacesFromMat(img)[0]
    lm = face.landmarks
    theta, cx, cy, size = geo.compute_embedding_alignment((lm[0].x, lm[0].y), (lm[1].x, lm[1].y))
    assert np.array_equal(d.embeddingCrop(face, img), co.extract_aligned_square(img, cx, cy, size, -theta, 112))
    gray = np.ascontiguousarray(img[..., 1])
    crops, ok = d.extractAlignedSquares(gray, rois[:4], 64)
    for i in range(4):
        assert np.array_equal(crops[i], co.extract_aligned_square(np.repeat(gray[..., None], 3, 2), *rois[i], 64))


def test_multi_device_handle(fdt, images):
    """One handle, several devices, one fdt_detect_batch call (fdt_create_ex device list)."""
    import torch
    from face_detection_tflite_b200 import synth
    ndev = torch.cuda.device_count()
    with pytest.raises(ValueError):
        fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, devices=[0, 0])
    with pytest.raises(ValueError):
        fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, devices=[0, ndev])
    one = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, devices=[0], maxBatch=8)
    assert one.numDevices() == 1
    frames = np.concatenate([synth.face_frames(21, 640, 360, max_side=300), synth.noise_frames(2, 640, 360)])
    want = one.detectFacesBatch(frames, count=23, width=640, height=360, mode=fdt.FaceDetectionMode.full)
    assert sum(len(x) for x in want) >= 10
    if ndev < 2:
        one.dispose()
        pytest.skip("needs >= 2 GPUs for the split itself (run with gpurun --gpus 2)")
    multi = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, devices=list(range(ndev)), maxBatch=8)
    assert multi.numDevices() == ndev
    for mode in (fdt.FaceDetectionMode.fast, fdt.FaceDetectionMode.full):
        a = multi.detectFacesBatch(frames, count=23, width=640, height=360, mode=mode)
        b = one.detectFacesBatch(frames, count=23, width=640, height=360, mode=mode)
        for fa, fb in zip(a, b):                                          # frame order, bitwise identical to one device
            assert [f.detectionData for f in fa] == [f.detectionData for f in fb]
            for x, y in zip(fa, fb):
                if mode == fdt.FaceDetectionMode.full:
                    assert np.array_equal(x.mesh.packed, y.mesh.packed) and np.array_equal(x.irisPacked, y.irisPacked)
    a1 = multi.detectFacesBatch(frames[:1], count=1, width=640, height=360)      # fewer frames than devices
    assert [f.detectionData for f in a1[0]] == [f.detectionData for f in want[0]]
    with pytest.raises(NotImplementedError):
        multi.detectBatchRaw(int(torch.from_numpy(frames).cuda().data_ptr()), count=23, width=640, height=360, memKind=1)
    one.dispose(); multi.dispose()


# ---- benchmark configs at their full frame sets / batch sizes ------------------------------------------------------
def _bench_frames(w, h):
    from face_detection_tflite_b200 import synth
    return np.concatenate([synth.face_frames(56, w, h), synth.noise_frames(8, w, h)])


@pytest.mark.parametrize("model,w,h", [("shortRange", 1280, 720), ("full", 1920, 1080)])
def test_all_unique_bench_frames_match_f64_oracle(fdt, model_bytes, iris_bytes, model, w, h):
    """All 64 unique frames of bench.py's C2 / C3 inputs: candidate index set bit-exact, same anchor per face,
    boxes / keypoints <= 1e-3, vs the f64 oracle."""
    frames = _bench_frames(w, h)
    d = detector(fdt, model, withMesh=False)
    o = oracle(model_bytes, iris_bytes, model)
    faces, counts, _ = d.detectBatchRaw(frames, count=64, width=w, height=h)
    S = o.in_h
    nfaces = 0
    for b in range(64):
        tensor, pad, _ = o.preprocess(frames[b])
        boxes, scores = o.raw_heads(tensor)
        idx, _ = dp.collect_candidates(scores)
        assert np.array_equal(d.debugCandidates(b), np.array(idx, np.int32)), b      # index set: bit-exact
        want = dp.postprocess(boxes, scores, o.anchors, S, pad)
        want = [x for x in want if co.dart_round(geo.compute_face_alignment(x.kp, float(w), float(h))[3]) > 0]
        assert counts[b] == len(want), b
        for j, wd in enumerate(want):
            g = faces[b * 100 + j]
            assert g.anchor_index == wd.anchor
            assert abs(g.score - wd.score) <= 1e-4
            assert np.abs(np.array([g.xmin, g.ymin, g.xmax, g.ymax]) - np.array([wd.xmin, wd.ymin, wd.xmax, wd.ymax])).max() <= COORD_TOL
            assert np.abs(np.array(list(g.keypoints)) - np.array(wd.kp)).max() <= COORD_TOL
            assert dp.iou((g.xmin, g.ymin, g.xmax, g.ymax), (wd.xmin, wd.ymin, wd.xmax, wd.ymax)) >= 0.99
        nfaces += len(want)
    assert nfaces >= 40
    assert counts[56:].sum() == 0


def test_c3_full_batch_2048(fdt):
    """BASELINE config 3 at its full batch (2048 x 1920x1080, full-range): periodic input -> periodic output."""
    import torch
    base = _bench_frames(1920, 1080)
    d = detector(fdt, "full", withMesh=False)
    ref, cref, _ = d.detectBatchRaw(base, count=64, width=1920, height=1080)
    dev = torch.from_numpy(base).cuda().repeat(2048 // 64, 1, 1, 1).contiguous()
    faces, counts, _ = d.detectBatchRaw(dev.data_ptr(), count=2048, width=1920, height=1080, memKind=1)
    assert np.array_equal(counts.reshape(-1, 64), np.tile(cref, (32, 1)))
    sz = C.sizeof(type(ref[0]))
    a = np.frombuffer(faces, np.uint8).reshape(2048, -1)
    r = np.frombuffer(ref, np.uint8).reshape(64, -1)
    for b in range(2048):
        n = counts[b] * sz
        assert np.array_equal(a[b, :n], r[b % 64, :n])


def test_c4_batch_1024_crosses_mesh_cap(fdt):
    """BASELINE config 4 at its full batch (1024 frames, up to 4 faces each) on a handle whose mesh pass capacity
    (64 faces) is below the faces of one chunk, so every chunk needs several mesh passes."""
    from face_detection_tflite_b200 import synth
    std = fdt.FaceDetectionMode.standard
    uniq = np.stack([synth.face_frame(5 * k + 4, 1280, 720) for k in range(32)])          # 4 face tiles per frame
    small = detector(fdt, "shortRange", maxBatch=32, withIris=False)                     # mesh_cap = 64 faces
    big = detector(fdt, "shortRange", withIris=False)
    ref = big.detectFacesBatch(uniq, count=32, width=1280, height=720, mode=std)
    per_chunk = sum(len(x) for x in ref)
    assert per_chunk > 64 and max(len(x) for x in ref) <= 4
    frames = np.tile(uniq, (32, 1, 1, 1))
    got = small.detectFacesBatch(frames, count=1024, width=1280, height=720, mode=std)
    assert len(got) == 1024
    for b in range(1024):
        r = ref[b % 32]
        assert [f.detectionData for f in got[b]] == [f.detectionData for f in r]
        for x, y in zip(got[b], r):
            assert x.meshScore == y.meshScore and np.array_equal(x.mesh.packed, y.mesh.packed)

"""GPU parity tests: the CUDA path, called through the C ABI (libfdt_cuda.so), against the oracle on
the same inputs.  Bars (BASELINE.json north_star):
  * bit-exact: letterbox pad offsets + u8 letterboxed image, anchor table, surviving-anchor index set;
  * raw head outputs: max|d| <= 1e-4 * max|ref| per head tensor (SURVEY.md 7.3 definition);
  * boxes / keypoints / mesh points: <= 1e-3 normalised, identical face counts, matched IoU >= 0.99.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import cv_ops as co, detect_post as dp, graph_exec, tflite_reader as tr
from oracle.pipeline import OraclePipeline

pytestmark = pytest.mark.gpu

HEAD_REL_TOL = 1e-4       # of max|ref| per head tensor
COORD_TOL = 1e-3          # normalised box / keypoint coordinates
MODEL_ENUM = {"shortRange": 2, "full": 3, "backCamera": 1}


@pytest.fixture(scope="module")
def fdt(lib):
    import face_detection_tflite_b200 as pkg
    return pkg


_dets = {}


def get_detector(fdt, model, fuse=-1, mesh=True, **kw):
    key = (model, fuse, mesh, tuple(sorted(kw.items())))
    if key not in _dets:
        _dets[key] = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], fuseLevel=fuse, withMesh=mesh, **kw)
    return _dets[key]


_oracles = {}


def get_oracle(model_bytes, model, backend="f64"):
    key = (model, backend)
    if key not in _oracles:
        _oracles[key] = OraclePipeline(model_bytes[model], model, model_bytes["mesh"], backend)
    return _oracles[key]


def iou(a, b):
    return dp.iou(a, b)


def assert_faces_match(got_faces, want_dets):
    assert len(got_faces) == len(want_dets)
    for g, w in zip(got_faces, want_dets):
        r = g.detectionData.boundingBox
        assert g.anchorIndex == w.anchor
        assert abs(g.score - w.score) <= 1e-4
        assert np.abs(np.array([r.xmin, r.ymin, r.xmax, r.ymax]) - np.array([w.xmin, w.ymin, w.xmax, w.ymax])).max() <= COORD_TOL
        assert np.abs(np.array(g.detectionData.keypointsXY) - np.array(w.kp)).max() <= COORD_TOL
        assert iou((r.xmin, r.ymin, r.xmax, r.ymax), (w.xmin, w.ymin, w.xmax, w.ymax)) >= 0.99


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["shortRange", "full", "backCamera"])
def test_anchor_table_bit_exact(fdt, model):
    d = get_detector(fdt, model)
    assert np.array_equal(d.anchors(), dp.generate_anchors(dp.ssd_options_for(model)))


LB_SHAPES = [(720, 1280, 3), (1080, 1920, 3), (853, 1280, 3), (100, 100, 3), (1, 1, 3), (10, 10, 3), (50, 37, 3),
             (128, 128, 3), (481, 641, 3), (300, 200, 4), (300, 200, 1), (64, 2000, 3), (2000, 64, 3)]


@pytest.mark.parametrize("shape", LB_SHAPES)
def test_letterbox_bit_exact_and_input_tensor(fdt, shape):
    h, w, ch = shape
    rng = np.random.default_rng(h * 7 + w)
    frames = rng.integers(0, 256, (3, h, w, ch), dtype=np.uint8)
    d = get_detector(fdt, "shortRange")
    mt = {1: 0, 3: 16, 4: 24}[ch]
    d.detectBatchRaw(frames, count=3, width=w, height=h, matType=mt)
    got = d.debugLetterboxed(3)
    t = d.debugInputTensor(3)
    for b in range(3):
        src = frames[b]
        bgr = src[..., :3] if ch >= 3 else np.repeat(src, 3, 2)
        want, p = co.letterbox_u8(np.ascontiguousarray(bgr), 128, 128)
        assert np.array_equal(got[b], want)                                  # u8 stage: bit-exact
        assert np.abs(t[b] - co.normalize_bgr_u8(want)).max() <= 1e-6       # fp32 stage: <= 2 ulp (Appendix B)


def test_letterbox_golden_cv2_and_row_stride(fdt, golden, sample_images):
    for model, S in (("shortRange", 128), ("full", 192), ("backCamera", 256)):
        d = get_detector(fdt, model)
        for name, img in sample_images.items():
            h, w = img.shape[:2]
            d.detectBatchRaw(img, count=1, width=w, height=h)
            assert np.array_equal(d.debugLetterboxed(1)[0], golden["%s/%s/letterboxed_cv2" % (model, name)])
    # padded rows (row_stride > width*3) give the same result
    d = get_detector(fdt, "shortRange")
    img = sample_images["landmark-ex1.jpg"]
    h, w = img.shape[:2]
    padded = np.zeros((h, w * 3 + 64), np.uint8)
    padded[:, :w * 3] = img.reshape(h, -1)
    d.detectBatchRaw(padded, count=1, width=w, height=h, rowStride=w * 3 + 64)
    assert np.array_equal(d.debugLetterboxed(1)[0], golden["shortRange/landmark-ex1.jpg/letterboxed_cv2"])


def _frames_for(model, sample_images, n_synth=3):
    from face_detection_tflite_b200 import synth
    S = {"shortRange": (1280, 720), "full": (1920, 1080), "backCamera": (1280, 720)}[model]
    fr = list(synth.face_frames(n_synth, S[0], S[1], start=3))
    fr.append(synth.noise_frames(1, S[0], S[1], seed=5)[0])
    return fr


@pytest.mark.parametrize("model,fuse", [("shortRange", 0), ("shortRange", 2), ("full", 0), ("full", 2), ("backCamera", 2)])
def test_every_materialised_tensor(fdt, model_bytes, sample_images, model, fuse):
    """Layer-by-layer parity: every activation the plan materialises vs the fp64 oracle."""
    d = get_detector(fdt, model, fuse=fuse, mesh=False, maxBatch=4)
    o = get_oracle(model_bytes, model)
    img = sample_images["landmark-ex1.jpg"]
    h, w = img.shape[:2]
    frames = np.stack([img, img[::-1].copy()])
    d.detectBatchRaw(frames, count=2, width=w, height=h)
    t0, _, _ = o.preprocess(frames[0])
    t1, _, _ = o.preprocess(frames[1])
    ref = o.det.exe.run(np.stack([t0, t1]), taps="all")
    checked, worst = 0, 0.0
    for tf_idx, want in ref.items():
        try:
            got = d.debugTensor(0, tf_idx, 2)
        except ValueError:
            continue                                  # fused away in this plan
        want = want.reshape(got.shape)
        scale = max(np.abs(want).max(), 1e-6)
        err = np.abs(got - want).max() / scale
        worst = max(worst, err)
        assert err <= HEAD_REL_TOL, "tensor %d (%s): rel err %.3e" % (tf_idx, o.det.model.tensors[tf_idx].name, err)
        checked += 1
    assert checked >= (20 if fuse else 60)
    print("%s fuse=%d: %d tensors, worst rel err %.2e" % (model, fuse, checked, worst))


@pytest.mark.parametrize("model", ["shortRange", "full", "backCamera"])
def test_raw_heads_and_candidates(fdt, model_bytes, sample_images, golden, model):
    d = get_detector(fdt, model)
    o = get_oracle(model_bytes, model)
    for name, img in sample_images.items():
        h, w = img.shape[:2]
        d.detectBatchRaw(img, count=1, width=w, height=h)
        boxes, scores = d.debugRawHeads(1)
        gb, gs = golden["%s/%s/boxes_f64" % (model, name)], golden["%s/%s/scores_f64" % (model, name)]
        assert np.abs(boxes[0] - gb).max() <= HEAD_REL_TOL * np.abs(gb).max()
        assert np.abs(scores[0] - gs).max() <= HEAD_REL_TOL * np.abs(gs).max()
        cand = d.debugCandidates(0)
        assert np.array_equal(cand, golden["%s/%s/candidates" % (model, name)])       # index set: bit-exact
        # margin: logits near the threshold (0.0) must sit far above their own numerical noise
        near = np.abs(gs) < 1.0
        if near.any():
            assert (np.abs(gs[near]) > 10 * np.abs(scores[0][near] - gs[near])).all()


@pytest.mark.parametrize("model", ["shortRange", "full", "backCamera"])
def test_detections_match_oracle(fdt, model_bytes, sample_images, golden, model):
    d = get_detector(fdt, model)
    o = get_oracle(model_bytes, model)
    expected = {("shortRange", "landmark-ex1.jpg"): 1, ("shortRange", "iris-detection-ex1.jpg"): 1,
                ("shortRange", "group-shot-bounding-box-ex1.jpeg"): 0, ("backCamera", "group-shot-bounding-box-ex1.jpeg"): 4,
                ("backCamera", "landmark-ex1.jpg"): 1, ("full", "landmark-ex1.jpg"): 1}
    for name, img in sample_images.items():
        h, w = img.shape[:2]
        faces = d.detectFacesFromMatBytes(img.tobytes(), width=w, height=h, mode=fdt.FaceDetectionMode.fast)
        g = golden["%s/%s/dets" % (model, name)]
        want = [dp.Detection(*row[:5], list(row[5:17]), int(row[17])) for row in g]
        assert_faces_match(faces, want)
        if (model, name) in expected:
            assert len(faces) == expected[(model, name)]               # pinned by the reference's integration tests
        for f in faces:
            assert 0.5 <= f.score <= 1.0
            bb = f.boundingBox
            assert bb.width == pytest.approx(f.detectionData.boundingBox.w * w) and len(f.landmarks) == 6
    for k, fr in enumerate(_frames_for(model, sample_images)):
        h, w = fr.shape[:2]
        faces = d.detectFacesFromMat(fr, mode=fdt.FaceDetectionMode.fast)
        assert_faces_match(faces, o.detect(fr))


def test_batch_equals_single_and_chunking(fdt, model_bytes):
    from face_detection_tflite_b200 import synth
    frames = np.concatenate([synth.face_frames(9, 640, 360, max_side=300), synth.noise_frames(2, 640, 360)])
    small = get_detector(fdt, "shortRange", mesh=False, maxBatch=4)       # 11 frames -> 3 chunks, both streams
    big = get_detector(fdt, "shortRange")
    a = small.detectFacesBatch(frames, count=11, width=640, height=360)
    b = big.detectFacesBatch(frames, count=11, width=640, height=360)
    c = [big.detectFacesFromMat(f, mode=fdt.FaceDetectionMode.fast) for f in frames]
    o = get_oracle(model_bytes, "shortRange", "cv2dnn")
    assert sum(len(x) for x in a) >= 5
    for fa, fb, fc, fr in zip(a, b, c, frames):
        assert len(fa) == len(fb) == len(fc) == len(o.detect(fr))
        for x, y, z in zip(fa, fb, fc):
            assert x.detectionData == y.detectionData == z.detectionData          # bitwise identical across batchings
    # permutation property: results follow their frames
    perm = np.random.default_rng(0).permutation(11)
    p = big.detectFacesBatch(frames[perm], count=11, width=640, height=360)
    for i, j in enumerate(perm):
        assert [f.detectionData for f in p[i]] == [f.detectionData for f in b[j]]
    # device-resident input gives the same answer as host input
    import torch
    dev = torch.from_numpy(frames).cuda()
    faces, counts, _ = big.detectBatchRaw(dev.data_ptr(), count=11, width=640, height=360, memKind=1)
    assert list(counts) == [len(x) for x in b]


def test_edge_cases(fdt):
    d = get_detector(fdt, "shortRange")
    fast = fdt.FaceDetectionMode.fast
    for h, w in ((1, 1), (10, 10), (50, 50)):                               # edge_cases_test.dart:44-83 -> []
        assert d.detectFacesFromMat(np.full((h, w, 3), 127, np.uint8), mode=fast) == []
    grad = np.tile(np.arange(256, dtype=np.uint8)[None, :, None], (200, 1, 3))
    assert d.detectFacesFromMat(grad, mode=fast) == []
    assert isinstance(d.detectFacesFromMat(np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8), mode=fast), list)
    for h, w in ((2160, 3840), (50, 2000), (2000, 50)):                      # 4K and extreme aspect ratios do not crash
        assert isinstance(d.detectFacesFromMat(np.zeros((h, w, 3), np.uint8), mode=fast), list)
    with pytest.raises(ValueError):                                          # helpers.dart:440-447 length check
        d.detectFacesFromMatBytes(b"\x00" * 10, width=4, height=4, mode=fast)
    assert d.detectFacesFromMatBytes(b"\x00" * 48, width=4, height=4) == []       # default mode (full): [] on a tiny image
    assert d.detectFacesBatch(np.zeros((0,), np.uint8), count=0, width=8, height=8) == []
    with pytest.raises(ValueError):
        fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, minScore=2.0)
    t = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withMesh=False)
    with pytest.raises(fdt.StateError):
        t.initialize(fdt.FaceDetectionModel.shortRange)                      # double initialise (face_detector.dart:315-317)
    with pytest.raises(fdt.StateError):
        t.detectFacesFromMatBytes(b"\x00" * 48, width=4, height=4, mode=fdt.FaceDetectionMode.standard)   # no mesh model
    t.dispose()
    with pytest.raises(fdt.StateError):
        t.detectFacesFromMatBytes(b"\x00" * 48, width=4, height=4, mode=fast)


def test_gates(fdt, sample_images):
    img = sample_images["group-shot-bounding-box-ex1.jpeg"]
    h, w = img.shape[:2]
    base = get_detector(fdt, "backCamera").detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    assert len(base) == 4
    thr = sorted(f.score for f in base)[1] + 1e-9
    g = get_detector(fdt, "backCamera", mesh=False, minScore=thr).detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    assert [f.detectionData for f in g] == [f.detectionData for f in base if f.score >= thr]
    widths = sorted(f.detectionData.boundingBox.w for f in base)
    cut = float(widths[1] + widths[2]) / 2.0
    g = get_detector(fdt, "backCamera", mesh=False, minFaceSize=cut).detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    assert len(g) == 2 and all(f.detectionData.boundingBox.w > cut for f in g)


# ---- mesh stage (C4) -----------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["backCamera", "shortRange", "full"])
def test_standard_mode_mesh(fdt, model_bytes, sample_images, golden, model):
    """Warp: BIT-EXACT against the oracle's cv::warpAffine restatement fed with the GPU's own detections
    (the ROI is a function of the fp32 keypoints, so it is compared like-for-like); mesh net: raw outputs
    vs the fp64 oracle on the same crop; end to end: mesh points / scores vs the all-oracle pipeline."""
    from oracle import geometry as geo
    d = get_detector(fdt, model, minFacePresenceConfidence=0.0)      # gate off: see every face's mesh
    gated = get_detector(fdt, model)
    o = get_oracle(model_bytes, model)
    std = fdt.FaceDetectionMode.standard
    for name, img in sample_images.items():
        h, w = img.shape[:2]
        faces = d.detectFacesFromMat(img, mode=std)
        want_all = OraclePipeline.detect_faces(_NoGate(o), img, "standard")
        assert len(faces) == len(want_all)
        crops, raw, flag = d.debugMeshStage(max(len(faces), 1))
        for i, (g, wf) in enumerate(zip(faces, want_all)):
            kp = g.detectionData.keypointsXY
            theta, cx, cy, size = geo.compute_face_alignment(kp, float(w), float(h))
            want_crop = co.extract_aligned_square(img, cx, cy, size, -theta, 192)
            assert np.array_equal(crops[i], want_crop)                    # the ROI matrix comes from the host libm: bit-exact
            ref = o.mesh.run(co.normalize_bgr_u8(crops[i])[None])
            li = int(np.argmax([x.shape[1] for x in ref]))
            assert np.abs(raw[i] - ref[li][0]).max() <= HEAD_REL_TOL * np.abs(ref[li][0]).max()
            assert abs(flag[i] - ref[1 - li][0][0]) <= HEAD_REL_TOL * max(1.0, abs(ref[1 - li][0][0]))
            # end to end vs the all-oracle pipeline: <= 1e-3 of the ROI side (normalised mesh coordinates)
            assert g.mesh is not None and len(g.mesh) == 468                # face_detection_integration_test.dart:124-135
            # (the ROI follows the fp32 keypoints, so a handful of crop pixels differ by a grey level and the
            #  face-flag logit moves with them: 5e-3 on the sigmoid end to end, 1e-4 like-for-like above)
            assert abs(g.meshScore - wf.mesh_score) <= 5e-3 and 0.0 <= g.meshScore <= 1.0
            assert np.abs(g.mesh.packed - wf.mesh_px).max() <= COORD_TOL * wf.align[3]
        if faces and "%s/%s/crop0_cv2" % (model, name) in golden:
            dg = np.abs(crops[0].astype(int) - golden["%s/%s/crop0_cv2" % (model, name)].astype(int))
            assert dg.max() <= 8 and (dg > 0).mean() <= 0.05                # the real cv2.warpAffine on the f64 ROI
        # presence gate (default 0.5) keeps exactly the faces the oracle keeps
        kept = gated.detectFacesFromMat(img, mode=std)
        want = o.detect_faces(img, "standard")
        assert len(kept) == len(want) == len(golden["%s/%s/mesh_scores" % (model, name)])
        for g, wf in zip(kept, want):
            assert g.anchorIndex == wf.det.anchor and g.meshScore >= 0.5
    # batched standard mode == per-frame standard mode
    img = sample_images["group-shot-bounding-box-ex1.jpeg"]
    h, w = img.shape[:2]
    two = gated.detectFacesBatch(np.stack([img, img[:, ::-1].copy()]), count=2, width=w, height=h, mode=std)
    one = gated.detectFacesFromMat(img, mode=std)
    assert [f.detectionData for f in two[0]] == [f.detectionData for f in one]
    assert all(np.array_equal(a.mesh.packed, b.mesh.packed) for a, b in zip(two[0], one))


class _NoGate:
    """The oracle pipeline with the presence gate off, to compare every crop the GPU produced."""
    def __init__(self, o):
        self.__dict__.update(o.__dict__)
        self.min_presence = 0.0
    detect = OraclePipeline.detect
    preprocess = OraclePipeline.preprocess
    raw_heads = OraclePipeline.raw_heads


def test_mesh_layers(fdt, model_bytes, sample_images):
    """Every materialised tensor of the face_landmark graph (fuse levels 0 and 2) vs the fp64 oracle."""
    img = sample_images["landmark-ex1.jpg"]
    h, w = img.shape[:2]
    o = get_oracle(model_bytes, "backCamera")
    crop = OraclePipeline.detect_faces(_NoGate(o), img, "standard")[0].crop
    ref = o.mesh.exe.run(co.normalize_bgr_u8(crop)[None], taps="all")
    for fuse in (0, 2):
        d = get_detector(fdt, "backCamera", fuse=fuse, maxBatch=4)
        d.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.standard)
        crops, _, _ = d.debugMeshStage(1)
        if not np.array_equal(crops[0], crop):
            ref_f = o.mesh.exe.run(co.normalize_bgr_u8(crops[0])[None], taps="all")
        else:
            ref_f = ref
        n = 0
        for tf_idx, want in ref_f.items():
            try:
                got = d.debugTensor(1, tf_idx, 1)
            except ValueError:
                continue
            want = want.reshape(got.shape)
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
            assert err <= HEAD_REL_TOL, "mesh tensor %d: %.3e (fuse %d)" % (tf_idx, err, fuse)
            n += 1
        assert n >= 10


# ---- full benchmark sizes: size-independent properties -------------------------------------------------
def test_full_size_c2_properties(fdt, model_bytes):
    """BASELINE config 2 at full size (4096 x 1280x720): determinism, tiling/periodicity, and agreement
    of a sampled subset with the oracle."""
    import torch
    from face_detection_tflite_b200 import synth
    d = get_detector(fdt, "shortRange")
    period = 64
    base = np.concatenate([synth.face_frames(period - 8, 1280, 720), synth.noise_frames(8, 1280, 720)])
    dev = torch.from_numpy(base).cuda().repeat(4096 // period, 1, 1, 1).contiguous()
    f1, c1, _ = d.detectBatchRaw(dev.data_ptr(), count=4096, width=1280, height=720, memKind=1)
    c1 = c1.copy()
    a1 = np.frombuffer(f1, np.uint8).copy()
    f2, c2, _ = d.detectBatchRaw(dev.data_ptr(), count=4096, width=1280, height=720, memKind=1)
    assert np.array_equal(c1, c2)
    rec = np.frombuffer(f2, np.uint8).reshape(4096, -1)
    a1 = a1.reshape(4096, -1)
    for b in range(4096):                                                     # determinism (valid slots only)
        n = c1[b] * C.sizeof(type(f1[0]))
        assert np.array_equal(a1[b, :n], rec[b, :n])
    assert np.array_equal(c1.reshape(-1, period), np.tile(c1[:period], (4096 // period, 1)))   # periodic input -> periodic output
    assert c1[period - 8:period].sum() == 0                                   # noise frames: no faces
    assert c1[:period].sum() >= 40
    o = get_oracle(model_bytes, "shortRange", "cv2dnn")
    for k in (1, 7, 18, 33, 49):
        assert c1[k] == len(o.detect(base[k]))


# ---- more coverage of the boundary ------------------------------------------------------------------
def test_mat_types_give_same_detections(fdt, sample_images):
    """matType 24 (BGRA) and 0 (GRAY) are colour-converted like bgrMatToSignedFloat32 does (helpers.dart:386-392)."""
    import cv2
    d = get_detector(fdt, "backCamera")
    img = sample_images["landmark-ex1.jpg"]
    fast = fdt.FaceDetectionMode.fast
    base = d.detectFacesFromMat(img, mode=fast)
    bgra = cv2.cvtColor(img, cv2.COLOR_BGR2BGRA)
    assert [f.detectionData for f in d.detectFacesFromMat(bgra, mode=fast)] == [f.detectionData for f in base]
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    g3 = cv2.cvtColor(gray, cv2.COLOR_GRAY2BGR)
    assert [f.detectionData for f in d.detectFacesFromMat(gray, mode=fast)] == [f.detectionData for f in d.detectFacesFromMat(g3, mode=fast)]
    # standard mode on BGRA: the warp reads the 4-channel frame directly
    a = d.detectFacesFromMat(bgra, mode=fdt.FaceDetectionMode.standard)
    b = d.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.standard)
    assert len(a) == len(b) == 1 and np.array_equal(a[0].mesh.packed, b[0].mesh.packed)


def test_random_frame_sizes_letterbox_bit_exact(fdt):
    rng = np.random.default_rng(42)
    d = get_detector(fdt, "full", mesh=False)
    for _ in range(12):
        h, w = int(rng.integers(1, 900)), int(rng.integers(1, 1400))
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        d.detectBatchRaw(frame, count=1, width=w, height=h)
        want, _ = co.letterbox_u8(frame, 192, 192)
        assert np.array_equal(d.debugLetterboxed(1)[0], want), (h, w)


def test_max_faces_truncation_and_device_frames_standard_mode(fdt, sample_images):
    import torch
    img = sample_images["group-shot-bounding-box-ex1.jpeg"]
    h, w = img.shape[:2]
    full = get_detector(fdt, "backCamera").detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    two = get_detector(fdt, "backCamera", maxFaces=2)
    t = two.detectFacesFromMat(img, mode=fdt.FaceDetectionMode.fast)
    assert [f.detectionData for f in t] == [f.detectionData for f in full[:2]]
    # device-resident frames in standard mode: same meshes as host frames
    frames = np.stack([img, img])
    dev = torch.from_numpy(frames).cuda()
    fa, ca, ma = two.detectBatchRaw(dev.data_ptr(), count=2, width=w, height=h, mode=fdt.FaceDetectionMode.standard, memKind=1)
    fb, cb, mb = two.detectBatchRaw(frames, count=2, width=w, height=h, mode=fdt.FaceDetectionMode.standard)
    assert list(ca) == list(cb) == [2, 2] and np.array_equal(ma, mb)


def test_two_handles_in_threads_are_independent_and_deterministic(fdt, sample_images):
    """concurrency_stress_test.dart:130-162 (parallel detectors) and
    face_detection_integration_test.dart:929-955 (identical results across instances)."""
    import threading
    img = sample_images["landmark-ex1.jpg"]
    dets = [fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, maxBatch=8) for _ in range(2)]
    out = [None, None]

    def work(i):
        res = []
        for _ in range(10):
            res.append(dets[i].detectFacesFromMat(img, mode=fdt.FaceDetectionMode.standard))
        out[i] = res

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    ref = out[0][0]
    assert len(ref) == 1
    for res in out:
        for r in res:
            assert [f.detectionData for f in r] == [f.detectionData for f in ref]
            assert np.array_equal(r[0].mesh.packed, ref[0].mesh.packed)
    [d.dispose() for d in dets]


def test_c_abi_status_codes(fdt, lib):
    import ctypes as C
    from face_detection_tflite_b200 import _ffi
    d = get_detector(fdt, "shortRange")
    h = d._h
    faces = (_ffi.FdtFace * 100)()
    cnt = C.c_int32()
    buf = np.zeros(48, np.uint8)
    assert lib.fdt_detect_one(h, buf.ctypes.data, 47, 4, 4, 16, 0, faces, C.byref(cnt), None, None) == _ffi.FDT_ERR_SIZE_MISMATCH
    assert b"length" in lib.fdt_last_error(h)
    assert lib.fdt_detect_one(h, buf.ctypes.data, 48, 4, 4, 17, 0, faces, C.byref(cnt), None, None) == _ffi.FDT_ERR_BAD_ARG      # unknown matType
    assert lib.fdt_detect_one(h, buf.ctypes.data, 48, 4, 4, 16, 2, faces, C.byref(cnt), None, None) == _ffi.FDT_OK and cnt.value == 0   # mode full
    assert lib.fdt_detect_one(h, buf.ctypes.data, 48, 4, 4, 16, 7, faces, C.byref(cnt), None, None) == _ffi.FDT_ERR_BAD_ARG
    assert lib.fdt_detect_batch(h, buf.ctypes.data, 1, 4, 4, 8, 16, 0, 0, faces, C.byref(cnt), None, None) == _ffi.FDT_ERR_SIZE_MISMATCH  # row_stride < w*3
    assert lib.fdt_detect_batch(h, buf.ctypes.data, 1, 4, 4, 12, 16, 0, 0, None, None, None, None) == _ffi.FDT_ERR_BAD_ARG
    assert lib.fdt_detect_one(h, buf.ctypes.data, 48, 4, 4, 16, 0, faces, C.byref(cnt), None, None) == _ffi.FDT_OK and cnt.value == 0
    iw, ih, na, mf, mb = (C.c_int32() for _ in range(5))
    assert lib.fdt_get_info(h, C.byref(iw), C.byref(ih), C.byref(na), C.byref(mf), C.byref(mb)) == 0
    assert (iw.value, ih.value, na.value, mf.value) == (128, 128, 896, 100)
    assert lib.fdt_last_h2d_bytes(h) == 48 and lib.fdt_last_launch_count(h) == 10   # letterbox + stem + 6 blocks + image-resident tail (10 blocks, 2 head pairs) + decode


# ---- alternative kernel paths (environment switches are read once per process -> subprocesses) ----------
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
_VARIANT_SCRIPT = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, cv2
import face_detection_tflite_b200 as fdt
from oracle.pipeline import OraclePipeline
from pathlib import Path
root = Path(%r)
img = cv2.imread(str(root / "assets/samples/group-shot-bounding-box-ex1.jpeg"))
for model, f in (("shortRange", "face_detection_short_range.tflite"), ("backCamera", "face_detection_back.tflite"),
                 ("full", "face_detection_full_range.tflite")):
    det_bytes = (root / "assets/models" / f).read_bytes()
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], withMesh=False)
    frames = np.stack([img, img[:, ::-1].copy(), img[::-1].copy()])
    d.detectBatchRaw(frames, count=3, width=img.shape[1], height=img.shape[0])
    boxes, scores = d.debugRawHeads(3)
    o = OraclePipeline(det_bytes, model, None, "f64")
    ref = o.det.run(np.stack([o.preprocess(fr)[0] for fr in frames]))
    for got, want in ((boxes, ref[0]), (scores, ref[1])):
        want = np.asarray(want).reshape(got.shape)
        err = float(np.abs(got - want).max() / np.abs(want).max())
        assert err <= 1e-4, (model, err)                     # north_star: 1e-4 relative on raw head outputs
    d.dispose()
print("variant ok")
"""


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"FDT_WS_NO": "2"},                 # k_block_ws with the TMA-store epilogue (output tile in shared memory + store warp)
                                 {"FDT_TS": "0"},                    # k_block_ws (TF32 hi/lo, operand in shared memory) for every BlazeBlock
                                 {"FDT_TS": "2"},                    # (historical switch: k_block_ts for the 24 -> 28 block too; the default since the taps come from the constant bank)
                                 {"FDT_TAIL": "0"},                  # one launch per 16x16 / 8x8 block and head pair instead of k_tail_ws
                                 {"FDT_CHAIN": "0"},                 # per-layer kernels instead of the image-resident chains (k_chain_wide, mesh trunk)
                                 {"FDT_TAIL": "0", "FDT_TS": "0"}])  # the round-1 plan: k_block_ws everywhere
def test_kernel_variants_match_the_oracle(env):
    """Every tuning switch selects code that must stay parity-green: raw heads of two models vs the fp64 oracle."""
    import subprocess
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % (str(ROOT), str(ROOT))], env=e, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "variant ok" in r.stdout, (env, r.stdout[-500:], r.stderr[-1500:])


@pytest.mark.gpu
def test_wide_chains_walk_several_images_per_cta(fdt, model_bytes):
    """k_chain_wide (full-range layers wider than 128 channels): 400 frames in one chunk, so that every CTA of the image-resident
    chains processes two or three images (buffer hand-over, W-block ring phases, loader warp).  The input repeats with period 16:
    the same frame must give bit-identical heads wherever it sits in the chunk, and the first frames match the fp64 oracle."""
    from face_detection_tflite_b200 import synth
    n, per = 400, 16
    base = np.concatenate([synth.face_frames(per - 2, 640, 360, start=2), synth.noise_frames(2, 640, 360)])
    frames = np.ascontiguousarray(np.concatenate([base] * (n // per))[:n])
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel.full, withMesh=False, maxBatch=n)
    faces, counts, _ = d.detectBatchRaw(frames, count=n, width=640, height=360)
    boxes, scores = d.debugRawHeads(n)
    for i in range(per, n):
        assert np.array_equal(boxes[i], boxes[i % per]) and np.array_equal(scores[i], scores[i % per]), i
    assert int(counts.sum()) >= n // 2
    o = get_oracle(model_bytes, "full")
    k = 4
    want = o.det.run(np.stack([o.preprocess(fr)[0] for fr in frames[:k]]))
    for got, w in ((boxes[:k], want[0]), (scores[:k], want[1])):
        w = np.asarray(w).reshape(got.shape)
        assert np.abs(got - w).max() <= HEAD_REL_TOL * np.abs(w).max()
    d.dispose()

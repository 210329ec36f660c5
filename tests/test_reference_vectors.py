"""The oracle (and the library's host-side geometry) against the reference's own unit-test vectors, verbatim
(tests/reference_vectors.py cites every Dart test).  The GPU suite feeds the same vectors to the CUDA kernels."""
import ctypes as C
import math

import numpy as np
import pytest

import reference_vectors as rv
from oracle import detect_post as dp, geometry as geo


def _dets(items, kp=None):
    return [dp.Detection(b[0], b[1], b[2], b[3], s, list(kp if kp is not None else []), i) for i, (b, s) in enumerate(items)]


def test_sigmoid_clipped_limit():
    for x, want in rv.SIGMOID_LIMIT2:
        assert dp.sigmoid_clipped(x, limit=2) == pytest.approx(want, abs=1e-4)


def test_letterbox_removal_vector():
    v = rv.LETTERBOX_REMOVAL
    d = dp.Detection(*v["box"], v["score"], list(v["kp"]), 0)
    out = dp.letterbox_removal([d], v["padding"])
    assert len(out) == 1
    assert out[0].xmin == pytest.approx(v["expect_xmin"], abs=v["tol"])
    assert out[0].kp[0] == pytest.approx(v["expect_kp0"], abs=v["tol"])


def test_unpack_landmarks_vector():
    v = rv.UNPACK_LANDMARKS
    out = geo.unpack_landmarks(np.array(v["flat"], np.float32), v["w"], v["h"], v["padding"], clamp=True)
    assert len(out) == v["expect"]["len"]
    for got, want in zip(out[0], v["expect"]["p0"]):
        assert got == pytest.approx(want, abs=v["tol"])


def test_nms_private_vector():
    v = rv.NMS_PRIVATE
    assert len(dp.weighted_nms(_dets(v["dets"]), v["iou"], v["score"])) == v["expect_len"]


@pytest.mark.parametrize("case", rv.NMS_COVERAGE, ids=[c["name"] for c in rv.NMS_COVERAGE])
def test_nms_coverage_vectors(case):
    out = dp.weighted_nms(_dets(case["dets"], rv.NMS_KP), case["iou"], case["score"])
    assert len(out) == case["expect_len"]
    if "expect_xmin_gt" in case:
        assert out[0].xmin > case["expect_xmin_gt"]
    if "expect_score0" in case:
        assert out[0].score == case["expect_score0"]


def oracle_decode(case):
    """decodeBlazeFaceCandidates (lib/src/web/detection_decode.dart:44-88) in terms of the oracle's native-path
    restatement: candidates (score >= kMinScore <=> raw >= 0, NaN rejected), decode, degenerate filter."""
    scores = np.array(case["scores"], np.float32)
    boxes = np.array(case["boxes"], np.float32)
    idx, sc = dp.collect_candidates(scores)
    dec = dp.decode_boxes(boxes, np.array(rv.DECODE_ANCHORS), idx, int(rv.DECODE_SCALE))
    return dp.to_detections_filtered(dec, sc, idx)


@pytest.mark.parametrize("case", rv.DECODE_CASES, ids=[c["name"] for c in rv.DECODE_CASES])
def test_web_decode_vectors(case):
    dets = oracle_decode(case)
    assert [d.anchor for d in dets] == case["expect_anchors"]
    for d in dets:                                           # each survivor carries its OWN score
        assert d.score == dp.sigmoid_clipped(case["scores"][d.anchor])
    if "expect_xmin1" in case:
        assert dets[1].xmin == pytest.approx(case["expect_xmin1"], abs=case["tol"])
    if case["name"] == "no degenerate box":
        # value-identical to the inline reference implementation of the test (f32 tmp, f64 arithmetic)
        for d in dets:
            row, (ax, ay) = case["boxes"][d.anchor], rv.DECODE_ANCHORS[d.anchor]
            tmp = [np.float32(np.float32(v) / rv.DECODE_SCALE) for v in row]
            tmp[0] = np.float32(float(tmp[0]) + ax); tmp[1] = np.float32(float(tmp[1]) + ay)
            for j in range(4, 16, 2):
                tmp[j] = np.float32(float(tmp[j]) + ax); tmp[j + 1] = np.float32(float(tmp[j + 1]) + ay)
            xc, yc, w, h = (float(t) for t in tmp[:4])
            assert (d.xmin, d.ymin, d.xmax, d.ymax) == (xc - w * 0.5, yc - h * 0.5, xc + w * 0.5, yc + h * 0.5)
            assert d.kp == [float(t) for t in tmp[4:]]


def _mesh_with(corners):
    m = np.zeros((468, 3))
    for k, (x, y) in corners.items():
        m[k, :2] = (x, y)
    return m


def _check_roi(roi, want):
    got = {"cx": roi[0], "cy": roi[1], "size": roi[2], "theta": roi[3]}
    for k, v in want.items():
        assert got[k] == pytest.approx(v, abs=1e-9)


@pytest.mark.parametrize("case", rv.EYE_ROIS)
def test_eye_rois_vectors(case, lib):
    rois = geo.eye_rois_from_mesh(_mesh_with(case["corners"]))
    assert len(rois) == 2
    _check_roi(rois[0], case["left"]); _check_roi(rois[1], case["right"])
    # the library's host restatement (fdt_host_eye_rois) is bit-identical to the oracle's
    c8 = (C.c_double * 8)(*[float(v) for k in (33, 133, 362, 263) for v in case["corners"][k]])
    out8 = (C.c_double * 8)()
    assert lib.fdt_host_eye_rois(c8, out8) == 0
    assert [tuple(out8[0:4]), tuple(out8[4:8])] == [tuple(rois[0]), tuple(rois[1])]


def test_iris_transform_vectors():
    roi = rv.IRIS_ROI
    for right in (False, True):                              # centre -> ROI centre for both eyes
        out = geo.transform_iris_norm_to_absolute([[0.5, 0.5, 0.0]], roi, right)
        assert out[0][0] == pytest.approx(roi[0], abs=1e-9) and out[0][1] == pytest.approx(roi[1], abs=1e-9)
    left = geo.transform_iris_norm_to_absolute([[0.25, 0.5, 0.0]], roi, False)
    right = geo.transform_iris_norm_to_absolute([[0.25, 0.5, 0.0]], roi, True)
    assert left[0][0] == pytest.approx(roi[0] - 0.25 * roi[2], abs=1e-9)
    assert right[0][0] == pytest.approx(roi[0] + 0.25 * roi[2], abs=1e-9)
    assert left[0][1] == pytest.approx(right[0][1], abs=1e-9)
    assert geo.transform_iris_norm_to_absolute([[0.5, 0.5, 0.75]], roi, False)[0][2] == pytest.approx(0.75, abs=1e-9)
    out = geo.transform_iris_norm_to_absolute([[1.0, 0.5, 0.0]], rv.IRIS_ROTATED, False)
    assert out[0][0] == pytest.approx(0.0, abs=1e-9) and out[0][1] == pytest.approx(50.0, abs=1e-9)


def test_iris_center_from_points():
    pts = [(0, 0, 0), (10, 0, 0), (10, 10, 0), (0, 10, 0), (5.5, 5.0, 1)]
    assert geo.iris_center_from_points(pts) == (5.5, 5.0, 1.0)
    assert geo.iris_center_from_points([]) == (0.0, 0.0, 0.0)
    assert geo.iris_center_from_points([(3, 4, 5)]) == (3.0, 4.0, 5.0)
    import face_detection_tflite_b200 as fdt
    P = fdt.Point
    assert fdt.irisCenterFromPoints([P(*p) for p in pts]) == P(5.5, 5.0, 1)


@pytest.mark.parametrize("case", rv.EMBED_ALIGN)
def test_embedding_alignment_vectors(case, lib):
    theta, cx, cy, size = geo.compute_embedding_alignment(case["l"], case["r"])
    got = {"theta": theta, "cx": cx, "cy": cy, "size": size}
    for k in ("theta", "cx", "size"):
        if k in case:
            assert got[k] == pytest.approx(case[k][0], abs=case[k][1])
    if "cy_gt" in case:
        assert cy > case["cy_gt"]
    if "theta_lt" in case:
        assert theta < case["theta_lt"]
    out4 = (C.c_double * 4)()
    assert lib.fdt_host_embedding_roi((C.c_double * 2)(*case["l"]), (C.c_double * 2)(*case["r"]), out4) == 0
    assert tuple(out4) == (theta, cx, cy, size)               # host restatement == oracle, bit for bit


def test_normalize_embedding_and_wire_maps():
    e = geo.normalize_embedding(np.array([3.0, 4.0], np.float32))
    assert np.allclose(e, [0.6, 0.8]) and e.dtype == np.float32
    assert np.array_equal(geo.normalize_embedding(np.zeros(4, np.float32)), np.zeros(4, np.float32))
    import face_detection_tflite_b200 as fdt
    det = fdt.Detection(fdt.RectF(0.1, 0.2, 0.3, 0.4), 0.9, [0.5] * 12, fdt.Size(640.0, 480.0))
    mesh = fdt.FaceMesh(np.arange(468 * 3, dtype=np.float32).reshape(468, 3), 0.75)
    iris = [fdt.Point(float(i), float(i) + 0.5, 1.0) for i in range(152)]
    f = fdt.Face(det, mesh, fdt.Size(640.0, 480.0), irisPoints=iris, trackingId=7)
    m = f.toMap()                                             # Face.toMap / fromMap (face_types.dart:1350-1378)
    assert set(m) == {"detection", "trackingId", "mesh", "irisPoints", "originalSize"}
    g = fdt.Face.fromMap(m)
    assert g.detectionData == f.detectionData and g.trackingId == 7 and g.irisPoints == iris
    assert np.array_equal(g.mesh.packed, mesh.packed) and g.mesh.score == 0.75
    assert "mesh" not in fdt.Face(det, None, fdt.Size(640.0, 480.0)).toMap()
    eyes = f.eyes                                             # Face._parseIris: 71 eye-mesh points + 5 iris points per eye
    assert len(eyes.leftEye.mesh) == 71 and len(eyes.leftEye.irisContour) == 4 and fdt.Face(det, None, fdt.Size(1.0, 1.0)).eyes is None
    # applyFaceGates (face_gates.dart:84-104)
    assert f.widthFraction == pytest.approx(0.2)
    assert fdt.applyFaceGates([f], minScore=0.0, minFaceSize=0.0) == [f]
    assert fdt.applyFaceGates([f], minScore=0.95, minFaceSize=0.0) == []
    assert fdt.applyFaceGates([f], minScore=0.0, minFaceSize=0.25) == []
    assert fdt.applyFaceGates([f], minScore=0.0, minFaceSize=0.0, minFacePresenceConfidence=0.8) == []
    nomesh = fdt.Face(det, None, fdt.Size(640.0, 480.0))      # a null mesh score always passes the presence gate
    assert fdt.applyFaceGates([nomesh], minScore=0.0, minFaceSize=0.0, minFacePresenceConfidence=0.8) == [nomesh]
    with pytest.raises(ValueError):
        fdt.validateFaceGates(minScore=1.5, minFaceSize=0.0)


def test_iris_parsing_vectors():
    """test/iris_parsing_test.dart:47-137: Face.eyes / _parseIris on 5-, 76- and odd-length iris lists."""
    import face_detection_tflite_b200 as fdt
    P = fdt.Point
    det = fdt.Detection(fdt.RectF(0.1, 0.1, 0.9, 0.9), 0.9, [0.5] * 12, fdt.Size(200.0, 200.0))
    mk = lambda pts: fdt.Face(det, None, fdt.Size(200.0, 200.0), irisPoints=pts)
    assert mk([]).eyes is None and mk([P(1.0, 1.0)] * 4).eyes is None
    five = [P(100.0, 100.0), P(95.0, 100.0), P(105.0, 100.0), P(100.0, 95.0), P(100.0, 105.0)]
    e = mk(five).eyes
    assert e.leftEye is not None and e.rightEye is None
    assert e.leftEye.irisCenter == P(100.0, 100.0) and len(e.leftEye.irisContour) == 4 and e.leftEye.mesh == []
    for q in (P(95.0, 100.0), P(105.0, 100.0), P(100.0, 95.0), P(100.0, 105.0)):
        assert q in e.leftEye.irisContour
    pts76 = [P(float(i), 50.0) for i in range(71)] + [P(85.0, 55.0), P(80.0, 55.0), P(90.0, 55.0), P(85.0, 50.0), P(85.0, 60.0)]
    e = mk(pts76).eyes
    assert e.rightEye is None and len(e.leftEye.mesh) == 71 and e.leftEye.irisCenter == P(85.0, 55.0)
    assert e.leftEye.mesh[0] == P(0.0, 50.0) and e.leftEye.mesh[70] == P(70.0, 50.0) and len(e.leftEye.irisContour) == 4
    e = mk(pts76 + pts76).eyes
    assert len(e.leftEye.mesh) == 71 and len(e.rightEye.mesh) == 71

"""Known-answer tests for the oracle's restatement of the Dart glue, taken from the reference's own
unit tests, plus the pinned sample-image face counts and the committed fixtures."""
import math

import numpy as np
import pytest

from oracle import detect_post as dp, geometry as geo
from oracle.pipeline import OraclePipeline


def det(xmin, ymin, xmax, ymax, score):
    return dp.Detection(xmin, ymin, xmax, ymax, score, [0.5] * 12)


# ---- anchors: test/helpers_coverage_test.dart:329-406 ------------------------------------------------
def test_anchor_counts_and_range():
    front = dp.generate_anchors(dp.ssd_options_for("frontCamera"))
    assert front.shape == (896, 2)
    assert np.array_equal(front, dp.generate_anchors(dp.ssd_options_for("shortRange")))
    assert dp.generate_anchors(dp.ssd_options_for("backCamera")).shape == (896, 2)
    full = dp.generate_anchors(dp.ssd_options_for("full"))
    assert full.shape == (48 * 48, 2)
    assert np.array_equal(full, dp.generate_anchors(dp.ssd_options_for("fullSparse")))
    for a in (front, full):
        assert (a > 0).all() and (a <= 1).all()


def test_anchor_layout_matches_head_reshape_order():
    a = dp.generate_anchors(dp.ssd_options_for("shortRange"))
    assert tuple(a[0]) == (0.5 / 16, 0.5 / 16) and tuple(a[1]) == tuple(a[0])       # 2 anchors per 16x16 cell
    assert tuple(a[2]) == (1.5 / 16, 0.5 / 16)
    assert tuple(a[512]) == (0.5 / 8, 0.5 / 8) and tuple(a[517]) == tuple(a[512])   # 6 per 8x8 cell
    f = dp.generate_anchors(dp.ssd_options_for("full"))
    assert tuple(f[49]) == (1.5 / 48, 1.5 / 48)


# ---- NMS: test/helpers_coverage_test.dart:220-291, test/helpers_private_test.dart:50-71 ---------------
def test_nms_empty_and_threshold():
    assert dp.weighted_nms([], 0.5, 0.5) == []
    assert dp.weighted_nms([det(0, 0, .5, .5, .3)], 0.5, 0.5) == []


def test_nms_keeps_disjoint():
    assert len(dp.weighted_nms([det(0, 0, .2, .2, .9), det(.8, .8, 1, 1, .8)], 0.5, 0.5)) == 2


def test_nms_suppresses_identical_and_keeps_top_score():
    r = dp.weighted_nms([det(0, 0, .5, .5, .95), det(0, 0, .5, .5, .6)], 0.3, 0.5)
    assert len(r) == 1 and r[0].score == 0.95


def test_nms_weighted_average():
    r = dp.weighted_nms([det(0, 0, .5, .5, .9), det(.05, .05, .55, .55, .8)], 0.3, 0.5)
    assert len(r) == 1 and r[0].score == 0.9
    assert 0.0 < r[0].xmin < 0.05
    assert r[0].xmin == pytest.approx((0 * .9 + .05 * .8) / 1.7)


def test_nms_more_than_8_disjoint():
    dets = [det(i * .1, 0, i * .1 + .08, .08, .9 - i * .01) for i in range(10)]
    assert len(dp.weighted_nms(dets, 0.5, 0.5)) == 10


def test_nms_three_clusters_and_strict_iou():
    dets = []
    for k, x in enumerate((0.0, 0.4, 0.8)):
        dets += [det(x, 0, x + .15, .15, .9 - k * .01), det(x + .01, 0, x + .16, .15, .7)]
    assert len(dp.weighted_nms(dets, 0.3, 0.5)) == 3
    # IoU exactly at the threshold is NOT merged (strict >, helpers.dart:176-179)
    a, b = det(0, 0, 1, 1, .9), det(0, 0, 1, .5, .8)      # IoU = 0.5
    assert len(dp.weighted_nms([a, b], 0.5, 0.5)) == 2
    assert len(dp.weighted_nms([a, b], 0.49, 0.5)) == 1


def test_nms_max_detections_and_keypoints_from_top():
    dets = [dp.Detection(i * .005, 0, i * .005 + .004, .004, .99 - i * .001, [float(i)] * 12, i) for i in range(150)]
    r = dp.weighted_nms(dets)
    assert len(r) == 100 and [d.anchor for d in r] == list(range(100))
    assert r[7].kp == [7.0] * 12


# ---- letterbox removal: helpers_private_test.dart:22-36, helpers_unit_test.dart:160-175 ----------------
def test_letterbox_removal_known_answers():
    d = dp.Detection(0.2, 0.2, 0.8, 0.8, 0.9, [0.5] * 12)
    r = dp.letterbox_removal([d], [0.05, 0.05, 0.05, 0.05])[0]
    assert r.xmin == pytest.approx((0.2 - 0.05) / 0.9) and r.xmin == pytest.approx(0.1667, abs=1e-4)
    r = dp.letterbox_removal([dp.Detection(0.3, 0.3, 0.6, 0.6, 0.9, [0.5] * 12)], [0.15, 0.15, 0.0, 0.0])[0]
    assert r.ymin == pytest.approx(0.2143, abs=1e-4) and r.xmin == pytest.approx(0.3)
    # no clamping on the native path
    r = dp.letterbox_removal([dp.Detection(0.0, 0.0, 1.0, 1.0, 0.9, [0.0, 1.0] * 6)], [0.2, 0.2, 0.0, 0.0])[0]
    assert r.ymin < 0 and r.ymax > 1 and r.kp[1] > 1


# ---- landmarks unpack: helpers_private_test.dart:38-48, helpers_coverage_test.dart:121-218 -------------
def test_unpack_landmarks_known_answers():
    lm = geo.unpack_landmarks([10, 20, 5], 100, 100, [0.1, 0.1, 0.1, 0.1], clamp=True)
    assert lm[0][0] == pytest.approx(0.0) and lm[0][1] == pytest.approx(0.125) and lm[0][2] == 5
    lm = geo.unpack_landmarks([-50, 500, 192], 192, 192, [0, 0, 0, 0], clamp=True, normalize_z=True)
    assert lm[0][0] == 0.0 and lm[0][1] == 1.0 and lm[0][2] == pytest.approx(1.0, abs=1e-4)
    lm = geo.unpack_landmarks([-50, 500, 1], 192, 192, [0, 0, 0, 0], clamp=False)
    assert lm[0][0] < 0 and lm[0][1] > 1


# ---- sigmoid: helpers_private_test.dart:15-20, helpers_coverage_test.dart:35-59 ------------------------
def test_sigmoid_clipped():
    assert dp.sigmoid_clipped(2.0) == pytest.approx(0.8808, abs=1e-4)
    assert dp.sigmoid_clipped(-2.0) == pytest.approx(0.1192, abs=1e-4)
    assert dp.sigmoid_clipped(10.0, limit=2.0) == pytest.approx(0.8808, abs=1e-4)
    assert dp.sigmoid_clipped(0.0) == 0.5
    assert dp.sigmoid_clipped(1e9) == 1.0 and 0.0 <= dp.sigmoid_clipped(-1e9) < 1e-40
    assert dp.raw_score_threshold() == 0.0


# ---- decode: test/web_detection_decode_test.dart:84-195 ------------------------------------------------
def test_decode_known_answers_and_f32_intermediate():
    anchors = np.array([[0.5, 0.5], [0.25, 0.75]], np.float64)
    raw = np.zeros((2, 16), np.float32)
    raw[0, :4] = [12.8, -6.4, 25.6, 51.2]
    raw[0, 4:6] = [6.4, 6.4]
    raw[1, :4] = [0.1, 0.2, 0.3, 0.7]
    b = dp.decode_boxes(raw, anchors, [0, 1], 128)
    xc, yc = np.float32(np.float32(12.8 / 128) + 0.5), np.float32(np.float32(np.float32(-6.4) / 128) + 0.5)
    w, h = np.float32(np.float32(25.6) / 128), np.float32(np.float32(51.2) / 128)
    assert b[0][0] == float(xc) - float(w) * 0.5 and b[0][3] == float(yc) + float(h) * 0.5
    assert b[0][4][0] == float(np.float32(np.float32(np.float32(6.4) / 128) + 0.5))
    # the f32 round trip is observable: a pure-f64 evaluation differs in the low bits
    assert b[1][0] != (0.1 / 128 + 0.25) - (0.3 / 128) * 0.5
    assert b[1][0] == pytest.approx((0.1 / 128 + 0.25) - (0.3 / 128) * 0.5, abs=1e-7)


def test_candidates_threshold_nan_and_degenerate():
    raw = np.array([-0.1, 0.0, np.nan, 3.0, -0.0], np.float32)
    idx, sc = dp.collect_candidates(raw)
    assert idx == [1, 3, 4] and sc[0] == 0.5                      # >= 0.0 kept, NaN rejected
    anchors = np.full((5, 2), 0.5)
    boxes = np.zeros((5, 16), np.float32)
    boxes[1, 2:4] = [10, 10]
    boxes[3, 2:4] = [0, 10]                                        # zero width -> dropped (:506)
    boxes[4, 2:4] = [10, -10]                                      # negative height -> dropped
    d = dp.to_detections_filtered(dp.decode_boxes(boxes, anchors, idx, 128), sc, idx)
    assert [x.anchor for x in d] == [1]


# ---- alignment / mesh transform: test/face_geometry_test.dart:57-270 -----------------------------------
def test_alignment_identities():
    kp = [0.4, 0.5, 0.6, 0.5, 0.5, 0.55, 0.5, 0.7, 0.3, 0.5, 0.7, 0.5]
    theta, cx, cy, size = geo.compute_face_alignment(kp, 1000.0, 1000.0)
    assert theta == 0.0 and cx == pytest.approx(500.0) and cy == pytest.approx(520.0)
    assert size == pytest.approx(max(200.0 * 3.6, 200.0 * 4.0))
    kp2 = list(kp)
    kp2[1], kp2[3] = 0.4, 0.6                                       # eyes on a diagonal
    theta2, *_ = geo.compute_face_alignment(kp2, 1000.0, 1000.0)
    assert theta2 == pytest.approx(math.atan2(200, 200))


def test_mesh_transform_identities():
    lm = np.array([[0.5, 0.5, 0.1], [0.0, 0.0, 0.0], [1.0, 1.0, -0.2]])
    out = geo.transform_mesh_to_absolute(lm, 300.0, 200.0, 100.0, 0.0)
    assert np.allclose(out[0], [300, 200, 10]) and np.allclose(out[1], [250, 150, 0]) and np.allclose(out[2], [350, 250, -20])
    out = geo.transform_mesh_to_absolute(lm, 300.0, 200.0, 100.0, math.pi / 2)
    assert np.allclose(out[0][:2], [300, 200]) and np.allclose(out[1][:2], [350, 150])   # centre is a fixed point


def test_gates():
    d = [dp.Detection(0.1, 0.1, 0.3, 0.3, 0.6), dp.Detection(-0.2, 0.1, 0.1, 0.3, 0.9)]
    assert dp.apply_detection_gates(d, 0.0, 0.0, 100.0) is d
    assert [x.score for x in dp.apply_detection_gates(d, 0.7, 0.0, 100.0)] == [0.9]
    assert dp.box_visible_width_fraction(d[1], 100.0) == pytest.approx(0.1)
    assert [x.score for x in dp.apply_detection_gates(d, 0.0, 0.15, 100.0)] == [0.6]


# ---- sample images: all_model_variants_test.dart:28-32, :297-356; face_detection_integration_test.dart:124-198
EXPECTED_DETS = {("shortRange", "landmark-ex1.jpg"): 1, ("shortRange", "iris-detection-ex1.jpg"): 1,
                 ("backCamera", "landmark-ex1.jpg"): 1, ("backCamera", "iris-detection-ex1.jpg"): 1,
                 ("backCamera", "group-shot-bounding-box-ex1.jpeg"): 4,
                 ("full", "landmark-ex1.jpg"): 1, ("full", "iris-detection-ex1.jpg"): 1}


@pytest.mark.parametrize("model", ["shortRange", "backCamera", "full"])
def test_sample_face_counts_and_fixture(model, model_bytes, sample_images, golden):
    p = OraclePipeline(model_bytes[model], model, model_bytes["mesh"], "cv2dnn")
    for name, img in sample_images.items():
        faces = p.detect_faces(img, "standard")
        if (model, name) in EXPECTED_DETS:
            assert len(faces) == EXPECTED_DETS[(model, name)]
        if model == "full" and name.startswith("group"):
            assert len(faces) >= 2                                  # all_model_variants_test.dart:346-356
        g = golden["%s/%s/dets" % (model, name)]
        dets = p.detect(img)
        assert len(dets) == len(g)
        for d, row in zip(dets, g):
            assert d.anchor == int(row[17])
            assert np.allclose(d.as_row(), row[:17], atol=1e-4)     # cv2.dnn fp32 vs the fp64 fixture
        for f in faces:
            assert 0.5 <= f.det.score <= 1.0 and f.mesh_px.shape == (468, 3) and 0.0 <= f.mesh_score <= 1.0


def test_f64_oracle_reproduces_fixture_heads(model_bytes, sample_images, golden):
    p = OraclePipeline(model_bytes["shortRange"], "shortRange", None, "f64")
    img = sample_images["landmark-ex1.jpg"]
    t, pad, _ = p.preprocess(img)
    b, s = p.raw_heads(t)
    assert np.array_equal(b, golden["shortRange/landmark-ex1.jpg/boxes_f64"])
    idx, _ = dp.collect_candidates(s)
    assert idx == list(golden["shortRange/landmark-ex1.jpg/candidates"]) and len(idx) == 10
    assert idx[np.argmax(s[idx])] == 241 and float(s[241]) == pytest.approx(1.5384, abs=1e-3)
    # an independent fp32 implementation (cv2.dnn) agrees to ~1e-6 of the head's range
    cv = golden["shortRange/landmark-ex1.jpg/scores_cv2dnn"]
    assert np.abs(cv - s).max() <= 1e-4 * np.abs(s).max()

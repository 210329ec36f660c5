"""k_decode_nms (threshold -> decode -> degenerate filter -> weighted NMS -> letterbox removal) on caller-supplied
tensors through fdt_debug_decode / fdt_debug_nms: the reference's own unit-test vectors, verbatim, and crafted edge
cases (NaN, degenerate boxes, score ties, more than maxDet clusters, IoU exactly at the threshold) — each compared
with the oracle on the same inputs.  Integer / index results are bit-exact; from decoded detections (fdt_debug_nms) boxes
are f64-identical because the kernel accumulates a cluster in the same order as the reference (sorted order, one
thread); from raw heads the score goes through exp(), whose last bit differs between CUDA and glibc."""
import math

import numpy as np
import pytest

import reference_vectors as rv
from oracle import detect_post as dp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det(lib):
    import face_detection_tflite_b200 as fdt
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withMesh=False, maxBatch=8)
    yield d
    d.dispose()


def rows17(items, kp):
    return np.array([[*b, s, *kp] for b, s in items], np.float64).reshape(-1, 17)


def oracle_nms(items, kp, iou, score, padding=None):
    dets = [dp.Detection(b[0], b[1], b[2], b[3], s, list(kp), i) for i, (b, s) in enumerate(items)]
    out = dp.weighted_nms(dets, iou, score)
    return dp.letterbox_removal(out, padding) if padding is not None else out


def assert_same(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["index"] == w.anchor
        assert g["box"] == (w.xmin, w.ymin, w.xmax, w.ymax)          # f64-identical
        assert g["score"] == w.score and g["kp"] == list(w.kp)


def assert_close(got, want):
    """From raw heads the score passes through exp(): CUDA's f64 exp is within 1 ulp of glibc's, not identical, and the
    score weights the cluster sums - indices stay exact, values agree to ~1e-15."""
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["index"] == w.anchor
        assert abs(g["score"] - w.score) <= 4e-16
        for a, b in zip(list(g["box"]) + g["kp"], [w.xmin, w.ymin, w.xmax, w.ymax] + list(w.kp)):
            assert abs(a - b) <= 1e-13 * max(1.0, abs(b))


# ---- the reference's vectors, verbatim ------------------------------------------------------------------
def test_nms_private_vector(det):
    v = rv.NMS_PRIVATE
    got = det.debugNms(rows17(v["dets"], [0.0] * 12), scoreThresh=v["score"], iouThresh=v["iou"])
    assert len(got) == v["expect_len"]
    assert_same(got, oracle_nms(v["dets"], [0.0] * 12, v["iou"], v["score"]))


@pytest.mark.parametrize("case", rv.NMS_COVERAGE, ids=[c["name"] for c in rv.NMS_COVERAGE])
def test_nms_coverage_vectors(det, case):
    got = det.debugNms(rows17(case["dets"], rv.NMS_KP), scoreThresh=case["score"], iouThresh=case["iou"])
    assert len(got) == case["expect_len"]
    if "expect_xmin_gt" in case:
        assert got[0]["box"][0] > case["expect_xmin_gt"]
    if "expect_score0" in case:
        assert got[0]["score"] == case["expect_score0"]
    assert_same(got, oracle_nms(case["dets"], rv.NMS_KP, case["iou"], case["score"]))


def test_letterbox_removal_vector(det):
    v = rv.LETTERBOX_REMOVAL
    kp = list(v["kp"]) + [0.0] * 6
    got = det.debugNms(rows17([(v["box"], v["score"])], kp), scoreThresh=0.0, iouThresh=0.3, padding=v["padding"])
    assert len(got) == 1
    assert got[0]["box"][0] == pytest.approx(v["expect_xmin"], abs=v["tol"])
    assert got[0]["kp"][0] == pytest.approx(v["expect_kp0"], abs=v["tol"])
    assert_same(got, oracle_nms([(v["box"], v["score"])], kp, 0.3, 0.0, v["padding"]))


@pytest.mark.parametrize("case", rv.DECODE_CASES, ids=[c["name"] for c in rv.DECODE_CASES])
def test_web_decode_vectors(det, case):
    scores = np.array(case["scores"], np.float32)
    boxes = np.array(case["boxes"], np.float32)
    faces, dec = det.debugDecode(boxes, scores, anchors=rv.DECODE_ANCHORS, scale=rv.DECODE_SCALE, iouThresh=1.0)
    dec = dec[0]
    kept = dec[dec[:, 17] == 1.0]
    idx, sc = dp.collect_candidates(scores)
    want = dp.to_detections_filtered(dp.decode_boxes(boxes, np.array(rv.DECODE_ANCHORS), idx, int(rv.DECODE_SCALE)), sc, idx)
    assert [d.anchor for d in want] == case["expect_anchors"]
    assert len(dec) == len(idx) and len(kept) == len(want)
    for row, w in zip(kept, want):                               # decoded candidates: boxes / keypoints bit-identical to the oracle
        assert tuple(row[:4]) == (w.xmin, w.ymin, w.xmax, w.ymax) and abs(row[4] - w.score) <= 4e-16 and list(row[5:17]) == w.kp
    if "expect_xmin1" in case:
        assert kept[1][0] == pytest.approx(case["expect_xmin1"], abs=case["tol"])
    assert sorted(f["index"] for f in faces[0]) == case["expect_anchors"]      # IoU threshold 1.0: nothing merges


# ---- crafted edge cases -------------------------------------------------------------------------------------
def test_nan_and_degenerate_inputs(det):
    rng = np.random.default_rng(3)
    N = 896
    scores = np.full(N, -9.0, np.float32)
    boxes = rng.normal(0, 20, (N, 16)).astype(np.float32)
    boxes[:, 2:4] = np.abs(boxes[:, 2:4]) + 10
    hot = rng.choice(N, 40, replace=False)
    scores[hot] = rng.uniform(0.0, 4.0, 40).astype(np.float32)
    scores[hot[0]] = np.nan                                    # NaN logit: rejected by `raw >= thr`
    scores[hot[1]] = 0.0                                       # exactly at the threshold: kept (>=)
    scores[hot[2]] = -0.0
    scores[hot[3]] = np.float32(-1e-30)                        # just below: rejected
    boxes[hot[4], 2] = 0.0                                     # zero width  -> dropped (_toDetectionsFiltered)
    boxes[hot[5], 3] = -3.0                                    # negative height -> dropped
    scores[hot[7]] = np.inf                                    # clipped by sigmoidClipped
    faces, dec = det.debugDecode(boxes, scores, scale=128.0)
    anchors = det.anchors()
    idx, sc = dp.collect_candidates(scores)
    assert hot[0] not in idx and hot[1] in idx and hot[2] in idx and hot[3] not in idx
    assert list(dec[0][:, 17].nonzero()[0]) == [k for k, i in enumerate(idx) if not (boxes[i, 2] <= 0 or boxes[i, 3] <= 0)]
    want = dp.weighted_nms(dp.to_detections_filtered(dp.decode_boxes(boxes, anchors, idx, 128), sc, idx))
    assert_close(faces[0], want)


def test_score_ties_keep_anchor_order(det):
    # equal scores: the oracle (and the kernel) break ties by ascending input order
    items = [((0.1 * i, 0.0, 0.1 * i + 0.08, 0.08), 0.75) for i in range(8)]
    items += [((0.1 * i + 0.005, 0.0, 0.1 * i + 0.085, 0.08), 0.75) for i in range(8)]     # overlapping twins, same score
    got = det.debugNms(rows17(items, rv.NMS_KP), scoreThresh=0.5, iouThresh=0.3)
    want = oracle_nms(items, rv.NMS_KP, 0.3, 0.5)
    assert [g["index"] for g in got] == list(range(8))
    assert_same(got, want)


def test_more_than_max_det_clusters(det):
    # 150 disjoint boxes -> weightedNms stops at maxDet = 100 (helpers.dart:187), best scores first
    items = [(((i % 15) * 0.06, (i // 15) * 0.09, (i % 15) * 0.06 + 0.05, (i // 15) * 0.09 + 0.08), 0.99 - 0.003 * i) for i in range(150)]
    perm = np.random.default_rng(1).permutation(150)
    shuffled = [items[i] for i in perm]
    got = det.debugNms(rows17(shuffled, rv.NMS_KP), scoreThresh=0.5, iouThresh=0.3)
    want = oracle_nms(shuffled, rv.NMS_KP, 0.3, 0.5)
    assert len(got) == 100
    assert_same(got, want)
    assert [g["score"] for g in got] == sorted((s for _, s in items), reverse=True)[:100]


def test_iou_exactly_at_threshold_is_not_merged(det):
    # two unit-height boxes overlapping by exactly 1/3 of their union: IoU == thr -> strict '>' keeps both (helpers.dart:176-179)
    a, b = (0.0, 0.0, 0.5, 0.25), (0.25, 0.0, 0.75, 0.25)        # inter 0.25*0.25, union 0.75*0.25 -> IoU = 1/3
    thr = dp.iou(a, b)
    items = [(a, 0.9), (b, 0.8)]
    got = det.debugNms(rows17(items, rv.NMS_KP), scoreThresh=0.5, iouThresh=thr)
    assert len(got) == 2
    assert_same(got, oracle_nms(items, rv.NMS_KP, thr, 0.5))
    got = det.debugNms(rows17(items, rv.NMS_KP), scoreThresh=0.5, iouThresh=np.nextafter(thr, 0.0))
    assert len(got) == 1
    assert_same(got, oracle_nms(items, rv.NMS_KP, float(np.nextafter(thr, 0.0)), 0.5))


def test_big_cluster_sum_is_sequential(det):
    # 300 heavily overlapping boxes in one cluster: the f64 weighted sum must equal the oracle's sequential accumulation
    rng = np.random.default_rng(7)
    items = []
    for _ in range(300):
        dx, dy = rng.uniform(-0.01, 0.01, 2)
        items.append(((0.3 + dx, 0.3 + dy, 0.6 + dx, 0.6 + dy), float(rng.uniform(0.5, 1.0))))
    got = det.debugNms(rows17(items, rv.NMS_KP), scoreThresh=0.5, iouThresh=0.3)
    want = oracle_nms(items, rv.NMS_KP, 0.3, 0.5)
    assert len(want) == 1
    assert_same(got, want)


def test_random_heads_full_range_anchor_count(lib):
    """2304 anchors (full-range layout), hundreds of candidates per image, 6 images in one launch."""
    import face_detection_tflite_b200 as fdt
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel.full, withMesh=False, maxBatch=8)
    rng = np.random.default_rng(11)
    B, N = 6, 2304
    scores = rng.normal(-3.0, 2.5, (B, N)).astype(np.float32)
    boxes = rng.normal(0, 12, (B, N, 16)).astype(np.float32)
    boxes[..., 2:4] = np.abs(boxes[..., 2:4]) * 2 + 4
    faces, dec = d.debugDecode(boxes, scores, scale=192.0, padding=[42 / 192, 42 / 192, 0.0, 0.0])
    anchors = d.anchors()
    for b in range(B):
        want = dp.postprocess(boxes[b], scores[b], anchors, 192, (42 / 192, 42 / 192, 0.0, 0.0))
        assert len(want) >= 20
        assert_close(faces[b], want)
    d.dispose()

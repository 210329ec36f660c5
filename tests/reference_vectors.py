"""The reference's own unit-test vectors for the detection hot path, reproduced VERBATIM (numbers and
expectations) from /root/reference/test/*.dart so that the oracle (CPU suite) and the CUDA kernels (GPU suite,
through fdt_debug_nms / fdt_debug_decode) are checked against exactly what the reference pins.

Each entry cites the Dart test it restates.  Data only — no oracle or product imports here."""
import math

# ---- test/helpers_private_test.dart:15-20  testSigmoidClipped(x, limit: 2) ---------------------------------
SIGMOID_LIMIT2 = [(1000.0, 0.8808), (-1000.0, 0.1192)]          # closeTo 1e-4

# ---- test/helpers_private_test.dart:22-36  testDetectionLetterboxRemoval ----------------------------------
LETTERBOX_REMOVAL = {
    "box": (0.2, 0.3, 0.6, 0.7), "score": 0.9, "kp": [0.2, 0.3, 0.4, 0.5, 0.6, 0.7],
    "padding": [0.1, 0.1, 0.05, 0.05],                           # top, bottom, left, right
    "expect_xmin": 0.1667, "expect_kp0": 0.1667, "tol": 1e-4,
}

# ---- test/helpers_private_test.dart:38-48  testUnpackLandmarks(clamp: true) -------------------------------
UNPACK_LANDMARKS = {"flat": [10, 20, 1, 30, 40, 2], "w": 100, "h": 100, "padding": [0.1, 0.1, 0.1, 0.1],
                    "expect": {"len": 2, "p0": (0.0, 0.125, 1.0)}, "tol": 1e-4}

# ---- test/helpers_private_test.dart:50-71  testNms(dets, 0.3, 0.0) -> 2 kept --------------------------------
NMS_PRIVATE = {"dets": [((0.1, 0.1, 0.4, 0.4), 0.9), ((0.15, 0.15, 0.45, 0.45), 0.8), ((0.7, 0.7, 0.9, 0.9), 0.7)],
               "iou": 0.3, "score": 0.0, "expect_len": 2}

# ---- test/helpers_coverage_test.dart:220-291  group('testNms'); det() uses keypointsXY = List.filled(12, 0.5) ----
NMS_COVERAGE = [
    {"name": "empty input", "dets": [], "iou": 0.5, "score": 0.5, "expect_len": 0},
    {"name": "score threshold", "dets": [((0.0, 0.0, 0.5, 0.5), 0.3)], "iou": 0.5, "score": 0.5, "expect_len": 0},
    {"name": "non-overlapping kept", "dets": [((0.0, 0.0, 0.2, 0.2), 0.9), ((0.8, 0.8, 1.0, 1.0), 0.8)], "iou": 0.5, "score": 0.5,
     "expect_len": 2},
    {"name": "identical suppressed", "dets": [((0.0, 0.0, 0.5, 0.5), 0.9), ((0.0, 0.0, 0.5, 0.5), 0.8)], "iou": 0.3, "score": 0.5,
     "expect_len": 1},
    {"name": "weighted average", "dets": [((0.0, 0.0, 0.5, 0.5), 0.9), ((0.05, 0.05, 0.55, 0.55), 0.8)], "iou": 0.3, "score": 0.5,
     "expect_len": 1, "expect_xmin_gt": 0.0, "expect_score0": 0.9},
    {"name": "score of the best", "dets": [((0.0, 0.0, 0.5, 0.5), 0.95), ((0.0, 0.0, 0.5, 0.5), 0.6)], "iou": 0.3, "score": 0.5,
     "expect_len": 1, "expect_score0": 0.95},
    {"name": "more than 8", "dets": [((i * 0.1, 0.0, i * 0.1 + 0.08, 0.08), 0.9 - i * 0.01) for i in range(10)], "iou": 0.5, "score": 0.5,
     "expect_len": 10},
]
NMS_KP = [0.5] * 12

# ---- test/web_detection_decode_test.dart:22-37, :84-195  decodeBlazeFaceCandidates ---------------------------


def box_row(xc, yc, w, h):
    row = [xc, yc, w, h]
    for j in range(6):
        row += [xc + j, yc - j]
    return row


DECODE_ANCHORS = [[0.1 * (i + 1), 0.2 * (i + 1)] for i in range(4)]
DECODE_SCALE = 128.0
DECODE_CASES = [
    {"name": "no degenerate box",                                   # :89-124: anchor 2 below the floor, the rest pass
     "scores": [2.0, 3.0, -5.0, 2.5],
     "boxes": [box_row(10, 12, 30, 32), box_row(20, 22, 40, 42), box_row(30, 32, 50, 52), box_row(40, 42, 60, 62)],
     "expect_anchors": [0, 1, 3]},
    {"name": "skipped degenerate box keeps scores paired",          # :126-166: anchor 1 has w == 0
     "scores": [2.0, 3.0, 2.5, -5.0],
     "boxes": [box_row(10, 12, 30, 32), box_row(20, 22, 0, 42), box_row(30, 32, 50, 52), box_row(40, 42, 60, 62)],
     "expect_anchors": [0, 2],
     "expect_xmin1": 30 / 128.0 + DECODE_ANCHORS[2][0] - 50 / 128.0 / 2, "tol": 1e-6},
    {"name": "all below the floor", "scores": [-5.0, -4.0, -3.0, -6.0],    # :168-178
     "boxes": [[0.0] * 16] * 4, "expect_anchors": []},
    {"name": "NaN rejected", "scores": [float("nan"), -5.0, -5.0, -5.0],   # :180-195
     "boxes": [box_row(10, 12, 30, 32), box_row(20, 22, 40, 42), box_row(30, 32, 50, 52), box_row(40, 42, 60, 62)],
     "expect_anchors": []},
]

# ---- test/face_geometry_test.dart:426-481  eyeRoisFromMesh --------------------------------------------------
EYE_ROIS = [
    {"corners": {33: (10, 50), 133: (30, 50), 362: (70, 50), 263: (90, 50)},
     "left": {"cx": 20.0, "cy": 50.0, "size": 20.0 * 2.3, "theta": 0.0}, "right": {"cx": 80.0, "size": 20.0 * 2.3}},
    {"corners": {33: (0, 0), 133: (10, 10), 362: (0, 0), 263: (0, 10)},
     "left": {"theta": math.pi / 4, "size": math.sqrt(200.0) * 2.3}, "right": {"theta": math.pi / 2}},
]

# ---- test/face_geometry_test.dart:272-329  transformIrisNormToAbsolute; roi = AlignedRoi(cx, cy, size, theta) -----
IRIS_ROI = (50.0, 60.0, 40.0, 0.0)
IRIS_ROTATED = (0.0, 0.0, 100.0, math.pi / 2)

# ---- test/face_embedding_test.dart:256-325  computeEmbeddingAlignment -----------------------------------------
EMBED_ALIGN = [
    {"l": (100.0, 100.0), "r": (200.0, 100.0), "theta": (0.0, 1e-4), "size": (250.0, 0.1), "cx": (150.0, 1.0), "cy_gt": 100.0},
    {"l": (100.0, 100.0), "r": (100.0 + 100.0 * math.cos(math.pi / 4), 100.0 + 100.0 * math.sin(math.pi / 4)),
     "theta": (math.pi / 4, 0.01), "size": (250.0, 0.1)},
    {"l": (100.0, 100.0), "r": (100.0, 200.0), "theta": (math.pi / 2, 0.01)},
    {"l": (100.0, 100.0), "r": (200.0, 50.0), "theta_lt": 0.0},
    {"l": (100.0, 100.0), "r": (101.0, 100.0), "size": (2.5, 0.1), "theta": (0.0, 0.01)},
]

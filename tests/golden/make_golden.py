"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the fixtures travel to the GPU box, /root/reference does not).

The reference itself (Dart + flutter_litert + opencv_dart) cannot run here, so the vectors come from
  * cv2 4.13.0 — the real OpenCV resize / copyMakeBorder / warpAffine the reference calls through
    opencv_dart (pins the integer stages independently of oracle/cv_ops.py),
  * cv2.dnn.readNetFromTFLite — an independent fp32 executor of the reference's .tflite graphs,
  * the oracle's fp64 graph executor and post-processing (regression pin).
"""
import math
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import cv_ops, detect_post as dp, geometry as geo  # noqa: E402
from oracle.pipeline import OraclePipeline  # noqa: E402

REF = Path("/root/reference/assets")
SAMPLES = ["landmark-ex1.jpg", "iris-detection-ex1.jpg", "group-shot-bounding-box-ex1.jpeg"]
MODELS = {"shortRange": "face_detection_short_range.tflite", "full": "face_detection_full_range.tflite",
          "backCamera": "face_detection_back.tflite"}


def main():
    out = {}
    mesh_bytes = (REF / "models/face_landmark.tflite").read_bytes()
    for mname, mfile in MODELS.items():
        det_bytes = (REF / "models" / mfile).read_bytes()
        p64 = OraclePipeline(det_bytes, mname, mesh_bytes, "f64")
        pcv = OraclePipeline(det_bytes, mname, mesh_bytes, "cv2dnn")
        S = p64.in_w
        for s in SAMPLES:
            img = cv2.imread(str(REF / "samples" / s))
            h, w = img.shape[:2]
            key = "%s/%s" % (mname, s)
            lp = cv_ops.compute_letterbox_params(w, h, S, S)
            # real OpenCV letterbox (helpers.dart:325-347)
            r = cv2.resize(img, (lp.new_w, lp.new_h), interpolation=cv2.INTER_LINEAR)
            lb = cv2.copyMakeBorder(r, lp.pad_top, lp.pad_bottom, lp.pad_left, lp.pad_right, cv2.BORDER_CONSTANT, value=(0, 0, 0))
            out[key + "/letterboxed_cv2"] = lb
            out[key + "/lbparams"] = np.array([lp.new_w, lp.new_h, lp.pad_top, lp.pad_bottom, lp.pad_left, lp.pad_right], np.int32)
            t, pad, _ = p64.preprocess(img)
            b64, s64 = p64.raw_heads(t)
            bcv, scv = pcv.raw_heads(t)
            out[key + "/boxes_f64"] = b64.astype(np.float32)
            out[key + "/scores_f64"] = s64.astype(np.float32)
            out[key + "/scores_cv2dnn"] = scv.astype(np.float32)
            out[key + "/boxes_cv2dnn_absmax"] = np.array([np.abs(bcv - b64).max(), np.abs(b64).max()], np.float64)
            idx, _ = dp.collect_candidates(s64)
            out[key + "/candidates"] = np.array(idx, np.int32)
            faces = p64.detect_faces(img, "standard")
            dets = p64.detect(img)
            out[key + "/dets"] = np.array([d.as_row() + [d.anchor] for d in dets], np.float64).reshape(-1, 18)
            out[key + "/mesh_scores"] = np.array([f.mesh_score for f in faces], np.float64)
            out[key + "/mesh_px"] = np.array([f.mesh_px for f in faces], np.float64).reshape(-1, 468, 3)
            if faces:
                f0 = faces[0]
                theta, cx, cy, size = f0.align
                # real OpenCV crop exactly as extractAlignedSquare does it (helpers.dart:583-625)
                si = cv_ops.dart_round(size)
                sc = 192 / si
                R = cv2.getRotationMatrix2D((float(np.float32(cx)), float(np.float32(cy))), theta * 180.0 / math.pi, sc)
                oc = 192 / 2.0 + 0.5 * (sc - 1.0)
                R[0, 2] += oc - cx
                R[1, 2] += oc - cy
                crop = cv2.warpAffine(img, R, (192, 192), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
                out[key + "/crop0_cv2"] = crop
                out[key + "/align0"] = np.array(f0.align, np.float64)
            print(key, "cands", len(idx), "dets", len(dets), "faces", len(faces))
    np.savez_compressed(Path(__file__).parent / "golden_samples.npz", **out)
    print("wrote", Path(__file__).parent / "golden_samples.npz")


if __name__ == "__main__":
    main()

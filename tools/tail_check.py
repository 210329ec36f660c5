"""k_tail_ws against the layer-by-layer kernels and the f64 oracle: raw heads of the sample images and synthetic frames.
Run twice (FDT_TAIL=0 writes the reference heads, FDT_TAIL=1 compares) or once with FDT_TAIL=1 (oracle only)."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cv2
import numpy as np
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth
from oracle.pipeline import OraclePipeline

out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
tail = os.environ.get("FDT_TAIL", "0") + "/ts" + os.environ.get("FDT_TS", "0")
base_run = os.environ.get("FDT_TAIL", "0") == "0" and os.environ.get("FDT_TS", "0") == "0"
for model, f in (("shortRange", "face_detection_short_range.tflite"), ("backCamera", "face_detection_back.tflite")):
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], withMesh=False, maxBatch=int(os.environ.get("CHUNK", "64")))
    img = cv2.imread(str(ROOT / "assets/samples/landmark-ex1.jpg"))
    img = cv2.resize(img, (1280, 720))
    frames = np.concatenate([np.stack([img, img[:, ::-1].copy(), img[::-1].copy()]), synth.face_frames(34, 1280, 720, start=2), synth.noise_frames(3, 1280, 720)])
    n = frames.shape[0]
    t = time.time()
    faces, counts, _ = d.detectBatchRaw(frames, count=n, width=1280, height=720)
    boxes, scores = d.debugRawHeads(min(n, d.maxBatch))
    print(model, "tail", tail, "launches", d.lastLaunchCount(), "faces", int(counts.sum()), "%.2fs" % (time.time() - t), flush=True)
    ref = out / ("heads_%s.npz" % model)
    if base_run:
        np.savez(ref, boxes=boxes, scores=scores, counts=counts)
    elif ref.exists():
        r = np.load(ref)
        eb = float(np.abs(boxes - r["boxes"]).max() / np.abs(r["boxes"]).max())
        es = float(np.abs(scores - r["scores"]).max() / np.abs(r["scores"]).max())
        print("  vs layer-by-layer kernels: boxes %.2e scores %.2e counts equal %s" % (eb, es, np.array_equal(counts, r["counts"])), flush=True)
    o = OraclePipeline((ROOT / "assets/models" / f).read_bytes(), model, None, "f64")
    k = min(6, boxes.shape[0])
    want = o.det.run(np.stack([o.preprocess(fr)[0] for fr in frames[:k]]))
    for got, w in ((boxes[:k], want[0]), (scores[:k], want[1])):
        w = np.asarray(w).reshape(got.shape)
        print("  vs f64 oracle: rel err %.2e" % float(np.abs(got - w).max() / np.abs(w).max()), flush=True)
    d.dispose()
print("tail check done")

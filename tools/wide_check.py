"""k_chain_wide (full-range BlazeFace below 24x24) against the layer-by-layer kernels (FDT_CHAIN=0 run first: writes the reference
heads) and the f64 oracle.  400 frames in one chunk, so every CTA walks several images."""
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
chain = os.environ.get("FDT_CHAIN", "1")
N = int(os.environ.get("NFRAMES", "400"))
d = fdt.FaceDetector.create(fdt.FaceDetectionModel.full, withMesh=False, maxBatch=N)
img = cv2.resize(cv2.imread(str(ROOT / "assets/samples/landmark-ex1.jpg")), (640, 360))
base = np.concatenate([np.stack([img, img[:, ::-1].copy(), img[::-1].copy()]), synth.face_frames(13, 640, 360, start=2), synth.noise_frames(2, 640, 360)])
frames = np.concatenate([base] * ((N + len(base) - 1) // len(base)))[:N]
frames = np.ascontiguousarray(frames)
t = time.time()
faces, counts, _ = d.detectBatchRaw(frames, count=N, width=640, height=360)
boxes, scores = d.debugRawHeads(N)
print("chain", chain, "launches", d.lastLaunchCount(), "faces", int(counts.sum()), "%.2fs" % (time.time() - t), flush=True)
ref = out / "wide_heads.npz"
if chain == "0":
    np.savez(ref, boxes=boxes, scores=scores, counts=counts)
elif ref.exists():
    r = np.load(ref)
    eb = np.abs(boxes - r["boxes"]).reshape(N, -1).max(1) / np.abs(r["boxes"]).max()
    es = np.abs(scores - r["scores"]).reshape(N, -1).max(1) / np.abs(r["scores"]).max()
    print("  vs layer-by-layer kernels: boxes %.2e scores %.2e (worst image %d) counts equal %s" % (eb.max(), es.max(), int(eb.argmax()), np.array_equal(counts, r["counts"])), flush=True)
    bad = np.nonzero((eb > 1e-4) | (es > 1e-4))[0]
    if len(bad): print("  bad images:", bad[:40], "of", len(bad))
# periodic input: image i and i + len(base) must agree bit for bit
per = len(base)
if N > per:
    same = all(np.array_equal(boxes[i], boxes[i % per]) for i in range(per, N))
    print("  periodic images bit-equal:", same, flush=True)
o = OraclePipeline((ROOT / "assets/models/face_detection_full_range.tflite").read_bytes(), "full", None, "f64")
k = 6
want = o.det.run(np.stack([o.preprocess(fr)[0] for fr in frames[:k]]))
for got, w in ((boxes[:k], want[0]), (scores[:k], want[1])):
    w = np.asarray(w).reshape(got.shape)
    print("  vs f64 oracle: rel err %.2e" % float(np.abs(got - w).max() / np.abs(w).max()), flush=True)
d.dispose()
print("wide check done")

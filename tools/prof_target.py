"""Small fixed workload for ncu: 3 chunks of synthetic frames through the fast-mode pipeline.
  python tools/prof_target.py [chunk=512] [model=shortRange|full|backCamera]
shortRange: 10 kernel launches per chunk (letterbox, stem, 6 BlazeBlocks, the image-resident tail, decode+NMS); full: 39."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = sys.argv[2] if len(sys.argv) > 2 else "shortRange"
w, h = (1920, 1080) if model == "full" else (1280, 720)
d = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], withMesh=False, maxBatch=chunk)
base = np.concatenate([synth.face_frames(56, w, h), synth.noise_frames(8, w, h)])
dev = torch.from_numpy(base).cuda().repeat((chunk + 63) // 64, 1, 1, 1)[:chunk].contiguous()
for _ in range(3):
    faces, counts, _ = d.detectBatchRaw(dev.data_ptr(), count=chunk, width=w, height=h, memKind=1)
print("faces", int(counts.sum()), "launches", d.lastLaunchCount())

"""Small fixed workload for ncu: 3 chunks of 512 synthetic 1280x720 frames through the fast-mode
pipeline (21 kernel launches per chunk: letterbox, stem, 16 BlazeBlocks, 2 head pairs, decode+NMS)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withMesh=False, maxBatch=chunk)
base = np.concatenate([synth.face_frames(56, 1280, 720), synth.noise_frames(8, 1280, 720)])
dev = torch.from_numpy(base).cuda().repeat(chunk // 64, 1, 1, 1).contiguous()
for _ in range(3):
    faces, counts, _ = d.detectBatchRaw(dev.data_ptr(), count=chunk, width=1280, height=720, memKind=1)
print("faces", int(counts.sum()), "launches", d.lastLaunchCount())

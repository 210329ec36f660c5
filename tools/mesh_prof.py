import sys; sys.path.insert(0,'.')
import numpy as np, torch
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth
det = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange)
base = np.concatenate([synth.face_frames(56, 1280, 720), synth.noise_frames(8, 1280, 720)])
dev = torch.from_numpy(base).cuda().repeat(16, 1, 1, 1).contiguous()
m = fdt.FaceDetectionMode.standard
for _ in range(2):
    faces, counts, mesh = det.detectBatchRaw(dev.data_ptr(), count=1024, width=1280, height=720, mode=m, memKind=1)
print("faces", int(counts.sum()))

"""Per-kernel device time of one model's detector plan (fdt_profile_chunk).  Usage: profile_model.py MODEL W H [chunk]"""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth
model, w, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 256
det = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], withMesh=False, maxBatch=chunk)
base = np.concatenate([synth.face_frames(24, w, h), synth.noise_frames(8, w, h)])
dev = torch.from_numpy(base).cuda().repeat(chunk // 32, 1, 1, 1).contiguous()
lib, hd = det._lib, det._h
det.detectBatchRaw(dev.data_ptr(), count=chunk, width=w, height=h, memKind=1)
arr = (C.c_float * 512)(); nl = C.c_int32()
rc = lib.fdt_profile_chunk(hd, dev.data_ptr(), chunk, w, h, w * 3, 16, 5, arr, 512, C.byref(nl))
assert rc == 0, lib.fdt_last_error(hd)
kn, tn = C.create_string_buffer(64), C.create_string_buffer(128)
macs, byt = C.c_double(), C.c_double()
tot = sum(arr[i] for i in range(nl.value))
for i in range(nl.value):
    lib.fdt_get_step_info(hd, i, kn, tn, 64, C.byref(macs), C.byref(byt))
    print("%3d %-14s %-28s %8.3f ms %5.1f%%  %6.2f TF" % (i, kn.value.decode(), tn.value.decode()[:28], arr[i], 100 * arr[i] / tot, 2 * macs.value * chunk / (arr[i] * 1e-3) / 1e12 if arr[i] > 0 else 0))
print("total %.3f ms for %d images -> %.0f img/s single stream" % (tot, chunk, chunk / tot * 1e3))

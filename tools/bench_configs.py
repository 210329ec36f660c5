"""Throughput of the other BASELINE.json configs (device-resident frames), for DESIGN.md / profiles:
C3 full-range 1920x1080 (fast mode), C4 detect + warp + mesh (standard mode) on C2-style frames, and the
default backCamera model.  Usage: python tools/bench_configs.py [batch]"""
import ctypes as C
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import face_detection_tflite_b200 as fdt
from face_detection_tflite_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def run(model, w, h, mode, batch, min_side, max_side):
    # standard mode: 4 result slots per frame (BASELINE config 4: "up to 4 faces/frame") keeps the mesh output buffer small
    det = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], withMesh=(mode == "standard"), maxFaces=(4 if mode == "standard" else 0))
    base = np.concatenate([synth.face_frames(56, w, h, min_side=min_side, max_side=max_side), synth.noise_frames(8, w, h)])
    dev = torch.from_numpy(base).cuda().repeat(batch // 64, 1, 1, 1).contiguous()
    m = fdt.FaceDetectionMode[mode]
    lib, hd = det._lib, det._h
    for _ in range(2):
        faces, counts, mesh = det.detectBatchRaw(dev.data_ptr(), count=batch, width=w, height=h, mode=m, memKind=1)
    torch.cuda.synchronize()
    t = time.perf_counter()
    n = 3
    for _ in range(n):
        faces, counts, mesh = det.detectBatchRaw(dev.data_ptr(), count=batch, width=w, height=h, mode=m, memKind=1)
    dt = (time.perf_counter() - t) / n
    nf = int(counts.sum())
    lib.fdt_set_stage_timing(hd, 1)
    det.detectBatchRaw(dev.data_ptr(), count=batch, width=w, height=h, mode=m, memKind=1)
    st = []
    for i, nm in enumerate(["letterbox", "conv", "decode", "warp", "mesh", "meshpost"]):
        ms, l = C.c_float(), C.c_int32()
        lib.fdt_get_stage_ms(hd, i, C.byref(ms), C.byref(l))
        st.append("%s %.2fms/%d" % (nm, ms.value, l.value))
    print("%-10s %dx%d %-8s batch %d: %.0f img/s, %d faces (%.0f faces/s) | %s" % (model, w, h, mode, batch, batch / dt, nf, nf / dt, "  ".join(st)), flush=True)
    det.dispose()


run("shortRange", 1280, 720, "fast", B, 160, 560)
run("full", 1920, 1080, "fast", B, 120, 700)
run("backCamera", 1280, 720, "fast", B, 160, 560)
run("shortRange", 1280, 720, "standard", B, 160, 560)
run("backCamera", 1280, 720, "standard", B, 160, 560)

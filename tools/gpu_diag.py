"""First-contact diagnostic for a GPU box: runs every model at every fuse level on a sample image,
compares each materialised tensor with the fp64 oracle and prints a table (never stops at the
first mismatch), then times the C2 pipeline.  Usage: python tools/gpu_diag.py [> gpurun_out/diag.log]"""
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cv2  # noqa: E402
import torch  # noqa: E402

import face_detection_tflite_b200 as fdt  # noqa: E402
from face_detection_tflite_b200 import build, synth  # noqa: E402
from oracle import cv_ops as co, detect_post as dp  # noqa: E402
from oracle.pipeline import OraclePipeline  # noqa: E402

build.build()
A = ROOT / "assets"
img = cv2.imread(str(A / "samples/landmark-ex1.jpg"))
grp = cv2.imread(str(A / "samples/group-shot-bounding-box-ex1.jpeg"))
mesh_bytes = (A / "models/face_landmark.tflite").read_bytes()
FILES = {"shortRange": "face_detection_short_range.tflite", "full": "face_detection_full_range.tflite", "backCamera": "face_detection_back.tflite"}


def section(t):
    print("\n==== " + t, flush=True)


def layer_table(d, which, ref, names, n):
    rows = []
    for tf_idx, want in sorted(ref.items()):
        try:
            got = d.debugTensor(which, tf_idx, n)
        except ValueError:
            continue
        want = want[:n].reshape(got.shape)
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
        rows.append((err, tf_idx, names[tf_idx], got.shape))
    bad = [r for r in rows if not (r[0] <= 1e-4)]
    print("  tensors compared: %d, worst rel err %.3e, failing: %d" % (len(rows), max(r[0] for r in rows) if rows else -1, len(bad)))
    for r in sorted(rows, key=lambda r: r[1]):
        if not (r[0] <= 1e-5):
            print("   tensor %4d %-40s %-18s rel err %.3e" % (r[1], r[2][:40], str(r[3]), r[0]))


for model, f in FILES.items():
    det_bytes = (A / "models" / f).read_bytes()
    o = OraclePipeline(det_bytes, model, mesh_bytes, "f64")
    h, w = img.shape[:2]
    frames = np.stack([img, img[::-1].copy()])
    t0 = o.preprocess(frames[0])[0]
    t1 = o.preprocess(frames[1])[0]
    ref = o.det.exe.run(np.stack([t0, t1]), taps="all")
    names = {i: t.name for i, t in enumerate(o.det.model.tensors)}
    for fuse in (0, 2, 1):
        section("%s fuse=%d" % (model, fuse))
        try:
            d = fdt.FaceDetector.create(fdt.FaceDetectionModel[model], fuseLevel=fuse, withMesh=True, maxBatch=(256 if fuse == 1 else 4))
            faces, counts, _ = d.detectBatchRaw(frames, count=2, width=w, height=h)
            lb = d.debugLetterboxed(2)
            want_lb = co.letterbox_u8(frames[0], o.in_w, o.in_h)[0]
            print("  letterbox mismatches:", int((lb[0] != want_lb).sum()), " launches:", d.lastLaunchCount())
            if fuse != 1:
                layer_table(d, 0, ref, names, 2)
            boxes, scores = d.debugRawHeads(2)
            for k, (g, nm) in enumerate(((boxes, "boxes"), (scores, "scores"))):
                wv = ref[o.det.model.outputs[k]].reshape(g.shape)
                print("  head %-6s rel err %.3e  (max|ref| %.2f)" % (nm, np.abs(g - wv).max() / np.abs(wv).max(), np.abs(wv).max()))
            cand = d.debugCandidates(0)
            print("  candidates gpu", cand.tolist(), "oracle", dp.collect_candidates(ref[o.det.model.outputs[1]][0].reshape(-1))[0])
            want = o.detect(frames[0])
            print("  faces gpu %d oracle %d" % (counts[0], len(want)))
            for j in range(counts[0]):
                fc = faces[j]
                print("   gpu    ", fc.anchor_index, "%.6f" % fc.score, ["%.5f" % v for v in (fc.xmin, fc.ymin, fc.xmax, fc.ymax)])
            for x in want:
                print("   oracle ", x.anchor, "%.6f" % x.score, ["%.5f" % v for v in (x.xmin, x.ymin, x.xmax, x.ymax)])
            # mesh
            std = d.detectFacesFromMat(grp if model != "shortRange" else img, mode=fdt.FaceDetectionMode.standard)
            ow = o.detect_faces(grp if model != "shortRange" else img, "standard")
            print("  standard mode faces gpu %d oracle %d" % (len(std), len(ow)))
            for g, x in zip(std, ow):
                print("   mesh score gpu %.5f oracle %.5f  max|dpx| %.4f (roi %.1f)" % (g.meshScore, x.mesh_score, np.abs(g.mesh.packed - x.mesh_px).max(), x.align[3]))
            if ow:
                crops, raw, flag = d.debugMeshStage(len(ow))
                print("   crop0 mismatching px: %d  raw mesh rel err %.3e" % (int((crops[0] != ow[0].crop).sum()), np.abs(raw[0] - ow[0].mesh_raw).max() / np.abs(ow[0].mesh_raw).max()))
                if fuse != 1:
                    mref = o.mesh.exe.run(co.normalize_bgr_u8(crops[0])[None], taps="all")
                    layer_table(d, 1, mref, {i: t.name for i, t in enumerate(o.mesh.model.tensors)}, 1)
            d.dispose()
        except Exception:
            traceback.print_exc(file=sys.stdout)

section("timing C2: shortRange, 1280x720, device-resident")
try:
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel.shortRange, withMesh=False)
    base = np.concatenate([synth.face_frames(56, 1280, 720), synth.noise_frames(8, 1280, 720)])
    dev = torch.from_numpy(base).cuda().repeat(16, 1, 1, 1).contiguous()     # 1024 frames
    lib = d._lib
    import ctypes as C
    for it in range(3):
        d.detectBatchRaw(dev.data_ptr(), count=1024, width=1280, height=720, memKind=1)
    pf, pc = C.c_void_p(), C.c_void_p()
    ms = C.c_float()
    lib.fdt_timer_begin(d._h)
    for it in range(5):
        lib.fdt_detect_batch_device(d._h, dev.data_ptr(), 1024, 1280, 720, 1280 * 3, 16, 0, C.byref(pf), C.byref(pc))
    lib.fdt_timer_end(d._h, C.byref(ms))
    print("  device-resident: %.3f ms per 1024 frames -> %.0f img/s" % (ms.value / 5, 1024 * 5 / ms.value * 1e3))
    lib.fdt_set_stage_timing(d._h, 1)
    d.detectBatchRaw(dev.data_ptr(), count=1024, width=1280, height=720, memKind=1)
    for st, nm in enumerate(["letterbox", "conv stack", "decode+nms"]):
        m, l = C.c_float(), C.c_int32()
        lib.fdt_get_stage_ms(d._h, st, C.byref(m), C.byref(l))
        print("  stage %-11s %.3f ms (%d launches)" % (nm, m.value, l.value))
    lib.fdt_set_stage_timing(d._h, 0)
    host = base
    t = time.time()
    d.detectBatchRaw(np.tile(host, (4, 1, 1, 1)), count=256, width=1280, height=720)
    print("  host pageable e2e: %.0f img/s" % (256 / (time.time() - t)))
except Exception:
    traceback.print_exc(file=sys.stdout)
print("\nDIAG DONE")

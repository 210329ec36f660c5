"""CPU-side checks of libfdt_cuda.so: it loads, exports every symbol include/fdt_api.h declares, and
its pure-host helpers (anchors, letterbox params, resize taps, decode arithmetic, ROI affine, the
TFLite -> kernel-plan lowering) agree with the oracle.  No device compute is called here."""
import ctypes as C
import math
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import cv_ops as co, detect_post as dp, geometry as geo

ROOT = Path(__file__).resolve().parents[1]


def test_every_declared_symbol_is_exported(lib):
    from face_detection_tflite_b200 import _ffi
    header = (ROOT / "include" / "fdt_api.h").read_text()
    declared = set(re.findall(r"FDT_EXPORT[^;(]*?\b(fdt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fdt_version().startswith(b"fdt-cuda")


def test_struct_layouts_match_header():
    from face_detection_tflite_b200 import _ffi
    assert C.sizeof(_ffi.FdtFace) == 18 * 8 + 16     # static_assert in csrc/fdt_api.cu
    assert C.sizeof(_ffi.FdtConfig) == 6 * 4 + 3 * 8


def test_create_fails_loudly_without_gpu(lib, model_bytes):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from face_detection_tflite_b200 import FaceDetector, FaceDetectionModel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FaceDetector.create(FaceDetectionModel.shortRange)


def test_create_argument_validation(lib, model_bytes):
    from face_detection_tflite_b200 import _ffi
    cfg = _ffi.FdtConfig()
    lib.fdt_default_config(C.byref(cfg))
    assert cfg.model == 1 and cfg.min_face_presence == 0.5            # backCamera default, presence 0.5
    h = C.c_void_p()
    d = model_bytes["shortRange"]
    cfg.model = 2
    cfg.min_score = 1.5                                                  # validateFaceGates -> ArgumentError
    assert lib.fdt_create(C.byref(cfg), d, len(d), None, 0, C.byref(h)) == _ffi.FDT_ERR_BAD_ARG
    assert b"minScore" in lib.fdt_last_error(None)
    cfg.min_score = 0.0
    cfg.min_face_presence = float("nan")
    assert lib.fdt_create(C.byref(cfg), d, len(d), None, 0, C.byref(h)) == _ffi.FDT_ERR_BAD_ARG
    cfg.min_face_presence = 0.5
    cfg.model = 4                                                        # fullSparse is refused
    assert lib.fdt_create(C.byref(cfg), d, len(d), None, 0, C.byref(h)) == _ffi.FDT_ERR_UNSUPPORTED
    cfg.model = 2
    assert lib.fdt_create(C.byref(cfg), None, 0, None, 0, C.byref(h)) == _ffi.FDT_ERR_BAD_ARG
    assert lib.fdt_detect_batch(None, None, 0, 1, 1, 3, 16, 0, 0, None, None, None, None) == _ffi.FDT_ERR_NOT_READY


@pytest.mark.parametrize("model,name", [(0, "frontCamera"), (1, "backCamera"), (2, "shortRange"), (3, "full")])
def test_anchors_bit_exact(lib, model, name):
    want = dp.generate_anchors(dp.ssd_options_for(name))
    n = lib.fdt_host_anchors(model, None, 0)
    assert n == want.shape[0]
    got = np.empty((n, 2), np.float64)
    lib.fdt_host_anchors(model, got.ctypes.data, n)
    assert np.array_equal(got, want)                                     # f64 bit-exact


def test_letterbox_params_match_oracle(lib):
    rng = np.random.default_rng(0)
    cases = [(1280, 720, 128), (1920, 1080, 192), (1280, 853, 128), (1, 1, 128), (10, 10, 128), (50, 50, 256),
             (3840, 2160, 192), (4000, 10, 128), (10, 4000, 128), (761, 764, 128), (6010, 4180, 192)]
    cases += [(int(rng.integers(1, 5000)), int(rng.integers(1, 5000)), int(rng.choice([128, 192, 256]))) for _ in range(200)]
    out = (C.c_int32 * 6)()
    for w, h, s in cases:
        assert lib.fdt_letterbox_params(w, h, s, s, out) == 0
        p = co.compute_letterbox_params(w, h, s, s)
        assert list(out) == [p.new_w, p.new_h, p.pad_top, p.pad_bottom, p.pad_left, p.pad_right], (w, h, s)
    assert lib.fdt_letterbox_params(0, 10, 128, 128, out) != 0


@pytest.mark.parametrize("src,dst", [(1280, 128), (720, 72), (853, 85), (1, 128), (2, 192), (37, 95), (4180, 134), (100, 100)])
def test_resize_taps_match_oracle(lib, src, dst):
    for is_x in (1, 0):
        i0 = np.empty(dst, np.int32); i1 = np.empty(dst, np.int32)
        w0 = np.empty(dst, np.int16); w1 = np.empty(dst, np.int16)
        assert lib.fdt_host_resize_taps(src, dst, is_x, i0.ctypes.data, i1.ctypes.data, w0.ctypes.data, w1.ctypes.data) == 0
        a0, a1, b0, b1 = co.resize_linear_coeffs(src, dst, bool(is_x))
        assert np.array_equal(i0, a0) and np.array_equal(i1, a1) and np.array_equal(w0, b0) and np.array_equal(w1, b1)


def test_decode_box_bit_exact(lib):
    rng = np.random.default_rng(1)
    raw = (rng.standard_normal((64, 16)) * 40).astype(np.float32)
    anchors = rng.uniform(0, 1, (64, 2))
    want = dp.decode_boxes(raw, anchors, range(64), 128)
    box = np.empty(4); kp = np.empty(12)
    for i in range(64):
        lib.fdt_host_decode_box(raw[i].ctypes.data, anchors[i, 0], anchors[i, 1], 128.0, box.ctypes.data, kp.ctypes.data)
        assert list(box) == list(want[i][:4]) and list(kp) == want[i][4]


def test_face_roi_matches_oracle(lib):
    rng = np.random.default_rng(2)
    out = np.empty(10)
    for _ in range(50):
        kp = rng.uniform(0.1, 0.9, 12)
        w, h = float(rng.integers(64, 4000)), float(rng.integers(64, 4000))
        ok = lib.fdt_host_face_roi(kp.ctypes.data, w, h, 192, out.ctypes.data)
        theta, cx, cy, size = geo.compute_face_alignment(list(kp), w, h)
        assert np.allclose(out[:4], [theta, cx, cy, size], rtol=1e-14, atol=1e-12)
        r = co.aligned_square_matrix(cx, cy, size, -theta, 192)
        assert bool(ok) == (r is not None)
        if r is not None:
            A = co.invert_affine(r[0]).reshape(-1)
            assert np.allclose(out[4:], A, rtol=1e-12, atol=1e-9)
    kp = np.full(12, 0.5)                                                 # coincident keypoints -> size 0 -> no ROI
    assert lib.fdt_host_face_roi(kp.ctypes.data, 100.0, 100.0, 192, out.ctypes.data) == 0


@pytest.mark.parametrize("model,macs,steps", [("shortRange", 30.761, 8), ("full", 105.671, None), ("backCamera", 188.750, None), ("mesh", 34.979, None)])
def test_plan_lowering(lib, model_bytes, model, macs, steps):
    buf = C.create_string_buffer(1 << 17)
    d = model_bytes[model]
    for fuse in (0, 1, 2):
        assert lib.fdt_host_plan_describe(d, len(d), fuse, buf, len(buf)) == 0, buf.value
        text = buf.value.decode()
        m = re.search(r"steps=(\d+)\s+MACs/image=([0-9.]+)M", text)
        assert abs(float(m.group(2)) - macs) < 0.002                      # SURVEY.md 2.3 MAC counts
        if fuse == 1 and steps:
            assert int(m.group(1)) == steps                               # stem + BlazeBlocks 1-6 + the image-resident tail (blocks 7-16 + both head pairs)
            assert text.count(" stem_ws ") == 1 and text.count(" block_ws ") + text.count(" block_ts ") == 6 and text.count(" tail_ws ") == 1
            assert len(re.findall(r"^\s+tail\s+\d+ ", text, re.M)) == 12     # 10 blocks + 2 head pairs inside the tail kernel
    assert lib.fdt_host_plan_describe(d[:1000], 1000, 1, buf, len(buf)) != 0   # truncated flatbuffer is rejected, not a crash
    assert lib.fdt_host_plan_describe(b"\x00" * 64, 64, 1, buf, len(buf)) != 0


@pytest.mark.parametrize("model", ["shortRange", "full", "backCamera", "mesh"])
def test_warp_specialised_plans_respect_the_hardware_limits(lib, model_bytes, model):
    """Every k_block_ws / k_stem_ws step must fit the 227 KB opt-in shared memory of an sm_100 CTA with room for the
    1 KB system reservation, use ring depths the kernel's barrier slots provide, and keep TMA boxes <= 256 per dim."""
    buf = C.create_string_buffer(1 << 17)
    d = model_bytes[model]
    assert lib.fdt_host_plan_describe(d, len(d), 1, buf, len(buf)) == 0, buf.value
    n_ws = 0
    for line in buf.value.decode().splitlines():
        m = re.search(r"^\s*\d+\s+(block_ws|block_ts|stem_ws|tail_ws)\s.*tile=(\d+)x(\d+)x(\d+) RS=(\d+) nd=(\d+) ns=(\d+) na=(\d+) no=(\d+) smem=(\d+)", line)
        if not m:
            continue
        n_ws += 1
        kind = m.group(1)
        th, tw, g, rs, nd, ns, na, no, smem = (int(x) for x in m.groups()[1:])
        assert smem <= 227 * 1024 - 1024, line
        if kind == "tail_ws":
            continue
        assert 1 <= ns <= 6 and 1 <= na <= 4 and 0 <= no <= 2, line
        if kind == "block_ws":
            assert nd in (8, 12) and rs in (1, 2, 4) and g * th * tw <= 128, line
            assert (th - 1) * 2 + 3 <= 256 and (tw - 1) * 2 + 3 <= 256 and g <= 256, line
    assert n_ws >= 8, "the conv stack should run on the warp-specialised kernels"


def test_detectors_and_mesh_run_on_tensor_core_kernels_only(lib, model_bytes):
    """no CUDA-core convolution step (dwpw / gemm_conv / naive_conv) is left in the production plans of the
    detectors and the face-landmark net; the full-range layers wider than 128 channels (12x12x144..256, 6x6x384) run as
    k_chain_wide programs whose W blocks respect the kernel's limits."""
    buf = C.create_string_buffer(1 << 17)
    for model in ("shortRange", "full", "backCamera", "mesh"):
        d = model_bytes[model]
        assert lib.fdt_host_plan_describe(d, len(d), 1, buf, len(buf)) == 0, buf.value
        text = buf.value.decode()
        kinds = re.findall(r"^\s*\d+ (\w+)\s+in=", text, re.M)
        assert kinds and not set(kinds) & {"dwpw", "gemm_conv", "naive_conv", "maxpool", "add", "act", "padc", "stem"}, (model, sorted(set(kinds)))
    d = model_bytes["full"]
    assert lib.fdt_host_plan_describe(d, len(d), 1, buf, len(buf)) == 0
    text = buf.value.decode()
    wide = re.findall(r"k_chain_wide: (\d+) W blocks \(ring (\d+) x (\d+) B\), act (\d+) floats, (\d+) HBM", text)
    assert len(wide) == 5                                                   # [18..21] [22] [23] [24..31] [32]
    for nblk, depth, wbuf, act, nres in wide:
        assert int(nblk) <= 128 and 2 <= int(depth) <= 4 and int(wbuf) <= 32768 and int(nres) <= 2
    layers = re.findall(r"C (\d+)->(\d+) K16=(\d+) Npad=(\d+) .* wide: groups=(\d+) blocks/group=(\d+) Kchunks=(\d+)", text)
    assert len(layers) == 15
    for cin, cout, k16, npad, ng, nbg, nk in (tuple(int(x) for x in l) for l in layers):
        assert nk == (k16 + 127) // 128 and ng * nbg == (npad + 127) // 128 and nbg <= 3
        assert max(cin, cout) > 128                                           # only layers no per-layer tensor-core kernel takes


def test_malformed_models_are_rejected_not_crashed(model_bytes):
    """The model bytes reach the flatbuffer reader through the public API (fdt_create / FaceDetector.create
    detectorBytes): truncated and bit-flipped buffers must come back as FDT_ERR_MODEL (or parse), never crash.
    Runs in a subprocess so that a wild read fails the test instead of killing the suite."""
    import subprocess
    import sys
    script = r"""
import sys, ctypes as C
sys.path.insert(0, %r)
import numpy as np
from face_detection_tflite_b200 import _ffi
lib = _ffi.load()
buf = C.create_string_buffer(1 << 16)
rng = np.random.default_rng(7)
n_ok = n_bad = 0
for name in ("face_detection_short_range.tflite", "face_landmark.tflite", "iris_landmark.tflite"):
    good = open(%r + "/assets/models/" + name, "rb").read()
    assert lib.fdt_host_plan_describe(good, len(good), 1, buf, len(buf)) == 0
    cases = [good[:k] for k in (0, 3, 16, 64, 1000, len(good) // 2, len(good) - 1)]
    head = min(len(good), 200000)
    for _ in range(250):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            # structural bytes live at the head (root, vtables of operators / tensors near the end): hit both regions
            pos = int(rng.integers(0, 4096)) if rng.random() < 0.3 else int(len(good) - 1 - rng.integers(0, min(len(good), 60000)))
            b[pos] = int(rng.integers(0, 256))
        cases.append(bytes(b))
    for c in cases:
        rc = lib.fdt_host_plan_describe(c, len(c), 1, buf, len(buf))
        assert rc in (0, _ffi.FDT_ERR_MODEL), rc
        n_ok += rc == 0
        n_bad += rc != 0
print("survived", n_ok, n_bad)
""" % (str(ROOT), str(ROOT))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "survived" in r.stdout, (r.returncode, r.stdout[-300:], r.stderr[-1500:])
    n_bad = int(r.stdout.split()[-1])
    assert n_bad >= 30            # the mutations do reach the validation paths


def test_incremental_tile_walk_equals_division(lib):
    """csrc/tile_walk.h: the persistent kernels advance (image, tile row, tile column) with adds and carries; the walk must
    visit exactly the tiles first + i * stride decomposed by division (what the kernels did before)."""
    import ctypes as C
    import random
    rng = random.Random(7)
    cases = [(0, 148, 4, 4, 500), (147, 148, 4, 4, 500), (3, 296, 4, 8, 300), (5, 1, 1, 1, 50), (0, 16, 4, 4, 64),
             (10, 148, 12, 12, 400), (1, 444, 2, 4, 200), (0, 7, 3, 5, 1000)]
    for _ in range(40):
        tx, ty = rng.randint(1, 16), rng.randint(1, 16)
        cases.append((rng.randint(0, 300), rng.randint(1, 600), tx, ty, rng.randint(1, 400)))
    for first, stride, tiles_x, tiles_y, n in cases:
        out = (C.c_int32 * (3 * n))()
        assert lib.fdt_host_tile_walk(first, stride, tiles_x, tiles_y, n, out) == 0
        tpi = tiles_x * tiles_y
        for i in range(n):
            t = first + i * stride
            b, r = divmod(t, tpi)
            assert (out[3 * i], out[3 * i + 1], out[3 * i + 2]) == (b, r // tiles_x, r % tiles_x), (first, stride, tiles_x, tiles_y, i)
    assert lib.fdt_host_tile_walk(0, 0, 4, 4, 1, (C.c_int32 * 3)()) != 0      # a zero stride is refused

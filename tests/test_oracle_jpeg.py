"""CPU: the host Huffman decoder (libfdt_cuda.so, no device needed) + the numpy restatement of libjpeg-turbo's IDCT /
upsampling / colour conversion (oracle/jpeg_ops.py) against the REAL cv2.imdecode of this container - the decoder the
reference calls (cv.imdecode, /root/reference/lib/src/face_detector.dart:477-485).  Byte equality."""
import ctypes as C
import struct
from pathlib import Path

import cv2
import numpy as np
import pytest

from face_detection_tflite_b200 import _ffi
from oracle import jpeg_ops

ROOT = Path(__file__).resolve().parents[1]
SAMPLES = sorted((ROOT / "assets" / "samples").glob("*.jp*g"))


def host_decode(lib, data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, np.uint8)
    info = (C.c_int32 * 8)()
    rc = lib.fdt_host_jpeg_info(buf.ctypes.data, buf.size, info)
    assert rc == 0, rc
    meta = dict(width=info[0], height=info[1], ncomp=info[2], progressive=info[3], orientation=info[4], hmax=info[5], vmax=info[6])
    comps = []
    for c in range(meta["ncomp"]):
        dims = (C.c_int32 * 6)()
        q = np.zeros(64, np.uint16)
        assert lib.fdt_host_jpeg_coefficients(buf.ctypes.data, buf.size, c, None, 0, dims, q.ctypes.data) == 0
        coef = np.zeros((dims[1], dims[0], 64), np.int16)
        assert lib.fdt_host_jpeg_coefficients(buf.ctypes.data, buf.size, c, coef.ctypes.data, coef.size, dims, q.ctypes.data) == 0
        comps.append(dict(coef=coef, q=q, dw=dims[2], dh=dims[3], h=dims[4], v=dims[5]))
    return jpeg_ops.decode_from_coefficients(meta, comps)


def synth_image(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 3) % 256], -1).astype(np.float64)
    img += rng.normal(0, 25, img.shape)
    img[h // 4: h // 2, w // 4: w // 2] = rng.integers(0, 256, 3)                # flat saturated patch: sharp chroma edges
    return np.clip(img, 0, 255).astype(np.uint8)


def with_exif_orientation(jpeg: bytes, orientation: int) -> bytes:
    tiff = b"MM\x00\x2a\x00\x00\x00\x08" + struct.pack(">H", 1) + struct.pack(">HHIHH", 0x0112, 3, 1, orientation, 0) + b"\x00\x00\x00\x00"
    payload = b"Exif\x00\x00" + tiff
    return jpeg[:2] + b"\xff\xe1" + struct.pack(">H", len(payload) + 2) + payload + jpeg[2:]


@pytest.fixture(scope="module")
def lib():
    return _ffi.load()


@pytest.mark.parametrize("path", SAMPLES, ids=[p.name for p in SAMPLES])
def test_sample_jpegs_decode_like_cv2(lib, path):
    data = path.read_bytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    got = host_decode(lib, data)
    assert got.shape == want.shape
    assert np.array_equal(got, want), "max diff %d" % np.abs(got.astype(int) - want).max()


SAMPLING = {"444": 0x111111, "422": 0x211111, "420": 0x221111, "440": 0x121111}
CASES = [(64, 48, "420", 90, False, 0), (17, 13, "420", 75, False, 0), (33, 31, "422", 60, False, 0), (50, 70, "440", 85, False, 0),
         (40, 40, "444", 95, False, 0), (129, 65, "420", 30, True, 0), (97, 55, "422", 80, True, 0), (200, 120, "420", 85, False, 7),
         (8, 8, "420", 90, False, 0), (3, 5, "420", 90, False, 0), (2, 2, "422", 90, False, 0), (256, 256, "420", 100, False, 0),
         (61, 47, "444", 50, True, 4)]


@pytest.mark.parametrize("w,h,samp,quality,progressive,rst", CASES)
def test_encoded_variants_decode_like_cv2(lib, w, h, samp, quality, progressive, rst):
    img = synth_image(w, h, w * 1000 + h)
    params = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[samp],
              cv2.IMWRITE_JPEG_PROGRESSIVE, int(progressive), cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    ok, enc = cv2.imencode(".jpg", img, params)
    assert ok
    data = enc.tobytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    got = host_decode(lib, data)
    assert np.array_equal(got, want), "max diff %d" % np.abs(got.astype(int) - want).max()


def test_greyscale_jpeg(lib):
    img = synth_image(75, 49, 5)[:, :, 0]
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 80])
    data = enc.tobytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(host_decode(lib, data), want)


@pytest.mark.parametrize("orientation", range(1, 9))
def test_exif_orientation_like_cv2(lib, orientation):
    img = synth_image(40, 24, 11)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    data = with_exif_orientation(enc.tobytes(), orientation)
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    got = host_decode(lib, data)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)


def test_undecodable_bytes_are_format_errors(lib):
    for data in (b"", b"not a jpeg", b"\xff\xd8\xff\xd9", SAMPLES[0].read_bytes()[:300]):
        buf = np.frombuffer(data + b"\x00", np.uint8)
        info = (C.c_int32 * 8)()
        assert lib.fdt_host_jpeg_info(buf.ctypes.data, len(data), info) == _ffi.FDT_ERR_FORMAT
        assert cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR) is None if data else True


def test_bit_flipped_streams_never_crash(lib):
    data = bytearray(SAMPLES[-1].read_bytes())
    rng = np.random.default_rng(3)
    for _ in range(40):
        d = bytearray(data)
        for pos in rng.integers(2, len(d), 8):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        cut = int(rng.integers(len(d) // 2, len(d)))
        buf = np.frombuffer(bytes(d[:cut]), np.uint8)
        info = (C.c_int32 * 8)()
        assert lib.fdt_host_jpeg_info(buf.ctypes.data, buf.size, info) in (0, _ffi.FDT_ERR_FORMAT, _ffi.FDT_ERR_UNSUPPORTED)

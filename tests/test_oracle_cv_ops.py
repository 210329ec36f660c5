"""Pins the oracle's integer restatements of OpenCV (oracle/cv_ops.py) bit-exactly against the real
cv2 of this image and against the committed cv2 golden vectors."""
import math

import cv2
import numpy as np
import pytest

from oracle import cv_ops as co

RESIZE_CASES = [((720, 1280), (128, 72)), ((1080, 1920), (192, 108)), ((853, 1280), (128, 85)),
                ((2160, 3840), (128, 72)), ((100, 100), (128, 128)), ((1, 1), (128, 128)), ((10, 10), (5, 5)),
                ((50, 37), (128, 95)), ((64, 64), (32, 32)), ((480, 640), (128, 96)), ((481, 641), (127, 95)),
                ((3, 500), (128, 1)), ((2, 2), (192, 192)), ((200, 300), (100, 100))]


@pytest.mark.parametrize("src_hw,dst_wh", RESIZE_CASES)
def test_resize_linear_bit_exact(src_hw, dst_wh):
    rng = np.random.default_rng(hash((src_hw, dst_wh)) & 0xffff)
    src = rng.integers(0, 256, (*src_hw, 3), dtype=np.uint8)
    want = cv2.resize(src, dst_wh, interpolation=cv2.INTER_LINEAR)
    got = co.resize_linear_u8(src, *dst_wh)
    assert np.array_equal(want, got)


def test_exact_tenfold_downscale_is_2x2_average():
    # for the benchmark geometry every output pixel is (a+b+c+d+2)>>2 of a 2x2 block
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    got = co.resize_linear_u8(src, 128, 72)
    s = src.astype(np.int32)
    want = (s[4::10, 4::10] + s[4::10, 5::10] + s[5::10, 4::10] + s[5::10, 5::10] + 2) >> 2
    assert np.array_equal(got, want.astype(np.uint8))


@pytest.mark.parametrize("trial", range(10))
def test_warp_affine_bit_exact(trial):
    rng = np.random.default_rng(100 + trial)
    sh, sw = [(720, 1280), (1080, 1920), (300, 200), (64, 64), (853, 1280)][trial % 5]
    src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    if trial % 2:
        src = cv2.GaussianBlur(src, (9, 9), 3)
    cx, cy = rng.uniform(-50, sw + 50), rng.uniform(-50, sh + 50)
    size, th = rng.uniform(20, 900), rng.uniform(-math.pi, math.pi)
    out = [192, 64, 112, None][trial % 4]
    M, o = co.aligned_square_matrix(cx, cy, size, th, out)
    sc = o / co.dart_round(size)
    R = cv2.getRotationMatrix2D((float(np.float32(cx)), float(np.float32(cy))), -th * 180 / math.pi, sc)
    oc = o / 2 + 0.5 * (sc - 1)
    R[0, 2] += oc - cx
    R[1, 2] += oc - cy
    assert np.abs(R - M).max() < 1e-9
    want = cv2.warpAffine(src, R, (o, o), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
    assert np.array_equal(want, co.warp_affine_u8(src, R, o, o))
    assert np.array_equal(want, co.warp_affine_u8(src, M, o, o))


def test_extract_aligned_square_degenerate_size():
    src = np.zeros((10, 10, 3), np.uint8)
    assert co.extract_aligned_square(src, 5, 5, 0.4, 0.0, 192) is None      # round(size) <= 0 -> null (helpers.dart:591-592)
    assert co.extract_aligned_square(src, 5, 5, 0.6, 0.0, 192).shape == (192, 192, 3)


def test_outsize_equal_to_round_size_is_plain_crop():
    # preprocessing_equivalence_test.dart:123-143: outSize == round(size) must be byte-identical to the unscaled warp
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    a = co.extract_aligned_square(src, 200, 150, 100.2, 0.3, None)
    b = co.extract_aligned_square(src, 200, 150, 100.2, 0.3, 100)
    assert np.array_equal(a, b)


def test_letterbox_params_known_cases():
    p = co.compute_letterbox_params(1280, 720, 128, 128)
    assert (p.new_w, p.new_h, p.pad_top, p.pad_bottom, p.pad_left, p.pad_right) == (128, 72, 28, 28, 0, 0)
    p = co.compute_letterbox_params(1920, 1080, 192, 192)
    assert (p.new_w, p.new_h, p.pad_top, p.pad_bottom) == (192, 108, 42, 42)
    p = co.compute_letterbox_params(1280, 853, 128, 128)
    assert (p.new_w, p.new_h, p.pad_top, p.pad_bottom) == (128, 85, 21, 22)
    p = co.compute_letterbox_params(1, 1, 128, 128)
    assert (p.new_w, p.new_h, p.pad_top, p.pad_left) == (128, 128, 0, 0)
    p = co.compute_letterbox_params(4000, 10, 128, 128)      # extreme aspect ratio: never a 0-pixel resize
    assert p.new_h >= 1 and p.pad_top + p.new_h + p.pad_bottom == 128


def test_letterbox_matches_cv2_golden(golden, sample_images):
    for model, S in (("shortRange", 128), ("full", 192), ("backCamera", 256)):
        for name, img in sample_images.items():
            got, p = co.letterbox_u8(img, S, S)
            assert np.array_equal(got, golden["%s/%s/letterboxed_cv2" % (model, name)])
            assert [p.new_w, p.new_h, p.pad_top, p.pad_bottom, p.pad_left, p.pad_right] == list(golden["%s/%s/lbparams" % (model, name)])


def test_normalize_range_and_equivalence():
    # preprocessing_equivalence_test.dart:35-56: the SIMD (f32 fma) and scalar formulations agree to 1e-5
    v = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2)
    t = co.normalize_bgr_u8(v)
    assert t.min() == -1.0 and abs(t.max() - 1.0) < 1e-6
    f32 = (v[..., ::-1].astype(np.float32) * np.float32(1 / 127.5) + np.float32(-1.0)).astype(np.float32)
    assert np.abs(t - f32).max() < 1e-5
    assert np.all(co.normalize_bgr_u8(np.zeros((2, 2, 3), np.uint8)) == -1.0)   # padding pixels are exactly -1


def test_crop_matches_cv2_golden(golden, sample_images):
    for model in ("shortRange", "full", "backCamera"):
        for name, img in sample_images.items():
            k = "%s/%s/crop0_cv2" % (model, name)
            if k not in golden:
                continue
            theta, cx, cy, size = golden["%s/%s/align0" % (model, name)]
            got = co.extract_aligned_square(img, cx, cy, size, -theta, 192)
            assert np.array_equal(got, golden[k])

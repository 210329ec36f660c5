"""TEST INFRASTRUCTURE (oracle): CPU restatement of libjpeg-turbo's default decompression path after the entropy decoder,
i.e. what cv::imdecode (the reference's decoder, /root/reference/lib/src/face_detector.dart:477-485 -> cv.imdecode) does to
the quantised DCT coefficients: dequantisation + jpeg_idct_islow (jidctint.c), fancy chroma upsampling (jdsample.c h2v1 /
h2v2 / h1v2 fancy upsample and their edge rules), YCbCr -> RGB with the 16-bit fixed-point tables (jdcolor.c), and OpenCV's
EXIF orientation transform (modules/imgcodecs/src/loadsave.cpp ExifTransform).  libjpeg-turbo is an un-vendored dependency of
opencv_dart 2.2.1+4 (OpenCV 4.x builds it from 3rdparty/libjpeg-turbo); parity is pinned on the real cv2.imdecode of this
container (4.13.0, libjpeg-turbo 3.1.2) in tests/test_oracle_jpeg.py.  numpy, integer arithmetic only."""
from __future__ import annotations

import numpy as np

CONST_BITS, PASS1_BITS = 13, 2
F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
         f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct8(v, shift):
    """v: [..., 8] int64 along the last axis -> [..., 8] (one pass of the LL&M inverse DCT, jidctint.c:233-330)."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F["f0_541"]
    tmp2 = z1 + z3 * (-F["f1_847"])
    tmp3 = z1 + z2 * F["f0_765"]
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F["f1_175"]
    tmp0 = tmp0 * F["f0_298"]; tmp1 = tmp1 * F["f2_053"]; tmp2 = tmp2 * F["f3_072"]; tmp3 = tmp3 * F["f1_501"]
    z1 = z1 * -F["f0_899"]; z2 = z2 * -F["f2_562"]; z3 = z3 * -F["f1_961"] + z5; z4 = z4 * -F["f0_390"] + z5
    tmp0 = tmp0 + z1 + z3; tmp1 = tmp1 + z2 + z4; tmp2 = tmp2 + z2 + z3; tmp3 = tmp3 + z1 + z4
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], -1)
    return _descale(out, shift)


def _range_limit(x):
    x = x & 1023
    return np.where(x < 128, x + 128, np.where(x < 512, 255, np.where(x < 896, 0, x - 896))).astype(np.uint8)


def idct_plane(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """coef [bh, bw, 64] int16 (natural order), q [64] -> u8 plane [bh * 8, bw * 8]."""
    bh, bw, _ = coef.shape
    d = coef.astype(np.int64) * q.astype(np.int64)
    blk = d.reshape(bh, bw, 8, 8)                          # [row, col]
    ws = _idct8(blk.transpose(0, 1, 3, 2), CONST_BITS - PASS1_BITS).transpose(0, 1, 3, 2)     # pass 1 along columns
    out = _idct8(ws, CONST_BITS + PASS1_BITS + 3)                                               # pass 2 along rows
    px = _range_limit(out)
    return px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def upsample_fancy(pl: np.ndarray, dw: int, dh: int, hs: int, vs: int, W: int, H: int) -> np.ndarray:
    """Chroma plane (u8, at least [dh, dw]) -> [H, W] int32, libjpeg's fancy upsampling for (hs, vs) in {1, 2}^2."""
    c = pl[:dh, :dw].astype(np.int32)
    if hs == 1 and vs == 1:
        return c[:H, :W]
    if vs == 2:
        # vertical neighbours: output row 2r -> row r - 1 (row 0 repeats), 2r + 1 -> row r + 1 (last row repeats)
        up = np.concatenate([c[:1], c[:-1]], 0)
        dn = np.concatenate([c[1:], c[-1:]], 0)
        if hs == 1:
            rows = np.empty((2 * dh, dw), np.int32)
            rows[0::2] = (3 * c + up + 1) >> 2
            rows[1::2] = (3 * c + dn + 2) >> 2
            return rows[:H, :W]
        sums = np.empty((2 * dh, dw), np.int32)
        sums[0::2] = 3 * c + up
        sums[1::2] = 3 * c + dn
        if dw <= 2:
            return np.repeat(np.repeat(c, 2, 0), 2, 1)[:H, :W]
        out = np.empty((2 * dh, 2 * dw), np.int32)
        last = np.concatenate([sums[:, :1], sums[:, :-1]], 1)
        nxt = np.concatenate([sums[:, 1:], sums[:, -1:]], 1)
        out[:, 0::2] = (sums * 3 + last + 8) >> 4
        out[:, 1::2] = (sums * 3 + nxt + 7) >> 4
        out[:, 0] = (sums[:, 0] * 4 + 8) >> 4
        out[:, -1] = (sums[:, -1] * 4 + 7) >> 4
        return out[:H, :W]
    # hs == 2, vs == 1
    if dw <= 2:
        return np.repeat(c, 2, 1)[:H, :W]
    out = np.empty((dh, 2 * dw), np.int32)
    last = np.concatenate([c[:, :1], c[:, :-1]], 1)
    nxt = np.concatenate([c[:, 1:], c[:, -1:]], 1)
    out[:, 0::2] = (3 * c + last + 1) >> 2
    out[:, 1::2] = (3 * c + nxt + 2) >> 2
    out[:, 0] = c[:, 0]
    out[:, -1] = c[:, -1]
    return out[:H, :W]


def ycc_to_bgr(Y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    y = Y.astype(np.int32)
    cb = cb.astype(np.int32) - 128
    cr = cr.astype(np.int32) - 128
    r = np.clip(y + ((91881 * cr + 32768) >> 16), 0, 255)
    b = np.clip(y + ((116130 * cb + 32768) >> 16), 0, 255)
    g = np.clip(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16), 0, 255)
    return np.stack([b, g, r], -1).astype(np.uint8)


def exif_transform(img: np.ndarray, orientation: int) -> np.ndarray:
    """OpenCV ExifTransform (loadsave.cpp): 1 none, 2 flip h, 3 flip both, 4 flip v, 5 transpose, 6 transpose + flip h,
    7 transpose + flip both, 8 transpose + flip v."""
    if orientation in (5, 6, 7, 8):
        img = img.transpose(1, 0, 2)
    if orientation in (2, 6):
        img = img[:, ::-1]
    elif orientation in (3, 7):
        img = img[::-1, ::-1]
    elif orientation in (4, 8):
        img = img[::-1]
    return np.ascontiguousarray(img)


def decode_from_coefficients(info, comps) -> np.ndarray:
    """info: dict(width, height, ncomp, orientation, hmax, vmax); comps: list of dict(coef [bh,bw,64], q [64], dw, dh, h, v)."""
    W, H = info["width"], info["height"]
    planes = [idct_plane(c["coef"], c["q"]) for c in comps]
    Y = planes[0][:H, :W]
    if info["ncomp"] == 1:
        img = np.stack([Y, Y, Y], -1)
    else:
        hs, vs = info["hmax"] // comps[1]["h"], info["vmax"] // comps[1]["v"]
        cb = upsample_fancy(planes[1], comps[1]["dw"], comps[1]["dh"], hs, vs, W, H)
        cr = upsample_fancy(planes[2], comps[2]["dw"], comps[2]["dh"], hs, vs, W, H)
        img = ycc_to_bgr(Y, cb, cr)
    return exif_transform(img, info["orientation"])

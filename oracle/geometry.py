"""ORACLE (test infrastructure, not product code).

Restates lib/src/shared/face_geometry.dart:17-73 (alignment, mesh back-projection) and
lib/src/util/helpers.dart:138-172 (_unpackLandmarks), all in float64 as Dart doubles.
"""
from __future__ import annotations

import math

import numpy as np

# FaceLandmarkType order (face_types.dart:19-37)
LEFT_EYE, RIGHT_EYE, NOSE_TIP, MOUTH, LEFT_TRAGION, RIGHT_TRAGION = range(6)


def compute_face_alignment(kp, img_w: float, img_h: float):
    lx, ly = kp[LEFT_EYE * 2] * img_w, kp[LEFT_EYE * 2 + 1] * img_h
    rx, ry = kp[RIGHT_EYE * 2] * img_w, kp[RIGHT_EYE * 2 + 1] * img_h
    mx, my = kp[MOUTH * 2] * img_w, kp[MOUTH * 2 + 1] * img_h
    ecx, ecy = (lx + rx) * 0.5, (ly + ry) * 0.5
    vex, vey = rx - lx, ry - ly
    vmx, vmy = mx - ecx, my - ecy
    theta = math.atan2(vey, vex)
    eye_dist = math.sqrt(vex * vex + vey * vey)
    mouth_dist = math.sqrt(vmx * vmx + vmy * vmy)
    size = max(mouth_dist * 3.6, eye_dist * 4.0)
    return theta, ecx + vmx * 0.1, ecy + vmy * 0.1, size


def clamp01(v: float) -> float:
    return 0.0 if v < 0.0 else (1.0 if v > 1.0 else v)


def unpack_landmarks(flat, in_w: int, in_h: int, padding, clamp=True, normalize_z=False):
    pt, pb, pl, pr = padding
    inv_sx = 1.0 / (1.0 - (pl + pr))
    inv_sy = 1.0 / (1.0 - (pt + pb))
    inv_w = 1.0 / in_w
    inv_h = 1.0 / in_h
    flat = np.asarray(flat, np.float32).reshape(-1)
    n = flat.shape[0] // 3
    out = np.empty((n, 3), np.float64)
    for i in range(n):
        x = (float(flat[3 * i]) * inv_w - pl) * inv_sx
        y = (float(flat[3 * i + 1]) * inv_h - pt) * inv_sy
        z = float(flat[3 * i + 2]) * inv_w * inv_sx if normalize_z else float(flat[3 * i + 2])
        if clamp:
            x, y = clamp01(x), clamp01(y)
        out[i] = (x, y, z)
    return out


def transform_mesh_to_absolute(lm_norm, cx, cy, size, theta):
    ct, st = math.cos(theta), math.sin(theta)
    sct, sst = size * ct, size * st
    tx = cx - 0.5 * sct + 0.5 * sst
    ty = cy - 0.5 * sst - 0.5 * sct
    lm = np.asarray(lm_norm, np.float64)
    out = np.empty_like(lm)
    out[:, 0] = tx + sct * lm[:, 0] - sst * lm[:, 1]
    out[:, 1] = ty + sst * lm[:, 0] + sct * lm[:, 1]
    out[:, 2] = lm[:, 2] * size
    return out

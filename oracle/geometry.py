"""ORACLE (test infrastructure, not product code).

Restates lib/src/shared/face_geometry.dart:17-73 (alignment, mesh back-projection), :109-125
(transformIrisNormToAbsolute), :155-168 (eyeRoisFromMesh), lib/src/util/helpers.dart:138-172
(_unpackLandmarks), lib/src/shared/face_types.dart:976-997 (irisCenterFromPoints) and
lib/src/models/face_embedding.dart:362-400 (embedding alignment + L2 normalisation), all in float64
as Dart doubles.
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


def eye_rois_from_mesh(mesh_abs):
    """eyeRoisFromMesh (face_geometry.dart:155-168): [(cx, cy, size, theta)] for the left (33/133) and right
    (362/263) eye from the absolute-pixel f64 mesh."""
    def from_corners(a, b):
        p0, p1 = mesh_abs[a], mesh_abs[b]
        cx = (float(p0[0]) + float(p1[0])) * 0.5
        cy = (float(p0[1]) + float(p1[1])) * 0.5
        dx = float(p1[0]) - float(p0[0])
        dy = float(p1[1]) - float(p0[1])
        eye_dist = math.sqrt(dx * dx + dy * dy)
        return (cx, cy, eye_dist * 2.3, math.atan2(dy, dx))
    return [from_corners(33, 133), from_corners(362, 263)]


def transform_iris_norm_to_absolute(lm_norm, roi, is_right: bool):
    """transformIrisNormToAbsolute (face_geometry.dart:109-125); roi = (cx, cy, size, theta)."""
    cx, cy, s, theta = roi
    ct, st = math.cos(theta), math.sin(theta)
    out = []
    for p in lm_norm:
        px = (1.0 - float(p[0])) if is_right else float(p[0])
        lx2 = (px - 0.5) * s
        ly2 = (float(p[1]) - 0.5) * s
        out.append([cx + lx2 * ct - ly2 * st, cy + lx2 * st + ly2 * ct, float(p[2])])
    return np.array(out, np.float64).reshape(-1, 3)


def iris_center_from_points(pts):
    """irisCenterFromPoints (face_types.dart:976-997): the point closest to the centroid (first on ties)."""
    pts = [tuple(float(v) for v in p) for p in pts]
    if not pts:
        return (0.0, 0.0, 0.0)
    if len(pts) == 1:
        return pts[0]
    cx = cy = 0.0
    for p in pts:
        cx += p[0]
        cy += p[1]
    cx /= len(pts)
    cy /= len(pts)
    best, best_d = 0, float("inf")
    for i, p in enumerate(pts):
        dx, dy = p[0] - cx, p[1] - cy
        d = dx * dx + dy * dy
        if d < best_d:
            best_d, best = d, i
    return pts[best]


def compute_embedding_alignment(left_eye, right_eye):
    """computeEmbeddingAlignment (face_embedding.dart:362-384) -> (theta, cx, cy, size)."""
    dx = float(right_eye[0]) - float(left_eye[0])
    dy = float(right_eye[1]) - float(left_eye[1])
    theta = math.atan2(dy, dx)
    eye_dist = math.sqrt(dx * dx + dy * dy)
    size = eye_dist * 2.5
    ecx = (float(left_eye[0]) + float(right_eye[0])) * 0.5
    ecy = (float(left_eye[1]) + float(right_eye[1])) * 0.5
    ct, st = math.cos(theta), math.sin(theta)
    oy = size * 0.15
    return theta, ecx - oy * st, ecy + oy * ct, size


def normalize_embedding(e):
    """_normalizeEmbeddingImpl (face_embedding.dart:386-400): f64 norm accumulated in order, f32 result."""
    e = np.asarray(e, np.float32)
    norm = 0.0
    for v in e:
        norm += float(v) * float(v)
    norm = math.sqrt(norm)
    if norm > 0:
        return np.array([float(v) / norm for v in e], np.float32)
    return e

"""Synthetic benchmark / parity inputs.

(N) noise frames: uniform u8 noise, expected 0 faces (exercises the threshold-reject path).
(F) composited frames: 8x8-block low-frequency background plus k mod 5 face tiles cut from the
    sample photos, pasted at non-overlapping positions — seeded, so every box reproduces them."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional

import numpy as np

ASSETS = Path(__file__).resolve().parent.parent / "assets" / "samples"
# face tiles: (file, x0, y0, side) squares around the faces of the sample photos
_TILES = [("landmark-ex1.jpg", 420, 60, 560), ("iris-detection-ex1.jpg", 120, 120, 560)]
_tile_cache: Optional[List[np.ndarray]] = None


def _tiles() -> List[np.ndarray]:
    global _tile_cache
    if _tile_cache is None:
        import cv2
        out = []
        for name, x0, y0, side in _TILES:
            img = cv2.imread(str(ASSETS / name))
            h, w = img.shape[:2]
            x0, y0 = max(0, min(x0, w - side)), max(0, min(y0, h - side))
            side = min(side, w - x0, h - y0)
            out.append(np.ascontiguousarray(img[y0:y0 + side, x0:x0 + side]))
        _tile_cache = out
    return _tile_cache


def noise_frames(count: int, width: int, height: int, seed: int = 0, channels: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (count, height, width, channels), dtype=np.uint8)


def face_frame(k: int, width: int, height: int, min_side: int = 160, max_side: int = 560) -> np.ndarray:
    """Frame k of the (F) distribution: k mod 5 face tiles on a low-frequency background."""
    import cv2
    rng = np.random.default_rng(1000 + k)
    bg = rng.integers(40, 216, ((height + 7) // 8, (width + 7) // 8, 3), dtype=np.uint8)
    frame = np.ascontiguousarray(np.repeat(np.repeat(bg, 8, 0), 8, 1)[:height, :width])
    tiles = _tiles()
    placed = []
    for _ in range(k % 5):
        side = int(rng.integers(min_side, min(max_side, height, width) + 1))
        for _try in range(50):
            x = int(rng.integers(0, width - side + 1))
            y = int(rng.integers(0, height - side + 1))
            if all(x + side <= px or px + ps <= x or y + side <= py or py + ps <= y for px, py, ps in placed):
                t = tiles[int(rng.integers(0, len(tiles)))]
                frame[y:y + side, x:x + side] = cv2.resize(t, (side, side), interpolation=cv2.INTER_AREA)
                placed.append((x, y, side))
                break
    return frame


def face_frames(count: int, width: int, height: int, start: int = 0, **kw) -> np.ndarray:
    return np.stack([face_frame(start + k, width, height, **kw) for k in range(count)])

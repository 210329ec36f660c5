import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ASSETS = ROOT / "assets"
GOLDEN = ROOT / "tests" / "golden" / "golden_samples.npz"
SAMPLES = ["landmark-ex1.jpg", "iris-detection-ex1.jpg", "group-shot-bounding-box-ex1.jpeg"]
MODEL_FILES = {"shortRange": "face_detection_short_range.tflite", "full": "face_detection_full_range.tflite",
               "backCamera": "face_detection_back.tflite"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def sample_images():
    import cv2
    return {s: cv2.imread(str(ASSETS / "samples" / s)) for s in SAMPLES}


@pytest.fixture(scope="session")
def model_bytes():
    d = {k: (ASSETS / "models" / v).read_bytes() for k, v in MODEL_FILES.items()}
    d["mesh"] = (ASSETS / "models" / "face_landmark.tflite").read_bytes()
    return d


@pytest.fixture(scope="session")
def lib():
    from face_detection_tflite_b200 import build, _ffi
    build.build()
    return _ffi.load()

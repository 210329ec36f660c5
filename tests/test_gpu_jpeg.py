"""GPU: the JPEG front end behind detectFacesFromBytes (reference: /root/reference/lib/src/face_detector.dart:477-485, cv.imdecode)
through the C ABI - device IDCT / upsampling / colour conversion / EXIF orientation against the real cv2.imdecode (byte equality),
and bytes -> faces against Mat -> faces."""
import cv2
import numpy as np
import pytest

from test_oracle_jpeg import CASES, SAMPLES, SAMPLING, synth_image, with_exif_orientation

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fdt():
    import face_detection_tflite_b200 as m
    return m


@pytest.fixture(scope="module")
def det(fdt):
    d = fdt.FaceDetector.create(fdt.FaceDetectionModel.backCamera)
    yield d
    d.dispose()


@pytest.mark.parametrize("path", SAMPLES, ids=[p.name for p in SAMPLES])
def test_sample_jpegs_bit_exact(det, path):
    data = path.read_bytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    got = det.decodeImage(data)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("w,h,samp,quality,progressive,rst", CASES)
def test_encoded_variants_bit_exact(det, w, h, samp, quality, progressive, rst):
    img = synth_image(w, h, w * 1000 + h)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[samp],
                                         cv2.IMWRITE_JPEG_PROGRESSIVE, int(progressive), cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    data = enc.tobytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(det.decodeImage(data), want)


def test_greyscale_and_orientations_bit_exact(det):
    g = synth_image(75, 49, 5)[:, :, 0]
    data = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes()
    assert np.array_equal(det.decodeImage(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))
    base = cv2.imencode(".jpg", synth_image(40, 24, 11), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes()
    for o in range(1, 9):
        data = with_exif_orientation(base, o)
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        got = det.decodeImage(data)
        assert got.shape == want.shape and np.array_equal(got, want), o


@pytest.mark.parametrize("path", SAMPLES, ids=[p.name for p in SAMPLES])
def test_bytes_to_faces_equals_mat_to_faces(fdt, det, path):
    """detectFacesFromBytes(bytes) == detectFacesFromMat(cv.imdecode(bytes)) - the reference's own definition (:487-520) - in the
    reference's default mode (full)."""
    data = path.read_bytes()
    mat = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    a = det.detectFacesFromBytes(data)
    b = det.detectFacesFromMat(mat)
    assert len(a) == len(b) and len(a) >= 1
    for fa, fb in zip(a, b):
        ba, bb = fa.detectionData.boundingBox, fb.detectionData.boundingBox
        assert (ba.xmin, ba.ymin, ba.xmax, ba.ymax) == (bb.xmin, bb.ymin, bb.xmax, bb.ymax)
        assert fa.detectionData.score == fb.detectionData.score
        assert np.array_equal(fa.mesh.packed, fb.mesh.packed)
        assert np.array_equal(fa.irisPacked, fb.irisPacked)
    assert det.lastLaunchCount() > 4                                  # the decode kernels are counted with the rest


def test_undecodable_bytes_raise_format_exception(fdt, det):
    for data in (b"", b"definitely not an image", SAMPLES[0].read_bytes()[:200]):
        with pytest.raises(fdt.FormatException):
            det.detectFacesFromBytes(data)
    # the detector stays usable
    assert len(det.detectFacesFromBytes(SAMPLES[-1].read_bytes(), mode=fdt.FaceDetectionMode.fast)) >= 1

#!/bin/bash
# round 2, GPU call J: JPEG front end + whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -m gpu -q -x > gpurun_out/pytest_jpeg.log 2>&1; echo "jpeg rc=$?"; tail -15 gpurun_out/pytest_jpeg.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python - <<'PY'
import time, numpy as np, cv2, sys
sys.path.insert(0, '.')
import face_detection_tflite_b200 as fdt
d = fdt.FaceDetector.create(fdt.FaceDetectionModel.backCamera)
for name in ["landmark-ex1.jpg", "group-shot-bounding-box-ex1.jpeg"]:
    data = open("assets/samples/" + name, "rb").read()
    d.decodeImage(data)
    t0 = time.perf_counter()
    for _ in range(10): d.decodeImage(data)
    t1 = time.perf_counter()
    for _ in range(10): cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    t2 = time.perf_counter()
    print("%s: ours %.2f ms (host Huffman + device + D2H), cv2.imdecode %.2f ms" % (name, (t1 - t0) * 100, (t2 - t1) * 100))
PY

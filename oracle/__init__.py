"""ORACLE — test infrastructure only.

CPU restatement of the reference's detection hot path (hugocornellier/face_detection_tflite
v6.8.0).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this package; the product (face_detection_tflite_b200) never does.
"""

"""SASS opcode histogram of libfdt_cuda.so per kernel (cuobjdump -sass): the tensor-core (UTCHMMA = tcgen05.mma), TMEM (LDTM/STTM =
tcgen05.ld/st), TMA (UTMALDG/UTMASTG), mbarrier (SYNCS) opcodes and the main CUDA-core ones.  Usage: python tools/sass_hist.py > profiles/<tag>_sass_opcodes.txt"""
import collections, re, subprocess, sys
from pathlib import Path
so = Path(__file__).resolve().parents[1] / "face_detection_tflite_b200" / "libfdt_cuda.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
fn = None
h = collections.defaultdict(collections.Counter)
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@\S+\s+)?([A-Z][A-Z0-9_]*)", l)
    if m and fn:
        h[fn][m.group(1)] += 1
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FFMA", "HFMA2", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR"]
tot = collections.Counter()
print("# SASS opcode histogram of libfdt_cuda.so (cuobjdump -sass, sm_100a)")
print("%-44s" % "kernel" + " ".join("%7s" % k for k in keys))
for f, c in sorted(h.items()):
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"fdt::\(anonymous namespace\)::", "", d).replace("void ", "").split("(")[0][:42]
    print("%-44s" % d + " ".join("%7d" % c.get(k, 0) for k in keys))
    tot.update(c)
print("%-44s" % "TOTAL" + " ".join("%7d" % tot.get(k, 0) for k in keys))

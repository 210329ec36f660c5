"""Prints registers / spills per kernel from `python face_detection_tflite_b200/build.py -v` output (stdin)."""
import re, sys, subprocess
txt = sys.stdin.read()
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'fdt::\(anonymous namespace\)::', '', cur).split('(')[0].replace('void ', '')
        spill = None
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores', line)
    if m: spill = m.group(2)
    m = re.search(r'Used (\d+) registers', line)
    if m and cur:
        print("%-40s regs %3s spill %s" % (cur, m.group(1), spill))

"""Histogram of warp-stall samples over the SASS of one kernel (ncu --page source --csv export)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 6 and r[2].strip().isdigit()]
def I(r, c):
    try: return int(r[c])
    except ValueError: return 0
tot = sum(I(r, 2) for r in data)
print("total samples", tot, "instrs", len(data))
def op(r):
    t = r[1].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for b in range(0, len(data), B):
    seg = data[b:b + B]
    s = sum(I(r, 2) for r in seg)
    ops = collections.Counter(op(r) for r in seg)
    ex = sum(I(r, 5) for r in seg)
    print("%5d-%5d samples %6d (%4.1f%%) exec %9d  ops %s" % (b, b + B, s, 100 * s / max(tot, 1), ex, ops.most_common(6)))
print()
top = sorted(range(len(data)), key=lambda i: -I(data[i], 2))[:30]
ex_i = idx.get('L1 Wavefronts Shared Excessive')
for i in sorted(top):
    r = data[i]
    print(i, r[2], r[5], r[1].strip()[:72], "| xs-wavefronts", r[ex_i] if ex_i else '')

"""Per-role stall summary of a k_block_ws SASS export (ncu --page source --csv): splits the listing at the
role boundaries (found from marker instructions) and prints samples / executed instructions / top stalls."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
name = rows[0][1] if rows and len(rows[0]) > 1 else ''
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 6 and r[2].strip().isdigit()]
# the listing is duplicated in the export: keep the first copy
first = data[0][1]
for k in range(1, len(data)):
    if data[k][1] == first and data[k][0] == data[0][0]: data = data[:k]; break
def I(r, c):
    try: return int(r[c])
    except ValueError: return 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(name[:90], 'instrs', len(data), 'samples', sum(I(r, 2) for r in data))
# role boundaries: epilogue = contains LDTM; dw = FFMA-heavy; producer = UTMALDG; mma = UTCHMMA
marks = {'LDTM': [], 'FFMA': [], 'UTMALDG': [], 'UTCHMMA': [], 'UTCQMMA': []}
for i, r in enumerate(data):
    for m in marks:
        if m in r[1]: marks[m].append(i)
def rng(m): return (min(marks[m]), max(marks[m])) if marks[m] else None
for m in marks: print(' ', m, rng(m), len(marks[m]))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if W:
    for b in range(0, len(data), W):
        seg = data[b:b + W]
        c = collections.Counter()
        for r in seg:
            for s in stall_cols: c[s[6:]] += I(r, idx[s])
        tot = sum(I(r, 2) for r in seg)
        if tot: print("%5d-%5d samples %5d exec %9d %s" % (b, b + W, tot, sum(I(r, 5) for r in seg), c.most_common(5)))
top = sorted(range(len(data)), key=lambda i: -I(data[i], 2))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for i in sorted(top):
    r = data[i]
    st = sorted(((I(r, idx[c]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(i, I(r, 2), I(r, 5), r[1].strip()[:64], st)

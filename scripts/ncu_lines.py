#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.
usage: ncu_lines.py <prof_source.csv[.gz]> <kernel-substring> [top_n] [instance]"""
import csv, gzip, io, sys, collections, os
p, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
f = io.TextIOWrapper(gzip.open(p), newline="") if p.endswith(".gz") else open(p, newline="")
rows = list(csv.reader(f))

def num(x):
    try: return int(float(x))
    except ValueError: return 0

# blocks: ("File Path", path) ("Function Name", fn) header body...
blocks = []
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path" and i + 2 < len(rows) and rows[i + 1][0] == "Function Name":
        path, fn, hdr = r[1], rows[i + 1][1], rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "File Name")):
            body.append(rows[j]); j += 1
        blocks.append((path, fn, hdr, body))
        i = j
    else:
        i += 1
# instances of a function: a new instance starts when the same (fn, path) repeats
seen, groups = {}, collections.OrderedDict()
for path, fn, hdr, body in blocks:
    if want not in fn: continue
    k = seen.get((fn, path), 0); seen[(fn, path)] = k + 1
    groups.setdefault((fn, k), []).append((path, hdr, body))
keys = list(groups)
if not keys:
    print("no function matches", want); sys.exit(0)
fn, k = keys[min(inst, len(keys) - 1)]
agg = collections.OrderedDict()
for path, hdr, body in groups[(fn, k)]:
    ci = {h: n for n, h in enumerate(hdr)}
    src_cols = [n for n, h in enumerate(hdr) if h == "Source"]
    cur = None
    for b in body:
        if len(b) < len(hdr): continue
        line, text = b[0], b[src_cols[0]]
        if b[ci["Address"]] == "-":
            cur = (os.path.basename(path), line, text.strip()); agg.setdefault(cur, [0, 0, 0]); continue
        a = agg.setdefault(cur, [0, 0, 0])
        a[0] += num(b[ci["Instructions Executed"]]); a[1] += num(b[ci["Warp Stall Sampling (All Samples)"]]); a[2] += 1
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("== %s  [instance %d of %d]\n   total warp-instructions %d, stall samples %d" % (fn[:100], k, len(keys), ti, ts))
for key, a in sorted(((k2, a2) for k2, a2 in agg.items() if k2), key=lambda kv: -kv[1][0])[:top]:
    path, line, text = key
    print("   %-14s L%-5s inst %5.1f%%  samples %5.1f%%  sass %3d | %s" % (path[:14], line, 100 * a[0] / ti, 100 * a[1] / ts, a[2], text[:100]))

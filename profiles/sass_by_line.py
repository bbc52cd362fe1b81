"""Join an `ncu --page source --csv` SASS table of one kernel with `nvdisasm -g` line markers of the same cubin and print
the source lines that executed the most warp instructions / collected the most stall samples.
    python profiles/sass_by_line.py <ncu_sass.csv> <nvdisasm.txt> <mangled-function-substring> [top]"""
import csv, re, sys
from collections import defaultdict
src_csv, sass_txt, func = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
inst = [(float(r[ii] or 0), float(r[si] or 0), r[1]) for r in rows[2:] if len(r) == len(hdr)]
lines, cur, on = [], None, False
for ln in open(sass_txt):
    if ln.startswith("//--------------------- .text."):
        on = func in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
print("sass instructions: ncu %d, nvdisasm %d" % (len(inst), len(lines)))
n = min(len(inst), len(lines))
agg = defaultdict(lambda: [0.0, 0.0])
for (e, s, _), l in zip(inst[:n], lines[:n]):
    agg[l][0] += e; agg[l][1] += s
te, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
srcs = {}
def text(l):
    if l is None: return ""
    f, n_ = l
    if f not in srcs:
        import glob
        c = glob.glob("/root/repo/open-vocabulary-3d-object-detection_b200/csrc/" + f)
        srcs[f] = open(c[0]).read().split("\n") if c else []
    return srcs[f][n_ - 1].strip()[:100] if 0 < n_ <= len(srcs[f]) else ""
print("total warp-inst %.3g, samples %d" % (te, ts))
for l, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * e / te, 100 * s / max(ts, 1), l[0] if l else "?", l[1] if l else "?", text(l)))

"""Aggregate ncu warp-stall samples per CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <object-or-cubin> <kernel-substring> [top]
Uses `ncu --page source --csv` (SASS rows in address order) and `nvdisasm -g` line markers of the same cubin."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
else:
    cubins = [obj]
line_of = {}
src_file = None
for cb in cubins:
    out = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    on, cur = False, None
    for ln in out.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kname in ln
            cur = None
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, bi = h.index("# Samples"), h.index("stall_barrier")
ii = h.index("Instructions Executed")
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
base = None
agg = {}
for r in rows[hi + 1:]:
    try:
        addr = int(r[0], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    a = agg.setdefault(key, {"all": 0, "inst": 0, **{c: 0 for c in cols}})
    a["all"] += int(r[si] or 0)
    a["inst"] += int(r[ii] or 0)
    for c in cols:
        a[c] += int(r[h.index(c)] or 0)
tot = sum(a["all"] for a in agg.values())
totnb = sum(a["all"] - a["stall_barrier"] for a in agg.values())
toti = sum(a["inst"] for a in agg.values())
print(f"samples: {tot} total, {totnb} outside barriers; warp instructions {toti}")
src = {}
def text(f, l):
    if f not in src:
        try:
            src[f] = open(f).read().splitlines()
        except OSError:
            src[f] = []
    return src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
for key, a in sorted(agg.items(), key=lambda kv: -(kv[1]["all"]))[:top]:
    dom = sorted(((a[c], c) for c in cols), reverse=True)[:3]
    doms = " ".join(f"{c[6:]}={v}" for v, c in dom if v)
    print(f"{a['all']:8d} ({100 * a['all'] / tot:5.1f}%) inst={a['inst']:9d} L{key[1]:<4d} {doms:45s} | {text(*key)}")

"""SASS instruction count per CUDA source line of one kernel (code-size hot spots).
usage: python tools/sass_lines.py <object> <kernel-substring> [top]"""
import os, re, subprocess, sys, tempfile, collections
obj, kname = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cnt = collections.Counter()
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    on, cur = False, None
    for ln in out.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kname in ln
            cur = None
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) and cur:
            cnt[cur] += 1
print("total", sum(cnt.values()))
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print(dict(byfile))
for (f, l), c in cnt.most_common(top):
    print(f"{c:6d} {f}:{l}")

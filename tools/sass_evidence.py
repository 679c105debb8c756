"""Static SASS counts per kernel of the built objects (profiles/r2_sass_evidence.md sections).
usage: python tools/sass_evidence.py collide planner   -> markdown on stdout"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"collide": ["DFMA", "DMUL", "DADD", "FFMA", "FADD", "LDG", "LDS", "MUFU", "ATOMG", "VOTE"],
        "planner": ["DFMA", "DADD", "ATOMG", "ATOMS", "SHFL", "VOTE", "BAR", "CALL", "NANOSLEEP"]}
for name in sys.argv[1:]:
    obj = os.path.join(ROOT, "hybrid_path_planning_b200", "csrc", "build", name + ".o")
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    print(f"## csrc/{name}.cu")
    fn, cnt = None, collections.Counter()

    def flush():
        if fn and cnt["_n"]:
            extra = " ".join(f"{m}={cnt[m]}" for m in WANT.get(name, []) if cnt[m])
            print(f"- `{fn}`: {cnt['_n']} SASS instructions; {extra}")
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            fn, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            cnt["_n"] += 1
            cnt[m.group(1)] += 1
    flush()
    print()

"""Static SASS opcode mix of one kernel of an object file (a proxy for the dynamic count: the feature kernel's hot
loops are fully unrolled).  usage: python scripts/sass_mix.py <obj> <kernel-substring> [top]"""
import collections
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, mix = None, collections.Counter()
for line in out.splitlines():
    s = line.strip()
    if s.startswith("Function :"):
        cur = s.split(":", 1)[1].strip()
        continue
    if cur is None or pat not in cur or not s.startswith("/*") or ";" not in s:
        continue
    body = s.split("*/", 1)[1].strip()
    toks = body.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.rstrip(";").split(".")[0]] += 1
tot = sum(mix.values())
print(f"{pat}: {tot} instructions ({tot * 16 / 1024:.1f} KB)")
for op, n in mix.most_common(top):
    print(f"  {op:10s} {n:7d}  {100 * n / tot:5.1f} %")

"""Static SASS instruction count of the feature kernel per PHASE of msa_features_body.cuh (nvdisasm -gi line
info, innermost features_cta line of every instruction).  The hot loops are unrolled, so the static count of a
loop body is its dynamic count per iteration (the rolled 2 x dft32 loop counts once).
usage: python scripts/sass_by_phase.py <obj> <kernel substring> edge1 edge2 ..."""
import collections, re, subprocess, sys, tempfile, os
obj, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
edges = [int(a) for a in sys.argv[3:]]
tmp = tempfile.mkdtemp()
cub = os.path.join(tmp, "k.cubin")
subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL)
cubs = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
out = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubs[0])], capture_output=True, text=True).stdout
inside, chain, last = False, [], 0
agg = collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    if ln.startswith(".text."):
        inside = kern in ln; chain = []; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        chain.append((m.group(1), int(m.group(2)))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        if chain:
            body = [l for f, l in chain if f.endswith("msa_features_body.cuh") and l >= 195]
            last = body[0] if body else last
        chain = []
        toks = m.group(2).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        if op in ("IMAD",) and ".MOV" in m.group(2): op = "IMAD.MOV"
        key = max([e for e in edges if e <= last], default=0)
        agg[key][op] += 1
for k in sorted(agg):
    c = agg[k]; tot = sum(c.values())
    print(f"line >= {k:4d}: {tot:6d} instr  " + "  ".join(f"{o} {n}" for o, n in c.most_common(12)))

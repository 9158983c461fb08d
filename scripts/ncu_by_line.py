"""Attribute the per-instruction counters of an `ncu --page source --csv` dump to the lines of
msa_features_body.cuh, using `nvdisasm -gi` of the same cubin (inline chains), so that the hot
PHASES of the fused feature kernel can be read off without a GPU.
usage: ncu_by_line.py <source.csv> <nvdisasm -gi output> <mangled kernel substring> [bucket edges ...]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1:4]
edges = [int(a) for a in sys.argv[4:]]
# --- nvdisasm: instruction offset -> body line
MIN_LINE = 195
chain, lines, inside, last = [], {}, False, (0, ('', 0))
for ln in open(sass):
    if ln.startswith('.text.'):
        inside = kern in ln
        chain = []
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        off = int(m.group(1), 16)
        if chain:                                   # markers are only printed when the line changes
            body = [l for f, l in chain if f.endswith('msa_features_body.cuh') and l >= MIN_LINE]
            last = (body[0] if body else last[0], chain[0])   # innermost line inside features_cta
        lines[off] = last
        chain = []
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix['Address']], 16) if data[0][ix['Address']].startswith('0x') else int(data[0][ix['Address']])
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
cols = ['Instructions Executed', '# Samples', 'stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_no_inst', 'stall_wait', 'stall_not_selected', 'stall_math', 'stall_lg']
for r in data:
    a = r[ix['Address']]
    off = (int(a, 16) if a.startswith('0x') else int(a)) - base
    bl = lines.get(off, (0, ('', 0)))[0]
    key = bl
    if edges:
        key = max([e for e in edges if e <= bl], default=0)
    for c in cols:
        v = r[ix[c]] if c in ix else '0'
        v = int(float(v)) if v not in ('', 'N/A') else 0
        agg[key][c] += v
        tot[c] += v
    agg[key]['sass'] += 1
print('total', dict(tot), 'sass instr', len(data))
print('%6s %6s %12s %6s %8s | %s' % ('line', 'sass', 'inst_exec', 'pct', 'samples', ' '.join(c.replace('stall_', '')[:8] for c in cols[2:])))
for k in sorted(agg):
    a = agg[k]
    print('%6d %6d %12d %5.1f%% %8d | %s' % (k, a['sass'], a['Instructions Executed'], 100.0 * a['Instructions Executed'] / max(1, tot['Instructions Executed']),
                                           a['# Samples'], ' '.join('%8d' % a[c] for c in cols[2:])))

"""Stall-reason totals, top instructions and the instruction mix of one kernel from an `ncu --page source --csv` export."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
h = rows[hi[0]]; col = {c: i for i, c in enumerate(h)}
body = rows[hi[0] + 1: (hi[1] - 1 if len(hi) > 1 else len(rows))]
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tot = collections.Counter(); samples = inst = 0
for r in body:
    try:
        samples += int(r[col['# Samples']] or 0); inst += int(r[col['Instructions Executed']] or 0)
        for c in stalls: tot[c] += int(r[col[c]] or 0)
    except (ValueError, IndexError): pass
print('samples', samples, 'warp-instructions', inst)
for c, v in tot.most_common(8): print(f'  {c:26s} {100 * v / max(samples, 1):5.1f}%')
def n(r, k):
    try: return int(r[col[k]] or 0)
    except (ValueError, IndexError): return 0
for r in sorted(body, key=lambda r: -n(r, '# Samples'))[:12]:
    print(f"  {n(r, '# Samples'):5d}  {r[col['Source']][:100]}")
mix = collections.Counter()
for r in body:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[col['Source']])
    if m: mix[m.group(2)] += n(r, 'Instructions Executed')
print('  mix:', ', '.join(f'{k} {100 * v / max(inst, 1):.0f}%' for k, v in mix.most_common(12)))

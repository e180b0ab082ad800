"""Opcode / stall histogram of one kernel from `ncu --page source --csv` output (SASS view)."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
ins = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = totS = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[idx['Instructions Executed']].isdigit():
        continue
    s = r[idx['Source']].strip()
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
    n = int(r[idx['Instructions Executed']]); ins[op] += n; tot += n
    sm = int(r[idx['# Samples']]); samp[op] += sm; totS += sm
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h:
            stalls[h] += int(r[idx[h]])
print('total warp inst', tot, 'samples', totS)
for op, n in ins.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f'{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(totS,1):5.1f}%')
print([(k, round(100 * v / max(totS, 1), 1)) for k, v in stalls.most_common(10)])

"""Opcode mix and stall samples of one kernel from `ncu --page source --csv` output (tools for reading profiles)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stall = collections.Counter(); total = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r == hdr: continue
    try: n = int(r[idx['Instructions Executed']] or 0)
    except ValueError: continue
    src = r[idx['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = (m.group(2) if m else src)
    key = op.split('.')[0] if len(sys.argv) < 3 else op
    ops[key] += n; total += n
    stall[key] += int(r[idx['Warp Stall Sampling (All Samples)']] or 0)
print("total warp instr", total)
for k, v in ops.most_common(30): print("%-22s %12d %5.1f%%   stall samples %d" % (k, v, 100 * v / total, stall[k]))

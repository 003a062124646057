"""Opcode-level histogram of an `ncu --page source --csv` export (one kernel): stall samples and executed instructions per opcode,
stall-reason totals, and the hottest SASS lines."""
import csv, collections, sys
allrows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which kernel section of the export
rows = allrows[starts[which]:starts[which + 1]]
print('sections:', [allrows[i][1][:60] for i in starts[:-1]])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('kernel', rows[0][1][:100]); print('total samples', tot, 'instr rows', len(data))
op_s = collections.Counter(); op_e = collections.Counter()
def opc(src):
    t = src.split()
    op = t[1] if t[0].startswith('@') else t[0]
    return '.'.join(op.split('.')[:2])
for r in data:
    op = opc(r[ix['Source']])
    op_s[op] += int(r[ix['# Samples']]); op_e[op] += int(r[ix['Instructions Executed']])
te = sum(op_e.values())
print('executed total', te)
for op, c in op_s.most_common(24):
    print(f"{op:14s} samples {c:7d} {100*c/tot:5.1f}%   exec {op_e[op]:10d} {100*op_e[op]/te:5.1f}%")
print('-- rest by exec')
for op, c in op_e.most_common(30):
    if op not in dict(op_s.most_common(24)): print(f"{op:14s} exec {c:10d} {100*c/te:5.1f}%")
for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h:
        s = sum(int(r[ix[h]] or 0) for r in data)
        if s > 0.02 * tot: print(h, s, f"{100*s/tot:.1f}%")
print('-- hottest lines')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(r[ix['# Samples']].rjust(6), r[ix['Source']].strip()[:110])

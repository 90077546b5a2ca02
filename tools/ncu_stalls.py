"""Per-instruction stall summary from an .ncu-rep source page: python tools/ncu_stalls.py rep kernel_regex [min_pct]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several kernels may follow each other: take the first block
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): break
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls: agg[s] += int(r[ix[s]])
print(rows[0][1][:120] if rows[0] else '')
print('total samples', tot, 'instructions', len(data), 'executed warp-instr', sum(int(r[ix['Instructions Executed']]) for r in data))
print({k[6:]: round(v / tot * 100, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# opcode classes
cls = {}
for r in data:
    op = r[ix['Source']].strip().split()
    op = [o for o in op if not o.startswith('@')][0].split('.')[0] if op else '?'
    c = cls.setdefault(op, [0, 0]); c[0] += int(r[ix['Instructions Executed']]); c[1] += int(r[ix['# Samples']])
print('opcode: executed / samples%')
for op, (n, s) in sorted(cls.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f'  {op:10s} {n:9d} {s / tot * 100:6.2f}')
for n, r in enumerate(data):
    smp = int(r[ix['# Samples']])
    if smp > tot * minpct / 100:
        top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(n, r[ix['Source']].strip()[:64].ljust(64), smp, round(smp / tot * 100, 2), top)

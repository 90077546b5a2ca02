"""Summarise an ncu launch-list CSV (--metrics gpu__time_duration.sum) into per-kernel totals and shares."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID": hdr = r; break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"void |adsp::", "", name)
    v = float(r[vi].replace(",", ""))
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
unit = rows[0][hdr.index("Metric Unit")]
total = sum(v[1] for v in tot.values())
print(f"# {len(rows)} launches, total {total:.1f} {unit}")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/total*100:6.2f}%  {v:12.1f} {unit}  launches={n:5d}  avg={v/n:10.2f}  {k}")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel the number of launches, the mean
duration and the share of the captured kernel time.  usage: ncu_launches_summary.py <launches.csv> ["header line"]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
t = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    t.setdefault(name, []).append(float(r[ix["Metric Value"]]) / 1e6)
tot = sum(sum(v) for v in t.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("kernel: launches, mean ms, share of the captured kernel time")
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {k:58s} {len(v):4d}  {sum(v) / len(v):8.4f} ms  {sum(v) / tot * 100:5.1f} %")

"""Per-kernel table from an `ncu --metrics ... --csv --log-file` launch list (one row per metric):
launches per step, mean duration, share of the summed durations, instructions, issue slots, DRAM
bytes.  Usage: python profiles/launch_table.py profiles/r1j_launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
per = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"].replace("void unnamed>::", "").replace("void ", ""))
    per.setdefault((d["ID"], name), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), m in per.items():
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    for k, v in m.items():
        a[k] += v
T = "gpu__time_duration.sum"
tot = sum(a[T] for a in agg.values())
steps = max(a["n"] for n, a in agg.items() if "fwd" in n)
print("| kernel | launches/step | ncu us | share | warp instr | issue %% | DRAM rd / wr MB |   (%d steps captured)" % steps)
print("|---|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][T]):
    c = a["n"]
    print("| %s | %.1f | %.1f | %.1f %% | %.2f M | %.0f | %.1f / %.1f |" % (
        n, c / steps, a[T] / c / 1e3, 100 * a[T] / tot, a["smsp__inst_executed.sum"] / c / 1e6,
        a["smsp__issue_active.avg.pct_of_peak_sustained_active"] / c,
        a["dram__bytes_read.sum"] / c / 1e6, a["dram__bytes_write.sum"] / c / 1e6))

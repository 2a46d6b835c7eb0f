"""Per-CUDA-source-line sampling / instruction counts from an ncu report (first matching launch).

    python profiles/ncu_lines.py report.ncu-rep <kernel regex> [top N]
"""
import collections
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
out = []
agg = {}
nfun = 0
for r in rows:
    if r and r[0] == "Function Name":
        nfun += 1
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "":
        continue           # SASS rows; the CUDA line row in front of them carries the sums
    d = dict(zip(hdr[4:], r[4:]))
    try:
        n = int(d["Instructions Executed"])
        s = int(d["# Samples"])
    except (KeyError, ValueError):
        continue
    key = (r[0], r[1].strip()[:90])
    e = agg.setdefault(key, [0, 0, collections.Counter()])
    e[0] += s
    e[1] += n
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            e[2][k.replace("stall_", "")] += int(v)
out = [(e[0], e[1], k[0], k[1], e[2].most_common(3)) for k, e in agg.items()]
tot_s = sum(o[0] for o in out)
tot_n = sum(o[1] for o in out)
print("samples %d, warp instructions %d" % (tot_s, tot_n))
for s, n, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  L%-4s %-90s %s" % (100.0 * s / max(tot_s, 1), 100.0 * n / max(tot_n, 1), ln, src,
                                                  " ".join("%s=%d" % (k, v) for k, v in st)))

"""Per-source-line and per-opcode instruction counts from an ncu report.

    python profiles/ncu_source.py report.ncu-rep <kernel regex> [top N]

Reads `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on).
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
ops = collections.Counter()
stall = collections.Counter()
total = 0
first = True
for r in rows:
    if r and r[0] == "Function Name":
        if not first:
            break          # first launch only
        first = False
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["Instructions Executed"])
    except (KeyError, ValueError):
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
    op = m.group(1) if m else "?"
    op = ".".join(op.split(".")[:2])
    ops[op] += n
    total += n
    try:
        stall[op] += int(d["# Samples"])
    except (KeyError, ValueError):
        pass
print("total warp instructions: %d" % total)
for op, n in ops.most_common(top):
    print("%-28s %12d  %5.1f%%   samples %d" % (op, n, 100.0 * n / total, stall[op]))

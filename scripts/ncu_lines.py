"""Per-source-line summary of an ncu capture made with --import-source on:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [top]
Prints the share of thread instructions and of warp-state samples per file and per CUDA source line,
and the stall mix."""
import csv
import os
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, hdr = None, None
lines = []          # (file, line, source, samples, thread_instr, stall dict)
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r) if h not in ("Source",)}
        hdr_list = r
        continue
    if hdr is None or not r or r[0] == "":
        continue
    def get(name):
        try:
            return int(float(r[hdr[name]] or 0)) if name in hdr and r[hdr[name]] not in ("", "-") else 0
        except (ValueError, IndexError):
            return 0
    try:
        st = {h[6:]: get(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        lines.append((cur, int(r[0]), r[1].strip(), get("# Samples"), get("Thread Instructions Executed"), st))
    except ValueError:
        continue
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
mix = defaultdict(int)
for l in lines:
    for k, v in l[5].items():
        mix[k] += v
print("# thread instructions %.3e, warp-state samples %d" % (tot_i, tot_s))
print("# stall mix, %% of samples: " + ", ".join("%s %.1f" % (k, 100 * v / tot_s) for k, v in sorted(mix.items(), key=lambda x: -x[1])[:10]))
byfile = defaultdict(lambda: [0, 0])
for l in lines:
    byfile[l[0]][0] += l[4]; byfile[l[0]][1] += l[3]
for f, (i, s) in sorted(byfile.items(), key=lambda x: -x[1][1]):
    print("# %-28s thread-instr %5.1f%%  samples %5.1f%%" % (f, 100 * i / tot_i, 100 * s / tot_s))
print("file line thread-instr% samples% top-stalls source")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    st = ",".join("%s:%.0f" % (k, 100 * v / max(1, l[3])) for k, v in sorted(l[5].items(), key=lambda x: -x[1])[:3])
    print("%-16s %5d %6.2f %6.2f  %-40s %s" % (l[0], l[1], 100 * l[4] / tot_i, 100 * l[3] / tot_s, st, l[2][:100]))

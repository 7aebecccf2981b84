"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals of the LAST step.
usage: launch_table.py launches.csv [first_kernel_of_step_regex]"""
import csv, re, sys
rows = []
for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
    rows.append(r)
hdr, rows = rows[0], rows[1:]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
# a step starts at the bow scoring kernel; keep the last one
starts = [i for i, r in enumerate(rows) if re.search(sys.argv[2] if len(sys.argv) > 2 else "bow_score", r[ik])]
rows = rows[starts[-1]:] if starts else rows
tot, seq = {}, []
for r in rows:
    name = re.sub(r"\(.*", "", r[ik])
    t = float(r[iv].replace(",", "")) / 1e6
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1; tot[name][1] += t
    seq.append((name, r[ig], t))
all_ms = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-45s n=%3d  %8.3f ms  %5.1f%%" % (k[:45], v[0], v[1], 100 * v[1] / all_ms))
print("total %.3f ms over %d launches" % (all_ms, len(rows)))
if len(sys.argv) > 3:
    for s in seq: print("   %-40s %-18s %.3f" % (s[0][:40], s[1], s[2]))

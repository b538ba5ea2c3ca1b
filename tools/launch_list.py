"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

usage: python tools/launch_list.py gpurun_out/launches.csv profiles/r01_launch_list_bench "<command>"
(copies the csv next to the .md)
"""
import collections, csv, shutil, sys

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ik] == "Kernel Name":
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * (1e3 if r[iu] in ("ms", "msecond") else 1.0)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list: `{cmd}`", "",
         f"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (first {sum(a[0] for a in agg.values())} "
         "launches of the process; per-launch times are cold-cache and serialised: compare SHARES).", "",
         "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / total:.1f}% |")
open(out + ".md", "w").write("\n".join(lines) + "\n")
shutil.copyfile(src, out + ".csv")
print("\n".join(lines))

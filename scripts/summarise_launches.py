"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table for profiles/.
usage: python scripts/summarise_launches.py gpurun_out/launches.csv 'title' 'command' > profiles/rN_launches_xxx.md"""
import collections
import csv
import re
import sys

path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
total = n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    total += ms
    n += 1
print(f"# {title}\n")
print(f"Command (1 GPU, under gpurun):\n`{command}`\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolute values.\n")
print(f"launches {n}, total {total:.2f} ms\n")
print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.2f}% | {ms / cnt:.4f} |")

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (profiles/)."""
import collections
import csv
import sys
from pathlib import Path

src, tag = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:80]
    t = float(r[-1].replace(",", ""))
    unit = r[-2]
    ns = t * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list — {tag}", "", f"`ncu --metrics gpu__time_duration.sum --clock-control none -c 80` around `python bench.py --steps 2 --warmup 3` "
       "(first 80 launches; per-launch times are cold-cache and serialised — compare shares).", "",
       "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{name}` | {c} | {ns/1e6:.3f} | {100*ns/tot:.2f}% |")
Path(f"profiles/{tag}_launches.md").write_text("\n".join(out) + "\n")
print("\n".join(out))

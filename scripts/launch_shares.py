"""Per-kernel time shares (and DRAM bytes when captured) from an ncu --csv launch list:
    python scripts/launch_shares.py profiles/r1_c9_launches.csv > profiles/r1_c9_launch_shares.txt"""
import csv, sys
from collections import defaultdict

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in rows:
    name = r["Kernel Name"].split("(")[0][:100]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    else:
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    per[name][m] += v
    ids[name].add(r["ID"])
tot = sum(p["gpu__time_duration.sum"] for p in per.values())
n = sum(len(s) for s in ids.values())
print(f"# {path}: total {tot:.1f} ms over {n} launches (ncu: cold-cache, serialised — compare SHARES)")
for name, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])[:28]:
    t = p["gpu__time_duration.sum"]
    extra = ""
    if "dram__bytes_read.sum" in p:
        extra = f"  dram rd {p['dram__bytes_read.sum']:9.0f} MB  wr {p['dram__bytes_write.sum']:9.0f} MB"
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={len(ids[name]):4d}{extra}  {name}")

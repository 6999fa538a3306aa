"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid).
    python tools/launch_summary.py gpurun_out/launches.csv [--grid]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
by_grid = "--grid" in sys.argv
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, gs, bs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("uc::<unnamed>::", "").replace("void ", "")[:60]
    key = (name, r[gs], r[bs]) if by_grid else name
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / 1e6:8.2f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d} avg {v[1] / v[0] / 1e3:7.1f} us  {k}")

"""Per-kernel DRAM traffic of one step from an ncu CSV (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
-> the JSON bench.py reads for `roofline.traffic` (newest profiles/r0*_dram_traffic_per_kernel.json).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py
    python tools/traffic_summary.py gpurun_out/traffic.csv profiles/r02n_dram_traffic_per_kernel.json
"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("uc::<unnamed>::", "").replace("void ", "")
    name = re.sub(r"gemm2_kernel<\(int\)(-?\d+), \(bool\)(\d), \(int\)(\d+), \(int\)(\d+)>", r"gemm2_kernel<\1, \2, \3, \4>", name)[:60]
    d = launch.setdefault(int(r[0]), {"k": name})
    d[r[mn]] = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in launch.values():
    a = agg[d["k"]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
gemm = [(k, v) for k, v in agg.items() if k.startswith("gemm")]
n_gemm = sum(v[0] for _, v in gemm)
out = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one fwd+bwd step (tools/profile_step.py), {src}",
       "gemm_launches": n_gemm,
       "gemm_dram_bytes_per_launch": sum(v[2] + v[3] for _, v in gemm) / max(n_gemm, 1),
       "kernels": {k: {"launches": v[0], "time_ms": v[1] / 1e6, "dram_read_bytes": v[2], "dram_write_bytes": v[3]}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
json.dump(out, open(dst, "w"), indent=1)
print(f"{n_gemm} GEMM launches, {out['gemm_dram_bytes_per_launch'] / 1e6:.1f} MB DRAM traffic per launch; total step "
      f"{sum(v[1] for v in agg.values()) / 1e6:.2f} ms, {sum(v[2] + v[3] for v in agg.values()) / 1e9:.1f} GB")

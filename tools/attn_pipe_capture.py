"""Turn an ncu CSV of tools/attn_once.py into the committed tensor-pipe summary that bench.py reports (`attn_tensor_pipe`).

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none -k regex:attn_ --csv \
        --log-file gpurun_out/attn_pipe.csv python tools/attn_once.py
    python tools/attn_pipe_capture.py gpurun_out/attn_pipe.csv profiles/r02h_attn_tensor_pipe.json

The JSON carries the sha256 of the attention sources it was captured from (bench.attn_source_hash): bench.py marks the capture
`current_build: false` as soon as the kernels change.  Per kernel the LAST launch of each shape is kept (warm caches)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, gs, mn, mv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Name"), hdr.index("Metric Value")
launch = {}
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.search(r"(attn_\w+)", r[kn])
    if not name:
        continue
    launch.setdefault(int(r[0]), {"kernel": name.group(1), "grid": r[gs]})[r[mn]] = float(r[mv].replace(",", ""))
# attn_once.py runs the encoder shape (16x16x1024) first, then the decoder shape (8x12x1024), two passes each
ids = sorted(launch)
per_kernel = {}
for i in ids:
    per_kernel.setdefault(launch[i]["kernel"], []).append(launch[i])
out = {"metric": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "how": "ncu --metrics ... --clock-control none -k regex:attn_ python tools/attn_once.py; per kernel and shape the second (warm) launch",
       "source": os.path.relpath(src, ROOT), "attn_source_hash": bench.attn_source_hash(),
       "encoder_16x16x1024": {}, "decoder_8x12x1024": {}}
for k, ls in per_kernel.items():
    half = len(ls) // 2
    for shape, l in (("encoder_16x16x1024", ls[half - 1]), ("decoder_8x12x1024", ls[-1])):
        out[shape][k] = {"tensor_pipe_pct": round(l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan")), 2),
                         "xu_pipe_pct": round(l.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", float("nan")), 2),
                         "us": round(l.get("gpu__time_duration.sum", float("nan")) / 1e3, 1), "grid": l["grid"]}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))

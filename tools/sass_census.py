"""Per-kernel census of the Blackwell-native SASS mnemonics in libuc_b200.so (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce,
UBLKCP = cp.async.bulk, plus MUFU.EX2 and the packed f32x2 ops of the attention kernels.  No GPU needed.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "uniception_b200", "libuc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "MUFU.EX2", "FFMA2", "FADD2", "FMUL2", "HMMA"]
print("libuc_b200.so SASS census (cuobjdump -sass; sm_100a); columns: " + " ".join(KEYS) + " | total instructions")
tot = collections.Counter()
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n")[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part)
    c = collections.Counter()
    for o in ops:
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                if o.startswith("UTCHMMA") and ".2CTA" in o:
                    c[k] += 1
            elif k == "UTCHMMA":
                if o.startswith("UTC") and "MMA" in o:
                    c[k] += 1
            elif o == k or o.startswith(k + "."):
                c[k] += 1
    tot.update(c)
    d = demangle(name).replace("(anonymous namespace)::", "").replace("void ", "")
    short = d.split("(")[0]
    print(f"{short[:58]:58s} " + " ".join(f"{c[k]:5d}" for k in KEYS) + f" | {len(ops)}")
print(f"{'TOTAL':58s} " + " ".join(f"{tot[k]:5d}" for k in KEYS))

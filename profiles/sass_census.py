#!/usr/bin/env python
"""Per-kernel SASS census of libeventful_b200.so: which kernels use tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMALDG /
UTMASTG / UBLKCP), cp.async (LDGSTS), legacy tensor cores (HMMA) or only CUDA cores (FFMA).  Runs without a GPU:
    python profiles/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eventful-transformer_b200", "lib", "libeventful_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "FFMA", "MUFU", "F2FP", "REDUX", "STSM", "LDSM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, current, it = {}, [], None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = next(it)
            current = re.sub(r"\(anonymous namespace\)::", "", current)
            current = re.sub(r"\(.*$", "", current)
            current = current.replace("__nv_bfloat16", "bf16").replace("__half", "f16").replace("void ", "")
            if current not in counts:
                counts[current] = collections.Counter()
                order.append(current)
            continue
        if current is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            for key in MNEMONICS:
                if op.startswith(key):
                    counts[current][key] += 1
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); counts of instructions per kernel")
    print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, cp.async -> LDGSTS, mma.sync -> HMMA")
    used = [k for k in MNEMONICS if any(c[k] for c in counts.values())]
    print("kernel".ljust(78) + "".join(k.rjust(9) for k in used))
    totals = collections.Counter()
    for name in sorted(order):
        c = counts[name]
        totals.update(c)
        print(name[:77].ljust(78) + "".join(str(c[k] or ".").rjust(9) for k in used))
    print("TOTAL".ljust(78) + "".join(str(totals[k]).rjust(9) for k in used))
    tc = sorted(n for n in order if counts[n]["UTCHMMA"] or counts[n]["UTCQMMA"])
    print(f"\n# {len(tc)} kernels issue tcgen05.mma; {sum(1 for n in order if counts[n]['HMMA'])} use mma.sync (general-shape fallbacks and rel-pos tables); "
          f"{sum(1 for n in order if not (counts[n]['UTCHMMA'] or counts[n]['HMMA']))} run on the CUDA cores (gates, scatter, fp32 path)")


if __name__ == "__main__":
    sys.exit(main())

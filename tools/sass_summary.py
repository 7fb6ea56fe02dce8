#!/usr/bin/env python
"""Per-kernel SASS evidence for libnndepth_b200.so: counts of the mnemonics that prove which hardware path a kernel uses
(tcgen05 tensor cores = UTCHMMA / UTCQMMA ..., TMEM loads = LDTM, TMA loads / stores = UTMALDG / UTMASTG, bulk copies =
UBLKCP, mbarrier waits = SYNCS, legacy tensor cores = HMMA / IMMA, async copies = LDGSTS), registers and spills.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nndepth_b200", "libnndepth_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "FFMA2", "FFMA",
        "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("# cuobjdump -sass nndepth_b200/libnndepth_b200.so (sm_100a), instruction counts per kernel")
    print("# columns: " + " ".join(KEYS) + " | registers, static shared bytes, local (spill) bytes")
    for (mangled, c), nice in zip(counts.items(), demangle):
        nice = re.sub(r"\(.*", "", nice.replace("void ", "").replace("nnd::", ""))
        reg = usage.get(mangled)
        cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{nice:70s} {cols}" + (f" | REG={reg[0]} SMEM={reg[1]} LOCAL={reg[2]}" if reg else ""))


if __name__ == "__main__":
    sys.exit(main())

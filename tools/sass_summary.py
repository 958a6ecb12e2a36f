#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native build (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG (TMA loads, .MULTICAST / .2CTA forms),
SYNCS (mbarrier), the peer-memory evidence of the fused exchange (STG/LDG .STRONG.SYS on mapped peer
pointers + MEMBAR.*.SYS in the scan kernel and the exchange kernel) and programmatic dependent launch
(ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents).

    python tools/sass_summary.py > profiles/round2/sass_summary.txt        # no GPU needed
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "picovdb_b200", "libpicovdb_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
            "UTCBAR", "HMMA", "ACQBULK", "PREEXIT", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "STG.E.64.STRONG.SYS",
            "LDG.E.64.STRONG.SYS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                kernels[cur][p] += 1
                for suffix in (".MULTICAST", ".2CTA"):
                    if suffix in op:
                        kernels[cur][p + suffix] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# kernel | instructions | mnemonic counts")
    groups = collections.OrderedDict()
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", pretty)
        base = re.sub(r"<.*", "", short)
        groups.setdefault(base, []).append((short, c))
        total.update({k: v for k, v in c.items() if k != "_total"})
    for base, items in groups.items():
        agg = collections.Counter()
        for _, c in items:
            agg.update(c)
        marks = ", ".join(f"{k} x{v}" for k, v in sorted(agg.items()) if k != "_total")
        print(f"{base}  [{len(items)} instantiation(s)] | {agg['_total']} | {marks or '-'}")
    print("# library totals: " + ", ".join(f"{k} x{v}" for k, v in sorted(total.items())))


if __name__ == "__main__":
    sys.exit(main())

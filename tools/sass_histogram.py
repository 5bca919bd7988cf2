#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass kat_b200/lib/libkatgpu.so):
the mnemonics that show what each kernel is made of — TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS),
shared / global atomics (ATOMS / ATOMG / RED), warp match (MATCH), wide global stores (STG.*128 / .256),
wide loads.  Writes profiles/r2_sass_opcodes.txt.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kat_b200", "lib", "libkatgpu.so")
KEYS = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "ATOMG.E.CAS.128", "RED", "MATCH", "STG.E.ENL2.256", "STG.E.128", "STG.E.64", "LDG.E.128", "LDG.E.ENL2.256",
        "LDS.128", "STS.64", "BAR.SYNC", "SHFL", "VOTE", "FLO", "POPC", "IMAD", "SHF", "LOP3", "S2R", "CALL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, name, counts, total = [], None, None, 0
    per = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("katgpu::", "")
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            per[name]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                    per[name][k] += 1
    lines = ["# SASS opcode histogram per kernel of kat_b200/lib/libkatgpu.so (sm_100a), tools/sass_histogram.py",
             "# columns: static instruction counts; UBLKCP = TMA bulk copy, SYNCS = mbarrier ops, ATOMS = shared atomics,",
             "# ATOMG/RED = global atomics (ATOMG.E.CAS.128 = the 16-byte compare-and-swap that claims a two-word-key slot), MATCH = match.any warp aggregation,",
             "# STG.E.ENL2.256 = 32-byte sector stores", ""]
    hdr = "%-62s %6s " % ("kernel", "instr") + " ".join("%7s" % k.replace("ATOMG.E.CAS.128", "CAS.128").replace("STG.E.", "STG.").replace("LDG.E.", "LDG.")[:7] for k in KEYS)
    lines.append(hdr)
    for name, c in per.items():
        lines.append("%-62s %6d " % (name[:62], c["_total"]) + " ".join("%7d" % c[k] for k in KEYS))
    path = os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
    print("->", path)


if __name__ == "__main__":
    sys.exit(main())

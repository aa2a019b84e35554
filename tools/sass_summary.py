#!/usr/bin/env python
"""Opcode evidence per kernel from the built library (no GPU needed): for every sm_100a kernel in libmultigrid_b200.so the number
of SASS instructions and the counts of the mnemonics that show Blackwell-native data movement and warp-level primitives -
UBLKCP (cp.async.bulk = TMA bulk copies, .G.S loads / .S.G stores), SYNCS (mbarrier), ACQBULK / fences of the async proxy,
MATCH (__match_any_sync), VOTE (ballots), REDUX, SHFL, RED / ATOM (fire-and-forget counters), BAR (CTA barriers: the warp-tile
kernel has none), plus registers from the ELF.       python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym-multigrid_b200", "libmultigrid_b200.so")
KEYS = ["UBLKCP", "SYNCS", "ACQBULK", "FENCE", "MATCH", "VOTE", "REDUX", "SHFL", "RED", "ATOM", "BAR", "LDS", "STS", "LDG", "STG", "PRMT", "IMAD", "LDTM", "UTC"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
        regs[m.group(1)] = int(m.group(2))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# {os.path.relpath(LIB, ROOT)}: cubin targets {arch}")
    print(f"# {'kernel':70s} {'instr':>6s} {'regs':>4s}  " + " ".join(f"{k:>7s}" for k in KEYS))
    cur, counts = None, None
    out = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                out.append((cur, counts))
            cur, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts["_n"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
            if op.startswith("UBLKCP"):
                counts["dir:" + (".".join(op.split(".")[1:3]))] += 1
    if cur:
        out.append((cur, counts))
    dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in out), capture_output=True, text=True).stdout.splitlines()
    for (name, c), d in sorted(zip(out, dem), key=lambda t: -t[0][1]["_n"]):
        short = re.sub(r"\(.*", "", d).replace("void mg::", "")[:70]
        dirs = ", ".join(f"{k[4:]} x{v}" for k, v in sorted(c.items()) if k.startswith("dir:"))
        print(f"  {short:70s} {c['_n']:6d} {regs.get(name, 0):4d}  " + " ".join(f"{c[k]:7d}" for k in KEYS) + (f"   [{dirs}]" if dirs else ""))


if __name__ == "__main__":
    main()

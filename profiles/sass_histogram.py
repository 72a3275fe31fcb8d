"""SASS instruction histogram per kernel of libmudg_sm100.so (evidence that the hot kernels are Blackwell-native:
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA; HMMA = legacy mma.sync).

    python profiles/sass_histogram.py > profiles/r2_sass_histogram.md      (CPU box; needs cuobjdump)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mudg_b200", "libmudg_sm100.so")
PAT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU", "LDGSTS", "SYNCS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
    chunks = re.split(r"\s+Function : \S+\n", out)[1:]
    print("# SASS instruction histogram of `mudg_b200/libmudg_sm100.so` (cuobjdump -sass, sm_100a)\n")
    print("`UTCHMMA` = `tcgen05.mma` (`.2CTA` = `cta_group::2`), `LDTM`/`STTM` = `tcgen05.ld`/`st`, `UTMALDG`/`UTMASTG` = TMA tensor load/store, "
          "`UTMAPF` = tensor-map prefetch, `SYNCS` = mbarrier ops, `HMMA` = `mma.sync` (legacy tensor path), `MUFU` = special-function unit.\n")
    print("| kernel | instr | " + " | ".join(PAT) + " | UTCHMMA.2CTA |")
    print("|---|---|" + "---|" * (len(PAT) + 1))
    rows = []
    for name, body in zip(names, chunks):
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
        c = collections.Counter()
        for op in ops:
            for p in PAT:
                if op.startswith(p):
                    c[p] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                c["2cta"] += 1
        short = re.sub(r"mudg::\(anonymous namespace\)::|mudg::|void ", "", name)
        short = re.sub(r"\(.*", "", short)
        rows.append((short, len(ops), c))
    for short, n, c in sorted(rows, key=lambda r: (-r[2]["UTCHMMA"], -r[2]["HMMA"], r[0])):
        print(f"| `{short}` | {n} | " + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + f" | {c['2cta'] or ''} |")


if __name__ == "__main__":
    main()

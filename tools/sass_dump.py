#!/usr/bin/env python
"""SASS listings of the hot kernels from the in-tree library (cuobjdump -sass), one file per kernel under profiles/.

    python tools/sass_dump.py r2h

Writes profiles/sass_<tag>_<kernel>.txt for the kernels DESIGN.md discusses and profiles/sass_<tag>_census.txt
(per-kernel static counts of the mnemonics that identify the mechanisms: UTMALDG / UBLKCP = TMA tensor / bulk copies,
UTMASTG / UBLKCP.S = bulk stores, SYNCS = mbarrier, LDGSTS = cp.async, IDP = dp2a, FFMA2 = packed fp32, ATOMG = claims).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "active_gym_b200", "lib", "libagym_b200.so")
HOT = {
    "k_ingest_gray_std_160_pc": r"k_ingest_gray_stdILi160ELb1ELi2E",
    "k_observe_peripheral_std_4_fs": r"k_observe_peripheral_stdILi4ELi0ELb1E",
    "k_ingest_atari_tma_rgb": r"k_ingest_atari_tmaILi480ELi84ELi3ELb1ELi2E",
    "k_observe_flexible_v3_mask": r"k_observe_flexible_v3ILi1E",
    "k_observe_fixed_crop_v3_30": r"k_observe_fixed_crop_v3ILi30E",
    "k_ingest_dmc": r"k_ingest_dmc",
}
MNEM = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "IDP", "FFMA2", "FADD2", "IMAD.HI", "PRMT", "ATOMG", "BAR.SYNC",
        "LDS", "STS", "LDG", "STG", "ACQBULK", "UTMACMDFLUSH")


def main():
    tag = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur, name = {}, None, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
        if cur is not None:
            cur.append(line)
    out_dir = os.path.join(ROOT, "profiles")
    census = []
    for fn, lines in funcs.items():
        c = collections.Counter()
        n = 0
        for ln in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            n += 1
            op = m.group(1)
            for k in MNEM:
                if op == k or op.startswith(k + "."):
                    c[k] += 1
        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"agym::\(anonymous namespace\)::", "", short).split("(")[0].replace("void ", "")
        census.append(f"{short}: {n} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(c.items()) if v))
    with open(os.path.join(out_dir, f"sass_{tag}_census.txt"), "w") as f:
        f.write("\n".join(sorted(census)) + "\n")
    for short, pat in HOT.items():
        for fn, lines in funcs.items():
            if re.search(pat, fn):
                with open(os.path.join(out_dir, f"sass_{tag}_{short}.txt"), "w") as f:
                    f.write("\n".join(lines) + "\n")
                print(short, len(lines), "lines")
                break


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that show what the kernels are built from (TMA tensor / bulk copies, mbarrier,
dp2a, saturating packs, votes, shuffles, PRMT), from `cuobjdump -sass` of the built objects.

    python tools/sass_summary.py > profiles/NAME_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ["video-coding_b200/lib/hcj_kernels.o", "video-coding_b200/lib/hcj_encode.o"]
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "IDP", "I2IP", "PRMT", "VOTE", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "IMAD", "IADD3", "SHF",
         "LOP3", "ATOMS", "ATOMG", "RED", "BAR", "MATCH", "HMMA", "UTCMMA", "TCGEN"]


def main():
    print("cuobjdump -sass of the sm_100a objects: static instruction counts per kernel (mnemonic prefix match)")
    print("(no HMMA / tcgen05: nothing on this path is a contraction; UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk,")
    print(" SYNCS = mbarrier, IDP = dp2a, I2IP = cvt.pack.sat)\n")
    for obj in OBJS:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, obj)], capture_output=True, text=True, check=True).stdout
        fn = None
        counts = collections.OrderedDict()
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                counts[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and fn:
                op = m.group(1)
                counts[fn]["total"] += 1
                for w in WATCH:
                    if op.startswith(w):
                        counts[fn][w] += 1
        print("== %s" % obj)
        for fn, c in counts.items():
            print("%-44s total %5d  %s" % (fn[:44], c["total"], "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
        print()


if __name__ == "__main__":
    main()

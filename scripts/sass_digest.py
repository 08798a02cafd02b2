"""Opcode histogram per kernel of libhdn.so (cuobjdump -sass, no GPU needed): python scripts/sass_digest.py > profiles/r02_sass_digest.txt
The mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, LDGSTS = cp.async, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, REDG / RED = red.global, ATOMG = atom.global."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "LDGSTS", "UTCBAR", "SYNCS", "ELECT", "BRA.U.ANY", "REDG", "ATOMG", "ATOMS", "PREFETCH", "CCTL",
       "LDG", "STG", "LDS", "STS", "FFMA", "HMMA", "BAR", "SHFL", "R2UR"]


def main():
    lib = os.path.join(ROOT, "h-denseunet_b200", "libhdn.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kern, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            hist[kern]["_total"] += 1
            for k in KEY:
                if op == k or op.startswith(k + "."):
                    hist[kern][k] += 1
    print("SASS opcode digest of h-denseunet_b200/libhdn.so (sm_100a), instructions per kernel")
    print("%-44s %7s  %s" % ("kernel", "total", "  ".join("%s" % k for k in KEY)))
    tot = collections.Counter()
    for k in sorted(hist):
        h = hist[k]
        tot.update(h)
        print("%-44s %7d  %s" % (k[:44], h["_total"], "  ".join("%*d" % (len(x), h[x]) for x in KEY)))
    print("%-44s %7d  %s" % ("ALL", tot["_total"], "  ".join("%*d" % (len(x), tot[x]) for x in KEY)))


if __name__ == "__main__":
    main()

"""SASS opcode histogram per kernel of the in-tree libpmp_b200.so (cuobjdump -sass), written as a markdown table.

    python tools/sass_histogram.py [out.md]

The tcgen05 / TMA / TMEM mnemonics that prove the hand-written path (B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTCBAR
(tcgen05.commit), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier),
USETMAXREG (setmaxnreg), UTCATOMSWS / UTCALLOC-family (tcgen05.alloc), ACQBULK / UCGABAR (cluster barrier).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pmp_vvc_tip2023_b200", "libpmp_b200.so")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "UTCATOMSWS", "UCGABAR", "HMMA", "FFMA", "LDG", "STG",
       "LDS", "STS", "LDL", "STL", "SHFL", "REDUX", "ATOM", "RED", "BAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "USETMAXREG"):
                cur[m.group(1) + m.group(2)] += 1
    lines = ["# SASS opcode histogram of libpmp_b200.so (%s), `cuobjdump -sass`" % arch, "",
             "| kernel | instructions | " + " | ".join(KEY) + " | tcgen05/TMA variants |", "|---|---|" + "---|" * (len(KEY) + 1)]
    for name, c in kernels.items():
        total = sum(v for k, v in c.items() if "." not in k)
        var = ", ".join("%s x%d" % (k, v) for k, v in sorted(c.items()) if "." in k)
        lines.append("| `%s` | %d | " % (name, total) + " | ".join(str(sum(v for k, v in c.items() if "." not in k and k.startswith(key))) for key in KEY) +
                     " | %s |" % var)
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()

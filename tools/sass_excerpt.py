#!/usr/bin/env python
"""SASS evidence per kernel of libb200geo.so (no GPU needed): for every kernel the counts of the mnemonics that show
Blackwell-native tile movement and synchronisation (UTMALDG / UBLKCP = TMA, SYNCS = mbarrier, SHFL, LDS/STS, the FP pipes),
registers, and the first lines in which the TMA / mbarrier instructions occur. Writes markdown to stdout.
usage: tools/sass_excerpt.py [path to libb200geo.so] > profiles/rNN_sass_excerpt.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "FADD", "FMUL", "FFMA",
         "HMMA", "UTCHMMA", "BAR", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def shorten(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void |b200geo::", "", name)
    name = re.sub(r"\(int\)", "", name)
    return re.sub(r"\((?!.*>).*$", "", name)      # drop the argument list (the last '(' behind the template arguments)


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "libgeodecomp_b200", "libb200geo.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = OrderedDict()
    arch = None
    for line in sass.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {"arch": arch, "ops": Counter(), "lines": [], "n": 0}
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?\s", line)
        if m and cur:
            k = kernels[cur]
            k["n"] += 1
            op = m.group(1)
            k["ops"][op] += 1
            if op in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS") and len(k["lines"]) < 4:
                k["lines"].append(re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", line)).strip().rstrip(";"))
    names = demangle(list(kernels))
    print("# SASS evidence per kernel of libb200geo.so (`cuobjdump -sass`, `tools/sass_excerpt.py`)\n")
    print("All cubins are `%s`. UTMALDG / UBLKCP = TMA tensor / bulk copies (`cp.async.bulk[.tensor]`), SYNCS = mbarrier, no tensor-core\n"
          "mnemonics (HMMA / UTC*MMA) anywhere: stencils and the pair loop are not contractions.\n" % ", ".join(sorted(set(k["arch"] for k in kernels.values() if k["arch"]))))
    print("| kernel | regs | SASS lines | " + " | ".join(WATCH) + " |")
    print("|---|---|---|" + "---|" * len(WATCH))
    for name, k in kernels.items():
        short = shorten(names.get(name, name))
        print("| `%s` | %s | %d | %s |" % (short[:80], regs.get(name, ("?",))[0], k["n"], " | ".join(str(k["ops"].get(w, 0) or "") for w in WATCH)))
    print("\n## Where the TMA and mbarrier instructions are (first occurrences per kernel)\n")
    for name, k in kernels.items():
        if k["lines"]:
            print("`%s`\n```" % shorten(names.get(name, name))[:100])
            for ln in k["lines"]:
                print(ln)
            print("```")


if __name__ == "__main__":
    main()

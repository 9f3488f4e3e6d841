#!/usr/bin/env python
"""Static summary of the compiled kernels (no GPU needed): registers, stack (spills), shared memory and the counts of the SASS
instructions that matter for an HBM-bound kernel -- 128-bit global loads/stores, shared-memory accesses, barriers, L2 prefetches,
bulk (TMA) copies.  python tools/sass_summary.py [object files...] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = sys.argv[1:] or [os.path.join(ROOT, "p3dfft_b200", "lib", "obj_p3dfft", f) for f in ("fft_fast.o", "fft_kernels.o")]
PAT = collections.OrderedDict([
    ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.64", r"\bLDG\.E(\.\w+)*\.64"), ("LDG.other", r"\bLDG\b"),
    ("STG.128", r"\bSTG\.E(\.\w+)*\.128"), ("STG.64", r"\bSTG\.E(\.\w+)*\.64"), ("STG.other", r"\bSTG\b"),
    ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("BAR", r"\bBAR\.SYNC"), ("PREFETCH", r"\bCCTL\b|\bPREFETCH"), ("UBLKCP", r"\bUBLKCP\b"),
    ("SHFL", r"\bSHFL\b"), ("DFMA+DADD+DMUL", r"\bD(FMA|ADD|MUL)\b"), ("FFMA+FADD+FMUL", r"\bF(FMA|ADD|MUL)\b"), ("LDL/STL", r"\b(LDL|STL)\b"),
])


def demangle(names):
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    print("# " + " | ".join(["kernel", "regs", "stack", "smem(static)"] + list(PAT) + ["instructions"]))
    for obj in OBJS:
        res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
        usage = {}
        for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
            usage[m.group(1)] = m.groups()[1:]
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        funcs, cur = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = []
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?)\s*/\*", line)
            if cur and m:
                funcs[cur].append(m.group(1))
        names = demangle(list(funcs))
        print(f"## {os.path.relpath(obj, ROOT)}")
        for f, ins in sorted(funcs.items(), key=lambda kv: names[kv[0]]):
            counts, seen = [], set()
            for key, pat in PAT.items():
                hits = [i for i, t in enumerate(ins) if re.search(pat, t) and (not key.endswith("other") or i not in seen)]
                seen.update(hits)
                counts.append(str(len(hits)))
            r = usage.get(f, ("?", "?", "?"))
            short = re.sub(r"p3d::(fast::)?", "", names[f]).replace("(p3d::FastStage)", "").replace("void ", "")
            print(" | ".join([short, *r, *counts, str(len(ins))]))


if __name__ == "__main__":
    main()

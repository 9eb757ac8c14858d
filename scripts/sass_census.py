"""SASS census of the shipped library: which Blackwell instructions each kernel contains.

    python scripts/sass_census.py > profiles/r02_sass_census.md

`cuobjdump -sass` of reazonspeech_b200/librs_engine.so, per kernel: tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2), TMA tensor
loads / stores / reductions (UTMALDG / UTMASTG / UTMAREDG) and bulk copies (UBLKCP), tensor-memory loads / stores (LDTM / STTM),
MMA-completion barriers (UTCBAR), legacy tensor-core instructions (HMMA = mma.sync) and the total instruction count.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "reazonspeech_b200", "librs_engine.so")
COLS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"),
        ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"^void |rs::|\(anonymous namespace\)::|\(.*$", "", n) for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        cur["total"] += 1
        for name, pat in COLS:
            if re.search(pat, ins):
                cur[name] += 1
    names = demangle(list(kernels))
    print("# r02 -- SASS census of `reazonspeech_b200/librs_engine.so` (sm_100a)\n")
    print("`python scripts/sass_census.py` = `cuobjdump -sass` of the library this commit builds, counted per kernel.  UTCHMMA = tcgen05.mma")
    print("(`.2CTA` = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP = 1-D bulk copy, LDTM / STTM =")
    print("tcgen05.ld / st, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync.  No cuBLAS / cuDNN / CUTLASS code is linked.\n")
    print("| kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    tot = collections.Counter()
    for n, (_, c) in sorted(zip(names, kernels.items()), key=lambda t: -t[1][1]["total"]):
        print(f"| `{n}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k, _ in COLS) + " |")
        tot.update(c)
    print(f"| **all {len(kernels)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k, _ in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())

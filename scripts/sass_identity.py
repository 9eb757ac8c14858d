#!/usr/bin/env python3
"""Compare the SASS of every kernel in two object files (or .so) by demangled kernel name.

Used to show that a change which adds env-gated or compile-time-gated experiments leaves the DEFAULT kernels
bit-identical to a build that was validated on the GPU:

    python scripts/sass_identity.py old/gemm_tcgen05.o reazonspeech_b200/csrc/build/gemm_tcgen05.o

Prints SAME / DIFF / NEW / GONE per kernel; exit status 1 if any kernel DIFFers.  Addresses and encodings are
stripped, so only the instruction stream (opcodes, registers, immediates, constant-bank offsets) is compared.
Parameter lists are dropped from the names: a kernel whose parameter TYPE was renamed still matches."""
import re
import subprocess
import sys


def kernel_name(mangled: str) -> str:
    d = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)", "{anon}")
    return d.split("(")[0]


def kernels(path: str):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, name, buf = {}, None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = buf
            name, buf = kernel_name(m.group(1)), []
        elif name:
            line = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)
            line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
            if line:
                buf.append(line)
    if name:
        out[name] = buf
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    diff = 0
    for k in sorted(set(a) | set(b)):
        if k not in a:
            print(f"NEW   {k}  ({len(b[k])} lines)")
        elif k not in b:
            print(f"GONE  {k}")
        elif a[k] == b[k]:
            print(f"SAME  {k}  ({len(a[k])} lines)")
        else:
            print(f"DIFF  {k}  ({len(a[k])} -> {len(b[k])} lines)")
            diff += 1
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())

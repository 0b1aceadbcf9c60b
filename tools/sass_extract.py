#!/usr/bin/env python
"""Opcode counts per kernel of the in-tree libvoxb200.so (cuobjdump -sass, sm_100a) -> profiles/rNN_sass_extract.txt.
Shows which Blackwell-specific instructions the kernels really contain (UBLKCP = cp.async.bulk shared->global,
FFMA2 / FADD2 / FMUL2 = packed FP32, REDUX, ...) without depending on the git-ignored binary."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "wgpu-voxel-engine_b200", "libvoxb200.so")
WATCH = ["UBLKCP", "UTMACMDFLUSH", "UTC", "LDTM", "FFMA2", "FADD2", "FMUL2", "REDUX", "FSET", "LDS", "STS", "LDG", "STG", "POPC", "VOTE",
         "SHFL", "BAR", "ATOMS", "REDG", "FRND", "F2I", "I2F", "IMAD", "LOP3", "FSEL", "FSETP", "FMNMX", "FMNMX3", "FFMA", "FMUL", "FADD"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
    arch = re.search(r"arch = (sm_\w+)", txt)
    print("# cuobjdump -sass wgpu-voxel-engine_b200/libvoxb200.so   (%s); static opcode counts per kernel" % (arch.group(1) if arch else "?"))
    print("# no UTCxMMA / LDTM / UTMA* anywhere: nothing on this path is a contraction or a tensor-map tile;")
    print("# UBLKCP = the emit kernel's cp.async.bulk shared->global store (UTMACMDFLUSH = its commit_group); FFMA2 / FADD2 / FMUL2 = packed FP32 in the fill")
    for name, c in kernels.items():
        tot = sum(c.values())
        watch = ["%s %d" % (k, c[k]) for k in WATCH if c.get(k)]
        print("\n== %s\n   %d instructions; %s" % (name[:150], tot, ", ".join(watch)))
    every = collections.Counter()
    for c in kernels.values():
        every.update(c)
    print("\n== whole library: " + ", ".join("%s %d" % (k, every[k]) for k in ("UBLKCP", "UTMACMDFLUSH", "FFMA2", "FADD2", "FMUL2", "REDUX", "FSET") if every.get(k)))
    bad = [k for k in every if k.startswith(("UTC", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "QMMA", "IMMA"))]
    print("   tensor-core / tensor-map opcodes: %s" % (", ".join(bad) if bad else "none"))


if __name__ == "__main__":
    main()

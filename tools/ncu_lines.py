#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins the SASS page of an ncu report (executed
instructions per address) with nvdisasm's line info of the in-tree cubin.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep kernels_mesh vx_mesh_emit_kernelILb0ELi1 [top]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(unit: str, func: str):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "wgpu-voxel-engine_b200", "libvoxb200.so")],
                   cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = glob.glob(os.path.join(tmp, unit + "*.cubin"))[0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout
    amap, cur, on = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//-----") and ".text." in ln:
            on = func in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return amap


def main():
    rep, unit, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(unit, func)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    per_line, per_line_s = collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in rows[rows.index(hdr) + 1:]:
        try:
            a, v = int(r[ia], 16), int(r[ie])
            sm = int(r[isamp] or 0)
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        key = amap.get(a - base, (("?", 0), ""))[0]
        per_line[key] += v
        per_line_s[key] += sm
        tot += v
        tots += sm
    src = {}
    print("# %s : %s   total warp-instructions %d, stall samples %d" % (os.path.basename(rep), func, tot, tots))
    print("%7s %7s  %-22s %s" % ("inst%", "samp%", "file:line", "source"))
    for key, v in per_line.most_common(top):
        f, l = key if key else ("?", 0)
        if f not in src:
            cands = glob.glob(os.path.join(ROOT, "wgpu-voxel-engine_b200", "csrc", f))
            src[f] = open(cands[0]).read().splitlines() if cands else []
        text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
        print("%6.2f%% %6.2f%%  %-22s %s" % (100.0 * v / tot, 100.0 * per_line_s[key] / max(tots, 1), "%s:%d" % (f, l), text))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Joins an `ncu --page source --csv` dump (SASS level) with `nvdisasm --print-line-info` of the same
cubin, and aggregates executed instructions / active threads / stall samples per CUDA source line.

  cuobjdump -xelf all vk-renderer_b200/libvkrt_cuda.so          # -> *.cubin
  nvdisasm --print-line-info vkrt_render.sm_100a.cubin > render.sass
  ncu -i prof.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py src.csv render.sass <mangled kernel name> [top_n]
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """offset -> (file, line) for the kernel's .text section"""
    out, cur, active = {}, None, False
    sec = re.compile(r"^\s*\.section\s+\.text\.(\S+?),")
    mark = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);")
    for ln in open(path, errors="replace"):
        m = sec.match(ln)
        if m:
            active = (m.group(1) == kernel)
            continue
        if not active:
            continue
        m = mark.search(ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = ins.match(ln)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = sass_lines(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    body = rows[hdr_i + 1:]
    base = int(body[0][col["Address"]], 16) if body[0][col["Address"]].startswith("0x") else int(body[0][col["Address"]])
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        a = r[col["Address"]]
        off = (int(a, 16) if a.startswith("0x") else int(a)) - base
        key = lines.get(off) or ("?", 0)
        ie = int(float(r[col["Instructions Executed"]] or 0))
        te = int(float(r[col["Thread Instructions Executed"]] or 0))
        smp = int(float(r[col["# Samples"]] or 0))
        g = agg[key]
        g[0] += ie; g[1] += te; g[2] += smp; g[3] += 1
        tot[0] += ie; tot[1] += te; tot[2] += smp
    print("total warp-inst %.3e  thread-inst %.3e  avg threads/inst %.2f  samples %d" % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
    print("%-22s %6s %8s %8s %7s %6s" % ("file:line", "sass", "inst%", "samples%", "thr/inst", ""))
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%-22s %6d %7.2f%% %7.2f%% %7.2f" % ("%s:%d" % key, g[3], 100.0 * g[0] / max(tot[0], 1), 100.0 * g[2] / max(tot[2], 1), g[1] / max(g[0], 1)))


if __name__ == "__main__":
    main()

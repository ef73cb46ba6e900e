#!/usr/bin/env python
"""Where the traversal kernel's issue slots go: executed warp instructions, stall samples and active lanes per PHASE of
k_wf_trace (ray set-up / inner node step / leaf test / loop control and votes / finish), from the SASS-level source page
of an ncu report joined with nvdisasm's line table (tools/ncu_lines.py does the join).

  cuobjdump -xelf all vk-renderer_b200/libvkrt_cuda.so && nvdisasm --print-line-info vkrt_wavefront.sm_100a.cubin > wf.sass
  ncu -i gpurun_out/<tag>_full_trace.ncu-rep --page source --csv --print-kernel-base mangled > trace_src.csv
  python tools/ncu_phases.py trace_src.csv wf.sass [launch index, default 1]

Phases are found from the sources themselves (function spans in csrc/vkrt_device.cuh / vkrt_arith.cuh, the section
comments of k_wf_trace in csrc/vkrt_wavefront.cu), so the table follows the code when lines move.
"""
import csv
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as N  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vk-renderer_b200", "csrc")

FUNC_PHASE = {
    # csrc/vkrt_device.cuh
    "trav_inner_step_q": "inner step", "qslab_test": "inner step", "qcode": "inner step", "ldg256u": "inner step",
    "push": "inner step", "pop": "inner step",
    "leaf_test": "leaf test", "s_consider": "leaf test", "sphere_intersect": "leaf test", "slab_test": "leaf test",
    "trav_leaf_step": "leaf test", "sphere_pad_radius": "leaf test",
    "slab_setup": "ray set-up", "safe_inv": "ray set-up", "qray_setup": "ray set-up", "trav_init": "ray set-up",
    "path_tmax": "ray set-up",
}


def function_spans(path):
    """line -> name of the (last) function defined at or above it"""
    out, cur = {}, None
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:VKRT_DEV|__device__|static|inline|__host__)[^;(]*?\b(\w+)\s*\(")
    for i, ln in enumerate(open(path, errors="replace"), 1):
        m = pat.match(ln)
        if m and not ln.rstrip().endswith(";"):
            cur = m.group(1)
        out[i] = cur
    return out


def trace_sections(path):
    """line -> section of k_wf_trace, from its '// ----' comments"""
    out, cur, inside = {}, None, False
    for i, ln in enumerate(open(path, errors="replace"), 1):
        if "k_wf_trace(" in ln and "__global__" in ln:
            inside, cur = True, "ray set-up"
        elif inside and ln.startswith("}"):
            inside = False
        if inside:
            if "// ---- refill" in ln:
                cur = "ray set-up"
            elif "// ---- traverse" in ln:
                cur = "loop control + votes"
            elif "// ---- finish" in ln:
                cur = "finish (result stores)"
            elif "trav_leaf_step" in ln or "s_cold" in ln and "const" in ln:
                out[i] = "leaf test"
                continue
            out[i] = cur
    return out


def main():
    src_csv, sass = sys.argv[1:3]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rows = list(csv.reader(open(src_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    # ncu prints every launch twice (one table per source view): keep one of each pair
    if len(starts) % 2 == 0 and all(rows[starts[i] + 2][:8] == rows[starts[i + 1] + 2][:8] for i in range(0, len(starts), 2)):
        starts = starts[0::2]
    s = starts[which]
    e = next((i for i in range(s + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    kernel = rows[s][1]
    block = rows[s:e]
    hdr_i = next(i for i, r in enumerate(block) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(block[hdr_i])}
    body = block[hdr_i + 1:]
    base = int(body[0][col["Address"]], 16)
    lines = N.sass_lines(sass, kernel)
    dev = function_spans(os.path.join(CSRC, "vkrt_device.cuh"))
    ari = function_spans(os.path.join(CSRC, "vkrt_arith.cuh"))
    wfs = trace_sections(os.path.join(CSRC, "vkrt_wavefront.cu"))
    wff = function_spans(os.path.join(CSRC, "vkrt_wavefront.cu"))
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        key = lines.get(int(r[col["Address"]], 16) - base)
        phase = "other"
        if key:
            f, l = key
            if f == "vkrt_device.cuh":
                phase = FUNC_PHASE.get(dev.get(l), "other (%s)" % dev.get(l))
            elif f == "vkrt_arith.cuh":
                phase = "IEEE division / sqrt, vector helpers (set-up and leaf)"
            elif f == "vkrt_wavefront.cu":
                phase = wfs.get(l) or ("loop control + votes" if wff.get(l) == "nested_inner_steps" else "ray set-up")
            else:
                phase = "loop control + votes" if "intrinsics" in f else "other (%s)" % f
        ie = int(float(r[col["Instructions Executed"]] or 0))
        te = int(float(r[col["Thread Instructions Executed"]] or 0))
        sm = int(float(r[col["# Samples"]] or 0))
        a = agg[phase]
        a[0] += ie; a[1] += te; a[2] += sm
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print("kernel %s, launch %d of the capture" % (kernel[:40], which))
    print("warp instructions %.3e, thread instructions %.3e, lanes / instruction %.2f" % (tot[0], tot[1], tot[1] / max(tot[0], 1)))
    print("%-58s %8s %9s %12s" % ("phase", "inst %", "samples %", "lanes / inst"))
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] * 1000 < tot[0]:
            continue
        print("%-58s %7.2f%% %8.2f%% %12.2f" % (ph, 100.0 * a[0] / tot[0], 100.0 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1)))


if __name__ == "__main__":
    main()

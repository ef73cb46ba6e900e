#!/usr/bin/env python
"""Turns one tools/evidence.sh run (gpurun_out/<tag>_*) into the committed round evidence under profiles/.

  python tools/make_profiles.py <tag> [round-prefix, default r01]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
G = lambda name: os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, name))
P = lambda name: os.path.join(ROOT, "profiles", "%s_%s" % (rnd, name))


def ncu_rows(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    out = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[ci["Metric Value"]].replace(",", ""))
        u = r[ci["Metric Unit"]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)
        out.setdefault((int(r[ci["ID"]]), r[ci["Kernel Name"]]), {})[r[ci["Metric Name"]]] = v * scale
    return out


# ---- bench lines ------------------------------------------------------------------------------------
for f in ("cfg4", "cfg4_serial", "cfg4_mega", "cfg3", "cfg2", "ref"):
    src = G("bench_%s.json" % f)
    if os.path.exists(src) and os.path.getsize(src):
        line = open(src).read().strip().splitlines()[-1]
        json.loads(line)
        open(P("bench_%s.json" % f), "w").write(line + "\n")

# ---- ncu launch list of the bench command, per-launch metrics ----------------------------------------
shutil.copy(G("launches_wf.csv"), P("launches_wavefront_cfg4.csv"))
shutil.copy(G("trace_metrics.csv"), P("metrics_k_wf_trace_cfg4.csv"))
if os.path.exists(G("frame_metrics.csv")) and os.path.getsize(G("frame_metrics.csv")) > 1000:
    shutil.copy(G("frame_metrics.csv"), P("metrics_frame_cfg4.csv"))
if os.path.exists(G("mega_cfg2_metrics.csv")):
    shutil.copy(G("mega_cfg2_metrics.csv"), P("metrics_k_path_mega_cfg2.csv"))

# kernel shares of one frame from the launch list (one frame = generate .. the second reduce after it)
launches = ncu_rows(G("launches_wf.csv"))
seq = [(k[1].split("(")[0].replace("void ", ""), m["gpu__time_duration.sum"]) for k, m in launches.items()]
gens = [i for i, (k, _) in enumerate(seq) if "k_wf_generate" in k]
reduces = [i for i, (k, _) in enumerate(seq) if "k_wf_reduce" in k]
first = gens[2]                                  # skip the first frame (cold)
last = [i for i in reduces if i > gens[3]][0]    # two waves = one 16-spp frame
frame = seq[first:last + 1]
share = collections.defaultdict(float)
for k, t in frame:
    share[k] += t
total = sum(share.values())
trace_share = sum(v for k, v in share.items() if "k_wf_trace" in k) / total

# dram traffic per launch of the dominant kernel over one frame's trace launches
tm = ncu_rows(G("trace_metrics.csv"))
trace_launches = [m for k, m in tm.items()]
n_per_frame = sum(1 for k, _ in frame if "k_wf_trace" in k)
sel = trace_launches[n_per_frame:2 * n_per_frame] if len(trace_launches) >= 2 * n_per_frame else trace_launches[:n_per_frame]
traffic = sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in sel) / max(len(sel), 1)
tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (k_wf_trace), mean over the %d "
                  "consecutive launches of one cfg4 frame of `python bench.py --steps 1 --warmup 3` under ncu --clock-control none; "
                  "source: profiles/%s_metrics_k_wf_trace_cfg4.csv" % (len(sel), rnd),
      "cfg4_wavefront": traffic}
old = {}
try:
    old = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass
old.update(tj)
json.dump(old, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

# ---- full capture -> text summary ----------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_global_loads", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.max"]
rep = G("full_trace.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(P("ncu_full_k_wf_trace_cfg4.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 2 (the first two trace launches of a cfg4 frame)\n")
        f.write("# command: python bench.py --steps 1 --warmup 3 --no-cpu-baseline   (100k spheres, 1920x1080, 16 spp, depth 8, wavefront)\n")
        for r in rows[2:]:
            f.write("---\nKernel Name = %s\n" % r[hdr.index("Kernel Name")])
            for w in WANT:
                if w in hdr:
                    f.write("%s = %s %s\n" % (w, r[hdr.index(w)], units[hdr.index(w)]))

# ---- summary -------------------------------------------------------------------------------------------
with open(P("SUMMARY.md"), "w") as f:
    f.write("# Round evidence (%s), generated by tools/make_profiles.py from tools/evidence.sh run `%s`\n\n" % (rnd, tag))
    for b in ("cfg4", "cfg4_serial", "cfg4_mega", "cfg3", "cfg2", "ref"):
        try:
            d = json.loads(open(P("bench_%s.json" % b)).read())
            rf = d.get("roofline") or {}
            f.write("* `%s_bench_%s.json`: %.1f %s, %.2f ms/step, e2e %.1f, roofline frac %s, share_of_step %s\n"
                    % (rnd, b, d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], rf.get("frac"), rf.get("share_of_step")))
            if "share_of_step_serial" in rf:
                f.write("  * waves serialised on one stream (each launch timed alone, comparable with the ncu launch list below): "
                        "traversal launches %.2f ms/frame, share_of_step_serial %.3f, frac_serial %.3f\n"
                        % (rf["kernel_ms_per_frame_serial"], rf["share_of_step_serial"], rf["frac_serial"]))
        except Exception as e:
            f.write("* %s: missing (%s)\n" % (b, e))
    f.write("\n## Kernel shares of one cfg4 frame in the ncu launch list (`%s_launches_wavefront_cfg4.csv`, serialised, cold caches)\n\n" % rnd)
    f.write("| kernel | launches | ms | share |\n|---|---|---|---|\n")
    cnt = collections.Counter(k for k, _ in frame)
    for k, v in sorted(share.items(), key=lambda kv: -kv[1]):
        f.write("| %s | %d | %.3f | %.1f %% |\n" % (k, cnt[k], v / 1e6, 100 * v / total))
    f.write("\nframe kernels: %d launches, %.3f ms summed; traversal kernels' share %.3f "
            "(bench.py's live `share_of_step_serial` must agree with this; `share_of_step` is taken while the two waves overlap)\n" % (len(frame), total / 1e6, trace_share))
    f.write("\ndram bytes per trace launch (mean over one frame): %.1f MB -> profiles/traffic.json\n" % (traffic / 1e6))
# other evidence files of the round, when present
extra = [("sweeps.txt", "tuning sweeps (one bench line per compile-time variant)"),
         ("ray_count_crosscheck.txt", "the CPU oracle's ray counts of the bench's full-size frames against the device counters (no GPU needed)"),
         ("spirv_fuzz.txt", "random-pose fuzz of the oracle against the reference's compiled shaders")]
with open(P("SUMMARY.md"), "a") as f:
    have = [(n, d) for n, d in extra if os.path.exists(P(n))]
    if have:
        f.write("\n## Other evidence\n\n")
        for n, d in have:
            f.write("* `%s_%s`: %s\n" % (rnd, n, d))
print(open(P("SUMMARY.md")).read())

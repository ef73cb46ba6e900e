#!/usr/bin/env python
"""Turns one tools/evidence_r2.sh run (gpurun_out/<tag>_*) into the committed round evidence under profiles/.

  python tools/make_profiles_r2.py <tag> [round-prefix, default r02] [commit the run was taken at, default HEAD]
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
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
G = lambda name: os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, name))
P = lambda name: os.path.join(ROOT, "profiles", "%s_%s" % (rnd, name))


def head():
    if len(sys.argv) > 3:            # the commit the evidence run was taken at, when it is not the current one
        return sys.argv[3]
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception:
        return "?"


def ncu_rows(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    out = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[ci["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        u = r[ci["Metric Unit"]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)
        out.setdefault((int(r[ci["ID"]]), r[ci["Kernel Name"]]), {})[r[ci["Metric Name"]]] = v * scale
    return out


def short(k):
    return k.split("(")[0].replace("void ", "").replace("vkrt::", "")


# ---- bench lines ------------------------------------------------------------------------------------
for f in ("bench", "bench_ref", "bench_mega"):
    src = G("%s.json" % f)
    if os.path.exists(src) and os.path.getsize(src):
        line = open(src).read().strip().splitlines()[-1]
        json.loads(line)
        open(P("%s.json" % f), "w").write(line + "\n")
if os.path.exists(G("pytest.log")):
    shutil.copy(G("pytest.log"), P("pytest_gpu.log"))

# ---- ncu launch list, per-launch metrics ---------------------------------------------------------------
shutil.copy(G("launches.csv"), P("launches_cfg4.csv"))
shutil.copy(G("frame_metrics.csv"), P("metrics_frame_cfg4.csv"))
for src, dst in (("mega_cfg2_metrics.csv", "metrics_k_path_mega_cfg2t.csv"), ("cfg3_metrics.csv", "metrics_frame_cfg3.csv")):
    if os.path.exists(G(src)) and os.path.getsize(G(src)) > 1000:
        shutil.copy(G(src), P(dst))

launches = ncu_rows(G("launches.csv"))
seq = [(short(k[1]), m["gpu__time_duration.sum"]) for k, m in launches.items()]
gens = [i for i, (k, _) in enumerate(seq) if "generate" in k]
reduces = [i for i, (k, _) in enumerate(seq) if "k_wf_reduce" in k]
first = gens[2]                                  # skip the cold frames
last = [i for i in reduces if i > first][0]      # one wave = one 16-spp frame
frame = seq[first:last + 1]
share = collections.defaultdict(float)
for k, t in frame:
    share[k] += t
total = sum(share.values())
trace_share = sum(v for k, v in share.items() if "k_wf_trace" in k) / total


def frame_of(metrics):
    """the launches of one warm frame (third generate .. its reduce is not in the filtered list: up to the next generate)"""
    ks = list(metrics.items())
    g = [i for i, (k, _) in enumerate(ks) if "generate" in k[1]]
    a = g[2] if len(g) > 3 else g[-2]
    b = g[3] if len(g) > 3 else g[-1]
    return ks[a:b]


def limiters(ms):
    """duration-weighted means of the per-launch metrics"""
    w = sum(m["gpu__time_duration.sum"] for m in ms)
    f = lambda name: sum(m.get(name, 0.0) * m["gpu__time_duration.sum"] for m in ms) / max(w, 1e-9)
    return {"alu_pipe_pct": f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "lanes_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
            "l2_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_bytes_per_launch": sum(m.get("lts__t_bytes.sum", 0.0) for m in ms) / max(len(ms), 1),
            "ms_summed": w / 1e6, "launches": len(ms)}


ev = {}
fm = frame_of(ncu_rows(G("frame_metrics.csv")))
tr = [m for k, m in fm if "k_wf_trace" in k[1]]
lg = [m for k, m in fm if "k_wfd_logic" in k[1] or "generate" in k[1]]
src = "profiles/%s_metrics_frame_cfg4.csv (ncu --clock-control none, `bench.py --steps 1 --warmup 3 --no-extras`, commit %s)" % (rnd, head())
ev["cfg4_wavefront"] = {"source": src, "kernel": "k_wf_trace<3,1,0> (the %d traversal launches of one cfg4 frame)" % len(tr),
                        "dram_bytes_per_launch": sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in tr) / max(len(tr), 1),
                        "limiters": limiters(tr), "logic_kernels": limiters(lg)}
if os.path.exists(G("cfg3_metrics.csv")) and os.path.getsize(G("cfg3_metrics.csv")) > 1000:
    rows3 = ncu_rows(G("cfg3_metrics.csv"))
    tr3 = [m for k, m in rows3.items() if "k_wf_trace" in k[1]][8:]
    if tr3:
        ev["cfg3_wavefront"] = {"source": src.replace("cfg4", "cfg3").replace("--no-extras", "--workload cfg3 --no-extras"),
                                "kernel": "k_wf_trace<3,1,0> (%d traversal launches of cfg3 waves)" % len(tr3),
                                "dram_bytes_per_launch": sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in tr3) / len(tr3),
                                "limiters": limiters(tr3)}
if os.path.exists(G("mega_cfg2_metrics.csv")) and os.path.getsize(G("mega_cfg2_metrics.csv")) > 1000:
    mg = [m for k, m in ncu_rows(G("mega_cfg2_metrics.csv")).items()][1:]
    if mg:
        ev["cfg2t_mega"] = {"source": "profiles/%s_metrics_k_path_mega_cfg2t.csv" % rnd, "kernel": "k_path_mega<0,0>",
                            "dram_bytes_per_launch": sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in mg) / len(mg),
                            "limiters": limiters(mg)}
        ev["cfg2_mega"] = ev["cfg2t_mega"]
json.dump(ev, open(os.path.join(ROOT, "profiles", "ncu_limiters.json"), "w"), indent=1)

# ---- microbenchmark peaks ---------------------------------------------------------------------------------
try:
    d = json.loads(open(P("bench.json")).read())
    json.dump({"fp32_ffma_tflops": d.get("fp32_peak_tflops_measured"), "l2_read_gbs": d.get("l2_read_gbs_measured"),
               "how": "vkrt_measure_fp32_peak / vkrt_measure_l2_bandwidth (csrc/vkrt_micro.cu) inside `bench.py --micro`, commit %s" % head(),
               "clocks": d.get("clocks")}, open(os.path.join(ROOT, "profiles", "peaks_micro.json"), "w"), indent=1)
except Exception as e:
    print("peaks_micro.json skipped:", e)

# ---- full captures -> text summaries ------------------------------------------------------------------------
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
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.max"]
for rep_name, out_name, what in (("full_trace.ncu-rep", "ncu_full_k_wf_trace_cfg4.txt", "k_wf_trace -c 3 (the first three traversal launches of a cfg4 frame)"),
                                 ("full_logic.ncu-rep", "ncu_full_k_wfd_logic_cfg4.txt", "k_wfd_logic -c 2 (the first two logic launches of a cfg4 frame)")):
    rep = G(rep_name)
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(P(out_name), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s\n" % what)
        f.write("# command: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras   (100k spheres, 1920x1080, 16 spp, depth 8, wavefront), commit %s\n" % head())
        for r in rows[2:]:
            f.write("---\nKernel Name = %s\n" % r[hdr.index("Kernel Name")])
            for w in WANT:
                if w in hdr:
                    f.write("%s = %s %s\n" % (w, r[hdr.index(w)], units[hdr.index(w)]))

# ---- where the traversal kernel's issue slots go (tools/ncu_phases.py on the full capture's source page) -----
phases_txt = ""
try:
    import tempfile
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vk-renderer_b200", "libvkrt_cuda.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sass = os.path.join(tmp, "wf.sass")
    open(sass, "w").write(subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "vkrt_wavefront.sm_100a.cubin")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)
    srcp = os.path.join(tmp, "trace_src.csv")
    open(srcp, "w").write(subprocess.run(["ncu", "-i", G("full_trace.ncu-rep"), "--page", "source", "--csv", "--print-kernel-base", "mangled"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)
    for which in (0, 1, 2):
        phases_txt += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), srcp, sass, str(which)],
                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout + "\n"
    with open(P("trace_phases.txt"), "w") as f:
        f.write("# tools/ncu_phases.py on %s_full_trace.ncu-rep (ncu --set full --import-source on, the first three traversal launches of a cfg4 frame),\n"
                "# library built from commit %s: executed warp instructions, stall samples and active lanes per phase of k_wf_trace\n" % (tag, head()))
        f.write(phases_txt)
except Exception as e:
    print("trace_phases.txt skipped:", e)

# ---- summary -------------------------------------------------------------------------------------------
with open(P("SUMMARY.md"), "w") as f:
    f.write("# Round evidence (%s), generated by tools/make_profiles_r2.py from tools/evidence_r2.sh run `%s` (commit %s)\n\n" % (rnd, tag, head()))
    for b in ("bench", "bench_mega", "bench_ref"):
        try:
            d = json.loads(open(P("%s.json" % b)).read())
            rf = d.get("roofline") or {}
            f.write("* `%s_%s.json`: %.1f %s, %.2f ms/step, e2e %.1f, traversed %.1f Mrays/s, roofline frac %s (kernel %.2f ms/frame, share_of_step %s)\n"
                    % (rnd, b, d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("traversed_mrays_per_s", 0.0), rf.get("frac"),
                       rf.get("kernel_ms_per_frame", 0.0), rf.get("share_of_step")))
            for k, v in (d.get("workloads") or {}).items():
                f.write("  * %s: %.3f ms/step, %.1f Mrays/s (%d steps)\n" % (k, v["ms_per_step"], v["value"], v["steps"]))
        except Exception as e:
            f.write("* %s: missing (%s)\n" % (b, e))
    f.write("\n## Kernel shares of one cfg4 frame in the ncu launch list (`%s_launches_cfg4.csv`, serialised, cold caches)\n\n" % rnd)
    f.write("| kernel | launches | ms | share |\n|---|---|---|---|\n")
    cnt = collections.Counter(k for k, _ in frame)
    for k, v in sorted(share.items(), key=lambda kv: -kv[1]):
        f.write("| %s | %d | %.3f | %.1f %% |\n" % (k, cnt[k], v / 1e6, 100 * v / total))
    f.write("\nframe kernels: %d launches, %.3f ms summed; traversal kernels' share %.3f (bench.py's live `roofline.share_of_step` must agree with this)\n"
            % (len(frame), total / 1e6, trace_share))
    L = ev["cfg4_wavefront"]["limiters"]
    f.write("\n## What bounds the traversal launches (duration-weighted over one frame, `%s_metrics_frame_cfg4.csv` -> profiles/ncu_limiters.json)\n\n" % rnd)
    for k in ("alu_pipe_pct", "issue_active_pct", "lsu_wavefronts_pct", "fma_pipe_pct", "lanes_per_instruction", "warps_active_pct", "l1_hit_pct",
              "l2_hit_pct", "l2_throughput_pct", "dram_throughput_pct"):
        f.write("* %s = %.1f\n" % (k, L[k]))
    f.write("* local / global load-store instruction counts: see the full captures (`%s_ncu_full_*.txt`, sass__inst_executed_*)\n" % rnd)
    f.write("* dram bytes per traversal launch: %.1f MB\n" % (ev["cfg4_wavefront"]["dram_bytes_per_launch"] / 1e6))
    # ---- multi-GPU lines (tools/r2_mgpu.sh, copied next to this file) and the other evidence of the round
    f.write("\n## Multi-GPU (`%s_bench_n*.json`: `torchrun ... bench.py --gpus N --steps 20 --warmup 5`; `%s_pytest_multi_gpu_N.log`)\n\n" % (rnd, rnd))
    try:
        one = json.loads(open(P("bench.json")).read())
        f.write("| N | scaling | ms/step | Mrays/s | vs N x one GPU | weak Mrays/s | cfg5 ms/frame | NCCL-gather ms/step |\n|---|---|---|---|---|---|---|---|\n")
        for n in (2, 4, 8):
            if not os.path.exists(P("bench_n%d.json" % n)):
                continue
            d = json.loads(open(P("bench_n%d.json" % n)).read())
            f.write("| %d | %s | %.3f | %.1f | %.3f | %.1f | %.1f | %.3f |\n" % (n, d["scaling"], d["ms_per_step"], d["value"], d["value"] / (n * one["value"]),
                    d["weak"]["value"], d["cfg5"]["ms_per_step"], d["nccl_gather"]["ms_per_step"]))
    except Exception as e:
        f.write("(multi-GPU lines missing: %s)\n" % e)
    f.write("\n## Other evidence in this directory\n\n"
            "* `%s_trace_phases.txt`: issue slots, stall samples and active lanes per phase of the traversal kernel (tools/ncu_phases.py)\n"
            "* `%s_sweeps.txt`: every tuning variant measured this round, taken or not\n"
            "* `%s_traversal_design_sim.txt`: CPU simulation of traversal schemes and tree builders (tests/tools/trav_sim.py)\n"
            "* `%s_sanitizer_*.txt`: compute-sanitizer memcheck / racecheck / synccheck (tools/sanitize.sh)\n"
            "* `%s_pytest_gpu.log`, `%s_pytest_multi_gpu_*.log`: the GPU test runs; `peaks_micro.json`, `ncu_limiters.json`: what bench.py reads\n"
            % ((rnd,) * 6))
print(open(P("SUMMARY.md")).read())

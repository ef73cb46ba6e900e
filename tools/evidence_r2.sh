#!/bin/bash
# usage (on the GPU box, via gpurun): tools/evidence_r2.sh <tag>
# Everything the round's profiles/ are made from (tools/make_profiles_r2.py): parity tests, the bench line with its
# sub-workloads, the reference arm, the ncu launch list of the bench command, per-launch metrics of the traversal and
# logic kernels, full captures of both, and the compute-bound small scenes.
T=${1:-ev2}; O=gpurun_out; mkdir -p $O
t0=$(date +%s); el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; el "smoke rc=$? $(tail -1 $O/${T}_smoke.log)"
timeout 900 python -m pytest tests -m gpu -q -rs > $O/${T}_pytest.log 2>&1; el "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
timeout 900 python bench.py --steps 20 --warmup 5 --micro > $O/${T}_bench.json 2> $O/${T}_bench.err; el bench
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; el reference
timeout 300 python bench.py --variant mega --no-cpu-baseline --no-extras > $O/${T}_bench_mega.json 2> $O/${T}_bench_mega.err; el mega
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $B > $O/${T}_n1.log 2>&1; el "launch list"
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sass__inst_executed_global_loads,sass__inst_executed_local_loads,sass__inst_executed_local_stores
timeout 300 ncu --metrics $M --clock-control none -k "regex:k_wf_trace|k_wfd_" -c 80 --csv --log-file $O/${T}_frame_metrics.csv $B > $O/${T}_n2.log 2>&1; el "frame metrics"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 3 -f -o $O/${T}_full_trace $B > $O/${T}_n3.log 2>&1; el "full trace"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wfd_logic -c 2 -f -o $O/${T}_full_logic $B > $O/${T}_n4.log 2>&1; el "full logic"
F=sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed
timeout 300 ncu --metrics $M,$F --clock-control none -k regex:k_path_mega -c 4 --csv --log-file $O/${T}_mega_cfg2_metrics.csv python bench.py --workload cfg2t --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_n5.log 2>&1; el "mega cfg2t metrics"
timeout 300 ncu --metrics $M,$F --clock-control none -k "regex:k_wf_trace|k_wfd_" -c 60 --csv --log-file $O/${T}_cfg3_metrics.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_n6.log 2>&1; el "cfg3 metrics"
python - <<P
import json
for f in ("bench", "bench_ref", "bench_mega"):
    try:
        d = json.loads(open("$O/${T}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.1f %s" % (d["value"], d["unit"]), "ms/step %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "frac", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
P
ls -la $O | grep ${T}_ | wc -l

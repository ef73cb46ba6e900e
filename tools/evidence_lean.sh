#!/bin/bash
# usage (on the GPU box, via gpurun): tools/evidence_lean.sh <tag>
# tools/evidence.sh without the CPU reference arm, most important artefacts first (for a call with little budget left).
T=${1:-ev}; O=gpurun_out; mkdir -p $O
t0=$(date +%s); el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 120 python -m pytest tests -m gpu -q -rs > $O/${T}_pytest.log 2>&1; el "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
timeout 120 python bench.py --micro > $O/${T}_bench_cfg4.json 2> $O/${T}_err1.txt; el cfg4
timeout 60 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_err4.txt; el cfg2
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_wf.csv $B > $O/${T}_n1.log 2>&1; el "launch list"
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
timeout 120 ncu --metrics $M --clock-control none -k regex:k_wf_trace -c 60 --csv --log-file $O/${T}_trace_metrics.csv $B > $O/${T}_n2.log 2>&1; el "trace metrics"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 2 -f -o $O/${T}_full_trace $B > $O/${T}_n4.log 2>&1; el "full capture"
timeout 60 python bench.py --variant mega --no-cpu-baseline > $O/${T}_bench_cfg4_mega.json 2> $O/${T}_err2.txt; el cfg4_mega
timeout 90 python bench.py --workload cfg3 --steps 3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_err3.txt; el cfg3
timeout 90 ncu --metrics $M,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed --clock-control none -k regex:k_path_mega -c 4 --csv --log-file $O/${T}_mega_cfg2_metrics.csv python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_n5.log 2>&1; el "mega metrics"
timeout 150 ncu --metrics $M --clock-control none -c 140 --csv --log-file $O/${T}_frame_metrics.csv $B > $O/${T}_n3.log 2>&1; el "frame metrics"
for f in cfg4 cfg4_mega cfg3 cfg2; do python - <<P
import json
try:
    d = json.loads(open("$O/${T}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.1f %s" % (d["value"], d["unit"]), "ms/step %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "frac", d.get("roofline", {}).get("frac"))
except Exception as e:
    print("$f FAILED", e)
P
done

#!/bin/bash
# usage (on the GPU box, via gpurun): tools/evidence.sh <tag>
# Everything the round's profiles/ are made from: parity tests, bench lines (cfg4 wavefront + megakernel, cfg3, cfg2,
# reference arm), ncu launch list of the bench command, per-launch metrics of the traversal kernels, one full capture.
T=${1:-ev}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python bench.py --micro > $O/${T}_bench_cfg4.json 2> $O/${T}_err1.txt
python bench.py --variant mega --no-cpu-baseline > $O/${T}_bench_cfg4_mega.json 2> $O/${T}_err2.txt
python bench.py --workload cfg3 --steps 3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_err3.txt
python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_err4.txt
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_err5.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_wf.csv $B > $O/${T}_n1.log 2>&1
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:k_wf_trace -c 60 --csv --log-file $O/${T}_trace_metrics.csv $B > $O/${T}_n2.log 2>&1
ncu --metrics $M --clock-control none -c 140 --csv --log-file $O/${T}_frame_metrics.csv $B > $O/${T}_n3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 2 -f -o $O/${T}_full_trace $B > $O/${T}_n4.log 2>&1
ncu --metrics $M,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed --clock-control none -k regex:k_path_mega -c 4 --csv --log-file $O/${T}_mega_cfg2_metrics.csv python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > $O/${T}_n5.log 2>&1
for f in cfg4 cfg4_mega cfg3 cfg2 ref; do python - <<P
import json
try:
    d = json.loads(open("$O/${T}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.1f %s" % (d["value"], d["unit"]), "ms/step %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "frac", d.get("roofline", {}).get("frac"))
except Exception as e:
    print("$f FAILED", e)
P
done
ls -la $O | grep ${T}_ | wc -l

#!/bin/bash
# One GPU call at the end of round 1 (little budget left): GPU tests of the new default build (inline PRMT + two-select
# child choice, the interop targets, slot rotation), a sweep of the tuning builds, a fresh bench line and launch list.
# Most important first: the call may be cut short.
O=gpurun_out; mkdir -p $O
S=$O/r1d_summary.txt; : > $S
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $S; }
timeout 330 python -m pytest tests -m gpu -q -rs > $O/r1d_pytest.log 2>&1; el "pytest rc=$? $(tail -1 $O/r1d_pytest.log)"
grep -E "FAILED|ERROR|SKIPPED|skipped" $O/r1d_pytest.log | head -20 >> $S
export SWEEP_ARGS="--steps 6 --warmup 3"
for v in new:- old:libvkrt_old.so pas6:libvkrt_pas6.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done
el "sweep 1 done"
VKRT_LIB=$PWD/vk-renderer_b200/libvkrt_pas6.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bvh or spheres or edge or shard or smoke" > $O/r1d_pytest_pas6.log 2>&1; el "pas6 parity rc=$? $(tail -1 $O/r1d_pytest_pas6.log)"
timeout 150 python bench.py > $O/r1d_bench_cfg4.json 2> $O/r1d_err_cfg4.txt; el "bench cfg4 rc=$?"
timeout 90 python bench.py --workload cfg2 --no-cpu-baseline > $O/r1d_bench_cfg2.json 2> $O/r1d_err_cfg2.txt; el "bench cfg2 rc=$?"
for v in pas6b10:libvkrt_pas6_b10.so pasu4:libvkrt_pas_u4.so l3:libvkrt_l3.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done
el "sweep 2 done"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1d_launches_wf.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r1d_n1.log 2>&1; el "ncu launch list rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $S
el "end"

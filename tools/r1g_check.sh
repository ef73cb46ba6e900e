#!/bin/bash
# GPU call 4: L1-bypass of the once-used ray records, reduced cold state
O=gpurun_out; mkdir -p $O
S=$O/r1g_summary.txt; : > $S
export SWEEP_ARGS="--steps 8 --warmup 3"
for v in new:- cg:libvkrt_cg.so c6:libvkrt_c6.so cgc6:libvkrt_cgc6.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done
VKRT_LIB=$PWD/vk-renderer_b200/libvkrt_cg.so timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/r1g_pytest_cg.log 2>&1; echo "cg parity rc=$? $(tail -1 $O/r1g_pytest_cg.log)" | tee -a $S
VKRT_LIB=$PWD/vk-renderer_b200/libvkrt_cgc6.so timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/r1g_pytest_cgc6.log 2>&1; echo "cgc6 parity rc=$? $(tail -1 $O/r1g_pytest_cgc6.log)" | tee -a $S

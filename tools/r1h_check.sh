#!/bin/bash
# GPU call 5: block sizes of the streaming kernels / the trace kernel, fetch chunk
O=gpurun_out; mkdir -p $O
S=$O/r1h_summary.txt; : > $S
export SWEEP_ARGS="--steps 8 --warmup 3"
for v in new:- sh128:libvkrt_sh128.so sh512:libvkrt_sh512.so tb64:libvkrt_tb64.so fc64:libvkrt_fc64.so fc256:libvkrt_fc256.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done

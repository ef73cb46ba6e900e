#!/bin/bash
# GPU call 3: shared-memory stack / occupancy sweep on top of the new default build
O=gpurun_out; mkdir -p $O
S=$O/r1f_summary.txt; : > $S
export SWEEP_ARGS="--steps 8 --warmup 3"
for v in new:- sm8:libvkrt_sm8.so sm12:libvkrt_sm12.so sm16:libvkrt_sm16.so b10:libvkrt_b10.so sm8b10:libvkrt_sm8b10.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done

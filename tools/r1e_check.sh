#!/bin/bash
# GPU call 2 of the round's tail: GPU tests (interop fd semantics), sweep of the sentinel-stack builds, parity of `sent`.
O=gpurun_out; mkdir -p $O
S=$O/r1e_summary.txt; : > $S
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $S; }
timeout 200 python -m pytest tests -m gpu -q -rs > $O/r1e_pytest.log 2>&1; el "pytest rc=$? $(tail -1 $O/r1e_pytest.log)"
grep -E "FAILED|ERROR|SKIPPED" $O/r1e_pytest.log | head -20 >> $S
export SWEEP_ARGS="--steps 8 --warmup 3"
for v in new:- sent:libvkrt_sent.so new2:- sent2:libvkrt_sent.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done
el "sweep 1 done"
VKRT_LIB=$PWD/vk-renderer_b200/libvkrt_sent.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/r1e_pytest_sent.log 2>&1; el "sent parity rc=$? $(tail -1 $O/r1e_pytest_sent.log)"
for v in r16:libvkrt_sent_r16.so r24:libvkrt_sent_r24.so lb6:libvkrt_sent_lb6.so lb12:libvkrt_sent_lb12.so u2:libvkrt_sent_u2.so; do timeout 90 tools/sweep.sh $v 2>&1 | tee -a $S; done
el "end"

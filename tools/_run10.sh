mkdir -p gpurun_out
export SWEEP_ARGS="--steps 20 --warmup 5"
tools/sweep.sh skey:- 2>&1 | tee gpurun_out/r2ad_sweep.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2ad_sweep.txt

mkdir -p gpurun_out
export SWEEP_ARGS="--steps 20 --warmup 5"
tools/sweep.sh base:- u2:libvkrt_u2.so 2>&1 | tee gpurun_out/r2ab_sweep.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pixel_major or waves" 2>&1 | tail -3 | tee -a gpurun_out/r2ab_sweep.txt

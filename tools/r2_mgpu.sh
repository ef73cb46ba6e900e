# usage: tools/r2_mgpu.sh <N> <tag>   -- N-GPU pass: multi-GPU parity tests, then the bench line the driver will run
N=${1:-2}; T=${2:-r2m}; O=gpurun_out
set -x
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -rs > $O/${T}_pytest_$N.log 2>&1; tail -8 $O/${T}_pytest_$N.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 900 $R --steps 20 --warmup 5 > $O/${T}_bench_$N.json 2> $O/${T}_bench_$N.err; tail -c 600 $O/${T}_bench_$N.err
python - <<PY
import json
try:
    d = json.loads(open("$O/${T}_bench_$N.json").read().strip().splitlines()[-1])
    print("N=$N", d["scaling"], "ms/step %.3f value %.1f e2e %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["exchange"])
    for k in ("weak", "strong", "cfg5", "nccl_gather"):
        if k in d:
            print("   ", k, "ms/step %.3f value %.1f" % (d[k]["ms_per_step"], d[k]["value"]), d[k]["exchange"], d[k]["config"]["parallelism"])
except Exception as e:
    print("bench FAILED", e)
PY
timeout 300 vk-renderer_b200/vkrt_headless --frames 200 --res 1024 --spp 4 --depth 4 --wavefront --seed 1 > $O/${T}_headless_1.txt 2>&1
timeout 300 vk-renderer_b200/vkrt_headless --frames 200 --res 1024 --spp 4 --depth 4 --wavefront --seed 1 --devices $(seq -s, 0 $((N-1))) > $O/${T}_headless_$N.txt 2>&1
tail -qn1 $O/${T}_headless_1.txt $O/${T}_headless_$N.txt

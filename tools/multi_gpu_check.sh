#!/bin/bash
# usage: tools/multi_gpu_check.sh <N> <tag>  -- weak / strong / cfg5 bench lines on N GPUs of one box
N=$1; T=${2:-mg}; O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
$R --steps 5 > $O/${T}_weak_$N.json 2> $O/${T}_weak_$N.err
$R --steps 5 --scaling strong > $O/${T}_strong_$N.json 2> $O/${T}_strong_$N.err
$R --workload cfg5 --steps 1 > $O/${T}_cfg5_$N.json 2> $O/${T}_cfg5_$N.err
for f in weak strong cfg5; do tail -1 $O/${T}_${f}_$N.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$f', d['n_gpus'], 'gpus', '%.1f' % d['value'], d['unit'], 'ms/step %.2f' % d['ms_per_step'], d['scaling'], d['config']['parallelism'])
except Exception as e:
    print('$f FAILED', e); print(open('$O/${T}_${f}_$N.err').read()[-1500:])
"; done

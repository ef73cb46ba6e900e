# round 2, first GPU pass on 1 GPU: parity tests, bench line, shard-of diagnostics
set -x
O=gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; tail -5 $O/r2b_pytest.log
timeout 600 python bench.py --steps 10 > $O/r2b_bench_cfg4.json 2> $O/r2b_bench_cfg4.err; tail -c 600 $O/r2b_bench_cfg4.err
for s in 8 4 2; do
  timeout 300 python bench.py --shard-of $s --steps 20 --no-cpu-baseline --no-extras > $O/r2b_shard_of_$s.json 2> $O/r2b_shard_of_$s.err
done
python - <<'PY'
import json
for s in (8, 4, 2):
    try:
        d = json.loads(open('gpurun_out/r2b_shard_of_%d.json' % s).read().strip().splitlines()[-1])
        print('shard-of', s, 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'trace serial %.3f' % d['roofline']['kernel_ms_per_frame'])
    except Exception as e:
        print('shard-of', s, 'FAILED', e)
try:
    d = json.loads(open('gpurun_out/r2b_bench_cfg4.json').read().strip().splitlines()[-1])
    print('cfg4 ms/step %.3f value %.1f frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))
    for k, v in d.get('workloads', {}).items():
        print(k, 'ms/step %.3f value %.1f' % (v['ms_per_step'], v['value']))
except Exception as e:
    print('bench FAILED', e)
PY
python tools/timeline.py 8 > $O/r2b_timeline8.txt 2>&1; tail -3 $O/r2b_timeline8.txt

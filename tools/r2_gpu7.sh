set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2h_pytest.log 2>&1; tail -5 $O/r2h_pytest.log
timeout 900 python bench.py > $O/r2h_bench.json 2> $O/r2h_bench.err; tail -c 400 $O/r2h_bench.err
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
timeout 300 $B --shard-of 8 > $O/r2h_s8.json 2> $O/r2h_s8.err
python - <<'PY'
import json
for f in ('gpurun_out/r2h_bench.json', 'gpurun_out/r2h_s8.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'e2e %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
        for k, v in d.get('workloads', {}).items():
            print('   ', k, 'ms/step %.3f value %.1f' % (v['ms_per_step'], v['value']))
    except Exception as e:
        print(f, 'FAILED', e)
PY

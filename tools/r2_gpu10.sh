O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "progressive" > $O/r2k_pytest.log 2>&1; tail -3 $O/r2k_pytest.log
B="python bench.py --no-cpu-baseline --no-extras"
for w in cfg2 cfg2t; do for v in mega wavefront; do
  timeout 300 $B --workload $w --variant $v > $O/r2k_${w}_$v.json 2> $O/r2k_${w}_$v.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2k_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), d.get('fp32', {}).get('frac'))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-500:])
PY

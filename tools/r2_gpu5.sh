set -x
O=gpurun_out
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
for k in 0 1 2 3 4; do
  VKRT_TUNE_STAGGER=$k timeout 300 $B --shard-of 8 > $O/r2f_s8_st$k.json 2> $O/r2f_s8_st$k.err
done
for k in 2 3 4; do
  VKRT_TUNE_STAGGER=$k timeout 300 $B > $O/r2f_full_st$k.json 2> $O/r2f_full_st$k.err
done
VKRT_TUNE_STAGGER=3 VKRT_TUNE_LANES=3 timeout 300 $B --shard-of 8 > $O/r2f_s8_st3_l3.json 2> $O/r2f_s8_st3_l3.err
VKRT_TUNE_STAGGER=2 VKRT_TUNE_LANES=4 timeout 300 $B --shard-of 8 > $O/r2f_s8_st2_l4.json 2> $O/r2f_s8_st2_l4.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or config5 or degenerate or multi_wave" > $O/r2f_pytest.log 2>&1; tail -3 $O/r2f_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'e2e %.3f' % d['e2e']['ms_per_step'])
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-800:])
PY

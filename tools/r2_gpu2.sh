# round 2, GPU pass 2 (1 GPU): dense pipeline parity + A/B against the indexed one, lanes at 1/8 shard
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; tail -5 $O/r2c_pytest.log
B="python bench.py --steps 10 --no-cpu-baseline --no-extras"
timeout 300 $B > $O/r2c_dense.json 2> $O/r2c_dense.err
VKRT_LIB=vk-renderer_b200/libvkrt_nodense.so timeout 300 $B > $O/r2c_nodense.json 2> $O/r2c_nodense.err
for l in 2 3 4; do
  VKRT_TUNE_LANES=$l timeout 300 $B --shard-of 8 --steps 20 > $O/r2c_s8_l$l.json 2> $O/r2c_s8_l$l.err
done
VKRT_TUNE_LANES=4 timeout 300 $B > $O/r2c_dense_l4.json 2> $O/r2c_dense_l4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f frac %.3f' % (r.get('kernel_ms_per_frame', 0), r.get('share_of_step', 0), r.get('frac', 0)))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-800:])
PY

O=gpurun_out
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
for n in tr11 tr10 tr9 sb128; do
  VKRT_LIB=vk-renderer_b200/libvkrt_$n.so timeout 300 $B > $O/r2n_$n.json 2> $O/r2n_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f' % (r['kernel_ms_per_frame'], r['share_of_step']))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-500:])
PY

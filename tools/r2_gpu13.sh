O=gpurun_out
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
VKRT_LIB=vk-renderer_b200/libvkrt_notri.so timeout 300 $B > $O/r2o_notri.json 2> $O/r2o_notri.err
timeout 300 $B > $O/r2o_base.json 2> $O/r2o_base.err
VKRT_LIB=vk-renderer_b200/libvkrt_notri.so timeout 300 $B --workload cfg2t > $O/r2o_notri_cfg2t.json 2> $O/r2o_notri_cfg2t.err
timeout 300 $B --workload cfg2t > $O/r2o_base_cfg2t.json 2> $O/r2o_base_cfg2t.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2o_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f' % (r.get('kernel_ms_per_frame',0), r.get('share_of_step',0)))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-500:])
PY

O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2j_pytest.log 2>&1; tail -3 $O/r2j_pytest.log
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
timeout 300 $B > $O/r2j_ps1.json 2> $O/r2j_ps1.err
VKRT_LIB=vk-renderer_b200/libvkrt_gps0.so timeout 300 $B > $O/r2j_ps0.json 2> $O/r2j_ps0.err
timeout 300 $B --shard-of 8 > $O/r2j_ps1_s8.json 2> $O/r2j_ps1_s8.err
timeout 300 $B --workload cfg3 --steps 3 > $O/r2j_ps1_cfg3.json 2> $O/r2j_ps1_cfg3.err
VKRT_LIB=vk-renderer_b200/libvkrt_gps0.so timeout 300 $B --workload cfg3 --steps 3 > $O/r2j_ps0_cfg3.json 2> $O/r2j_ps0_cfg3.err
python tools/timeline.py 1 > $O/r2j_timeline1.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f' % (r['kernel_ms_per_frame'], r['share_of_step']))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-500:])
PY

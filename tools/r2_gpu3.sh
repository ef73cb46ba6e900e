set -x
O=gpurun_out
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
VKRT_TUNE_LANES=1 timeout 300 $B --shard-of 8 > $O/r2d_s8_l1.json 2> $O/r2d_s8_l1.err
VKRT_TUNE_LANES=1 timeout 300 $B --shard-of 4 > $O/r2d_s4_l1.json 2> $O/r2d_s4_l1.err
timeout 300 $B --shard-of 4 > $O/r2d_s4_l2.json 2> $O/r2d_s4_l2.err
VKRT_LIB=vk-renderer_b200/libvkrt_fc32.so timeout 300 $B --shard-of 8 > $O/r2d_s8_fc32.json 2> $O/r2d_s8_fc32.err
VKRT_LIB=vk-renderer_b200/libvkrt_fc32.so timeout 300 $B > $O/r2d_full_fc32.json 2> $O/r2d_full_fc32.err
python tools/timeline.py 1 > $O/r2d_timeline1.txt 2>&1
python tools/timeline.py 8 > $O/r2d_timeline8.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2d_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f' % (r.get('kernel_ms_per_frame', 0), r.get('share_of_step', 0)))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-800:])
PY

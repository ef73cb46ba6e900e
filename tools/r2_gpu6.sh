set -x
O=gpurun_out
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
VKRT_TUNE_WAVE_SPP=16 timeout 300 $B --shard-of 8 > $O/r2g_s8_w16.json 2> $O/r2g_s8_w16.err
VKRT_TUNE_WAVE_SPP=16 timeout 300 $B --shard-of 4 > $O/r2g_s4_w16.json 2> $O/r2g_s4_w16.err
VKRT_TUNE_WAVE_SPP=16 timeout 300 $B --shard-of 2 > $O/r2g_s2_w16.json 2> $O/r2g_s2_w16.err
VKRT_TUNE_WAVE_SPP=16 timeout 300 $B > $O/r2g_full_w16.json 2> $O/r2g_full_w16.err
VKRT_TUNE_WAVE_SPP=16 VKRT_TUNE_LANES=3 timeout 300 $B --shard-of 8 > $O/r2g_s8_w16_l3.json 2> $O/r2g_s8_w16_l3.err
timeout 300 $B --shard-of 2 > $O/r2g_s2.json 2> $O/r2g_s2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2g_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'e2e %.3f' % d['e2e']['ms_per_step'])
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-800:])
PY

O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2p_pytest.log 2>&1; tail -3 $O/r2p_pytest.log
B="python bench.py --steps 20 --no-cpu-baseline --no-extras"
timeout 300 $B > $O/r2p_full.json 2> $O/r2p_full.err
timeout 300 $B --workload cfg2t > $O/r2p_cfg2t.json 2> $O/r2p_cfg2t.err
timeout 300 $B --workload cfg2 > $O/r2p_cfg2.json 2> $O/r2p_cfg2.err
timeout 300 $B --workload cfg2t --variant wavefront > $O/r2p_cfg2t_wf.json 2> $O/r2p_cfg2t_wf.err
timeout 300 $B --workload cfg3 --steps 3 > $O/r2p_cfg3.json 2> $O/r2p_cfg3.err
timeout 300 $B --shard-of 8 > $O/r2p_s8.json 2> $O/r2p_s8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2p_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'ms/step %.3f value %.1f' % (d['ms_per_step'], d['value']), 'trace serial %.3f share %.3f' % (r.get('kernel_ms_per_frame',0), r.get('share_of_step',0)))
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-500:])
PY

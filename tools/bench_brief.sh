#!/bin/bash
# usage: tools/bench_brief.sh <label> [bench.py args...]  -> one short line per run
label=$1; shift
python bench.py --no-cpu-baseline --no-extras "$@" 2>gpurun_out/err_$label.txt | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$label', 'Mrays/s %.1f' % d['value'], 'ms/frame %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'], 'trace-serial ms', round((d.get('roofline') or {}).get('kernel_ms_per_frame') or 0, 3), 'share', round((d.get('roofline') or {}).get('share_of_step') or 0, 3))
except Exception as e:
    print('$label FAILED', e); print(open('gpurun_out/err_$label.txt').read()[-800:])
"

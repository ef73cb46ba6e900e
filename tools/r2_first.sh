set -x
nvidia-smi -L
for s in 8 4 2; do
  python bench.py --shard-of $s --steps 20 --no-cpu-baseline > gpurun_out/r2a_shard_of_$s.json 2> gpurun_out/r2a_shard_of_$s.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2a_shard_of_$s.json').read().strip().splitlines()[-1])
print('shard-of', $s, 'ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'roof', d['roofline']['kernel_ms_per_frame'], d['roofline'].get('kernel_ms_per_frame_serial'))
PY
done
python tools/timeline.py 8 > gpurun_out/r2a_timeline8.txt 2>&1
tail -3 gpurun_out/r2a_timeline8.txt

#!/usr/bin/env python
"""Per-kernel breakdown of one wavefront frame from an `ncu --csv` launch list."""
import collections
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ci = {n: i for i, n in enumerate(hdr)}
data = collections.OrderedDict()
for r in rows[1:]:
    k = (r[ci['ID']], r[ci['Kernel Name']][:44])
    data.setdefault(k, {})[r[ci['Metric Name']]] = float(r[ci['Metric Value']].replace(',', ''))
seq = [(i, k, m) for (i, k), m in data.items()]
start = next(j for j, (i, k, m) in enumerate(seq) if 'k_wf_generate' in k)
end = next(j for j, (i, k, m) in enumerate(seq) if 'k_wf_reduce' in k and j > start)
frame = seq[start:end + 1]
T = sum(m['gpu__time_duration.sum'] for _, _, m in frame)
print("frame kernels", len(frame), "sum ms", T / 1e6)
n_show = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for i, k, m in frame[:n_show]:
    print("%-46s %8.3f ms thr/inst %5.2f warps%% %5.1f issue%% %5.1f l1wf%% %5.1f" % (
        k, m['gpu__time_duration.sum'] / 1e6, m.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0),
        m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0), m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0),
        m.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 0)))
agg = collections.defaultdict(float)
for i, k, m in frame:
    agg[k.split('(')[0]] += m['gpu__time_duration.sum'] / 1e6
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%-44s %8.3f ms %5.1f%%" % (k, v, 100 * v / (T / 1e6)))

#!/usr/bin/env python
"""Summarise an ncu per-launch CSV (gpu__time_duration.sum) of tools/train_bench.py --no-graph: one training step
(from one stem_tc_kernel launch to the next), split into forward / loss / backward / optimizer, per kernel."""
import collections, csv, json, re, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith('==')]))
names = [(re.sub(r'\(.*', '', re.sub(r'void |sedt::|\(anonymous namespace\)::|<unnamed>::|at::native::', '', r['Kernel Name'])),
          float(r['Metric Value'].replace(',', '')) / 1e3) for r in rows]
idx = [i for i, (n, t) in enumerate(names) if n.startswith('stem_tc_kernel')]
step = names[idx[-2]:idx[-1]]
# phase boundaries inside the step
def first(pred, start=0):
    return next(i for i in range(start, len(step)) if pred(step[i][0]))
i_loss = first(lambda n: n.startswith('set_criterion_kernel'))
i_bwd = first(lambda n: n.startswith('fill_zero_kernel'), i_loss)
i_opt = first(lambda n: n.startswith('grad_sumsq_kernel'), i_bwd)
i_pack = first(lambda n: n.startswith('pack_jobs_kernel') or n.startswith('stem_pack_kernel'), i_opt)
phases = {'forward': step[:i_loss], 'loss': step[i_loss:i_bwd], 'backward': step[i_bwd:i_opt], 'optimizer': step[i_opt:i_pack],
          'pack (next step)': step[i_pack:]}
out = {}
for ph, ks in phases.items():
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in ks:
        agg[n[:60]][0] += 1; agg[n[:60]][1] += t
    out[ph] = {'launches': len(ks), 'us': round(sum(t for _, t in ks), 1),
               'kernels': {k: {'n': c, 'us': round(t, 1)} for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])}}
out['total'] = {'launches': len(step), 'us': round(sum(t for _, t in step), 1)}
json.dump(out, open(sys.argv[2], 'w'), indent=1) if len(sys.argv) > 2 else None
for ph, v in out.items():
    if ph == 'total': print('total', v); continue
    print(f"== {ph}: {v['launches']} launches, {v['us']} us")
    for k, kv in list(v['kernels'].items())[:14]:
        print(f"   {kv['us']:9.1f} us {kv['n']:4d}  {k}")

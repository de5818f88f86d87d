#!/usr/bin/env python
"""Summarise an ncu per-launch CSV of ONE forward (gpu__time_duration.sum) by network section."""
import csv, re, sys
def load(fn):
    return list(csv.DictReader([l for l in open(fn) if not l.startswith('==')]))
rows = load(sys.argv[1])
sec_bounds = None
out = []
names = []
for r in rows:
    t = float(r['Metric Value'].replace(',', '')) / 1e3
    m = re.search(r'(conv_tc\d?_kernel|attention\w*|stem\w*|layernorm\w*|conv_simt\w*|heads\w*|pos_table\w*|fill_zero\w*)<?([^>(]*)', r['Kernel Name'])
    names.append((m.group(1), m.group(2).replace('__nv_bfloat16', 'bf16').strip(), r['Grid Size'], t))
# sections: stem(1) l1(10) l2(13) l3(19) l4(10) then pos_table... encoder until fill_zero, decoder until conv_simt, heads
sections = [('stem', 1), ('l1', 10), ('l2', 13), ('l3', 19), ('l4', 10)]
i = 0; summary = {}
for name, n in sections:
    summary[name] = sum(x[3] for x in names[i:i + n]); i += n
enc_end = next(j for j, x in enumerate(names) if x[0].startswith('fill_zero'))
summary['enc'] = sum(x[3] for x in names[i:enc_end])
dec_end = next(j for j, x in enumerate(names) if x[0].startswith('conv_simt'))
summary['dec'] = sum(x[3] for x in names[enc_end:dec_end])
summary['heads'] = sum(x[3] for x in names[dec_end:])
summary['total'] = sum(x[3] for x in names)
print({k: round(v) for k, v in summary.items()})
if len(sys.argv) > 2:
    for j, x in enumerate(names):
        print(j, x[0][:18], x[1][:16], x[2], f"{x[3]:.1f}")

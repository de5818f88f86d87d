#!/usr/bin/env python
"""Top stall-sample source lines per kernel of an .ncu-rep (needs -lineinfo and --import-source on).
    python tools/ncu_hotlines.py gpurun_out/x.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern = None; cur_file = None; hdr = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'Function Name': kern = r[1][:90]; data.setdefault(kern, []); continue
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 1 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and kern:
        d = dict(zip(hdr, r))
        try: s = int(d['# Samples'])
        except Exception: continue
        data[kern].append((s, cur_file, d['Line No'], d['Source'][:100]))
for k, v in data.items():
    tot = sum(s for s, *_ in v) or 1
    print(f"== {k}  ({tot} samples)")
    for s, f, l, src in sorted(v, key=lambda x: -x[0])[:topn]:
        print(f"  {100 * s / tot:5.1f}%  {f}:{l}  {src}")

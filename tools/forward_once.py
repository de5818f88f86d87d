#!/usr/bin/env python
"""Three eager forwards of bench.py's workload (config 2, B = 256, bf16) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file launches.csv python tools/forward_once.py
    python tools/launch_summary.py launches.csv 1 --csv profiles/x.csv --json profiles/x.json"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model
args = spec.config_args("c2")
model, _, _ = build_model(args)
model.load_state_dict(synth.synth_state_dict(args, 12))
model = model.cuda().eval()
x = synth.synth_clips(int(sys.argv[1]) if len(sys.argv) > 1 else 256, 496, 64, seed=2).cuda()
with torch.no_grad():
    for _ in range(3):
        model(x)
torch.cuda.synchronize()

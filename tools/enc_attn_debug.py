#!/usr/bin/env python
"""Phase timeline of the fused encoder attention kernel (SEDT_EA_DEBUG=1): python tools/enc_attn_debug.py [B] [S]"""
import os, sys
os.environ["SEDT_EA_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_enc_attn import make, run
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 124
args = make(B, S, 1, False)
for _ in range(3):
    run(*args, B, S)

#!/usr/bin/env python
"""Times ops.attention at the bench shape (B=16, 12 heads, 2049 tokens, randn*2 data as in bench.py) for the knob
combination given in the environment (the knobs are read once per process).  usage: python scripts/time_attn_variants.py [iters]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denseclip_vit_multimodal_b200 import ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B, H, N = 16, 12, 2049
D = H * 64
torch.manual_seed(0)
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 2).to(torch.bfloat16)
out = torch.empty(B, N, D, dtype=torch.bfloat16, device="cuda")
run = lambda: ops.attention(qkv, qkv, qkv, B=B, H=H, Nq=N, Nk=N, q_col0=0, k_col0=D, v_col0=2 * D, scale=0.125, out=out)
for _ in range(5):
    run()
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / iters)
knobs = {k: v for k, v in os.environ.items() if k.startswith("DCLIP_ATTN")}
print(knobs, " ".join(f"{r:.4f}" for r in res), "ms")

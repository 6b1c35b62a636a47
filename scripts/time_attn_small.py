"""Cross-attention of the ContextDecoder in isolation: 19 queries x 2048 keys, 4 heads, fp32 K/V with row stride 1536 (B = 16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denseclip_vit_multimodal_b200 import ops

def t(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B, K, N, Wd, H = 16, 19, 2048, 256, 4
q = torch.randn(B, K, Wd, device="cuda")
kv = torch.randn(B, N, 1536, device="cuda")
out = torch.empty(B, K, 2 * Wd, device="cuda", dtype=torch.bfloat16)
ms = t(lambda: ops.attention_small(q, kv, kv, B=B, H=H, q_first=0, q_count=K, Nk=N, q_col0=0, k_col0=512, v_col0=768, scale=0.125, out=out, out_split_off=Wd))
print(f"cross attention 19 x {N}, B={B}, H={H}: {ms*1e3:.1f} us   (DCLIP_ATTN_SMALL_NO_SPLIT={os.environ.get('DCLIP_ATTN_SMALL_NO_SPLIT', '0')})")

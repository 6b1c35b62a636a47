"""Latency of the ContextDecoder's small-M GEMMs (M = 16 images x 19 classes = 304 rows) for different BLOCK_N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denseclip_vit_multimodal_b200 import ops

def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

M = 304
for name, N, K, kw in (("out-proj (split, resid f32)", 256, 256, dict(resid=True)), ("qkv (split, f32 out)", 768, 256, dict()),
                       ("fc (split, gelu, split bf16 out)", 1024, 256, dict(gelu=True)), ("proj (split, resid)", 256, 1024, dict(resid=True)),
                       ("out_proj 256->512 f32", 512, 256, dict())):
    a = torch.randn(M, 2 * K, device="cuda").bfloat16(); w = torch.randn(N, 2 * K, device="cuda").bfloat16(); bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    for bn in (0, 128, 64):
        if kw.get("resid"):
            fn = lambda: ops.gemm(a, w, split_in=True, bias=bias, residual=x, out_f32=x, block_n=bn)
        elif kw.get("gelu"):
            ob = torch.empty(M, 2 * N, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, split_in=True, bias=bias, act="gelu", out_bf16=ob, split_out=True, block_n=bn)
        else:
            of = torch.empty(M, N, device="cuda")
            fn = lambda: ops.gemm(a, w, split_in=True, bias=bias, out_f32=of, block_n=bn)
        print(f"{name:36s} N={N:4d} K={K:4d} bn={bn:3d}: {t(fn)*1e3:6.1f} us (in a CUDA graph, back to back)")

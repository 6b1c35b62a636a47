"""Times the split-precision / generic-epilogue GEMMs of the DenseCLIP tail at the bench shape (B=16, 2048 tokens/image)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denseclip_vit_multimodal_b200 import ops

def t(fn, n=20):
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

M = 32784
def case(name, N, K, split, **kw):
    a = torch.randn(M, K * (2 if split else 1), device="cuda").bfloat16()
    w = torch.randn(N, K * (2 if split else 1), device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    of = torch.empty(M, N, device="cuda") if kw.pop("f32", False) else None
    ob = torch.empty(M, N * (2 if kw.get("split_out") else 1), device="cuda", dtype=torch.bfloat16) if kw.pop("bf16", False) else None
    ms = t(lambda: ops.gemm(a, w, K=K, split_in=split, bias=bias, out_f32=of, out_bf16=ob, **kw))
    fl = 2.0 * M * N * K * (3 if split else 1)
    print(f"{name:44s} N={N:5d} K={K:5d} split={int(split)}  {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TF/s (issued)")

case("vis_proj (split, f32 out)", 512, 768, True, f32=True)
case("vis_proj plain bf16 in, f32 out", 512, 768, False, f32=True)
case("vis_proj plain bf16 in, bf16 out", 512, 768, False, bf16=True)
case("memory_proj (split, f32 out)", 256, 512, True, f32=True)
case("ca_kv (split, f32 out)", 512, 256, True, f32=True)
case("ca_kv plain, f32 out", 512, 256, False, f32=True)
case("ca_kv plain, bf16 out", 512, 256, False, bf16=True)
case("ca_kv (split, f32 out) bn=128", 512, 256, True, f32=True, block_n=128)
case("fusion 1x1 (bf16, relu, f32+bf16 out)", 256, 1536, False, f32=True, bf16=True, act="relu")
case("fusion 1x1 (bf16, relu, bf16 out)", 256, 1536, False, bf16=True, act="relu")

#!/usr/bin/env python
"""CUDA-event timings of every hot kernel of the forward at the bench shape (ViT-B/16, B = 16, 512x1024: M = 32784 token rows,
32768 pixels), each launched alone through the C ABI in a loop (warm L2 is unavoidable for the small ones; the big ones stream
far more than the 126 MB L2).  Prints one line per kernel with the roofline figure that bounds it (TF/s or GB/s).
usage: python scripts/time_hot_kernels.py [filter-substring]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denseclip_vit_multimodal_b200 import models as Mo  # noqa: E402
from denseclip_vit_multimodal_b200 import ops  # noqa: E402

flt = sys.argv[1] if len(sys.argv) > 1 else ""
B, Nt, D = 16, 2049, 768
M = B * Nt
P = 2048
dev = "cuda"


def t(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best


def report(name, ms, flops=None, nbytes=None):
    s = f"{name:58s} {ms * 1e3:8.1f} us"
    if flops:
        s += f"  {flops / ms / 1e9:8.1f} TF/s ({flops / ms / 1e9 / 1674.0 * 100:4.1f}% of 1674)"
    if nbytes:
        s += f"  {nbytes / ms / 1e6:8.1f} GB/s ({nbytes / ms / 1e6 / 6541.8 * 100:4.1f}% of 6542)"
    print(s, flush=True)


def want(name):
    return flt in name


g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731

if want("gemm"):
    x = rnd(M, D)
    h = rnd(M, D).bfloat16()
    for name, N, K, kw in (("gemm qkv (bias, bf16 out, TMA store)", 2304, 768, dict(want_bf16=True)),
                           ("gemm c_fc (bias + QuickGELU, bf16 out)", 3072, 768, dict(want_bf16=True, act="quickgelu")),
                           ("gemm out_proj (bias + fp32 residual in place)", 768, 768, dict(residual=x, out_f32=x)),
                           ("gemm c_proj (bias + residual in place + bf16 tap)", 768, 3072, dict(residual=x, out_f32=x, out_bf16=torch.empty(M, D, device=dev, dtype=torch.bfloat16)))):
        a = rnd(M, K).bfloat16()
        w = (rnd(N, K) * K ** -0.5).bfloat16()
        b = rnd(N)
        ms = t(lambda: ops.gemm(a, w, bias=b, **kw))
        report(name, ms, flops=2.0 * M * N * K)
        del a, w

if want("conv"):
    G, C, N = 12, 768, 128
    taps = (rnd(G, B, 1 + P, C) * 0.5).bfloat16()
    w_all = (rnd(G * N, 9 * C) * (9 * C) ** -0.5).bfloat16()
    bias = rnd(G * N)
    cat = torch.empty(B * P, G * N, dtype=torch.bfloat16, device=dev)
    a = taps[0, :, 1:, :]
    a2 = a.as_strided((B * P, C), (a.stride(1), 1), a.storage_offset())
    ms = t(lambda: ops.gemm(a2, w_all, K=9 * C, bias=bias, act="relu", out_bf16=cat, M=B * P, block_n=N,
                            conv=dict(C=C, gw=64, gh=32, B=B, a_bs=taps.stride(1), G=G, a_gs=taps.stride(0))), n=10)
    report("conv3x3 grouped neck (12 taps, 768->128, ReLU, bf16 out)", ms, flops=2.0 * B * P * 9 * C * N * G)
    wf = (rnd(256, 1536) * 1536 ** -0.5).bfloat16()
    fb = rnd(256)
    ms = t(lambda: ops.gemm(cat, wf, bias=fb, act="relu", want_f32=True, want_bf16=True))
    report("conv1x1 fusion (1536->256, ReLU, f32 + bf16 out)", ms, flops=2.0 * B * P * 1536 * 256)
    del taps, cat

if want("layernorm"):
    x = rnd(M, D)
    gm, bt = rnd(D), rnd(D)
    ob = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    ms = t(lambda: ops.layernorm(x, gm, bt, out_bf16=ob))
    report("layernorm (fp32 in, bf16 out)", ms, nbytes=M * D * 6)

if want("attention"):
    qkv = (rnd(B, Nt, 3 * D) * 2).bfloat16()
    out = torch.empty(B, Nt, D, dtype=torch.bfloat16, device=dev)
    ms = t(lambda: ops.attention(qkv, qkv, qkv, B=B, H=12, Nq=Nt, Nk=Nt, q_col0=0, k_col0=D, v_col0=2 * D, scale=0.125, out=out))
    report("flash attention (bf16)", ms, flops=4.0 * B * 12 * Nt * Nt * 64)
    sp = ops.split_bf16((rnd(B, Nt, 3 * D) * 2).view(B * Nt, 3 * D)).view(B, Nt, 6 * D)
    out2 = torch.empty(B, Nt, 2 * D, dtype=torch.bfloat16, device=dev)
    ms = t(lambda: ops.attention_split(sp, sp, sp, B=B, H=12, Nq=Nt, Nk=Nt, q_col0=0, k_col0=D, v_col0=2 * D, lo_off=3 * D, scale=0.125,
                                       out=out2, out_lo_off=D), n=10)
    report("flash attention (fp32-class split; algorithmic flops)", ms, flops=4.0 * B * 12 * Nt * Nt * 64)
    del qkv, sp

if want("score"):
    V = rnd(B, Nt, 512)
    text = rnd(B, 19, 512)
    ms = t(lambda: ops.score_map(V, 1, P, text))
    report("score map (normalise x2 + 19-class einsum)", ms, nbytes=B * P * 512 * 4 + B * 19 * P * 4)

if want("upsample"):
    lr = rnd(B, P, 20)
    ms = t(lambda: ops.upsample_bilinear(lr, (512, 1024), tokens_hw=(32, 64), channels=19))
    report("upsample seg logits (19 ch, x16)", ms, nbytes=B * 19 * 512 * 1024 * 4)
    ms = t(lambda: ops.upsample_argmax(lr, (512, 1024), tokens_hw=(32, 64), channels=19))
    report("upsample + argmax (uint8 map)", ms, nbytes=B * 512 * 1024)
    d = rnd(B, P, 4)
    ms = t(lambda: ops.upsample_bilinear(d, (512, 1024), tokens_hw=(32, 64), channels=1))
    report("upsample depth (1 ch)", ms, nbytes=B * 512 * 1024 * 4)

if want("patch"):
    img = rnd(B, 3, 512, 1024)
    ms = t(lambda: ops.im2col_patches(img, 16))
    report("im2col patches (fp32 image -> bf16 [M, 768])", ms, nbytes=B * 3 * 512 * 1024 * 4 + B * P * 768 * 2)

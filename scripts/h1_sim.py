#!/usr/bin/env python
"""H1 design study (CPU only, test infrastructure): how many of the LAST k ViT blocks must run in fp32-class precision
for the bf16 path to reach >= 99.9% per-pixel argmax agreement on the score map?

The oracle's ViT is re-run with bf16 rounding emulated at exactly the points where the CUDA path rounds (weights, LayerNorm
output, fused QKV, un-normalised P, attention output, QuickGELU output; fp32 residual stream, fp32 accumulation), layer by
layer, and compared with the plain fp32 oracle.  Nothing here is shipped; the script only decides the hybrid mode's shape.

    python scripts/h1_sim.py [--images 2] [--gamma 1e-4] [--height 512 --width 1024]
"""
import argparse
import copy
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import denseclip_oracle as O  # noqa: E402


def r16(t):
    return t.to(torch.bfloat16).float()


def split3(a, w):
    """3-pass split-bf16 product a @ w.T (hi*hi + lo*hi + hi*lo) with fp32 accumulation."""
    ah, wh = r16(a), r16(w)
    al, wl = r16(a - ah), r16(w - wh)
    return ah @ wh.t() + al @ wh.t() + ah @ wl.t()


def block(x, sd, p, heads, mode):
    """x [B, N, D] fp32 residual stream.  mode: 'fp32' | 'bf16' | 'split' (3-pass GEMMs + fp32 attention) |
    'splitA' (activations hi+lo, weights bf16) | 'bf16_attn32' (bf16 GEMMs, fp32 attention)."""
    B, N, D = x.shape
    hd = D // heads

    def lin(a, wk, bk):
        w, b = sd[wk], sd[bk]
        if mode == 'fp32':
            return a @ w.t() + b
        if mode == 'split':
            return split3(a, w) + b
        if mode == 'splitA':
            ah = r16(a)
            return ah @ r16(w).t() + r16(a - ah) @ r16(w).t() + b
        return r16(a) @ r16(w).t() + b

    h = F.layer_norm(x, (D,), sd[p + '.ln_1.weight'], sd[p + '.ln_1.bias'], 1e-5)
    qkv = lin(h, p + '.attn.in_proj_weight', p + '.attn.in_proj_bias')
    attn_bf16 = mode in ('bf16', 'splitA')
    if attn_bf16:
        qkv = r16(qkv)
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, N, heads, hd).transpose(1, 2)
    k = k.reshape(B, N, heads, hd).transpose(1, 2)
    v = v.reshape(B, N, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(2, 3)) * hd ** -0.5
    m = s.max(dim=-1, keepdim=True).values
    pexp = torch.exp(s - m)
    l = pexp.sum(-1, keepdim=True)
    o = (r16(pexp) if attn_bf16 else pexp) @ v / l
    o = o.transpose(1, 2).reshape(B, N, D)
    x = x + lin(o, p + '.attn.out_proj.weight', p + '.attn.out_proj.bias')
    h = F.layer_norm(x, (D,), sd[p + '.ln_2.weight'], sd[p + '.ln_2.bias'], 1e-5)
    g = O.quick_gelu(lin(h, p + '.mlp.c_fc.weight', p + '.mlp.c_fc.bias'))
    return x + lin(g, p + '.mlp.c_proj.weight', p + '.mlp.c_proj.bias')


def vit_last(sd, bcfg, img, modes, patch_mode='bf16'):
    """Final-layer feature map [B, D, gh, gw] with per-layer precision modes."""
    ps, layers, heads = bcfg['patch_size'], bcfg['layers'], bcfg['heads']
    w = sd['backbone.conv1.weight']
    if patch_mode == 'bf16':
        x = F.conv2d(r16(img), r16(w), stride=ps)
    else:
        x = F.conv2d(img, w, stride=ps)
    B, D, gh, gw = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd['backbone.class_embedding'].expand(B, 1, -1), x], dim=1)
    x = x + O.interpolate_pos_encoding(sd['backbone.positional_embedding'], x.shape[1], gh, gw)
    x = O.layer_norm(x, sd, 'backbone.ln_pre')
    for i in range(layers):
        x = block(x, sd, f'backbone.transformer.resblocks.{i}', heads, modes[i])
    seq = O.layer_norm(x, sd, 'backbone.ln_post')
    return seq[:, 1:, :].permute(0, 2, 1).reshape(B, D, gh, gw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=1)
    ap.add_argument('--height', type=int, default=512)
    ap.add_argument('--width', type=int, default=1024)
    ap.add_argument('--gamma', type=float, default=None, help="override gamma (reference init 1e-4)")
    ap.add_argument('--ks', default="0,1,2,4,6,8,10,11,12")
    ap.add_argument('--variants', default="split")
    ap.add_argument('--weights', default="bench", choices=["bench", "seeded"])
    ap.add_argument('--first', action='store_true', help="make the FIRST k blocks (and the patch embed) precise instead of the last k")
    args = ap.parse_args()
    import bench
    import denseclip_vit_multimodal_b200 as D
    torch.manual_seed(0)
    cfg = O.model_config("vit_b16", 3)
    model = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()))
    if args.weights == "bench":
        bench.init_uninitialised(model)
        sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    else:
        sd = O.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 11)
    if args.gamma is not None:
        sd['gamma'] = torch.full_like(sd['gamma'], args.gamma)
    img = O.synthetic_images(args.images, args.height, args.width, seed=5)
    L = cfg['backbone']['layers']
    with torch.no_grad():
        ref_last = vit_last(sd, cfg['backbone'], img, ['fp32'] * L, 'fp32')
        _, ref = O.process_features(sd, cfg, [ref_last])
        top2 = ref.topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1]).flatten()
        print(f"fp32 top-2 gap: median {gap.median():.2e}  p1 {gap.quantile(0.01):.2e}  p0.1 {gap.quantile(0.001):.2e}")
        for variant in args.variants.split(','):
            for k in [int(s) for s in args.ks.split(',')]:
                modes = [variant] * k + ['bf16'] * (L - k) if args.first else ['bf16'] * (L - k) + [variant] * k
                last = vit_last(sd, cfg['backbone'], img, modes, 'fp32' if (k == L or (args.first and k > 0)) else 'bf16')
                _, sc = O.process_features(sd, cfg, [last])
                agree = float((sc.argmax(1) == ref.argmax(1)).float().mean())
                err = float((sc - ref).abs().max())
                ferr = float((last - ref_last).abs().max() / ref_last.abs().max())
                print(f"variant={variant:12s} last k={k:2d} precise: score max-abs {err:.2e}  feat rel {ferr:.2e}  argmax agree {agree:.5f}", flush=True)


if __name__ == '__main__':
    main()

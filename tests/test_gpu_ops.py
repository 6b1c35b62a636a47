"""GPU: each native kernel, through the C ABI, against a plain PyTorch fp32 evaluation of the same op on the same
(bf16-rounded where applicable) inputs.  Covers ragged / tail / single-row edge cases."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from denseclip_vit_multimodal_b200 import ops as o
    return o


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (130, 200, 192), (1000, 768, 768), (257, 2304, 768), (4099, 512, 1024)])
def test_gemm_bf16_bias_act_residual(ops, M, N, K):
    a, w, b, r = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    ab, wb = a.bfloat16(), w.bfloat16()
    ref = ab.float() @ wb.float().t() + b
    out, _ = ops.gemm(ab, wb, bias=b, want_f32=True)
    assert rel_err(out, ref) < 2e-5
    _, ob = ops.gemm(ab, wb, bias=b, act="quickgelu", want_bf16=True)
    assert rel_err(ob, ref * torch.sigmoid(1.702 * ref)) < 1.2e-2            # bf16 output rounding + tanh.approx
    x = r.clone()
    ops.gemm(ab, wb, bias=b, residual=x, out_f32=x)                          # in-place residual stream update
    assert rel_err(x, ref + r) < 2e-5
    o32, o16 = ops.gemm(ab, wb, bias=b, act="relu", want_f32=True, want_bf16=True)
    assert rel_err(o32, F.relu(ref)) < 2e-5 and rel_err(o16, F.relu(ref)) < 1e-2


@pytest.mark.parametrize("M,N,K", [(19, 256, 256), (300, 512, 768), (2049, 256, 512)])
def test_gemm_split_bf16_is_fp32_class(ops, M, N, K):
    a, w, b = _rand(M, K, seed=5), _rand(N, K, scale=K ** -0.5, seed=6), _rand(N, seed=7)
    ref = (a.double() @ w.double().t() + b.double()).float()
    out, _ = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), split_in=True, bias=b, want_f32=True)
    assert rel_err(out, ref) < 3e-5
    _, osplit = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), split_in=True, bias=b, act="gelu", want_bf16=True, split_out=True)
    g = F.gelu(ref)
    assert rel_err(osplit[:, :N].float() + osplit[:, N:].float(), g) < 3e-5


@pytest.mark.parametrize("M,D", [(1, 128), (33, 256), (4098, 768), (77, 1024)])
def test_layernorm(ops, M, D):
    x, g, b = _rand(M, D, scale=3, seed=8) + 0.5, _rand(D, seed=9) * 0.1 + 1, _rand(D, seed=10) * 0.1
    ref = F.layer_norm(x, (D,), g, b, 1e-5)
    of, ob = ops.layernorm(x, g, b, want_f32=True, want_bf16=True, split=True)
    assert rel_err(of, ref) < 1e-5
    assert rel_err(ob[:, :D].float() + ob[:, D:].float(), ref) < 2e-5


# N % 256 in 1..4: the last query block takes the kernel's CUDA-core tail path (attn_tail_rows); 5: tensor-core path again
@pytest.mark.parametrize("B,H,N", [(1, 1, 1), (2, 2, 129), (1, 4, 1025), (2, 12, 2049), (1, 2, 2629), (2, 3, 260), (1, 2, 261),
                                   (1, 2, 515)])
def test_flash_attention_vs_sdpa(ops, B, H, N):
    D = H * 64
    qkv = _rand(B, N, 3 * D, scale=1.5, seed=11).bfloat16()
    out = torch.empty(B, N, D, dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, qkv, qkv, B=B, H=H, Nq=N, Nk=N, q_col0=0, k_col0=D, v_col0=2 * D, scale=0.125, out=out)
    q, k, v = (t.float().view(B, N, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, D)
    assert rel_err(out, ref) < 1.5e-2
    tail = N % 256
    if tail:  # the rows of the ragged last block on their own (they are few: a global max would hide them)
        assert rel_err(out[:, N - tail:], ref[:, N - tail:]) < 1.5e-2


def test_flash_attention_item_boundaries_bit_exact(ops):
    """Persistent kernel: with 288 regular work items every CTA walks two items (shared TMEM / barriers / KV ring, Q double
    buffer, output tile staged in the finished Q buffer).  Each image computed alone (48 items, one per CTA) must give the
    same bits: a row's arithmetic does not depend on which CTA / item slot serves it."""
    B, H, N = 6, 12, 1025
    D = H * 64
    qkv = _rand(B, N, 3 * D, scale=1.5, seed=21).bfloat16()
    out = torch.empty(B, N, D, dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, qkv, qkv, B=B, H=H, Nq=N, Nk=N, q_col0=0, k_col0=D, v_col0=2 * D, scale=0.125, out=out)
    for b in (0, 3, 5):
        one = torch.empty(1, N, D, dtype=torch.bfloat16, device="cuda")
        x = qkv[b:b + 1].contiguous()
        ops.attention(x, x, x, B=1, H=H, Nq=N, Nk=N, q_col0=0, k_col0=D, v_col0=2 * D, scale=0.125, out=one)
        assert torch.equal(one[0], out[b])
    q, k, v = (t.float().view(B, N, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, D)
    assert rel_err(out, ref) < 1.5e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("causal,Nq,Nk", [(False, 19, 19), (False, 19, 2049), (True, 22, 22), (False, 1, 300), (False, 19, 1025),
                                          (False, 5, 4100), (False, 1, 2049)])   # the long-key cases run key-split + combine
def test_attention_small(ops, dtype, causal, Nq, Nk):
    B, H = 3, 4
    D = H * 64
    if causal:
        x = _rand(B, Nq, 3 * D, seed=12).to(dtype)
        q, k, v, cols = x, x, x, (0, D, 2 * D)
    else:
        q, kv = _rand(B, Nq, D, seed=13).to(dtype), _rand(B, Nk, 2 * D, seed=14).to(dtype)
        k, v, cols = kv, kv, (0, 0, D)
    out = torch.empty(B, Nq, D, dtype=torch.float32, device="cuda")
    ops.attention_small(q, k, v, B=B, H=H, q_first=0, q_count=Nq, Nk=Nk, q_col0=cols[0], k_col0=cols[1], v_col0=cols[2],
                        scale=0.125, out=out, causal=causal)
    qq = q[..., cols[0]:cols[0] + D].float().view(B, Nq, H, 64).transpose(1, 2)
    kk = k[..., cols[1]:cols[1] + D].float().view(B, Nk, H, 64).transpose(1, 2)
    vv = v[..., cols[2]:cols[2] + D].float().view(B, Nk, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qq, kk, vv, is_causal=causal).transpose(1, 2).reshape(B, Nq, D)
    assert rel_err(out, ref) < 2e-5
    # hi|lo bf16 output written by the kernel == split of the fp32 result
    out2 = torch.empty(B, Nq, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.attention_small(q, k, v, B=B, H=H, q_first=0, q_count=Nq, Nk=Nk, q_col0=cols[0], k_col0=cols[1], v_col0=cols[2],
                        scale=0.125, out=out2, causal=causal, out_split_off=D)
    hi = out.bfloat16()
    assert torch.equal(out2[..., :D], hi) and torch.equal(out2[..., D:], (out - hi.float()).bfloat16())


@pytest.mark.parametrize("g0,gh,gw", [(14, 32, 64), (2, 2, 4), (14, 14, 14), (7, 3, 5)])
def test_posemb_interp(ops, g0, gh, gw):
    D = 256
    pos = _rand(1 + g0 * g0, D, seed=15)
    got = ops.posemb_interp(pos, g0, gh, gw)
    ref = F.interpolate(pos[1:].reshape(1, g0, g0, D).permute(0, 3, 1, 2), size=(gh, gw), mode="bilinear", align_corners=False)
    ref = torch.cat([pos[:1], ref.permute(0, 2, 3, 1).reshape(-1, D)])
    assert rel_err(got, ref) < 1e-6


def test_layout_mean_score_upsample(ops):
    B, gh, gw, C, K = 2, 8, 16, 128, 19
    tok = _rand(B, 1 + gh * gw, C, seed=16)
    nchw = ops.tap_nchw(tok, gh, gw)
    ref_nchw = tok[:, 1:].permute(0, 2, 1).reshape(B, C, gh, gw)
    assert torch.equal(nchw, ref_nchw)
    back, _ = ops.nchw_to_tokens(nchw, row_off=1, rows=1 + gh * gw)
    assert torch.equal(back[:, 1:], tok[:, 1:])
    assert rel_err(ops.token_mean(tok, 1, gh * gw), tok[:, 1:].mean(1)) < 1e-6
    text = _rand(B, K, C, seed=17)
    sm = ops.score_map(tok, 1, gh * gw, text).view(B, K, gh, gw)
    ref = torch.einsum("bchw,bkc->bkhw", F.normalize(ref_nchw, dim=1), F.normalize(text, dim=2))
    assert float((sm - ref).abs().max()) < 1e-6
    low = _rand(B, gh * gw, 20, seed=18)
    up = ops.upsample_bilinear(low, (gh * 16, gw * 16), tokens_hw=(gh, gw), channels=K)
    ref_up = F.interpolate(low[..., :K].permute(0, 2, 1).reshape(B, K, gh, gw), size=(gh * 16, gw * 16), mode="bilinear",
                           align_corners=False)
    assert rel_err(up, ref_up) < 1e-6
    assert torch.equal(ops.upsample_argmax(low, (gh * 16, gw * 16), tokens_hw=(gh, gw), channels=K).long(), ref_up.argmax(1))
    assert rel_err(ops.upsample_bilinear(ref_nchw, (13 * 4, 20)), F.interpolate(ref_nchw, size=(52, 20), mode="bilinear", align_corners=False)) < 1e-6


@pytest.mark.parametrize("K,ld", [(1, 4), (5, 8), (19, 20), (19, 24), (21, 24)])   # K4 = 1, 2, 5, 5 + padding, generic fallback
@pytest.mark.parametrize("hw,HW", [((5, 7), (37, 52)), ((3, 4), (48, 64)), ((9, 6), (9, 8))])   # ragged ratio / x16 / ~identity
def test_upsample_strips_ragged(ops, K, ld, hw, HW):
    """The strip kernels (4 px x 16 rows per thread, source rows cached while (y0, y1) is unchanged) against
    F.interpolate for sizes that are not a multiple of the strip height and scales that are not integers."""
    (h, w), (H, W) = hw, HW
    B = 2
    low = _rand(B, h * w, ld, seed=31 + K)
    ref = F.interpolate(low[..., :K].permute(0, 2, 1).reshape(B, K, h, w), size=(H, W), mode="bilinear", align_corners=False)
    up = ops.upsample_bilinear(low, (H, W), tokens_hw=(h, w), channels=K)
    assert up.shape == ref.shape and rel_err(up, ref) < 1e-6
    am = ops.upsample_argmax(low, (H, W), tokens_hw=(h, w), channels=K).long()
    # ties / last-bit differences: accept a different class only where the two top scores are within fp32 rounding
    bad = am != ref.argmax(1)
    if bad.any():
        top2 = ref.topk(2, dim=1).values if K > 1 else None
        assert K > 1 and float((top2[:, 0] - top2[:, 1])[bad].abs().max()) < 1e-6


@pytest.mark.parametrize("gh,gw", [(8, 16), (4, 8), (2, 64)])   # implicit-conv TMA path / gather fallback / one-row box
def test_conv3x3_implicit_gemm(ops, gh, gw):
    from denseclip_vit_multimodal_b200 import models as M
    B, C, N = 2, 128, 64
    x = _rand(B, C, gh, gw, seed=19)
    w, b = _rand(N, C, 3, 3, scale=(9 * C) ** -0.5, seed=20), _rand(N, seed=21)
    tok, tokb = ops.nchw_to_tokens(x, row_off=1, rows=1 + gh * gw, f32=True, bf16=True)
    wp = ops.pack_weight(M.conv3x3_weight_to_gemm(w), False)
    y, _ = M.conv3x3_tokens(tokb, 1, gh, gw, C, wp, split_in=False, bias=b, act="relu", want_f32=True)
    ref = F.relu(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1)).permute(0, 2, 3, 1).reshape(-1, N)
    assert rel_err(y, ref) < 2e-5


def test_ops_reject_bad_inputs(ops):
    from denseclip_vit_multimodal_b200 import DclipError
    a = torch.zeros(4, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(DclipError):
        ops.gemm(a, torch.zeros(6, 64, dtype=torch.bfloat16, device="cuda"), want_f32=True)      # N % 4 != 0
    with pytest.raises(DclipError):
        ops.gemm(a.float(), a, want_f32=True)                                                     # wrong dtype
    with pytest.raises(DclipError):
        ops.layernorm(torch.zeros(2, 100, device="cuda"), torch.ones(100, device="cuda"), torch.zeros(100, device="cuda"), want_f32=True)


@pytest.mark.parametrize("K,tdtype", [(19, torch.int64), (19, torch.uint8), (5, torch.int64), (64, torch.uint8)])
def test_eval_stats_kernel(ops, K, tdtype):
    """Native evaluation statistics against the torch definitions: confusion matrix with ignore_index
    (bit-exact, integer) and masked squared depth error / count (float64)."""
    from denseclip_vit_multimodal_b200.distributed import confusion_matrix, reduce_eval_stats, shard_eval_stats
    g = torch.Generator().manual_seed(40 + K)
    B, H, W = 3, 37, 52
    pred = torch.randint(0, K, (B, H, W), generator=g).to(torch.uint8).cuda()
    target = torch.randint(0, K, (B, H, W), generator=g)
    target[torch.rand(B, H, W, generator=g) < 0.1] = 255        # ignore_index
    target = target.to(tdtype).cuda()
    dp = torch.randn(B, 1, H, W, generator=g).cuda()
    dg = (torch.rand(B, 1, H, W, generator=g) * 80).cuda()
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.7).cuda()
    conf, ds = ops.eval_stats(pred, target, K, 255, dp, dg, mask)
    assert torch.equal(conf, confusion_matrix(pred, target, K, 255))
    err = (dp.double() - dg.double())[mask]
    assert float(ds[1]) == float(mask.sum())
    assert abs(float(ds[0]) - float((err * err).sum())) <= 1e-9 * float((err * err).sum())
    # accumulation into caller buffers, segmentation-only and depth-only calls, unmasked depth
    ops.eval_stats(pred, target, K, 255, conf=conf, depth_stats=ds)
    assert torch.equal(conf, 2 * confusion_matrix(pred, target, K, 255)) and float(ds[1]) == float(mask.sum())
    _, ds2 = ops.eval_stats(depth_pred=dp, depth_gt=dg)
    assert float(ds2[1]) == dp.numel()
    # all pixels ignored / empty mask leave the buffers untouched
    c0, d0 = ops.eval_stats(pred, torch.full_like(target, 255), K, 255, dp, dg, torch.zeros_like(mask))
    assert int(c0.sum()) == 0 and float(d0.sum()) == 0.0
    c, se, n = shard_eval_stats(pred, target, K, 255, dp, dg, mask)
    _, miou, acc, rmse = reduce_eval_stats(c, se, n)
    tp = conf.diag().double() / 2
    assert abs(acc - float(tp.sum() / (conf.sum() / 2))) < 1e-12 and 0.0 <= miou <= 1.0
    assert abs(rmse - float(err.pow(2).mean().sqrt())) < 1e-9
    with pytest.raises(Exception):
        ops.eval_stats(pred, target.float(), K)


@pytest.mark.parametrize("ps,H,W", [(16, 64, 96), (16, 40, 52), (14, 50, 76), (8, 32, 40)])   # vector path, ragged, scalar path (ps 14), ps 8
@pytest.mark.parametrize("split", [False, True])
def test_im2col_patches(ops, ps, H, W, split):
    """Patch-embed operand gather against F.unfold (K order = conv1.weight flattening; trailing pixels dropped)."""
    img = _rand(2, 3, H, W, seed=50 + ps)
    got = ops.im2col_patches(img, ps, split=split)
    gh, gw, K = H // ps, W // ps, 3 * ps * ps
    ref = F.unfold(img[:, :, :gh * ps, :gw * ps], kernel_size=ps, stride=ps).transpose(1, 2).reshape(-1, K)
    hi = ref.bfloat16()
    assert torch.equal(got[:, :K], hi)
    if split:
        assert torch.equal(got[:, K:], (ref - hi.float()).bfloat16())


# ---------------------------------------------------------------------------------------------------------------------
# fp32-class (hi|lo split) tensor-core attention: against an fp64 evaluation of softmax(QK^T/8)V on the SAME fp32 inputs.
# (2, 12, 2049): the BASELINE ViT-B/16 sequence (17th KV tile holds one key, ninth query block one row); 2629: ViT-L/14.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,N", [(1, 1, 1), (2, 2, 129), (1, 2, 300), (2, 12, 2049), (1, 2, 2629), (1, 3, 257)])
def test_attention_split_is_fp32_class(ops, B, H, N):
    D = H * 64
    qkv = _rand(B, N, 3 * D, scale=1.5, seed=31)
    sp = ops.split_bf16(qkv.view(B * N, 3 * D)).view(B, N, 6 * D)            # [q k v]_hi | [q k v]_lo
    out = torch.empty(B, N, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.attention_split(sp, sp, sp, B=B, H=H, Nq=N, Nk=N, q_col0=0, k_col0=D, v_col0=2 * D, lo_off=3 * D, scale=0.125, out=out,
                        out_lo_off=D)
    got = out[..., :D].float() + out[..., D:].float()
    q, k, v = (t.double().view(B, N, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    ref = (torch.softmax(q @ k.transpose(2, 3) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(B, N, D).float()
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 5e-5      # hi+lo operands carry 16 mantissa bits (2^-17 = 7.6e-6 each), ex2.approx 2^-22
    tail = N % 256
    if tail:
        assert rel_err(got[:, N - tail:], ref[:, N - tail:]) < 5e-5


# ---------------------------------------------------------------------------------------------------------------------
# Production tile paths at the BASELINE shape M = 16 x 2049 = 32784 rows: CTA-pair (cta_group::2) 256x256 tiles, the
# 256x192 residual tile, the TMA-store epilogue, the second bf16 output (feature tap) and the patch-embed row remap.
# None of these is reached by the small shapes above (pair mode needs >= 148 pair units).
# ---------------------------------------------------------------------------------------------------------------------
M_PROD = 16 * 2049


@pytest.mark.parametrize("N,K,act", [(2304, 768, None), (3072, 768, "quickgelu")])
def test_gemm_production_tma_store_pair(ops, N, K, act):
    a, w, b = _rand(M_PROD, K, seed=41).bfloat16(), _rand(N, K, scale=K ** -0.5, seed=42).bfloat16(), _rand(N, seed=43)
    _, ob = ops.gemm(a, w, bias=b, act=act, want_bf16=True)
    for r0 in (0, 16000, M_PROD - 1200):    # fp32 reference in row slabs (the full [M, N] fp32 product is 400 MB)
        ref = a[r0:r0 + 1200].float() @ w.float().t() + b
        if act:
            ref = ref * torch.sigmoid(1.702 * ref)
        assert rel_err(ob[r0:r0 + 1200], ref) < 1.2e-2


@pytest.mark.parametrize("K,tap", [(768, False), (3072, False), (3072, True)])
def test_gemm_production_residual_pair_and_bn192(ops, K, tap):
    """out-proj (K = 768 -> 256x192 pair tiles) and c_proj (K = 3072 -> 256x256 pair tiles, optional bf16 tap output):
    fp32 residual stream updated in place."""
    N = 768
    a, w, b = _rand(M_PROD, K, seed=44).bfloat16(), _rand(N, K, scale=K ** -0.5, seed=45).bfloat16(), _rand(N, seed=46)
    x = _rand(M_PROD, N, seed=47)
    x0 = x.clone()
    tap_out = torch.empty(M_PROD, N, dtype=torch.bfloat16, device="cuda") if tap else None
    ops.gemm(a, w, bias=b, residual=x, out_f32=x, out_bf16=tap_out)
    ref = a.float() @ w.float().t() + b + x0
    assert rel_err(x, ref) < 2e-5
    if tap:
        assert rel_err(tap_out, ref) < 1e-2
        assert torch.equal(tap_out, x.bfloat16())      # the tap is exactly the rounded residual stream


def test_gemm_production_patch_embed_remap(ops):
    """patch rows m = b*P + p -> token rows b*Ntok + 1 + p, + positional embedding row 1 + p (models.py:546-556)."""
    Bb, P, Nt, D, K = 16, 2048, 2049, 768, 768
    a, w = _rand(Bb * P, K, seed=48).bfloat16(), _rand(D, K, scale=K ** -0.5, seed=49).bfloat16()
    pos = _rand(Nt, D, seed=50)
    x = torch.zeros(Bb * Nt, D, device="cuda")
    from denseclip_vit_multimodal_b200 import _lib
    import ctypes as C
    g = _lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw = a.data_ptr(), K, w.data_ptr(), K
    g.M, g.N, g.K, g.out_scale = Bb * P, D, K, 1.0
    g.residual, g.ldr, g.res_mod, g.remap_P, g.remap_Nt = pos.data_ptr(), D, 1, P, Nt
    g.out_f32, g.ldc = x.data_ptr(), D
    ops._call(a, _lib.lib().dclip_gemm, C.byref(g), ops._stream(a))
    ref = (a.float() @ w.float().t()).view(Bb, P, D) + pos[1:]
    got = x.view(Bb, Nt, D)
    assert rel_err(got[:, 1:], ref) < 2e-5
    assert float(got[:, 0].abs().max()) == 0.0          # CLS rows are not touched by the GEMM


def test_gemm_production_split_qkv(ops):
    """fp32-class fused QKV at the production shape: 3-pass split product, bf16 hi|lo output (pair tiles)."""
    M, N, K = 8 * 2049, 2304, 768
    a, w, b = _rand(M, K, seed=51), _rand(N, K, scale=K ** -0.5, seed=52), _rand(N, seed=53)
    _, o = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), split_in=True, bias=b, want_bf16=True, split_out=True)
    for r0 in (0, M - 900):
        ref = (a[r0:r0 + 900].double() @ w.double().t() + b.double()).float()
        assert rel_err(o[r0:r0 + 900, :N].float() + o[r0:r0 + 900, N:].float(), ref) < 3e-5


def test_two_streams_small_attention_scratch_is_per_stream(ops):
    """the key-split scratch of the few-query attention is per (handle, stream) and never freed while the
    handle lives: two streams running different problem sizes concurrently must not disturb each other."""
    B, H, D = 4, 4, 256
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    q1, kv1 = _rand(B, 19, D, seed=61), _rand(B, 2049, 2 * D, seed=62)
    q2, kv2 = _rand(2 * B, 19, D, seed=63), _rand(2 * B, 4097, 2 * D, seed=64)
    o1 = torch.empty(B, 19, D, device="cuda")
    o2 = torch.empty(2 * B, 19, D, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(s1):
            ops.attention_small(q1, kv1, kv1, B=B, H=H, q_first=0, q_count=19, Nk=2049, q_col0=0, k_col0=0, v_col0=D, scale=0.125, out=o1)
        with torch.cuda.stream(s2):
            ops.attention_small(q2, kv2, kv2, B=2 * B, H=H, q_first=0, q_count=19, Nk=4097, q_col0=0, k_col0=0, v_col0=D, scale=0.125, out=o2)
    torch.cuda.synchronize()
    for q, kv, o, n in ((q1, kv1, o1, 2049), (q2, kv2, o2, 4097)):
        bb = q.shape[0]
        qq = q.view(bb, 19, H, 64).transpose(1, 2)
        kk = kv[..., :D].reshape(bb, n, H, 64).transpose(1, 2)
        vv = kv[..., D:].reshape(bb, n, H, 64).transpose(1, 2)
        ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(bb, 19, D)
        assert rel_err(o, ref) < 2e-5


def test_grouped_conv3x3_production_shape_pair_tiles(ops):
    """The neck's 12 x (conv3x3 768 -> 128 + folded BN + ReLU) as ONE grouped launch at the BASELINE shape (B = 16, 32 x 64
    grid): the CTA-pair implicit-GEMM tiles (256 pixels x 128 filters per SM pair, 5-D TMA gather, TMA-store epilogue) are
    only reached at this size.  Checked against F.conv2d on the same bf16-rounded operands for three of the taps."""
    from denseclip_vit_multimodal_b200 import models as M
    G, B, gh, gw, C, N = 12, 16, 32, 64, 768, 128
    P = gh * gw
    taps = (_rand(G, B, 1 + P, C, seed=71) * 0.5).bfloat16()                  # token-major, CLS row first (as the encoder writes them)
    w = _rand(G, N, C, 3, 3, scale=(9 * C) ** -0.5, seed=72)
    b = _rand(G * N, seed=73)
    w_all = torch.cat([ops.pack_weight(M.conv3x3_weight_to_gemm(w[g]), False) for g in range(G)], 0).contiguous()
    cat = torch.empty(B * P, G * N, dtype=torch.bfloat16, device="cuda")
    a = taps[0, :, 1:, :]
    a2 = a.as_strided((B * P, C), (a.stride(1), 1), a.storage_offset())
    ops.gemm(a2, w_all, K=9 * C, bias=b, act="relu", out_bf16=cat, M=B * P, block_n=N,
             conv=dict(C=C, gw=gw, gh=gh, B=B, a_bs=taps.stride(1), G=G, a_gs=taps.stride(0)))
    for g in (0, 5, 11):
        x = taps[g, :, 1:, :].float().reshape(B, gh, gw, C).permute(0, 3, 1, 2)
        ref = F.relu(F.conv2d(x, w[g].bfloat16().float(), b[g * N:(g + 1) * N], padding=1)).permute(0, 2, 3, 1).reshape(B * P, N)
        assert rel_err(cat[:, g * N:(g + 1) * N], ref) < 1e-2                 # bf16 output rounding


@pytest.mark.parametrize("B,P,C,K", [(2, 2048, 512, 19), (1, 2628, 512, 19), (3, 37, 256, 5), (1, 130, 1024, 33), (2, 5, 128, 1)])
def test_score_map_ragged_pixels_and_classes(ops, B, P, C, K):
    """F.normalize x2 + einsum('bchw,bkc->bkhw') (denseclip.py:672-675): 4 pixels per warp and classes in rounds of 8, so pixel
    counts that are not a multiple of 4 / 32 and class counts that are not a multiple of 8 hit the masked paths."""
    tok = _rand(B, 1 + P, C, seed=81) * 3
    text = _rand(B, K, C, seed=82)
    sm = ops.score_map(tok, 1, P, text)
    ref = torch.einsum("bpc,bkc->bkp", F.normalize(tok[:, 1:], dim=2), F.normalize(text, dim=2))
    assert sm.shape == (B, K, P)
    assert float((sm - ref).abs().max()) < 2e-6

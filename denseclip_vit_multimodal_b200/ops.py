"""Tensor-level wrappers over the C ABI.  PyTorch is used here only for device memory and the current stream; every
op below launches hand-written sm_100a kernels from libdenseclip_b200.so and raises ``DclipError`` on any failure
(there is no eager/CPU fallback)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import DclipError, GemmArgs

ACT_NONE, ACT_QUICKGELU, ACT_QUICKGELU_PRECISE, ACT_GELU_ERF, ACT_RELU = range(5)
_ACT_NAMES = {None: ACT_NONE, "none": ACT_NONE, "quickgelu": ACT_QUICKGELU, "quickgelu_precise": ACT_QUICKGELU_PRECISE,
              "gelu": ACT_GELU_ERF, "relu": ACT_RELU}


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise DclipError("denseclip_b200 ops need CUDA tensors (no CPU fallback); got a tensor on %s" % t.device)
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(t: torch.Tensor):
    _dev(t)   # (CPU tensors fail here with DclipError, before torch is asked for a CUDA stream)
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if t.dtype != dtype:
        raise DclipError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise DclipError(f"{name}: last dim must be contiguous")
    return t


def _call(t, fn, *args):
    h = _lib.handle(_dev(t))
    _lib.check(h, fn(h, *args))


# ---------------------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor, scale: float = 1.0, pad_cols_to: int = 1) -> torch.Tensor:
    """fp32 [R, Cc] -> bf16 [R, 2*Cp] = hi | lo (Cp = Cc rounded up to pad_cols_to; padding is zero)."""
    x = _req(x.contiguous(), torch.float32, "x")
    R, Cc = x.shape
    Cp = (Cc + pad_cols_to - 1) // pad_cols_to * pad_cols_to
    out = torch.zeros(R, 2 * Cp, dtype=torch.bfloat16, device=x.device) if Cp != Cc else torch.empty(
        R, 2 * Cp, dtype=torch.bfloat16, device=x.device)
    _call(x, _lib.lib().dclip_cast_bf16, _ptr(x), Cc, _ptr(out), 2 * Cp, R, Cc, 1, Cp, float(scale), _stream(x))
    return out


def cast_bf16(x: torch.Tensor, scale: float = 1.0, pad_cols_to: int = 1) -> torch.Tensor:
    x = _req(x.contiguous(), torch.float32, "x")
    R, Cc = x.shape
    Cp = (Cc + pad_cols_to - 1) // pad_cols_to * pad_cols_to
    out = (torch.zeros if Cp != Cc else torch.empty)(R, Cp, dtype=torch.bfloat16, device=x.device)
    _call(x, _lib.lib().dclip_cast_bf16, _ptr(x), Cc, _ptr(out), Cp, R, Cc, 0, 0, float(scale), _stream(x))
    return out


def pack_weight(w: torch.Tensor, precise: bool, pad_cols_to: int = 8) -> torch.Tensor:
    """Linear-style weight fp32 [N, K] -> bf16 [N, Kp] (or [N, 2*Kp] hi|lo when precise)."""
    w2 = w.detach().reshape(w.shape[0], -1).float().contiguous()
    return split_bf16(w2, pad_cols_to=pad_cols_to) if precise else cast_bf16(w2, pad_cols_to=pad_cols_to)


def gemm(a: torch.Tensor, w: torch.Tensor, *, K: int | None = None, split_in: bool = False, bias=None, act=None,
         out_scale: float = 1.0, residual=None, out_f32=None, out_bf16=None, want_f32: bool = False,
         want_bf16: bool = False, split_out: bool = False, block_n: int = 0, M: int | None = None, conv=None):
    """C = A @ W^T (+bias, act, *out_scale, +residual).  a: bf16 [M, K] (split_in: [M, 2K] hi|lo); w: bf16 [N, K(2K)].
    Returns (out_f32, out_bf16) (either may be None).  bf16 output with split_out is [M, 2N] = hi|lo."""
    a = _req(a, torch.bfloat16, "a")
    w = _req(w, torch.bfloat16, "w")
    assert a.dim() == 2 and w.dim() == 2
    M = a.shape[0] if M is None else M
    N = w.shape[0]
    if K is None:
        K = a.shape[1] // 2 if split_in else a.shape[1]
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(M, N, dtype=torch.float32, device=a.device)
    if want_bf16 and out_bf16 is None:
        out_bf16 = torch.empty(M, N * (2 if split_out else 1), dtype=torch.bfloat16, device=a.device)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
    g.M, g.N, g.K, g.split_in = M, N, K, int(split_in)
    g.bias = None if bias is None else _req(bias, torch.float32, "bias").data_ptr()
    g.act = _ACT_NAMES[act] if not isinstance(act, int) else act
    g.out_scale = out_scale
    if residual is not None:
        g.residual, g.ldr = _req(residual, torch.float32, "residual").data_ptr(), residual.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ldc = _req(out_f32, torch.float32, "out_f32").data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        g.out_bf16, g.ldcb = _req(out_bf16, torch.bfloat16, "out_bf16").data_ptr(), out_bf16.stride(0)
        g.split_out, g.split_out_off = int(split_out), N
    g.block_n = block_n
    if conv is not None:  # implicit 3x3 conv: a is token-major [B, rows, ld], conv = dict(C, gh, gw, row0)
        g.conv_C, g.conv_gw, g.conv_gh, g.conv_B = conv["C"], conv["gw"], conv["gh"], conv["B"]
        g.a_bs = conv["a_bs"]
        g.conv_G, g.a_gs = conv.get("G", 0), conv.get("a_gs", 0)
    _call(a, _lib.lib().dclip_gemm, C.byref(g), _stream(a))
    return out_f32, out_bf16


def im2col_patches(img: torch.Tensor, ps: int, *, split: bool = False) -> torch.Tensor:
    """Patch-embed operand gather (Conv2d(3, D, ps, stride=ps), models.py:407, 546-548): fp32 [B, 3, H, W] ->
    bf16 [B*gh*gw, K (2K = hi|lo with split)], K index = c*ps*ps + ky*ps + kx; trailing pixels that do not fill a patch are dropped."""
    img = _req(img.contiguous(), torch.float32, "img")
    B, Cc, H, W = img.shape
    if Cc != 3 or ps % 2:
        raise DclipError("im2col_patches: 3 input channels and an even patch size are required")
    gh, gw, K = H // ps, W // ps, 3 * ps * ps
    out = torch.empty(B * gh * gw, K * (2 if split else 1), dtype=torch.bfloat16, device=img.device)
    _call(img, _lib.lib().dclip_im2col_patches, _ptr(img), B, H, W, ps, _ptr(out), out.stride(0), int(split), K, _stream(img))
    return out


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, *, want_f32=False, want_bf16=False, split=False,
              out_f32=None, out_bf16=None):
    """LayerNorm over the last dim of fp32 [M, D]. bf16 output is [M, D] or, with split, [M, 2D] = hi|lo."""
    x = _req(x, torch.float32, "x")
    assert x.dim() == 2
    M, D = x.shape
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(M, D, dtype=torch.float32, device=x.device)
    if want_bf16 and out_bf16 is None:
        out_bf16 = torch.empty(M, D * (2 if split else 1), dtype=torch.bfloat16, device=x.device)
    _call(x, _lib.lib().dclip_layernorm, _ptr(x), x.stride(0), _ptr(_req(gamma, torch.float32, "gamma")),
          _ptr(_req(beta, torch.float32, "beta")), float(eps), M, D, _ptr(out_f32), 0 if out_f32 is None else out_f32.stride(0),
          _ptr(out_bf16), 0 if out_bf16 is None else out_bf16.stride(0), int(split), D, _stream(x))
    return out_f32, out_bf16


def attention(qkv_q, qkv_k, qkv_v, *, B, H, Nq, Nk, q_col0, k_col0, v_col0, scale, out, q_start=0):
    """Tensor-core flash attention (head_dim 64). q/k/v: bf16 [B, N, ld] views (may be the same fused tensor)."""
    for t in (qkv_q, qkv_k, qkv_v, out):
        _req(t, torch.bfloat16, "attention operand")
        assert t.dim() == 3
    _call(qkv_q, _lib.lib().dclip_attention, _ptr(qkv_q), _ptr(qkv_k), _ptr(qkv_v), qkv_q.stride(1), qkv_k.stride(1),
          qkv_v.stride(1), qkv_q.stride(0), qkv_k.stride(0), qkv_v.stride(0), q_col0, k_col0, v_col0, B, H, Nq, q_start, Nk,
          float(scale), _ptr(out), out.stride(1), out.stride(0), _stream(qkv_q))
    return out


def attention_split(qkv_q, qkv_k, qkv_v, *, B, H, Nq, Nk, q_col0, k_col0, v_col0, lo_off, scale, out, out_lo_off):
    """fp32-class tensor-core flash attention (head_dim 64): operands are bf16 [B, N, ld] views whose rows hold a hi half at
    column col0 + 64h and the matching lo half ``lo_off`` columns further; three tcgen05 passes per product, fp32 softmax.
    ``out`` bf16 [B, Nq, ldo]: hi at column 64h, lo at ``out_lo_off`` + 64h."""
    for t in (qkv_q, qkv_k, qkv_v, out):
        _req(t, torch.bfloat16, "attention operand")
        assert t.dim() == 3
    _call(qkv_q, _lib.lib().dclip_attention_split, _ptr(qkv_q), _ptr(qkv_k), _ptr(qkv_v), qkv_q.stride(1), qkv_k.stride(1),
          qkv_v.stride(1), qkv_q.stride(0), qkv_k.stride(0), qkv_v.stride(0), q_col0, k_col0, v_col0, lo_off, B, H, Nq, Nk,
          float(scale), _ptr(out), out.stride(1), out.stride(0), out_lo_off, _stream(qkv_q))
    return out


def attention_small(q, k, v, *, B, H, q_first, q_count, Nk, q_col0, k_col0, v_col0, scale, out, causal=False,
                    out_split_off=0):
    """fp32 CUDA-core attention for a handful of query rows. q/k/v: [B, N, ld] fp32 or bf16 (all the same dtype)."""
    is_f32 = q.dtype == torch.float32
    for t in (q, k, v):
        _req(t, torch.float32 if is_f32 else torch.bfloat16, "attention_small operand")
        assert t.dim() == 3
    _call(q, _lib.lib().dclip_attention_small, _ptr(q), _ptr(k), _ptr(v), int(is_f32), q.stride(1), k.stride(1), v.stride(1),
          q.stride(0), k.stride(0), v.stride(0), q_col0, k_col0, v_col0, B, H, q_first, q_count, Nk, float(scale), int(causal),
          _ptr(out), int(out.dtype == torch.float32), out.stride(1), out.stride(0), out_split_off, _stream(q))
    return out


def posemb_interp(pos: torch.Tensor, g0: int, gh: int, gw: int) -> torch.Tensor:
    pos = _req(pos.contiguous(), torch.float32, "pos")
    D = pos.shape[1]
    out = torch.empty(1 + gh * gw, D, dtype=torch.float32, device=pos.device)
    _call(pos, _lib.lib().dclip_posemb_interp, _ptr(pos), g0, gh, gw, D, _ptr(out), _stream(pos))
    return out


def tap_nchw(tokens: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """fp32 tokens [B, 1+P, D] -> NCHW fp32 [B, D, gh, gw] (CLS dropped)."""
    tokens = _req(tokens.contiguous(), torch.float32, "tokens")
    B, Ntok, D = tokens.shape
    out = torch.empty(B, D, gh, gw, dtype=torch.float32, device=tokens.device)
    _call(tokens, _lib.lib().dclip_tap_nchw, _ptr(tokens), B, Ntok, D, _ptr(out), _stream(tokens))
    return out


def nchw_to_tokens(x: torch.Tensor, *, row_off: int = 0, rows: int | None = None, f32=True, bf16=False):
    """NCHW fp32 [B, C, h, w] -> token-major [B, rows, C] (pixel p at row row_off + p); other rows are left zero."""
    x = _req(x.contiguous(), torch.float32, "x")
    B, Cc, hh, ww = x.shape
    P = hh * ww
    rows = row_off + P if rows is None else rows
    of = torch.zeros(B, rows, Cc, dtype=torch.float32, device=x.device) if f32 else None
    ob = torch.zeros(B, rows, Cc, dtype=torch.bfloat16, device=x.device) if bf16 else None
    _call(x, _lib.lib().dclip_nchw_to_tokens, _ptr(x), B, Cc, P, _ptr(of), _ptr(ob), Cc, rows * Cc, row_off, _stream(x))
    return of, ob


def token_mean(x: torch.Tensor, row0: int, P: int) -> torch.Tensor:
    """mean over rows [row0, row0+P) of fp32 [B, rows, D] -> [B, D]."""
    x = _req(x, torch.float32, "x")
    B, _, D = x.shape
    out = torch.empty(B, D, dtype=torch.float32, device=x.device)
    _call(x, _lib.lib().dclip_token_mean, _ptr(x), B, row0, P, x.stride(1), x.stride(0), D, _ptr(out), _stream(x))
    return out


def score_map(vis: torch.Tensor, row0: int, P: int, text: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """vis fp32 token-major [B, rows, C]; text fp32 [B, K, C] -> cosine score map [B, K, P]."""
    vis = _req(vis, torch.float32, "vis")
    text = _req(text.contiguous(), torch.float32, "text")
    B, K, Cc = text.shape
    out = torch.empty(B, K, P, dtype=torch.float32, device=vis.device)
    _call(vis, _lib.lib().dclip_score_map, _ptr(vis), vis.stride(1), vis.stride(0), row0, _ptr(text), B, K, Cc, P, float(eps),
          _ptr(out), _stream(vis))
    return out


def upsample_bilinear(x: torch.Tensor, size, *, tokens_hw=None, channels=None) -> torch.Tensor:
    """Bilinear (align_corners=False) resize to `size`. x: NCHW fp32, or token-major fp32 [B, h*w, ld] with tokens_hw."""
    H, W = size
    if tokens_hw is None:
        x = _req(x.contiguous(), torch.float32, "x")
        B, Cc, hh, ww = x.shape
        out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device)
        _call(x, _lib.lib().dclip_upsample_bilinear, _ptr(x), 1, 0, 0, B, Cc, hh, ww, H, W, _ptr(out), _stream(x))
    else:
        x = _req(x, torch.float32, "x")
        hh, ww = tokens_hw
        B = x.shape[0]
        Cc = channels if channels is not None else x.shape[2]
        out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device)
        _call(x, _lib.lib().dclip_upsample_bilinear, _ptr(x), 0, x.stride(1), x.stride(0), B, Cc, hh, ww, H, W, _ptr(out),
              _stream(x))
    return out


def upsample_argmax(x: torch.Tensor, size, *, tokens_hw, channels) -> torch.Tensor:
    """token-major fp32 low-res logits [B, h*w, ld] -> uint8 class map [B, H, W] (bilinear upsample + argmax, fused)."""
    x = _req(x, torch.float32, "x")
    H, W = size
    hh, ww = tokens_hw
    out = torch.empty(x.shape[0], H, W, dtype=torch.uint8, device=x.device)
    _call(x, _lib.lib().dclip_upsample_argmax, _ptr(x), x.stride(1), x.stride(0), x.shape[0], channels, hh, ww, H, W, _ptr(out),
          _stream(x))
    return out


def eval_stats(pred=None, target=None, num_classes: int = 19, ignore_index: int = 255, depth_pred=None, depth_gt=None,
               depth_mask=None, conf=None, depth_stats=None):
    """Accumulate one shard's evaluation statistics on the device: int64 confusion matrix [K, K]
    (rows = target, cols = prediction, `ignore_index` skipped) and float64 [sum squared depth error, count] over
    `depth_mask`.  `conf` / `depth_stats` are created zeroed when not passed, accumulated into when passed."""
    ref = pred if pred is not None else depth_pred
    if ref is None:
        raise DclipError("eval_stats: pass pred/target and/or depth_pred/depth_gt")
    dev = ref.device
    if conf is None:
        conf = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=dev)
    if depth_stats is None:
        depth_stats = torch.zeros(2, dtype=torch.float64, device=dev)
    _req(conf, torch.int64, "conf")
    _req(depth_stats, torch.float64, "depth_stats")
    p_ptr = t_ptr = dp_ptr = dg_ptr = dm_ptr = None
    n = nd = 0
    t_i64 = 0
    keep = []
    if pred is not None:
        pred = _req(pred.contiguous(), torch.uint8, "pred")
        if target is None or target.numel() != pred.numel():
            raise DclipError("eval_stats: target must have as many pixels as pred")
        if target.dtype not in (torch.uint8, torch.int64):
            raise DclipError(f"eval_stats: target must be uint8 or int64, got {target.dtype}")
        target = _req(target.contiguous(), target.dtype, "target")
        p_ptr, t_ptr, n, t_i64 = _ptr(pred), _ptr(target), pred.numel(), int(target.dtype == torch.int64)
        keep += [pred, target]
    if depth_pred is not None:
        depth_pred = _req(depth_pred.contiguous(), torch.float32, "depth_pred")
        if depth_gt is None or depth_gt.numel() != depth_pred.numel():
            raise DclipError("eval_stats: depth_gt must have as many pixels as depth_pred")
        depth_gt = _req(depth_gt.contiguous(), torch.float32, "depth_gt")
        dp_ptr, dg_ptr, nd = _ptr(depth_pred), _ptr(depth_gt), depth_pred.numel()
        keep += [depth_pred, depth_gt]
        if depth_mask is not None:
            if depth_mask.numel() != nd:
                raise DclipError("eval_stats: depth_mask must have as many pixels as depth_pred")
            depth_mask = depth_mask.contiguous()
            depth_mask = depth_mask.view(torch.uint8) if depth_mask.dtype == torch.bool else _req(depth_mask, torch.uint8, "depth_mask")
            dm_ptr = _ptr(depth_mask)
            keep.append(depth_mask)
    _call(ref, _lib.lib().dclip_eval_stats, p_ptr, t_ptr, t_i64, n, int(conf.shape[0]), int(ignore_index), dp_ptr, dg_ptr, dm_ptr, nd,
          _ptr(conf), _ptr(depth_stats), _stream(ref))
    return conf, depth_stats


def gamma_residual(a: torch.Tensor, gamma: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    a = _req(a.contiguous(), torch.float32, "a")
    d = _req(d.contiguous(), torch.float32, "d")
    out = torch.empty_like(a)
    _call(a, _lib.lib().dclip_gamma_residual, _ptr(a), _ptr(_req(gamma, torch.float32, "gamma")), _ptr(d), _ptr(out), a.numel(),
          gamma.numel(), _stream(a))
    return out


def conv3x3_gather(x: torch.Tensor, *, row0: int, hh: int, ww: int, channels: int) -> torch.Tensor:
    """token-major [B, rows, ld] (fp32 or bf16) -> bf16 [B*h*w, 9*C] 3x3/pad-1 patches, K order (ky, kx, c)."""
    assert x.dim() == 3 and x.stride(2) == 1
    B = x.shape[0]
    out = torch.empty(B * hh * ww, 9 * channels, dtype=torch.bfloat16, device=x.device)
    _call(x, _lib.lib().dclip_conv3x3_gather, _ptr(x), int(x.dtype == torch.float32), x.stride(1), x.stride(0), row0, B, hh, ww,
          channels, _ptr(out), out.stride(0), _stream(x))
    return out

"""Training mode of the trainable tail: forward with BatchNorm on batch statistics / Dropout, and the
backward of everything ``loss.backward()`` reaches in the reference's training step (train_denseclip.py:1226-1330).

What that is: the backbone and the text tower are frozen (train_denseclip.py:1040-1044), and ``DenseCLIP.forward`` feeds the
heads with the neck output of the ORIGINAL backbone features while the score map / ContextDecoder branch is computed and dropped
(denseclip.py:755-812, ``aux_losses = {}``), so gradients reach exactly

    ViTFeatureFusionNeck  (models.py:717-782)    12 x [conv3x3 -> BN -> ReLU] -> concat -> conv1x1 -> BN -> ReLU
    FCNHead x 2           (denseclip.py:305-349) conv3x3 -> BN -> ReLU -> Dropout(0.1) -> conv1x1 -> classifier conv1x1
    F.interpolate(bilinear, align_corners=False) to the ground-truth size (denseclip.py:838, 849)

Every matrix product is the tcgen05 GEMM of ``libdenseclip_b200.so``: the forward convs and the input gradients as implicit
3x3 convs (input gradient = conv of dY with the transposed, spatially flipped filter), the weight gradients in the GEMM's shifted-K
mode over channel-major operands (``dclip_transpose_pad``).  BatchNorm statistics / ReLU / Dropout masks, the reductions for
dgamma / dbeta / bias gradients and the adjoint of the bilinear resize are the kernels of ``csrc/train_tail.cuh``.  GEMM operands
are bf16 with fp32 accumulation (activations and gradients are kept in fp32 between layers), like the inference path.
``torch.autograd`` is only the tape: each ``Function`` below calls native kernels in both directions; the Dropout keep-mask is
drawn with ``torch.rand`` (so ``torch.manual_seed`` governs it, as in the reference)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import DclipError


def _ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=like.device)


def _p(t):
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ---------------------------------------------------------------------------------------------------------------
def col_stats(x: torch.Tensor, eps: float, run_mean=None, run_var=None, momentum: float = 0.1):
    """nn.BatchNorm2d training statistics of token-major fp32 [M, N]: (mean, biased var, rstd); running stats updated in place."""
    M, N = x.shape
    mean, var, rstd = (torch.empty(N, dtype=torch.float32, device=x.device) for _ in range(3))
    a = _lib.ColReduceArgs()
    a.a, a.lda, a.M, a.N, a.mode = x.data_ptr(), x.stride(0), M, N, 0
    ws = _ws(_lib.lib().dclip_col_reduce_workspace(M, N), x)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.out0, a.out1, a.out2, a.eps = mean.data_ptr(), var.data_ptr(), rstd.data_ptr(), eps
    a.run_mean, a.run_var, a.momentum = _p(run_mean), _p(run_var), momentum
    ops._call(x, _lib.lib().dclip_col_reduce, C.byref(a), ops._stream(x))
    return mean, var, rstd


def col_grad_sums(g: torch.Tensor, x=None, mean=None, rstd=None, gamma=None, beta=None, relu=False, mask=None, mask_scale=1.0):
    """(sum_rows g', sum_rows g' * xhat) with g' = g [* mask * mask_scale] [* (BN(x) > 0)]: dbeta / dgamma, or a bias gradient."""
    M, N = g.shape
    s0 = torch.empty(N, dtype=torch.float32, device=g.device)
    s1 = torch.empty(N, dtype=torch.float32, device=g.device) if x is not None else None
    a = _lib.ColReduceArgs()
    a.a, a.lda, a.M, a.N, a.mode, a.relu = g.data_ptr(), g.stride(0), M, N, 1, int(relu)
    if x is not None:
        a.x, a.ldx = x.data_ptr(), x.stride(0)
        a.mean, a.rstd, a.gamma, a.beta = mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    if mask is not None:
        a.mask, a.ldm, a.mask_scale = mask.data_ptr(), mask.stride(0), mask_scale
    ws = _ws(_lib.lib().dclip_col_reduce_workspace(M, N), g)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.out0, a.out1 = s0.data_ptr(), _p(s1)
    ops._call(g, _lib.lib().dclip_col_reduce, C.byref(a), ops._stream(g))
    return s0, s1


def bn_apply(mode: int, M: int, N: int, *, x=None, g=None, mean=None, rstd=None, gamma=None, beta=None, sum_g=None, sum_gx=None,
             relu=False, mask=None, mask_scale=1.0, want_f32=False, want_bf16=False, pad_cols_to: int = 8):
    """mode 0: dropout(relu(BN(x))); mode 1: BatchNorm(+ReLU, +dropout) input gradient; mode 2: masks only.
    Returns (fp32 [M, N] or None, bf16 [M, Np] or None) with Np = N rounded up (zero padding: a GEMM K dimension)."""
    ref = x if x is not None else g
    of = torch.empty(M, N, dtype=torch.float32, device=ref.device) if want_f32 else None
    Np = (N + pad_cols_to - 1) // pad_cols_to * pad_cols_to
    ob = (torch.zeros if Np != N else torch.empty)(M, Np, dtype=torch.bfloat16, device=ref.device) if want_bf16 else None
    a = _lib.BnApplyArgs()
    a.mode, a.M, a.N, a.relu = mode, M, N, int(relu)
    if g is not None:
        a.a, a.lda = g.data_ptr(), g.stride(0)
    if x is not None:
        a.x, a.ldx = x.data_ptr(), x.stride(0)
    if mean is not None:
        a.mean, a.rstd, a.gamma, a.beta = mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.sum_g, a.sum_gx = _p(sum_g), _p(sum_gx)
    if mask is not None:
        a.mask, a.ldm, a.mask_scale = mask.data_ptr(), mask.stride(0), mask_scale
    if of is not None:
        a.out_f32, a.ldo = of.data_ptr(), of.stride(0)
    if ob is not None:
        a.out_bf16, a.ldb = ob.data_ptr(), ob.stride(0)
    ops._call(ref, _lib.lib().dclip_bn_apply, C.byref(a), ops._stream(ref))
    return of, ob


def transpose_pad(t: torch.Tensor, B: int, gh: int, gw: int, channels: int, pad: int, *, ld=None, bs=None, pitch=None, lead: int = 0,
                  shift: int = 0, ldk=None, out=None, nshift: int = 1, plane: int = 0) -> torch.Tensor:
    """token-major [B][gh*gw][C] (fp32 or bf16; ``t`` points at the first pixel row) -> bf16 [C, ldk] channel-major over
    zero-padded images: out[c][k] = padded[c][k - lead + shift], padded index (b*(gh+pad) + y)*pitch + x (dclip_transpose_pad).
    nshift = 3: one pass writes the copies shift = -1, 0, +1 to out, out + plane, out + 2*plane (elements)."""
    pitch = gw + pad if pitch is None else pitch
    K = B * (gh + pad) * pitch
    if ldk is None:
        ldk = (K + lead + 63) // 64 * 64
    if out is None:
        out = torch.empty(channels, ldk, dtype=torch.bfloat16, device=t.device)
    ld = t.stride(-2) if ld is None else ld
    bs = gh * gw * ld if bs is None else bs
    ops._call(t, _lib.lib().dclip_transpose_pad, C.c_void_p(t.data_ptr()), int(t.dtype == torch.float32), ld, bs, B, gh, gw,
              channels, pad, pitch, lead, shift, nshift, plane, C.c_void_p(out.data_ptr()), ldk, ops._stream(t))
    return out


def _wgrad_geometry(geo):
    """(pitch, K): row pitch of the padded pixel axis (gw + 1 rounded up to 8: 16-byte aligned TMA box starts) and its length
    rounded up to the 64-column K block."""
    pitch = (geo.gw + 1 + 7) // 8 * 8
    return pitch, (geo.B * (geo.gh + 1) * pitch + 63) // 64 * 64


def conv3x3_wgrad_x_operand(x_tokens: torch.Tensor, geo, channels: int, *, x_ld=None, x_bs=None, out=None):
    """The three horizontally shifted copies of X^T, bf16 [3, C, K + 2*pitch] (see dclip_gemm_args.wg_*), from ONE pass over x.
    ``out``: optional [3, Ctot, ldx] buffer slice view whose planes are ``out.stride(0)`` elements apart (grouped launch)."""
    pitch, K = _wgrad_geometry(geo)
    ldx = K + 2 * pitch
    if out is None:
        out = torch.empty(3, channels, ldx, dtype=torch.bfloat16, device=x_tokens.device)
    transpose_pad(x_tokens, geo.B, geo.gh, geo.gw, channels, 1, ld=x_ld, bs=x_bs, pitch=pitch, lead=pitch, ldk=ldx, out=out[0],
                  nshift=3, plane=out.stride(0))
    return out


def conv3x3_wgrad_dy_operand(dy_b: torch.Tensor, geo, filters: int, *, dy_ld=None):
    """dY^T bf16 [F, K] over the same padded pixel axis."""
    pitch, K = _wgrad_geometry(geo)
    return transpose_pad(dy_b, geo.B, geo.gh, geo.gw, filters, 1, ld=dy_ld, pitch=pitch, ldk=K)


def upsample_bilinear_bwd(dout: torch.Tensor, gh: int, gw: int, ldc: int) -> torch.Tensor:
    dout = ops._req(dout.contiguous(), torch.float32, "dout")
    B, K, H, W = dout.shape
    dtok = (torch.zeros if ldc != K else torch.empty)(B * gh * gw, ldc, dtype=torch.float32, device=dout.device)
    ops._call(dout, _lib.lib().dclip_upsample_bilinear_bwd, C.c_void_p(dout.data_ptr()), B, K, H, W, gh, gw,
              C.c_void_p(dtok.data_ptr()), ldc, ops._stream(dout))
    return dtok


def conv3x3_wgrad(dyT: torch.Tensor, xT3: torch.Tensor, pitch: int, accumulate_into=None, groups: int = 1) -> torch.Tensor:
    """dW of a 3x3 / pad-1 conv as ONE GEMM over the padded pixel axis: dyT bf16 [F, K], xT3 bf16 [3, C, K + 2*pitch] -> fp32
    [F, 9*C] in the (ky, kx, c) order of the forward operand: output block t = (ky, kx) multiplies dyT with copy kx of X^T read
    ky*pitch columns further right.  ``groups`` = G > 1: G independent convs in one launch (dyT [G*128, K], xT3 [3, G*C, ...]:
    filter rows 128g.. pair with channel rows g*C..).  ``accumulate_into``: add to an earlier partial product."""
    F_, K = dyT.shape
    channels = xT3.shape[1] // groups
    if channels % 64:
        raise DclipError(f"conv3x3 weight gradient: input channels ({channels}) must be a multiple of 64")
    if groups > 1 and F_ != groups * 128:
        raise DclipError("grouped conv3x3 weight gradient needs 128 filters per group")
    out = accumulate_into if accumulate_into is not None else torch.empty(F_, 9 * channels, dtype=torch.float32, device=dyT.device)
    g = _lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw = dyT.data_ptr(), dyT.stride(0), xT3.data_ptr(), xT3.stride(1)
    g.M, g.N, g.K = F_, 9 * channels, K
    g.out_scale = 1.0
    g.out_f32, g.ldc = out.data_ptr(), out.stride(0)
    if accumulate_into is not None:
        g.residual, g.ldr = out.data_ptr(), out.stride(0)
    g.wg_C, g.wg_pitch, g.wg_grouped, g.wg_rows = channels, pitch, int(groups > 1), xT3.shape[1]
    g.block_n = 256 if channels % 256 == 0 else (128 if channels % 128 == 0 else 64)
    ops._call(dyT, _lib.lib().dclip_gemm, C.byref(g), ops._stream(dyT))
    return out


# ---------------------------------------------------------------------------------------------------------------
# operands in 1 (bf16) or 2 (hi + lo, fp32-class) bf16 parts
# ---------------------------------------------------------------------------------------------------------------
# precision "fp32" (default: the reference trains in fp32, train_denseclip.py has no autocast): every product is the three-pass
# split  A W^T ~= Ah Wh^T + Al Wh^T + Ah Wl^T  (x = hi + lo, both bf16; ~4e-6 relative error) accumulated in fp32 through the GEMM's
# residual input; the frozen backbone taps are exactly bf16, so their lo part vanishes (two passes).  precision "bf16": one pass.
def _parts(x2d: torch.Tensor, split: bool, pad_cols_to: int = 8):
    """fp32 [M, C] -> list of bf16 [M, Cp] operands (views of one hi|lo buffer when split)."""
    if x2d.shape[1] % 2:   # (the cast kernel converts column pairs: one zero column of layout padding)
        x2d = torch.nn.functional.pad(x2d, (0, 1))
    x2d = ops._req(x2d.contiguous(), torch.float32, "x")
    if not split:
        return [ops.cast_bf16(x2d, pad_cols_to=pad_cols_to)]
    hl = ops.split_bf16(x2d, pad_cols_to=pad_cols_to)
    Cp = hl.shape[1] // 2
    return [hl[:, :Cp], hl[:, Cp:]]


def _pairs(a_parts, w_parts):
    """(a, w) products of the split expansion: hi*hi, lo*hi, hi*lo."""
    out = [(a_parts[0], w_parts[0])]
    if len(a_parts) > 1:
        out.append((a_parts[1], w_parts[0]))
    if len(w_parts) > 1:
        out.append((a_parts[0], w_parts[1]))
    return out


def _mm(a_parts, w_parts, bias=None, out=None):
    """fp32 [M, N] = A W^T (+ bias) from operand parts; ``out``: optional destination (a column slice is fine)."""
    acc = None
    for a, w in _pairs(a_parts, w_parts):
        first = acc is None
        acc, _ = ops.gemm(a, w, bias=bias if first else None, residual=acc, out_f32=out if first else acc,
                          want_f32=first and out is None)
    return acc


def _conv3x3_mm(tok_parts, row0, geo, Cp, w_parts, out=None):
    """3x3 / pad-1 conv of token-major activations (parts: bf16 [B, rows, Cp] views) with filter parts bf16 [F, 9*Cp] in
    (ky, kx, c) order -> fp32 [M, F]: implicit GEMM when the grid tiles by 128 pixels, explicit gather otherwise."""
    tileable = (geo.gh * geo.gw) % 128 == 0 and (128 % geo.gw == 0 or geo.gw % 128 == 0) and Cp % 64 == 0
    if not tileable:
        gathered = [ops.conv3x3_gather(t, row0=row0, hh=geo.gh, ww=geo.gw, channels=Cp) for t in tok_parts]
        return _mm(gathered, w_parts, out=out)
    acc = None
    for t, w in _pairs(tok_parts, w_parts):
        first = acc is None
        a = t[:, row0:, :]
        a2 = a.as_strided((geo.M, Cp), (a.stride(1), 1), a.storage_offset())   # row view: pointer / pitch only
        acc, _ = ops.gemm(a2, w, K=9 * Cp, residual=acc, out_f32=out if first else acc, want_f32=first and out is None, M=geo.M,
                          conv=dict(C=Cp, gw=geo.gw, gh=geo.gh, B=geo.B, a_bs=t.stride(0)))
    return acc


class _Geom:
    def __init__(self, B, gh, gw, split=True):
        self.B, self.gh, self.gw, self.M, self.split = B, gh, gw, B * gh * gw, split


def _rows2d(t: torch.Tensor, row0: int, geo: _Geom):
    """token tensor [B, rows, Cp] -> its pixel rows as a 2-D [M, Cp] operand (a view whenever the images are back to back)."""
    v = t[:, row0:, :]
    if row0 == 0 and t.stride(0) == t.shape[1] * t.stride(1):
        return v.as_strided((geo.M, v.shape[2]), (t.stride(1), 1), v.storage_offset())
    return v.contiguous().view(geo.M, v.shape[2])


def _conv_forward(tok_parts, row0: int, geo: _Geom, weight: torch.Tensor, bias, out=None):
    """tok_parts: bf16 token tensors [B, rows, Cp]; weight [F, C, k, k] -> pre-activation fp32 [M, F] (bias of a plain conv
    fused; for the 1x1 convs F is padded to a multiple of 4 with zero filters: the GEMM epilogue is 4-wide)."""
    k, Cc = weight.shape[-1], weight.shape[1]
    Cp = tok_parts[0].shape[2]
    w = weight.detach().float()
    if k == 3:
        if Cp != Cc:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, Cp - Cc))
        w9 = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9 * Cp)      # (ky, kx, c) order of the implicit-conv operand
        return _conv3x3_mm(tok_parts, row0, geo, Cp, _parts(w9, geo.split), out=out)
    w2 = w.reshape(w.shape[0], Cc)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    F_ = w2.shape[0]
    if F_ % 4:
        w2 = torch.nn.functional.pad(w2, (0, 0, 0, 4 - F_ % 4))
        if bias32 is not None:
            bias32 = torch.nn.functional.pad(bias32, (0, 4 - F_ % 4))
    if Cp != Cc:
        w2 = torch.nn.functional.pad(w2, (0, Cp - Cc))
    return _mm([_rows2d(t, row0, geo) for t in tok_parts], _parts(w2, geo.split), bias=bias32, out=out)


def _conv_backward(dpre: torch.Tensor, tok_parts, row0: int, geo: _Geom, weight: torch.Tensor, need_dx: bool):
    """dpre: fp32 [M, >= F] gradient of the conv output (only the first F columns are read); tok_parts: the input operand parts
    saved by the forward.  Returns (dW like weight, dX fp32 [M, C] or None)."""
    F_, Cc, k = weight.shape[0], weight.shape[1], weight.shape[-1]
    dy_parts = _parts(dpre[:, :F_] if dpre.shape[1] != F_ else dpre, geo.split)      # bf16 [M, Fp] (zero columns up to Fp)
    Fp = dy_parts[0].shape[1]
    x_views = [t[:, row0:, :] for t in tok_parts]
    if k == 3:
        pitch, _ = _wgrad_geometry(geo)
        dyT = [conv3x3_wgrad_dy_operand(d, geo, F_, dy_ld=d.stride(0)) for d in dy_parts]
        xT3 = [conv3x3_wgrad_x_operand(v, geo, Cc, x_ld=t.stride(1), x_bs=t.stride(0)) for v, t in zip(x_views, tok_parts)]
        dw9 = None
        for a_t, x_t in _pairs(dyT, xT3):
            dw9 = conv3x3_wgrad(a_t, x_t, pitch, accumulate_into=dw9)
        dw = dw9.view(F_, 3, 3, Cc).permute(0, 3, 1, 2)
    else:
        dyT = [transpose_pad(d, geo.B, geo.gh, geo.gw, F_, 0, ld=d.stride(0)) for d in dy_parts]
        xT = [transpose_pad(v, geo.B, geo.gh, geo.gw, Cc, 0, ld=t.stride(1), bs=t.stride(0)) for v, t in zip(x_views, tok_parts)]
        dw = _mm(dyT, xT).view(F_, Cc, 1, 1)
    dx = None
    if need_dx:
        w = weight.detach().float()
        P = geo.gh * geo.gw
        if k == 3:
            wt = w.flip(2, 3).permute(1, 2, 3, 0)                 # W'[c, ky', kx', f] = W[f, c, 2 - ky', 2 - kx']
            if Fp != F_:
                wt = torch.nn.functional.pad(wt, (0, Fp - F_))
            dy_tok = [d.as_strided((geo.B, P, Fp), (P * d.stride(0), d.stride(0), 1), d.storage_offset()) for d in dy_parts]
            dx = _conv3x3_mm(dy_tok, 0, geo, Fp, _parts(wt.reshape(Cc, 9 * Fp), geo.split))
        else:
            wt = torch.zeros(Cc, Fp, dtype=torch.float32, device=weight.device)
            wt[:, :F_] = w.reshape(F_, Cc).t()
            dx = _mm(dy_parts, _parts(wt, geo.split))
    return dw.contiguous(), dx


def _tok_parts_of(x: torch.Tensor, geo: _Geom):
    """fp32 [M, C] activations -> operand parts as token tensors [B, P, Cp] (views of one buffer)."""
    P = geo.gh * geo.gw
    return [p2.as_strided((geo.B, P, p2.shape[1]), (P * p2.stride(0), p2.stride(0), 1), p2.storage_offset())
            for p2 in _parts(x, geo.split)]


# ---------------------------------------------------------------------------------------------------------------
# autograd tape
# ---------------------------------------------------------------------------------------------------------------
def _grouped_wgrad(dpre: torch.Tensor, flat_parts, nparts: int, row0: int, geo: _Geom, G: int, F_: int, Cc: int) -> torch.Tensor:
    """Weight gradients of G independent 3x3 convs (128 filters each, same C) in ONE launch per operand pair: dpre fp32 [M, G*F],
    flat_parts = the taps' operand parts (tap-major).  A per-tap launch is only 9*C/256 = 27 tiles on 148 SMs.  -> fp32 [G*F, 9*C]."""
    pitch, K = _wgrad_geometry(geo)
    dyT = [conv3x3_wgrad_dy_operand(d, geo, G * F_, dy_ld=d.stride(0)) for d in _parts(dpre, geo.split)]
    xT3 = []
    for part in range(nparts):
        buf = torch.empty(3, G * Cc, K + 2 * pitch, dtype=torch.bfloat16, device=dpre.device)
        for i in range(G):
            t = flat_parts[i * nparts + part]
            conv3x3_wgrad_x_operand(t[:, row0:, :], geo, Cc, x_ld=t.stride(1), x_bs=t.stride(0), out=buf[:, i * Cc:(i + 1) * Cc])
        xT3.append(buf)
    dw_all = None
    for a_t, x_t in _pairs(dyT, xT3):
        dw_all = conv3x3_wgrad(a_t, x_t, pitch, accumulate_into=dw_all, groups=G)
    return dw_all


class _ConvBlock(torch.autograd.Function):
    """conv (k = 1 | 3, pad = k // 2) [+ BatchNorm(batch statistics) + ReLU] [+ Dropout] on token-major activations:
    fp32 [M, C] -> fp32 [M, F] (F rounded up to a multiple of 4 with zero columns for a plain 1x1 conv)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, bias, geo, bn, relu, drop_p):
        if x.shape[1] != weight.shape[1]:
            raise DclipError(f"conv block: input has {x.shape[1]} channels, the filter expects {weight.shape[1]}")
        toks = _tok_parts_of(x, geo)
        F_ = weight.shape[0]
        pre = _conv_forward(toks, 0, geo, weight, None if bn is not None else bias)
        mask, scale = None, 1.0
        if drop_p > 0.0:
            mask = (torch.rand(geo.M, F_, device=pre.device) >= drop_p).to(torch.uint8)
            scale = 1.0 / (1.0 - drop_p)
        ctx.geo, ctx.relu, ctx.scale, ctx.has_bn = geo, relu, scale, bn is not None
        if bn is not None:
            g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
            track = bn.track_running_stats
            mean, _, rstd = col_stats(pre, bn.eps, bn.running_mean if track else None, bn.running_var if track else None,
                                      bn.momentum if bn.momentum is not None else 0.1)
            if track and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            y, _ = bn_apply(0, geo.M, F_, x=pre, mean=mean, rstd=rstd, gamma=g32, beta=b32, relu=relu, mask=mask, mask_scale=scale,
                            want_f32=True)
            ctx.save_for_backward(weight, pre, mean, rstd, g32, b32, mask, *toks)
        elif relu or mask is not None:
            raise DclipError("conv + activation without BatchNorm has no native training path (not used by the reference's tail)")
        else:
            y = pre
            ctx.save_for_backward(weight, None, None, None, None, None, None, *toks)
        return y

    @staticmethod
    def backward(ctx, gy):
        weight, pre, mean, rstd, g32, b32, mask = ctx.saved_tensors[:7]
        toks = list(ctx.saved_tensors[7:])
        geo = ctx.geo
        F_ = weight.shape[0]
        gy = ops._req(gy.contiguous(), torch.float32, "grad")
        dgamma = dbeta = dbias = None
        if ctx.has_bn:
            sg, sgx = col_grad_sums(gy, pre, mean, rstd, g32, b32, relu=ctx.relu, mask=mask, mask_scale=ctx.scale)
            dgamma, dbeta = sgx, sg
            dpre, _ = bn_apply(1, geo.M, F_, x=pre, g=gy, mean=mean, rstd=rstd, gamma=g32, beta=b32, sum_g=sg, sum_gx=sgx,
                               relu=ctx.relu, mask=mask, mask_scale=ctx.scale, want_f32=True)
        else:
            if ctx.needs_input_grad[4]:
                dbias, _ = col_grad_sums(gy)
                dbias = dbias[:F_]
            dpre = gy
        dw, dx = _conv_backward(dpre, toks, 0, geo, weight, ctx.needs_input_grad[0])
        return dx, dw.to(weight.dtype), dgamma, dbeta, dbias, None, None, None, None


class _NeckTaps(torch.autograd.Function):
    """The G per-tap ConvBNReLU layers of ViTFeatureFusionNeck (models.py:766-770) + the channel concat: every conv writes its
    column slice of one [M, G*F] pre-activation matrix, ONE pass computes the batch statistics of all G BatchNorms, one pass
    applies them.  The taps are frozen backbone features: no input gradient.  ``tap_parts[i]``: the operand parts of tap i
    (one exactly-bf16 token tensor from the bf16 encoder, or hi + lo of the fp32-class encoder's output)."""

    @staticmethod
    def forward(ctx, geo, row0, bns, tap_parts, *params):
        G = len(tap_parts)
        ws, gammas, betas = params[:G], params[G:2 * G], params[2 * G:]
        F_ = ws[0].shape[0]
        dev = tap_parts[0][0].device
        pre = torch.empty(geo.M, G * F_, dtype=torch.float32, device=dev)
        for i in range(G):
            _conv_forward(tap_parts[i], row0, geo, ws[i], None, out=pre[:, i * F_:(i + 1) * F_])
        g32 = torch.cat([g.detach().float() for g in gammas]).contiguous()
        b32 = torch.cat([b.detach().float() for b in betas]).contiguous()
        track = all(bn.track_running_stats for bn in bns)
        rm = torch.cat([bn.running_mean for bn in bns]) if track else None
        rv = torch.cat([bn.running_var for bn in bns]) if track else None
        mean, _, rstd = col_stats(pre, bns[0].eps, rm, rv, bns[0].momentum if bns[0].momentum is not None else 0.1)
        if track:
            for i, bn in enumerate(bns):   # (parameter bookkeeping: scatter the updated running statistics back)
                bn.running_mean.copy_(rm[i * F_:(i + 1) * F_])
                bn.running_var.copy_(rv[i * F_:(i + 1) * F_])
                bn.num_batches_tracked += 1
        y, _ = bn_apply(0, geo.M, G * F_, x=pre, mean=mean, rstd=rstd, gamma=g32, beta=b32, relu=True, want_f32=True)
        ctx.geo, ctx.row0, ctx.G, ctx.counts = geo, row0, G, [len(p) for p in tap_parts]
        ctx.save_for_backward(pre, mean, rstd, g32, b32, *ws, *[t for p in tap_parts for t in p])
        return y

    @staticmethod
    def backward(ctx, gy):
        G, geo = ctx.G, ctx.geo
        pre, mean, rstd, g32, b32 = ctx.saved_tensors[:5]
        ws, flat = ctx.saved_tensors[5:5 + G], list(ctx.saved_tensors[5 + G:])
        F_ = ws[0].shape[0]
        gy = ops._req(gy.contiguous(), torch.float32, "grad")
        sg, sgx = col_grad_sums(gy, pre, mean, rstd, g32, b32, relu=True)
        dpre, _ = bn_apply(1, geo.M, G * F_, x=pre, g=gy, mean=mean, rstd=rstd, gamma=g32, beta=b32, sum_g=sg, sum_gx=sgx, relu=True,
                           want_f32=True)
        Cc = ws[0].shape[1]
        same = all(w.shape == ws[0].shape for w in ws) and len(set(ctx.counts)) == 1
        if same and F_ == 128 and Cc % 64 == 0 and G > 1:
            dw_all = _grouped_wgrad(dpre, flat, ctx.counts[0], ctx.row0, geo, G, F_, Cc)
            dws = [dw_all[i * F_:(i + 1) * F_].view(F_, 3, 3, Cc).permute(0, 3, 1, 2).contiguous().to(ws[i].dtype) for i in range(G)]
        else:
            dws, pos = [], 0
            for i in range(G):
                parts = flat[pos:pos + ctx.counts[i]]
                pos += ctx.counts[i]
                dw, _ = _conv_backward(dpre[:, i * F_:(i + 1) * F_].contiguous(), parts, ctx.row0, geo, ws[i], need_dx=False)
                dws.append(dw.to(ws[i].dtype))
        dg = [sgx[i * F_:(i + 1) * F_] for i in range(G)]
        db = [sg[i * F_:(i + 1) * F_] for i in range(G)]
        return (None, None, None, None, *dws, *dg, *db)


class _UpsampleBilinear(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=False) of token-major logits [M, ld] (first ``channels`` columns) to NCHW
    [B, channels, H, W] (denseclip.py:838, 849); backward = the adjoint kernel."""

    @staticmethod
    def forward(ctx, y, geo, channels, out_hw):
        ctx.geo, ctx.channels, ctx.ld = geo, channels, y.shape[1]
        return ops.upsample_bilinear(y.view(geo.B, geo.gh * geo.gw, -1), out_hw, tokens_hw=(geo.gh, geo.gw), channels=channels)

    @staticmethod
    def backward(ctx, gout):
        return upsample_bilinear_bwd(gout, ctx.geo.gh, ctx.geo.gw, ctx.ld), None, None, None


# ---------------------------------------------------------------------------------------------------------------
# module-level entry points (called by DenseCLIP.forward in training mode)
# ---------------------------------------------------------------------------------------------------------------
def neck_forward_train(neck, taps, row0: int, gh: int, gw: int, split: bool = True) -> torch.Tensor:
    """ViTFeatureFusionNeck.forward in training mode (models.py:761-782) -> fused fp32 [M, out] with a tape.  ``taps``: per tap
    either a bf16 token tensor [B, rows, C] whose pixels start at ``row0`` (bf16 encoder) or fp32 [M, C] pixels (fp32-class encoder)."""
    B = taps[0].shape[0] if taps[0].dim() == 3 else taps[0].shape[0] // (gh * gw)
    geo = _Geom(B, gh, gw, split)
    exact = taps[0].dtype == torch.bfloat16
    tap_parts = [[t] if exact else _tok_parts_of(t, geo) for t in taps]
    convs = [layer[0] for layer in neck.process_layers]
    bns = [layer[1] for layer in neck.process_layers]
    cat = _NeckTaps.apply(geo, row0 if exact else 0, bns, tap_parts, *[c.weight for c in convs], *[b.weight for b in bns],
                          *[b.bias for b in bns])
    fconv, fbn = neck.fusion_layer[0], neck.fusion_layer[1]
    return _ConvBlock.apply(cat, fconv.weight, fbn.weight, fbn.bias, None, geo, fbn, True, 0.0)


def head_forward_train(head, fused: torch.Tensor, B: int, gh: int, gw: int, split: bool = True):
    """FCNHead (+ appended classifier) in training mode on token-major fp32 [M, C] -> (logits fp32 [M, n_out padded to 4], n_out)."""
    geo = _Geom(B, gh, gw, split)
    mods = list(head.children())
    conv0, bn0 = mods[0], mods[1]
    drop_p = next((m.p for m in mods if isinstance(m, nn.Dropout)), 0.0)
    y = _ConvBlock.apply(fused, conv0.weight, bn0.weight, bn0.bias, None, geo, bn0, True, float(drop_p))
    n_out = conv0.out_channels
    for c in [m for m in mods[2:] if isinstance(m, nn.Conv2d)]:   # the 1x1 convs: [4] and the appended classifier
        y = _ConvBlock.apply(y, c.weight, None, None, c.bias, geo, None, False, 0.0)
        n_out = c.out_channels
    return y, n_out   # (y may carry zero columns up to a multiple of 4)


def upsample_train(y: torch.Tensor, B: int, gh: int, gw: int, channels: int, out_hw) -> torch.Tensor:
    return _UpsampleBilinear.apply(y, _Geom(B, gh, gw), channels, tuple(out_hw))

"""Training losses of the reference (train_denseclip.py:1086-1096), native on the B200:

* ``SILogLoss`` -- same class name, constructor and ``forward(prediction, target, mask=None)`` as ``denseclip/losses.py:7-79``
  (``loss = mean(d^2) - lambd * mean(d)^2`` over the masked pixels, ``d = log(clamp(pred, eps)) - log(clamp(target, eps))``,
  0 when no pixel is valid);
* ``CrossEntropyLoss`` -- ``torch.nn.CrossEntropyLoss(ignore_index=...)`` for ``[B, K, H, W]`` logits and ``[B, H, W]`` int64 targets,
  mean over the non-ignored pixels (the reference's ``criterion_seg``).

Forward and backward are the deterministic two-stage reductions / elementwise kernels of ``csrc/train_tail.cuh``; autograd is
only the tape.  CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops


def _stats_ws(like: torch.Tensor):
    stats = torch.empty(4, dtype=torch.float32, device=like.device)
    ws = torch.empty(_lib.lib().dclip_loss_workspace(), dtype=torch.uint8, device=like.device)
    return stats, ws


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        logits = ops._req(logits.contiguous(), torch.float32, "logits")
        target = ops._req(target.contiguous(), torch.int64, "target")
        B, K = logits.shape[:2]
        HW = logits[0, 0].numel()
        if target.numel() != B * HW:
            raise ValueError(f"CrossEntropyLoss: target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
        stats, ws = _stats_ws(logits)
        ops._call(logits, _lib.lib().dclip_ce_loss, C.c_void_p(logits.data_ptr()), C.c_void_p(target.data_ptr()), B, K, HW,
                  int(ignore_index), C.c_void_p(ws.data_ptr()), C.c_void_p(stats.data_ptr()), ops._stream(logits))
        ctx.save_for_backward(logits, target, stats)
        ctx.ignore_index = int(ignore_index)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, gout):
        logits, target, stats = ctx.saved_tensors
        B, K = logits.shape[:2]
        HW = logits[0, 0].numel()
        grad = torch.empty_like(logits)
        g = gout.reshape(1).float().contiguous()
        ops._call(logits, _lib.lib().dclip_ce_loss_bwd, C.c_void_p(logits.data_ptr()), C.c_void_p(target.data_ptr()), B, K, HW,
                  ctx.ignore_index, C.c_void_p(stats.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(grad.data_ptr()),
                  ops._stream(logits))
        return grad, None, None


class CrossEntropyLoss(nn.Module):
    """torch.nn.CrossEntropyLoss(ignore_index) restricted to what the reference uses: mean reduction, no class weights."""

    def __init__(self, ignore_index: int = -100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits, target):
        return _CrossEntropyFn.apply(logits, target, self.ignore_index)


class _SILogFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, mask, lambd, eps):
        pred = ops._req(prediction.contiguous(), torch.float32, "prediction")
        tgt = ops._req(target.contiguous(), torch.float32, "target")
        m8 = None if mask is None else mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous().to(torch.uint8)
        stats, ws = _stats_ws(pred)
        ops._call(pred, _lib.lib().dclip_silog_loss, C.c_void_p(pred.data_ptr()), C.c_void_p(tgt.data_ptr()),
                  None if m8 is None else C.c_void_p(m8.data_ptr()), pred.numel(), float(lambd), float(eps), C.c_void_p(ws.data_ptr()),
                  C.c_void_p(stats.data_ptr()), ops._stream(pred))
        ctx.save_for_backward(pred, tgt, m8, stats)
        ctx.lambd, ctx.eps = float(lambd), float(eps)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, gout):
        pred, tgt, m8, stats = ctx.saved_tensors
        grad = torch.empty_like(pred)
        g = gout.reshape(1).float().contiguous()
        ops._call(pred, _lib.lib().dclip_silog_loss_bwd, C.c_void_p(pred.data_ptr()), C.c_void_p(tgt.data_ptr()),
                  None if m8 is None else C.c_void_p(m8.data_ptr()), pred.numel(), ctx.lambd, ctx.eps, C.c_void_p(stats.data_ptr()),
                  C.c_void_p(g.data_ptr()), C.c_void_p(grad.data_ptr()), ops._stream(pred))
        return grad, None, None, None, None


class SILogLoss(nn.Module):
    """Scale-invariant log loss, reference denseclip/losses.py:7-79 (variance form, no square root)."""

    def __init__(self, lambd=0.5, eps=1e-6, reduction='mean'):
        super().__init__()
        self.lambd = lambd
        self.eps = eps
        self.reduction = reduction
        if reduction not in ['mean', 'sum']:
            raise ValueError(f"Invalid reduction type: {reduction}. Must be 'mean' or 'sum'.")

    def forward(self, prediction, target, mask=None):
        if mask is not None and mask.shape != prediction.shape:
            if mask.dim() == prediction.dim() - 1 and tuple(mask.shape) == tuple(prediction.shape[:1] + prediction.shape[2:]):
                mask = mask.unsqueeze(1)   # [B, H, W] -> [B, 1, H, W] (losses.py:43-44)
            else:
                raise ValueError(f"Mask shape {mask.shape} incompatible with log_diff shape {prediction.shape}")
        return _SILogFn.apply(prediction, target, mask, self.lambd, self.eps)

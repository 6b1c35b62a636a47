"""Decode-head surface of the reference (segmentation/denseclip/heads.py:64-107; torchvision FCNHead used at
denseclip.py:305-309, 343-349): the ``HEADS`` registry, ``IdentityHead`` and an ``FCNHead`` parameter container whose
layer indices (0 conv3x3, 1 BN, 2 ReLU, 3 Dropout, 4 conv1x1, + appended ``classifier``) match torchvision's, so
``decode_head.{0,1,4,classifier}.*`` checkpoints load unchanged.  The forward is native (implicit-GEMM conv with
the BatchNorm folded, ReLU in the epilogue, 1x1 convs as GEMMs)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib, ops
from .models import _param_versions, conv3x3_tokens, conv3x3_weight_to_gemm, fold_bn, default_precision


class Registry:
    _registry = {}

    @classmethod
    def register_module(cls, name=None):
        def decorator(module_class):
            cls._registry[name if name is not None else module_class.__name__] = module_class
            return module_class
        return decorator

    @classmethod
    def get(cls, name):
        return cls._registry.get(name)


HEADS = Registry()


class BaseDecodeHead(nn.Module):
    def __init__(self, input_transform=None, **kwargs):
        super().__init__()
        self.input_transform = input_transform

    def forward(self, inputs):
        raise NotImplementedError


@HEADS.register_module()
class IdentityHead(BaseDecodeHead):
    """Reference heads.py:91-107: returns its input unchanged."""

    def __init__(self, **kwargs):
        super().__init__(input_transform=None, **kwargs)
        self.conv_seg = None

    def forward(self, inputs):
        return inputs


def _pad_rows(w: torch.Tensor, b: torch.Tensor, mult: int = 4):
    n = w.shape[0]
    npad = (n + mult - 1) // mult * mult
    if npad == n:
        return w, b
    w2 = torch.zeros(npad, w.shape[1], dtype=w.dtype, device=w.device)
    w2[:n] = w
    b2 = torch.zeros(npad, dtype=b.dtype, device=b.device)
    b2[:n] = b
    return w2, b2


@HEADS.register_module()
class FCNHead(nn.Sequential):
    """Same layer list as torchvision.models.segmentation.fcn.FCNHead(in_channels, channels); DenseCLIP appends a
    ``classifier`` 1x1 conv afterwards (assigning a module attribute on an nn.Sequential appends it to the run order)."""

    def __init__(self, in_channels: int, channels: int, precision: str = None):
        inter_channels = in_channels // 4
        super().__init__(
            nn.Conv2d(in_channels, inter_channels, 3, padding=1, bias=False),
            nn.BatchNorm2d(inter_channels),
            nn.ReLU(),
            nn.Dropout(0.1),
            nn.Conv2d(inter_channels, channels, 1),
        )
        self.in_channels = in_channels
        self.precision = precision or default_precision()
        self._packed = None

    def _pack(self):
        precise = self.precision == "fp32"
        ver = (_param_versions(self), precise)
        if self._packed is not None and self._packed["ver"] == ver:
            return self._packed
        convs = [m for m in self.children() if isinstance(m, nn.Conv2d)]
        w0, b0 = fold_bn(self[0].weight, self[1])
        p = dict(ver=ver, w0=ops.pack_weight(conv3x3_weight_to_gemm(w0), precise), b0=b0, tail=[])
        for c in convs[1:]:  # 1x1 convs after the 3x3: [4] and the appended classifier
            w = c.weight.detach().float().reshape(c.out_channels, -1)
            b = c.bias.detach().float() if c.bias is not None else torch.zeros(c.out_channels, device=w.device)
            w, b = _pad_rows(w, b, 4)
            p["tail"].append((ops.pack_weight(w, True), b.contiguous(), c.out_channels))
        self._packed = p
        return p

    def forward_tokens(self, tokens_b: torch.Tensor, gh: int, gw: int):
        """tokens_b: bf16 [B, gh*gw, C(x2 hi|lo in fp32 mode)] -> fp32 [B*gh*gw, n_out_padded], n_out."""
        if self.training:
            raise _lib.DclipError("FCNHead native path is inference-only (BatchNorm/Dropout); call .eval()")
        precise = self.precision == "fp32"
        pk = self._pack()
        y, _ = conv3x3_tokens(tokens_b, 0, gh, gw, self.in_channels, pk["w0"], split_in=precise, bias=pk["b0"], act="relu",
                              want_f32=True)
        n_out = y.shape[1]
        for w, b, n in pk["tail"]:  # fp32-accurate (split) 1x1 convs: tiny, and they produce the logits
            y, _ = ops.gemm(ops.split_bf16(y), w, split_in=True, bias=b, want_f32=True)
            n_out = n
        return y, n_out

    def forward(self, x: torch.Tensor):
        """API-compatible entry: NCHW fp32 [B, C, h, w] -> NCHW fp32 [B, n_out, h, w]."""
        precise = self.precision == "fp32"
        B, _, gh, gw = x.shape
        tf, tb = ops.nchw_to_tokens(x, f32=precise, bf16=not precise)
        tok = ops.split_bf16(tf.view(-1, tf.shape[2])).view(B, gh * gw, -1) if precise else tb
        y, n_out = self.forward_tokens(tok, gh, gw)
        return ops.upsample_bilinear(y.view(B, gh * gw, -1), (gh, gw), tokens_hw=(gh, gw), channels=n_out)

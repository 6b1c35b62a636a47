"""Host-side mirror of the reference model components (segmentation/denseclip/models.py) for the DenseCLIP forward path.

Every class keeps the reference's name, constructor kwargs and ``state_dict`` keys, so checkpoints and configs are
interchangeable, but ``forward`` runs the hand-written sm_100a kernels behind the C ABI (``ops`` / ``_lib``).  The
``nn.Module`` tree only *owns parameters*; no torch compute op runs on the forward path, and there is no CPU fallback.

Precision: ``precision="bf16"`` (default) is the tensor-core path (bf16 operands, fp32 accumulation, fp32 residual
stream and LayerNorm/softmax statistics); ``precision="fp32"`` is the fp32-class path: every GEMM and both attention
products run as three-pass split-bf16 tensor-core products (hi*hi + lo*hi + hi*lo, ~2^-16 relative error), softmax in fp32.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import os
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import _lib, ops

logger = logging.getLogger(__name__)


def default_precision() -> str:
    return os.environ.get("DENSECLIP_B200_PRECISION", "bf16")


def _native_only(name):
    raise RuntimeError(f"{name} is a parameter container of the B200-native path; call the owning model's forward")


# ================ registry stand-in (reference models.py:48-67) ================ #
class Registry:
    """Same surface as the reference's mmseg stand-in: ``register_module(name=None)`` decorator, ``build(cfg, **kw)``."""
    _registry = {}

    @classmethod
    def register_module(cls, name=None):
        def decorator(module_class):
            cls._registry[name if name is not None else module_class.__name__] = module_class
            return module_class
        return decorator

    @classmethod
    def build(cls, cfg, **kwargs):
        if isinstance(cfg, dict):
            cfg = dict(cfg)
            obj_type = cfg.pop('type')
            return cls._registry[obj_type](**cfg, **kwargs)
        return cls._registry[cfg](**kwargs)

    @classmethod
    def get(cls, name):
        return cls._registry.get(name)


BACKBONES = Registry()


# ================ parameter containers ================ #
class LayerNorm(nn.LayerNorm):
    """Reference models.py:243-249. ``forward`` runs the native LayerNorm kernel (fp32 statistics)."""

    def forward(self, x: torch.Tensor):
        shp = x.shape
        y, _ = ops.layernorm(x.reshape(-1, shp[-1]).float().contiguous(), self.weight, self.bias, self.eps, want_f32=True)
        return y.reshape(shp).to(x.dtype)


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) (reference models.py:252-254); fused into the c_fc GEMM epilogue on the native path."""

    def forward(self, x):
        _native_only("QuickGELU")


class ConvBNReLU(nn.Sequential):
    """Conv-BatchNorm-ReLU container (reference models.py:13-20); folded into one GEMM epilogue natively."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, stride=1):
        super().__init__(
            nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )


class ResidualAttentionBlock(nn.Module):
    """Reference models.py:271-294: x += MHA(ln_1(x)); x += c_proj(QuickGELU(c_fc(ln_2(x))))."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None, drop_path=0.):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.drop_path_rate = drop_path

    def forward(self, x):
        _native_only("ResidualAttentionBlock")


class Transformer(nn.Module):
    """Reference models.py:297-307 (parameter container). NB the reference ``forward`` applies the stack twice; the
    ViT iterates ``resblocks`` itself (one pass), the text towers go through ``forward`` (two passes)."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None, drop_path_rate=0.):
        super().__init__()
        self.width = width
        self.layers = layers
        self.heads = heads
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, layers)]
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask, dpr[i]) for i in range(layers)])

    def forward(self, x):
        _native_only("Transformer")


def _param_versions(module: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters()) + tuple(
        (b.data_ptr(), b._version) for b in module.buffers())


def _f32(p: torch.Tensor) -> torch.Tensor:
    t = p.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _fold_layernorm(ln: nn.LayerNorm, weight: torch.Tensor, bias: torch.Tensor):
    """LayerNorm folded into the Linear that follows it (parameter preprocessing, once per weight version):
    LN(x) W^T + b = rstd (x Wf^T) - rstd mean c + bf  with  Wf = bf16(W * gamma), c = rowsum(Wf) (of the ROUNDED weight, so the
    mean term cancels exactly what the GEMM accumulates) and bf = b + W beta.  Returns (Wf bf16 [N, K], bf fp32 [N], c fp32 [N])."""
    w = weight.detach().float()
    wf = ops.pack_weight(w * ln.weight.detach().float()[None, :], False)
    c = wf.float().sum(dim=1).contiguous()
    bf = (bias.detach().float() + w @ ln.bias.detach().float()).contiguous()
    return wf, bf, c


class _PackedBlocks:
    """bf16 (or hi|lo split) copies of a ``Transformer``'s weights in the layout the GEMM kernels read (derived cache).
    ``fold_ln``: ln_1 / ln_2 are folded into in_proj / c_fc (bf16 ViT path, see ``_fold_layernorm``)."""

    def __init__(self, transformer: Transformer, precise: bool, fold_ln: bool = False):
        self.layers = []
        for blk in transformer.resblocks:
            d = dict(
                ln1_g=_f32(blk.ln_1.weight), ln1_b=_f32(blk.ln_1.bias), ln2_g=_f32(blk.ln_2.weight), ln2_b=_f32(blk.ln_2.bias),
                out_w=ops.pack_weight(blk.attn.out_proj.weight, precise), out_b=_f32(blk.attn.out_proj.bias),
                pj_w=ops.pack_weight(blk.mlp.c_proj.weight, precise), pj_b=_f32(blk.mlp.c_proj.bias),
            )
            if fold_ln and not precise:
                d["in_w"], d["in_b"], d["ln1_c"] = _fold_layernorm(blk.ln_1, blk.attn.in_proj_weight, blk.attn.in_proj_bias)
                d["fc_w"], d["fc_b"], d["ln2_c"] = _fold_layernorm(blk.ln_2, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias)
            else:
                d.update(in_w=ops.pack_weight(blk.attn.in_proj_weight, precise), in_b=_f32(blk.attn.in_proj_bias),
                         fc_w=ops.pack_weight(blk.mlp.c_fc.weight, precise), fc_b=_f32(blk.mlp.c_fc.bias))
            self.layers.append(d)


# ================ CLIP ViT image encoder ================ #
@BACKBONES.register_module()
class CLIPVisionTransformer(nn.Module):
    """CLIP ViT backbone, drop-in for reference models.py:378-597 (same kwargs, state_dict keys, list-of-NCHW output).

    Extra opt-in kwarg: ``precision`` ("bf16" | "fp32").  ``forward`` requires a CUDA (sm_100) input.
    """

    def __init__(self, input_resolution: int = 224, patch_size: int = 16, width: int = 768, layers: int = 12,
                 heads: int = 12, output_dim: int = 768, drop_path_rate: float = 0.0, out_indices=None,
                 pretrained: str = None, precision: str = None, **kwargs):
        super().__init__()
        self.pretrained = pretrained
        self.input_resolution = input_resolution
        self.patch_size = patch_size
        self.width = width
        self.heads = heads
        self.output_dim = width
        self.layers = layers
        self.precision = precision or default_precision()
        # bf16 path: fold ln_1 / ln_2 into the QKV / c_fc GEMMs (no stand-alone LayerNorm pass inside the blocks; encoder
        # 10.55 -> 10.05 ms at B = 16 on the same box).  DENSECLIP_B200_LN_FOLD=0 keeps the separate LayerNorm kernels (A/B only)
        self.ln_fold = os.environ.get("DENSECLIP_B200_LN_FOLD", "1") != "0"
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.grid_size = input_resolution // patch_size
        seq_len = self.grid_size ** 2 + 1
        self.positional_embedding = nn.Parameter(scale * torch.randn(seq_len, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, drop_path_rate=drop_path_rate)
        self.ln_post = LayerNorm(width)
        self._clip_proj_dim = 512
        self.proj = nn.Parameter(scale * torch.randn(width, self._clip_proj_dim))  # unused in forward (checkpoint key)
        if out_indices is None:
            self.out_indices = [layers - 1]
        else:
            if not isinstance(out_indices, (list, tuple)):
                raise TypeError("out_indices must be list or tuple")
            for i in out_indices:
                if not 0 <= i < layers:
                    raise ValueError(f"Index {i} in out_indices is out of range for {layers} layers.")
            self.out_indices = sorted(list(set(out_indices)))
        self._native = {}  # per-device native encoder objects + packed weights
        self.init_weights()

    # ---- initialisation / checkpoint ingestion (reference models.py:447-512) ----
    def _init_weights_default(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def init_weights(self, pretrained=None):
        pretrained = pretrained or self.pretrained
        if isinstance(pretrained, str):
            checkpoint = torch.jit.load(pretrained, map_location='cpu').float().state_dict()
            state_dict = OrderedDict((k[len('visual.'):], v) for k, v in checkpoint.items() if k.startswith('visual.'))
            pe = state_dict.get('positional_embedding')
            if pe is not None and pe.shape != self.positional_embedding.shape:
                n0 = pe.shape[0] - 1
                g_old = int(np.sqrt(n0))
                if g_old * g_old != n0:
                    state_dict.pop('positional_embedding')
                else:
                    # checkpoint-time resize to this module's stored grid (bilinear, align_corners=False), done once
                    # on the host with numpy so that no torch compute op is involved
                    state_dict['positional_embedding'] = torch.from_numpy(
                        _resize_pos_embed_np(pe.numpy(), g_old, self.grid_size))
            if 'proj' in state_dict and self.proj.shape != state_dict['proj'].shape:
                state_dict.pop('proj')
            msg = self.load_state_dict(state_dict, strict=False)
            if msg.missing_keys:
                logger.warning("ViT missing keys: %s", msg.missing_keys)
        else:
            self.apply(self._init_weights_default)

    # ---- native plumbing ----
    def _native_state(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        precise = self.precision == "fp32"
        fold = bool(self.ln_fold) and not precise
        st = self._native.get(idx)
        ver = (_param_versions(self), precise, fold)
        if st is not None and st["ver"] == ver:
            return st
        if st is not None:
            _lib.lib().dclip_vit_destroy(st["vit"])
        h = _lib.handle(idx)
        cfg = _lib.VitConfig(self.width, self.layers, self.heads, self.patch_size, self.grid_size, int(precise), int(fold))
        vit = C.c_void_p()
        _lib.check(h, _lib.lib().dclip_vit_create(h, C.byref(cfg), C.byref(vit)))
        packed = _PackedBlocks(self.transformer, precise, fold_ln=fold)
        keep = dict(
            conv1=ops.pack_weight(self.conv1.weight, precise, pad_cols_to=64 if precise else 8),   # = VitEncoder::kp()
            cls=_f32(self.class_embedding), pos=_f32(self.positional_embedding),
            lpg=_f32(self.ln_pre.weight), lpb=_f32(self.ln_pre.bias), log=_f32(self.ln_post.weight), lob=_f32(self.ln_post.bias),
        )
        w = _lib.VitWeights()
        w.conv1_w, w.class_embedding, w.positional_embedding = keep["conv1"].data_ptr(), keep["cls"].data_ptr(), keep["pos"].data_ptr()
        w.ln_pre_g, w.ln_pre_b, w.ln_post_g, w.ln_post_b = (keep[k].data_ptr() for k in ("lpg", "lpb", "log", "lob"))
        arrays = {}

        def arr(field, key):
            a = (C.c_void_p * self.layers)(*[l[key].data_ptr() for l in packed.layers])
            arrays[field] = a
            setattr(w, field, C.cast(a, C.POINTER(C.c_void_p)))

        for field, key in (("ln1_g", "ln1_g"), ("ln1_b", "ln1_b"), ("ln2_g", "ln2_g"), ("ln2_b", "ln2_b"),
                           ("in_proj_w", "in_w"), ("in_proj_b", "in_b"), ("out_proj_w", "out_w"), ("out_proj_b", "out_b"),
                           ("fc_w", "fc_w"), ("fc_b", "fc_b"), ("proj_w", "pj_w"), ("proj_b", "pj_b")):
            arr(field, key)
        if fold:
            arr("ln1_c", "ln1_c")
            arr("ln2_c", "ln2_c")
        _lib.check(h, _lib.lib().dclip_vit_set_weights(vit, C.byref(w)))
        st = dict(ver=ver, vit=vit, h=h, packed=packed, keep=keep, arrays=arrays, ws={}, tok={})
        self._native[idx] = st
        return st

    def forward_native(self, x: torch.Tensor, *, taps_nchw=True, taps_tokens_bf16=False, last_tokens=False):
        """Run the encoder. Returns dict(nchw=[...], tokens_bf16=[...], last_tokens=fp32 [B,Ntok,D] or None, grid=(gh,gw)).
        ``nchw`` follows the reference contract: one fp32 [B, width, gh, gw] per out_index, ln_post on the last layer."""
        if not x.is_cuda:
            raise _lib.DclipError("CLIPVisionTransformer (B200-native) needs a CUDA input; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected [B,3,H,W] input, got {tuple(x.shape)}")
        x = x.float().contiguous() if (x.dtype != torch.float32 or not x.is_contiguous()) else x
        B, _, H, W = x.shape
        st = self._native_state(x.device)
        gh, gw = H // self.patch_size, W // self.patch_size
        ntok = gh * gw + 1
        if gh * gw != self.grid_size ** 2 and int(np.sqrt(self.positional_embedding.shape[0] - 1)) ** 2 != self.positional_embedding.shape[0] - 1:
            raise ValueError("stored positional embedding is not a square grid; cannot interpolate")
        key = (B, H, W)
        ws = st["ws"].get(key)
        if ws is None:
            nbytes = C.c_size_t()
            _lib.check(st["h"], _lib.lib().dclip_vit_workspace_bytes(st["vit"], B, H, W, C.byref(nbytes)))
            # a few workspaces stay cached (alternating input shapes); a captured CUDA graph keeps its own reference to the
            # one it was recorded with (returned as "workspace"), so dropping the oldest entry here never frees memory a
            # live graph still writes to
            while len(st["ws"]) >= 4:
                old = next(iter(st["ws"]))
                st["ws"].pop(old)
                st["tok"].pop(old, None)
            ws = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device=x.device)
            st["ws"][key] = ws
        ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        n = len(self.out_indices)
        nchw = [torch.empty(B, self.width, gh, gw, dtype=torch.float32, device=x.device) for _ in range(n)] if taps_nchw else []
        # one contiguous [n_taps, B, Ntok, D] buffer: the neck convolves all taps in a single grouped launch.  It is cached per
        # input shape (like the workspace) and OVERWRITTEN by the next call: with ln_fold the tap of layer i is also the A
        # operand of layer i+1's QKV GEMM, so its address is baked into that GEMM's tensor map (plans are built once)
        tok_all = None
        if taps_tokens_bf16:
            tok_all = st["tok"].get(key)
            if tok_all is None or tok_all.device != x.device:
                tok_all = torch.empty(n, B, ntok, self.width, dtype=torch.bfloat16, device=x.device)
                st["tok"][key] = tok_all
        tok = [tok_all[i] for i in range(n)] if taps_tokens_bf16 else []
        last = torch.empty(B, ntok, self.width, dtype=torch.float32, device=x.device) if last_tokens else None
        o = _lib.VitOutputs()
        layers_arr = (C.c_int * n)(*self.out_indices)
        nchw_arr = (C.c_void_p * n)(*[t.data_ptr() for t in nchw]) if nchw else None
        tok_arr = (C.c_void_p * n)(*[t.data_ptr() for t in tok]) if tok else None
        o.n_taps = n
        o.tap_layers = C.cast(layers_arr, C.POINTER(C.c_int))
        o.taps_nchw = C.cast(nchw_arr, C.POINTER(C.c_void_p)) if nchw else None
        o.taps_tokens_bf16 = C.cast(tok_arr, C.POINTER(C.c_void_p)) if tok else None
        o.last_tokens_f32 = last.data_ptr() if last is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(st["h"], _lib.lib().dclip_vit_forward(st["vit"], C.c_void_p(x.data_ptr()), B, H, W, C.c_void_p(ws_ptr),
                                                          C.c_size_t(nbytes_of(ws, ws_ptr)), C.byref(o), stream))
        return dict(nchw=nchw, tokens_bf16=tok, tokens_bf16_stacked=tok_all, last_tokens=last, grid=(gh, gw), workspace=(ws, tok_all))

    def forward(self, x: torch.Tensor):
        """[B,3,H,W] -> list of fp32 [B, width, H//ps, W//ps], one per out_index (reference models.py:543-597)."""
        return self.forward_native(x, taps_nchw=True)["nchw"]

    def interpolate_pos_encoding(self, x, H, W):
        """Reference models.py:514-540, as a native kernel. Returns [1+H*W, C]."""
        n_loaded = self.positional_embedding.shape[0] - 1
        if x.shape[1] - 1 == n_loaded:
            return self.positional_embedding.to(x.dtype)
        g0 = int(np.sqrt(n_loaded))
        if g0 * g0 != n_loaded:
            return self.positional_embedding.to(x.dtype)
        return ops.posemb_interp(_f32(self.positional_embedding), g0, H, W).to(x.dtype)


def nbytes_of(ws: torch.Tensor, aligned_ptr: int) -> int:
    return ws.numel() - (aligned_ptr - ws.data_ptr())


def _resize_pos_embed_np(pe: np.ndarray, g_old: int, g_new: int) -> np.ndarray:
    """Bilinear (align_corners=False) resize of a [1+g_old^2, D] positional embedding on the host (checkpoint load)."""
    cls, grid = pe[:1], pe[1:].reshape(g_old, g_old, -1).astype(np.float32)

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5), 0).astype(np.float32)
        i0 = np.minimum(src.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0).astype(np.float32)
        return i0, i1, np.float32(1) - l1, l1

    y0, y1, ly0, ly1 = axis(g_old, g_new)
    x0, x1, lx0, lx1 = axis(g_old, g_new)
    top = grid[y0][:, x0] * lx0[None, :, None] + grid[y0][:, x1] * lx1[None, :, None]
    bot = grid[y1][:, x0] * lx0[None, :, None] + grid[y1][:, x1] * lx1[None, :, None]
    out = top * ly0[:, None, None] + bot * ly1[:, None, None]
    return np.concatenate([cls, out.reshape(g_new * g_new, -1)], 0).astype(pe.dtype)


# ================ text towers ================ #
def _text_transformer_native(x: torch.Tensor, packed: _PackedBlocks, heads: int, n_seq: int, seq_len: int, passes: int):
    """CLIP text transformer with the additive causal mask (reference models.py:836-842), on fp32 tokens x [n_seq*seq_len, D].
    Runs in split-bf16 precision (weights-only, computed once and cached)."""
    D = x.shape[1]
    for _ in range(passes):
        for l in packed.layers:
            _, h = ops.layernorm(x, l["ln1_g"], l["ln1_b"], want_bf16=True, split=True)
            qkv, _ = ops.gemm(h, l["in_w"], split_in=True, bias=l["in_b"], want_f32=True)
            q3 = qkv.view(n_seq, seq_len, 3 * D)
            att = torch.empty(n_seq, seq_len, 2 * D, dtype=torch.bfloat16, device=x.device)   # hi|lo written by the kernel
            ops.attention_small(q3, q3, q3, B=n_seq, H=heads, q_first=0, q_count=seq_len, Nk=seq_len, q_col0=0, k_col0=D,
                                v_col0=2 * D, scale=64 ** -0.5, out=att, causal=True, out_split_off=D)
            ops.gemm(att.view(-1, 2 * D), l["out_w"], split_in=True, bias=l["out_b"], residual=x, out_f32=x)
            _, h = ops.layernorm(x, l["ln2_g"], l["ln2_b"], want_bf16=True, split=True)
            _, g = ops.gemm(h, l["fc_w"], split_in=True, bias=l["fc_b"], act="quickgelu_precise", want_bf16=True, split_out=True)
            ops.gemm(g, l["pj_w"], split_in=True, bias=l["pj_b"], residual=x, out_f32=x)
    return x


class _TextTowerBase(nn.Module):
    def _build(self, context_length, vocab_size, transformer_width, transformer_heads, transformer_layers, embed_dim):
        self.context_length = context_length
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.embed_dim = embed_dim
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self._cache = {}

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def _load_clip_text(self, pretrained):
        checkpoint = torch.jit.load(pretrained, map_location='cpu').float().state_dict()
        state_dict = {}
        for k, v in checkpoint.items():
            if k.startswith('transformer.'):
                state_dict[k] = v
            if k in ('positional_embedding', 'text_projection') or k.startswith('token_embedding') or k.startswith('ln_final'):
                if k == 'positional_embedding' and v.size(0) > self.context_length:
                    v = v[:self.context_length]
                if k == 'text_projection' and v.shape != self.text_projection.shape:
                    continue
                state_dict[k] = v
        return self.load_state_dict(state_dict, strict=False)

    def _encode(self, tokens_f32: torch.Tensor, n_seq: int, seq_len: int, eos_index: torch.Tensor):
        """tokens_f32 [n_seq*seq_len, D] (embeddings + positional) -> [n_seq, embed_dim]."""
        if self.transformer.heads * 64 != self.transformer.width:
            raise _lib.DclipError("native text tower needs head_dim 64")
        packed = _PackedBlocks(self.transformer, True)
        # reference quirk: Transformer.forward applies the whole stack twice (models.py:305-307)
        x = _text_transformer_native(tokens_f32, packed, self.transformer.heads, n_seq, seq_len, passes=2)
        y, _ = ops.layernorm(x, _f32(self.ln_final.weight), _f32(self.ln_final.bias), want_f32=True)
        rows = (torch.arange(n_seq, device=y.device) * seq_len + eos_index.to(y.device)).to(torch.int64)
        sel = y.index_select(0, rows)  # gather of n_seq rows (indexing, not arithmetic)
        wp = ops.pack_weight(self.text_projection.detach().t().contiguous(), True)
        out, _ = ops.gemm(ops.split_bf16(sel), wp, split_in=True, want_f32=True)
        return out


@BACKBONES.register_module()
class CLIPTextEncoder(_TextTowerBase):
    """Reference models.py:600-714. ``forward(text[K, L]) -> [K, embed_dim]``."""

    def __init__(self, context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                 transformer_layers=12, embed_dim=512, pretrained=None, **kwargs):
        super().__init__()
        self.pretrained = pretrained
        self._build(context_length, vocab_size, transformer_width, transformer_heads, transformer_layers, embed_dim)
        self._output_dim = embed_dim
        self.init_weights()

    def _init_weights_default(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, std=0.02)

    def init_weights(self, pretrained=None):
        pretrained = pretrained or self.pretrained
        if isinstance(pretrained, str):
            self._load_clip_text(pretrained)
        else:
            self.apply(self._init_weights_default)

    def forward(self, text):
        key = (_param_versions(self), tuple(text.cpu().flatten().tolist()))
        if self._cache.get("key") == key:
            return self._cache["val"]
        K, L = text.shape
        dev = self.token_embedding.weight.device
        emb = self.token_embedding.weight.detach().index_select(0, text.reshape(-1).to(dev))  # row gather
        pos = _f32(self.positional_embedding)[:L]
        x = ops.gamma_residual(pos.repeat(K, 1), torch.ones(1, device=dev), _f32(emb))  # emb + pos
        out = self._encode(x, K, L, text.cpu().argmax(dim=-1))
        self._cache = {"key": key, "val": out}
        return out


@BACKBONES.register_module()
class CLIPTextContextEncoder(_TextTowerBase):
    """Reference models.py:785-864. ``forward(text[K, N1], context[Bc, N2, C]) -> [Bc, K, embed_dim]``; the learnable
    context is spliced after the SOT token. Input-independent at inference, so the result is cached per weight version."""

    def __init__(self, context_length=22, vocab_size=49408, transformer_width=512, transformer_heads=8,
                 transformer_layers=12, embed_dim=512, out_dim=256, pretrained=None, **kwargs):
        super().__init__()
        self.pretrained = pretrained
        self._build(context_length, vocab_size, transformer_width, transformer_heads, transformer_layers, embed_dim)

    def init_weights(self, pretrained=None):
        pretrained = pretrained or self.pretrained
        if isinstance(pretrained, str):
            self._load_clip_text(pretrained)

    def forward(self, text, context):
        key = (_param_versions(self), context.data_ptr(), context._version, tuple(text.cpu().flatten().tolist()))
        if self._cache.get("key") == key:
            return self._cache["val"]
        dev = self.token_embedding.weight.device
        K, N1 = text.shape
        Bc, N2, Cw = context.shape
        L = N1 + N2
        emb = self.token_embedding.weight.detach().index_select(0, text.reshape(-1).to(dev)).view(K, N1, Cw)
        # sequence = [SOT, context(N2), rest of text]: pure row placement (no arithmetic)
        seq = torch.empty(Bc, K, L, Cw, dtype=torch.float32, device=dev)
        seq[:, :, 0:1] = emb[None, :, 0:1]
        seq[:, :, 1:1 + N2] = _f32(context)[:, None]
        seq[:, :, 1 + N2:] = emb[None, :, 1:]
        pos = _f32(self.positional_embedding)
        x = ops.gamma_residual(pos.repeat(Bc * K, 1), torch.ones(1, device=dev), seq.view(-1, Cw))
        eos = (text.cpu().argmax(dim=-1) + N2).reshape(1, K).expand(Bc, K).reshape(-1)
        out = self._encode(x, Bc * K, L, eos).view(Bc, K, self.embed_dim)
        self._cache = {"key": key, "val": out}
        return out


# ================ ContextDecoder ================ #
class Attention(nn.Module):
    """Reference models.py:311-344 (container): q/k/v projections without bias, output projection with bias."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q_proj = nn.Linear(dim, dim, bias=qkv_bias)
        self.k_proj = nn.Linear(dim, dim, bias=qkv_bias)
        self.v_proj = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, q, k, v):
        _native_only("Attention")


class TransformerDecoderLayer(nn.Module):
    """Reference models.py:346-375 (container)."""

    def __init__(self, d_model, nhead, dropout=0.1):
        super().__init__()
        self.self_attn = Attention(d_model, nhead, proj_drop=dropout)
        self.cross_attn = Attention(d_model, nhead, proj_drop=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.mlp = nn.Sequential(nn.Linear(d_model, d_model * 4), nn.GELU(), nn.Dropout(dropout), nn.Linear(d_model * 4, d_model))

    def forward(self, x, mem):
        _native_only("TransformerDecoderLayer")


def _cat_bias(*linears):
    if all(l.bias is None for l in linears):
        return None
    return torch.cat([_f32(l.bias) if l.bias is not None else torch.zeros(l.out_features, device=l.weight.device) for l in linears])


@BACKBONES.register_module()
class ContextDecoder(nn.Module):
    """Text-class queries cross-attend to visual tokens (reference models.py:867-916).

    ``forward(text[B,K,C], visual[B,N,C]) -> [B,K,C]``.  Always computed with split-bf16 GEMMs and fp32 attention: it is
    <0.5% of the FLOPs and feeds the ill-conditioned score-map argmax.  Dropout is inference-mode identity.
    """

    def __init__(self, transformer_width=256, transformer_heads=4, transformer_layers=6, visual_dim=1024, dropout=0.1,
                 **kwargs):
        super().__init__()
        self.visual_dim = visual_dim
        self.transformer_width = transformer_width
        self.memory_proj = nn.Sequential(nn.LayerNorm(visual_dim), nn.Linear(visual_dim, transformer_width),
                                         nn.LayerNorm(transformer_width))
        self.text_proj = nn.Sequential(nn.LayerNorm(visual_dim), nn.Linear(visual_dim, transformer_width))
        self.decoder = nn.ModuleList([TransformerDecoderLayer(transformer_width, transformer_heads, dropout)
                                      for _ in range(transformer_layers)])
        self.out_proj = nn.Sequential(nn.LayerNorm(transformer_width), nn.Linear(transformer_width, visual_dim))
        self._packed = None
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _pack(self):
        ver = _param_versions(self)
        if self._packed is not None and self._packed["ver"] == ver:
            return self._packed
        pw = lambda w: ops.pack_weight(w, True)  # noqa: E731
        p = dict(ver=ver, mem_w=pw(self.memory_proj[1].weight), txt_w=pw(self.text_proj[1].weight),
                 out_w=pw(self.out_proj[1].weight), layers=[])
        for l in self.decoder:
            p["layers"].append(dict(
                sa_qkv=pw(torch.cat([l.self_attn.q_proj.weight, l.self_attn.k_proj.weight, l.self_attn.v_proj.weight], 0)),
                sa_qkv_b=_cat_bias(l.self_attn.q_proj, l.self_attn.k_proj, l.self_attn.v_proj),
                sa_o=pw(l.self_attn.proj.weight),
                ca_q=pw(l.cross_attn.q_proj.weight), ca_q_b=_cat_bias(l.cross_attn.q_proj),
                ca_kv=pw(torch.cat([l.cross_attn.k_proj.weight, l.cross_attn.v_proj.weight], 0)),
                ca_kv_b=_cat_bias(l.cross_attn.k_proj, l.cross_attn.v_proj),
                ca_o=pw(l.cross_attn.proj.weight), fc=pw(l.mlp[0].weight), pj=pw(l.mlp[3].weight)))
        if len(self.decoder):
            # the memory does not change across layers: K/V projections of ALL layers run as one GEMM (N = layers * 2 * Wd)
            p["ca_kv_all"] = pw(torch.cat([torch.cat([l.cross_attn.k_proj.weight, l.cross_attn.v_proj.weight], 0) for l in self.decoder], 0))
            kvb = [lp["ca_kv_b"] for lp in p["layers"]]
            p["ca_kv_all_b"] = None if any(b is None for b in kvb) else torch.cat(kvb).contiguous()
        self._packed = p
        return p

    def forward(self, text, visual):
        if self.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules()):
            logger.debug("ContextDecoder native path is inference-only: dropout is not applied")
        B, N, Cv = visual.shape
        K = text.shape[1]
        Wd = self.transformer_width
        heads = self.decoder[0].self_attn.num_heads if len(self.decoder) else 4
        if Wd != heads * 64:
            raise _lib.DclipError("native ContextDecoder needs head_dim 64")
        pk = self._pack()
        ln = lambda m: (_f32(m.weight), _f32(m.bias), m.eps)  # noqa: E731
        vis2 = visual.reshape(B * N, Cv)
        vis2 = vis2 if vis2.dtype == torch.float32 and vis2.stride(1) == 1 else vis2.float().contiguous()
        # memory_proj: LN -> Linear -> LN
        g, b, e = ln(self.memory_proj[0])
        _, hm = ops.layernorm(vis2, g, b, e, want_bf16=True, split=True)
        m1, _ = ops.gemm(hm, pk["mem_w"], split_in=True, bias=_f32(self.memory_proj[1].bias), want_f32=True)
        g, b, e = ln(self.memory_proj[2])
        _, mem = ops.layernorm(m1, g, b, e, want_bf16=True, split=True)          # [B*N, 2*Wd] hi|lo
        # text_proj: LN -> Linear
        t2 = text.reshape(B * K, Cv).float().contiguous()
        g, b, e = ln(self.text_proj[0])
        _, ht = ops.layernorm(t2, g, b, e, want_bf16=True, split=True)
        x, _ = ops.gemm(ht, pk["txt_w"], split_in=True, bias=_f32(self.text_proj[1].bias), want_f32=True)  # [B*K, Wd]
        dev = x.device
        L = len(self.decoder)
        kv_all = None
        if L and (pk["ca_kv_all_b"] is not None or all(lp["ca_kv_b"] is None for lp in pk["layers"])):
            kv_all, _ = ops.gemm(mem, pk["ca_kv_all"], split_in=True, bias=pk["ca_kv_all_b"], want_f32=True)   # [B*N, L*2*Wd]
            kv_all = kv_all.view(B, N, L * 2 * Wd)
        att = torch.empty(B, K, 2 * Wd, dtype=torch.bfloat16, device=dev)   # attention output, written as hi|lo by the kernel
        for li, (l, lp) in enumerate(zip(self.decoder, pk["layers"])):
            sc = l.self_attn.scale
            # self attention on the K text tokens
            g, b, e = ln(l.norm1)
            _, h1 = ops.layernorm(x, g, b, e, want_bf16=True, split=True)
            qkv, _ = ops.gemm(h1, lp["sa_qkv"], split_in=True, bias=lp["sa_qkv_b"], want_f32=True)
            q3 = qkv.view(B, K, 3 * Wd)
            ops.attention_small(q3, q3, q3, B=B, H=heads, q_first=0, q_count=K, Nk=K, q_col0=0, k_col0=Wd, v_col0=2 * Wd,
                                scale=sc, out=att, out_split_off=Wd)
            ops.gemm(att.view(-1, 2 * Wd), lp["sa_o"], split_in=True, bias=_f32(l.self_attn.proj.bias), residual=x, out_f32=x)
            # cross attention: text queries over the visual memory
            sc = l.cross_attn.scale
            g, b, e = ln(l.norm2)
            _, h2 = ops.layernorm(x, g, b, e, want_bf16=True, split=True)
            q, _ = ops.gemm(h2, lp["ca_q"], split_in=True, bias=lp["ca_q_b"], want_f32=True)
            if kv_all is not None:
                kv3, kc0 = kv_all, li * 2 * Wd
            else:
                kv, _ = ops.gemm(mem, lp["ca_kv"], split_in=True, bias=lp["ca_kv_b"], want_f32=True)  # [B*N, 2*Wd]
                kv3, kc0 = kv.view(B, N, 2 * Wd), 0
            ops.attention_small(q.view(B, K, Wd), kv3, kv3, B=B, H=heads, q_first=0, q_count=K, Nk=N, q_col0=0, k_col0=kc0,
                                v_col0=kc0 + Wd, scale=sc, out=att, out_split_off=Wd)
            ops.gemm(att.view(-1, 2 * Wd), lp["ca_o"], split_in=True, bias=_f32(l.cross_attn.proj.bias), residual=x, out_f32=x)
            # MLP with exact (erf) GELU
            g, b, e = ln(l.norm3)
            _, h3 = ops.layernorm(x, g, b, e, want_bf16=True, split=True)
            _, gg = ops.gemm(h3, lp["fc"], split_in=True, bias=_f32(l.mlp[0].bias), act="gelu", want_bf16=True, split_out=True)
            ops.gemm(gg, lp["pj"], split_in=True, bias=_f32(l.mlp[3].bias), residual=x, out_f32=x)
        g, b, e = ln(self.out_proj[0])
        _, ho = ops.layernorm(x, g, b, e, want_bf16=True, split=True)
        out, _ = ops.gemm(ho, pk["out_w"], split_in=True, bias=_f32(self.out_proj[1].bias), want_f32=True)
        return out.view(B, K, Cv)


# ================ ViTFeatureFusionNeck ================ #
def fold_bn(conv_w: torch.Tensor, bn: nn.BatchNorm2d, conv_b=None):
    """Fold an eval-mode BatchNorm into the preceding conv: returns (weight [N, ...], bias [N]) in fp32.
    Parameter preprocessing (done once per weight version), not part of the forward data path."""
    w = conv_w.detach().double()
    inv = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b0 = conv_b.detach().double() if conv_b is not None else torch.zeros_like(inv)
    wf = w * inv.view(-1, *([1] * (w.dim() - 1)))
    bf = bn.bias.detach().double() + (b0 - bn.running_mean.detach().double()) * inv
    return wf.float(), bf.float().contiguous()


def conv3x3_weight_to_gemm(w: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] -> [N, 9*C] with K order (ky, kx, c), matching the implicit-conv operand order."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def conv3x3_tokens(tokens: torch.Tensor, row0: int, gh: int, gw: int, C_in: int, w_packed: torch.Tensor, *, split_in: bool,
                   bias, act, out_bf16=None, out_f32=None, want_f32=False, want_bf16=False, split_out=False):
    """3x3/pad-1 conv over token-major activations [B, rows, ld] (bf16, hi|lo when split_in) as an implicit GEMM; falls
    back to an explicit gather when the grid cannot be tiled by 128 pixels."""
    B = tokens.shape[0]
    M = B * gh * gw
    tileable = (gh * gw) % 128 == 0 and (128 % gw == 0 or gw % 128 == 0) and C_in % 64 == 0
    if tileable:
        a = tokens[:, row0:, :]
        a2 = a.as_strided((M, a.shape[2]), (a.stride(1), 1), a.storage_offset())  # row view for pointer/ld only
        return ops.gemm(a2, w_packed, K=9 * C_in, split_in=split_in, bias=bias, act=act, out_bf16=out_bf16, out_f32=out_f32,
                        want_f32=want_f32, want_bf16=want_bf16, split_out=split_out, M=M,
                        conv=dict(C=C_in, gw=gw, gh=gh, B=B, a_bs=tokens.stride(0)))
    if split_in:
        hi = ops.conv3x3_gather(tokens[:, :, :C_in], row0=row0, hh=gh, ww=gw, channels=C_in)
        lo = ops.conv3x3_gather(tokens[:, :, C_in:], row0=row0, hh=gh, ww=gw, channels=C_in)
        a2 = torch.cat([hi, lo], 1)
    else:
        a2 = ops.conv3x3_gather(tokens, row0=row0, hh=gh, ww=gw, channels=C_in)
    return ops.gemm(a2, w_packed, K=9 * C_in, split_in=split_in, bias=bias, act=act, out_bf16=out_bf16, out_f32=out_f32,
                    want_f32=want_f32, want_bf16=want_bf16, split_out=split_out)


@BACKBONES.register_module()
class ViTFeatureFusionNeck(nn.Module):
    """Reference models.py:717-782: per-tap 3x3 conv+BN+ReLU, channel concat, 1x1 conv+BN+ReLU -> ``[fused]``.

    Natively each 3x3 conv is an implicit GEMM (TMA-gathered operand, BN folded into weight/bias, ReLU in the epilogue)
    that writes straight into its slice of the concatenated buffer; the fusion conv is one more GEMM.  Eval-mode only.
    """

    def __init__(self, in_channels_list, out_channels, inter_channels=None, precision: str = None):
        super().__init__()
        if not isinstance(in_channels_list, (list, tuple)):
            raise TypeError("in_channels_list must be a list or tuple")
        if inter_channels is None:
            inter_channels = out_channels
        self.num_inputs = len(in_channels_list)
        self.in_channels_list = list(in_channels_list)
        self.inter_channels = inter_channels
        self.out_channels = out_channels
        self.precision = precision or default_precision()
        self.process_layers = nn.ModuleList(ConvBNReLU(c, inter_channels, kernel_size=3, padding=1) for c in in_channels_list)
        self.fusion_layer = ConvBNReLU(inter_channels * self.num_inputs, out_channels, kernel_size=1, padding=0)
        self._packed = None
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def _pack(self):
        precise = self.precision == "fp32"
        ver = (_param_versions(self), precise)
        if self._packed is not None and self._packed["ver"] == ver:
            return self._packed
        p = dict(ver=ver, w=[], b=[])
        for layer in self.process_layers:
            wf, bf = fold_bn(layer[0].weight, layer[1])
            p["w"].append(ops.pack_weight(conv3x3_weight_to_gemm(wf), precise))
            p["b"].append(bf)
        wf, bf = fold_bn(self.fusion_layer[0].weight, self.fusion_layer[1])
        p["fw"], p["fb"] = ops.pack_weight(wf.reshape(wf.shape[0], -1), precise), bf
        if not precise and len(set(self.in_channels_list)) == 1:   # all taps alike: one grouped conv launch
            p["w_all"] = torch.cat(p["w"], 0).contiguous()
            p["b_all"] = torch.cat(p["b"], 0).contiguous()
        self._packed = p
        return p

    def forward_tokens(self, tokens_list, row0: int, gh: int, gw: int):
        """tokens_list: bf16 token-major [B, rows, C(x2 hi|lo in fp32 mode)] per tap. Returns fused fp32 [B*gh*gw, out]
        plus its bf16 (hi|lo) copy for the heads."""
        if self.training:
            raise _lib.DclipError("ViTFeatureFusionNeck native path is inference-only (BatchNorm uses running stats); call .eval()")
        if len(tokens_list) != self.num_inputs:
            raise ValueError(f"Fusion neck received {len(tokens_list)} inputs, expected {self.num_inputs}")
        precise = self.precision == "fp32"
        pk = self._pack()
        B = tokens_list[0].shape[0]
        M = B * gh * gw
        s = 2 if precise else 1
        ic = self.inter_channels
        tot = ic * self.num_inputs
        cat = torch.empty(M, tot * s, dtype=torch.bfloat16, device=tokens_list[0].device)
        stacked = _as_stacked(tokens_list)
        C_in = self.in_channels_list[0]
        grouped = ("w_all" in pk and stacked is not None and ic in (64, 128, 256) and (gh * gw) % 128 == 0
                   and (128 % gw == 0 or gw % 128 == 0) and C_in % 64 == 0)
        if grouped:
            a = stacked[0, :, row0:, :]
            a2 = a.as_strided((M, a.shape[2]), (a.stride(1), 1), a.storage_offset())
            ops.gemm(a2, pk["w_all"], K=9 * C_in, bias=pk["b_all"], act="relu", out_bf16=cat, M=M, block_n=ic,
                     conv=dict(C=C_in, gw=gw, gh=gh, B=B, a_bs=stacked.stride(1), G=self.num_inputs, a_gs=stacked.stride(0)))
        for i, t in enumerate(tokens_list if not grouped else []):
            dst = cat[:, i * ic:]  # column slice; the lo half (fp32 mode) lands `tot` columns further right
            if precise:
                _gemm_split_out_at(t, row0, gh, gw, self.in_channels_list[i], pk["w"][i], pk["b"][i], dst, tot)
            else:
                conv3x3_tokens(t, row0, gh, gw, self.in_channels_list[i], pk["w"][i], split_in=False, bias=pk["b"][i], act="relu",
                               out_bf16=dst)
        fused, fused_b = ops.gemm(cat, pk["fw"], K=tot, split_in=precise, bias=pk["fb"], act="relu", want_f32=True,
                                  want_bf16=True, split_out=precise)
        return fused, fused_b

    def forward(self, features):
        """API-compatible entry: list of NCHW fp32 -> [fused NCHW fp32]."""
        if len(features) != self.num_inputs:
            logger.error("Fusion Neck received %d inputs, expected %d", len(features), self.num_inputs)
            return [features[0]] if features else []
        precise = self.precision == "fp32"
        B, _, gh, gw = features[0].shape
        toks = []
        for f in features:
            tf, tb = ops.nchw_to_tokens(f, f32=precise, bf16=not precise)
            toks.append(ops.split_bf16(tf.view(-1, tf.shape[2])).view(B, gh * gw, -1) if precise else tb)
        fused, _ = self.forward_tokens(toks, 0, gh, gw)
        return [ops.tap_nchw(_with_dummy_cls(fused.view(B, gh * gw, -1)), gh, gw)]


def _as_stacked(tokens_list):
    """If the per-tap tensors are equally spaced views of one buffer, return it as [G, B, rows, C] (no copy), else None."""
    t0 = tokens_list[0]
    if len(tokens_list) < 2 or not all(t.shape == t0.shape and t.stride() == t0.stride() and t.dtype == t0.dtype for t in tokens_list):
        return None
    try:
        base = t0.untyped_storage().data_ptr()
        if any(t.untyped_storage().data_ptr() != base for t in tokens_list):
            return None
    except Exception:
        return None
    step = tokens_list[1].storage_offset() - t0.storage_offset()
    if step <= 0 or any(t.storage_offset() != t0.storage_offset() + i * step for i, t in enumerate(tokens_list)):
        return None
    return t0.as_strided((len(tokens_list),) + tuple(t0.shape), (step,) + tuple(t0.stride()), t0.storage_offset())


def _with_dummy_cls(tok: torch.Tensor) -> torch.Tensor:
    """[B, P, C] -> [B, 1+P, C] with an unused row 0 (layout helper for tap_nchw, which skips the CLS row)."""
    B, P, Cc = tok.shape
    out = torch.empty(B, P + 1, Cc, dtype=tok.dtype, device=tok.device)
    out[:, 1:] = tok
    return out


def _gemm_split_out_at(t, row0, gh, gw, C_in, w, bias, dst, lo_off):
    B = t.shape[0]
    M = B * gh * gw
    tileable = (gh * gw) % 128 == 0 and (128 % gw == 0 or gw % 128 == 0) and C_in % 64 == 0
    if not tileable:
        hi = ops.conv3x3_gather(t[:, :, :C_in], row0=row0, hh=gh, ww=gw, channels=C_in)
        lo = ops.conv3x3_gather(t[:, :, C_in:], row0=row0, hh=gh, ww=gw, channels=C_in)
        a2, conv = torch.cat([hi, lo], 1), None
    else:
        a = t[:, row0:, :]
        a2 = a.as_strided((M, a.shape[2]), (a.stride(1), 1), a.storage_offset())
        conv = dict(C=C_in, gw=gw, gh=gh, B=B, a_bs=t.stride(0))
    g = _lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw = a2.data_ptr(), a2.stride(0), w.data_ptr(), w.stride(0)
    g.M, g.N, g.K, g.split_in = M, w.shape[0], 9 * C_in, 1
    g.bias, g.act, g.out_scale = bias.data_ptr(), ops.ACT_RELU, 1.0
    g.out_bf16, g.ldcb, g.split_out, g.split_out_off = dst.data_ptr(), dst.stride(0), 1, lo_off
    if conv:
        g.conv_C, g.conv_gw, g.conv_gh, g.conv_B, g.a_bs = conv["C"], conv["gw"], conv["gh"], conv["B"], conv["a_bs"]
    ops._call(a2, _lib.lib().dclip_gemm, C.byref(g), ops._stream(a2))


# Not part of the ViT hot path: kept as importable names that fail loudly.
class CLIPResNet(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("CLIPResNet is outside the B200-native scope (ViT path only); use the reference for ResNet backbones")


class CLIPResNetWithAttention(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("CLIPResNetWithAttention is outside the B200-native scope (ViT path only)")

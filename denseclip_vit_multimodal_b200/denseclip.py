"""DenseCLIP segmentor, drop-in for the reference ``segmentation/denseclip/denseclip.py`` (class ``DenseCLIP`` :60):
same constructor kwargs / config keys / type strings, same ``state_dict`` keys, same ``forward`` / ``extract_feat`` /
``_process_features`` signatures and return structures -- with the whole forward running on hand-written sm_100a
kernels (no torch compute op, no CPU fallback).

Hot-path layout: the backbone hands over token-major activations (``[B, 1+P, C]``, CLS at row 0), never NCHW:
  * ``_process_features``: token mean -> global_proj; vis_proj as one GEMM whose output buffer *is* the visual
    context (row 0 of every image receives the projected global feature); ContextDecoder; gamma residual;
    fused L2-normalise + ``einsum('bchw,bkc->bkhw')`` score-map kernel.  This tail always runs with split-bf16
    (fp32-class) GEMMs: it is <1% of the FLOPs and the score-map argmax is ill-conditioned.
  * neck / heads: implicit-GEMM 3x3 convs with folded BatchNorm, 1x1 convs as GEMMs, vectorised bilinear upsample.
NCHW fp32 tensors are only materialised when a caller asks for them through the reference's list-of-NCHW API.
"""
from __future__ import annotations

import logging
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .heads import FCNHead, IdentityHead
from .models import (CLIPTextContextEncoder, CLIPTextEncoder, CLIPVisionTransformer, ContextDecoder, ViTFeatureFusionNeck,
                     _f32, _param_versions, default_precision)
from .utils import tokenize

# run the score-map / ContextDecoder branch on a side stream, concurrently with the neck / heads branch (DCLIP_OVERLAP_TAIL=0: off)
_OVERLAP_TAIL = os.environ.get("DCLIP_OVERLAP_TAIL", "1") != "0"

logger = logging.getLogger(__name__)


class DenseCLIP(nn.Module):
    """See module docstring. Extra opt-in kwargs (defaults keep reference behaviour): ``precision`` ("bf16"|"fp32"),
    ``token_ids`` (int64 [K, context_length]; skips the built-in token table)."""

    def __init__(self, backbone, text_encoder, decode_head, class_names, context_length, depth_head=None,
                 context_decoder=None, neck=None, context_feature='attention', score_concat_index=3, text_head=False,
                 tau=0.07, auxiliary_head=None, identity_head=None, train_cfg=None, test_cfg=None, token_embed_dim=512,
                 text_dim=512, clip_pretrained_path=None, precision=None, token_ids=None, **kwargs):
        super().__init__()
        self.precision = precision or default_precision()
        self.class_names = class_names
        self.num_classes = len(class_names)
        self.fixed_text_context_length = context_length
        self.context_feature = context_feature
        self.score_concat_index = score_concat_index
        self.text_head = text_head
        self.tau = tau  # stored and never used, exactly like the reference (denseclip.py:101)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.align_corners = False
        self.text_dim = text_dim

        # --- backbone (denseclip.py:111-126) ---
        bcfg = dict(backbone)
        btype = bcfg.pop('type')
        if btype == "CLIPVisionTransformer":
            bcfg.setdefault('precision', self.precision)
            self.backbone = CLIPVisionTransformer(**bcfg)
            backbone_out_channels = backbone.get('width', 768)
        elif btype in ("CLIPResNet", "CLIPResNetWithAttention"):
            raise NotImplementedError(f"{btype}: ResNet backbones are outside the B200-native scope (ViT path only)")
        else:
            raise ValueError(f"Unsupported backbone type: {btype}")

        # --- text encoder (denseclip.py:130-152) ---
        tcfg = dict(text_encoder)
        ttype = tcfg.pop('type')
        enc_dim = text_encoder.get('embed_dim')
        if enc_dim is not None and enc_dim != self.text_dim:
            self.text_dim = enc_dim
        tcfg['embed_dim'] = self.text_dim
        self.is_context_encoder = False
        if ttype == "CLIPTextEncoder":
            tcfg['context_length'] = self.fixed_text_context_length
            self.text_encoder = CLIPTextEncoder(**tcfg)
        elif ttype == "CLIPTextContextEncoder":
            if text_encoder.get('context_length') is None:
                raise ValueError("`context_length` required in CLIPTextContextEncoder config.")
            self.text_encoder = CLIPTextContextEncoder(**tcfg)
            self.is_context_encoder = True
        else:
            raise ValueError(f"Unsupported text_encoder type: {ttype}")

        # --- CLIP checkpoint ingestion (denseclip.py:156-191) ---
        if clip_pretrained_path:
            self.backbone.init_weights(clip_pretrained_path)
            self.text_encoder._load_clip_text(clip_pretrained_path)

        # --- projections (denseclip.py:195-200) ---
        self.vis_proj = None
        self.global_proj = None
        if backbone_out_channels != self.text_dim:
            self.vis_proj = nn.Conv2d(backbone_out_channels, self.text_dim, kernel_size=1)
            self.global_proj = nn.Linear(backbone_out_channels, self.text_dim)

        # --- context decoder (denseclip.py:204-211) ---
        self.context_decoder = None
        if context_decoder:
            ccfg = dict(context_decoder)
            ctype = ccfg.pop('type')
            if ctype != "ContextDecoder":
                raise ValueError(f"Unsupported context_decoder type: {ctype}")
            ccfg['visual_dim'] = self.text_dim
            self.context_decoder = ContextDecoder(**ccfg)

        # --- neck (denseclip.py:215-287) ---
        self.neck = None
        self._neck_out_keys = None
        head_in_channels = backbone_out_channels
        if neck:
            ntype = dict(neck).pop('type')
            if ntype == "ViTFeatureFusionNeck":
                out_idx = backbone.get('out_indices', [])
                if not out_idx:
                    raise ValueError("Backbone config must specify 'out_indices' when using ViTFeatureFusionNeck.")
                out_channels = neck.get('out_channels')
                if not isinstance(out_channels, int) or out_channels <= 0:
                    raise ValueError(f"Neck 'out_channels' must be a positive integer, got: {out_channels}")
                self.neck = ViTFeatureFusionNeck(in_channels_list=[backbone.get('width', 768)] * len(out_idx),
                                                 out_channels=out_channels, inter_channels=neck.get('inter_channels'),
                                                 precision=self.precision)
                head_in_channels = out_channels
            elif ntype == "FPN":
                raise NotImplementedError("FPN neck (ResNet path) is outside the B200-native scope")
            else:
                raise ValueError(f"Unsupported neck type: {ntype}")

        # --- decode head (denseclip.py:291-323) ---
        self.decode_head = None
        self._decode_head_cfg = None
        if decode_head:
            dtype_ = dict(decode_head).pop('type')
            self.align_corners = decode_head.get('align_corners', False)
            self.num_classes = decode_head.get('num_classes', self.num_classes)
            in_ch = decode_head.get('in_channels') or head_in_channels
            if dtype_ == "FPNHead":
                channels = decode_head.get('channels', 256)
                self.decode_head = FCNHead(in_channels=in_ch, channels=channels, precision=self.precision)
                self.decode_head.classifier = nn.Conv2d(channels, self.num_classes, kernel_size=1)
            elif dtype_ == "IdentityHead":
                self.decode_head = IdentityHead()
            else:
                raise ValueError(f"Unsupported/unavailable decode_head type: {dtype_}")
        self.with_decode_head = self.decode_head is not None

        # --- depth head (denseclip.py:327-357) ---
        self.depth_head = None
        self.with_depth_head = False
        if depth_head:
            dht = dict(depth_head).pop('type')
            din = depth_head.get('in_channels') or head_in_channels
            if dht == "FCNHeadDepth":
                channels = depth_head.get('channels', 128)
                self.depth_head = FCNHead(in_channels=din, channels=channels, precision=self.precision)
                self.depth_head.classifier = nn.Conv2d(channels, 1, kernel_size=1)
                self.with_depth_head = True
            else:
                logger.warning("Unsupported or unavailable depth_head type: %s", dht)

        self.auxiliary_head = None
        self.with_auxiliary_head = False
        self.identity_head = None
        self.with_identity_head = False

        # --- tokenisation and learnable prompt parameters (denseclip.py:374-408) ---
        if token_ids is not None:
            self.texts = torch.as_tensor(token_ids, dtype=torch.long)
        else:
            self.texts = torch.cat([tokenize(c, context_length=self.fixed_text_context_length) for c in self.class_names])
        self.contexts = None
        self.gamma = None
        if self.is_context_encoder:
            n_learn = getattr(self.text_encoder, 'context_length', 77) - self.fixed_text_context_length
            if n_learn > 0:
                self.contexts = nn.Parameter(torch.randn(1, n_learn, token_embed_dim))
                nn.init.trunc_normal_(self.contexts, std=.02)
            self.gamma = nn.Parameter(torch.ones(self.text_dim) * 1e-4)
        self._packed = None
        self.last_score_map = None
        self.last_text_embeddings = None
        self._init_non_clip_weights()

    # ---- initialisation of the non-CLIP modules (denseclip.py:416-513) ----
    def _init_weights_fn(self, m):
        classname = m.__class__.__name__
        if classname.find('Conv') != -1:
            try:
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            except AttributeError:
                pass
        elif classname.find('Linear') != -1:
            try:
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            except AttributeError:
                pass
        elif classname.find('BatchNorm') != -1 or classname.find('GroupNorm') != -1:
            try:
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            except AttributeError:
                pass

    def _init_non_clip_weights(self):
        mods = [(n, m) for n, m in (('vis_proj', self.vis_proj), ('global_proj', self.global_proj),
                                    ('context_decoder', self.context_decoder), ('neck', self.neck),
                                    ('decode_head', self.decode_head), ('depth_head', self.depth_head)) if m is not None]
        for name, module in mods:
            module.apply(self._init_weights_fn)
            if name in ('decode_head', 'depth_head') and hasattr(module, 'classifier'):
                cl = module.classifier
                if isinstance(cl, (nn.Conv2d, nn.Linear)):
                    nn.init.normal_(cl.weight, mean=0, std=0.01)
                    if cl.bias is not None:
                        nn.init.constant_(cl.bias, 0)

    # ---- backbone ----
    def extract_feat(self, img):
        """Reference denseclip.py:516-565: list of NCHW fp32 feature maps, one per backbone out_index."""
        features = self.backbone(img)
        if not isinstance(features, (list, tuple)) or not features:
            logger.error("Backbone returned an empty or invalid feature list.")
            return []
        return list(features)

    # ---- _process_features ----
    def _pack_proj(self):
        ver = tuple(_param_versions(m) for m in (self.vis_proj, self.global_proj) if m is not None)
        if self._packed is not None and self._packed["ver"] == ver:
            return self._packed
        p = dict(ver=ver)
        if self.vis_proj is not None:
            p["vis_w"] = ops.pack_weight(self.vis_proj.weight, True)
            p["vis_b"] = _f32(self.vis_proj.bias)
            p["glob_w"] = ops.pack_weight(self.global_proj.weight, True)
            p["glob_b"] = _f32(self.global_proj.bias)
        self._packed = p
        return p

    def _text_embeddings(self, B: int, device):
        """Class text embeddings [B, K, C].  Input-independent (weights-only), so the [1, K, C] result is cached per
        parameter version; the per-forward cost is an expand."""
        ver = (_param_versions(self.text_encoder), None if self.contexts is None else (self.contexts.data_ptr(), self.contexts._version),
               str(device))
        if getattr(self, "_text_cache", None) is None or self._text_cache[0] != ver:
            if isinstance(self.text_encoder, CLIPTextContextEncoder) and self.contexts is not None:
                t = self.text_encoder(self.texts, self.contexts)          # [1, K, C]
            elif isinstance(self.text_encoder, CLIPTextEncoder):
                t = self.text_encoder(self.texts).unsqueeze(0)            # [1, K, C]
            else:
                raise TypeError("unsupported text encoder")
            self._text_cache = (ver, t)
        return self._text_cache[1].expand(B, -1, -1).contiguous()

    def _tail_native(self, tokens: torch.Tensor, gh: int, gw: int):
        """tokens: fp32 [B, 1+P, Cb] (final-layer features, CLS at row 0). Returns (text_embeddings [B,K,C],
        score_map [B,K,gh,gw], visual tokens fp32 [B, 1+P, C] whose row 0 is the projected global feature)."""
        B, Ntok, Cb = tokens.shape
        P = gh * gw
        glob = ops.token_mean(tokens, 1, P)                                            # denseclip.py:596
        if self.vis_proj is not None:
            pk = self._pack_proj()
            a = ops.split_bf16(tokens.reshape(B * Ntok, Cb))
            V, _ = ops.gemm(a, pk["vis_w"], split_in=True, bias=pk["vis_b"], want_f32=True)  # denseclip.py:616
            Ct = V.shape[1]
            V = V.view(B, Ntok, Ct)
            ops.gemm(ops.split_bf16(glob), pk["glob_w"], split_in=True, bias=pk["glob_b"],     # denseclip.py:607
                     out_f32=V.view(B, Ntok * Ct)[:, :Ct])
        else:
            V = tokens.clone()
            V[:, 0] = glob
        text = self._text_embeddings(B, tokens.device)
        if self.context_decoder is not None:
            if self.context_feature == 'attention':
                ctx = V                                                                  # [global, spatial...] denseclip.py:632
            elif self.context_feature == 'backbone':
                ctx = V[:, 1:]
            else:
                raise ValueError(f"Invalid context_feature type: {self.context_feature}")
            if self.gamma is None:
                raise AttributeError("context decoder configured but gamma is missing")
            diff = self.context_decoder(text, ctx)
            text = ops.gamma_residual(text, _f32(self.gamma), diff)                      # denseclip.py:665
        score = ops.score_map(V, 1, P, text, eps=1e-12).view(B, text.shape[1], gh, gw)   # denseclip.py:672-675
        return text, score, V

    def _process_features(self, x):
        """Reference denseclip.py:570-698. x: list of NCHW fp32 maps -> (text_embeddings, features_for_head, score_map, _x_orig)."""
        if not isinstance(x, (list, tuple)) or not x:
            raise ValueError(f"Expected _process_features input 'x' to be a non-empty list/tuple. Got: {type(x)}")
        vis = x[-1]
        if vis.ndim != 4:
            raise ValueError(f"Expected last backbone feature map to be 4D, got {vis.ndim}D")
        B, _, gh, gw = vis.shape
        tokens, _ = ops.nchw_to_tokens(vis, row_off=1, rows=1 + gh * gw, f32=True, bf16=False)
        text, score, _ = self._tail_native(tokens, gh, gw)
        # denseclip.py:586 clones the inputs and :678 binds features_for_head to that SAME list, so the 2nd and 4th return
        # values are one object and both carry the score-map concat at score_concat_index (copies are layout, not compute)
        x_orig = [f.clone() for f in x]
        feats = x_orig
        if 0 <= self.score_concat_index < len(feats):
            tgt = feats[self.score_concat_index]
            sm = score if score.shape[2:] == tgt.shape[2:] else ops.upsample_bilinear(score, tuple(tgt.shape[2:]))
            feats[self.score_concat_index] = torch.cat([tgt, sm], dim=1)  # layout only
        elif self.score_concat_index != -1:
            logger.warning("score_concat_index %s invalid. Score map not concatenated.", self.score_concat_index)
        return text, feats, score, x_orig

    # ---- forward ----
    def _heads_native(self, feat_b: torch.Tensor, gh: int, gw: int, out_hw, class_map: bool = False):
        B = feat_b.shape[0]
        seg = depth = None
        if self.with_decode_head:
            if isinstance(self.decode_head, IdentityHead):
                raise NotImplementedError("IdentityHead has no native token-major path; use the reference API entry points")
            y, n = self.decode_head.forward_tokens(feat_b, gh, gw)
            seg = (y.view(B, gh * gw, -1), n)
        if self.with_depth_head:
            y, n = self.depth_head.forward_tokens(feat_b, gh, gw)
            depth = (y.view(B, gh * gw, -1), n)
        up = lambda t: None if t is None else ops.upsample_bilinear(t[0], out_hw, tokens_hw=(gh, gw), channels=t[1])  # noqa: E731
        if class_map:
            seg_map = None if seg is None else ops.upsample_argmax(seg[0], out_hw, tokens_hw=(gh, gw), channels=seg[1])
            return seg_map, up(depth)
        return up(seg), up(depth)

    @torch.no_grad()
    def predict(self, img):
        """Batched on-device counterpart of ``simple_test`` (denseclip.py:987-1000): {'seg': uint8 class map [B,H,W],
        'depth': fp32 [B,1,H,W]}.  The argmax is fused with the bilinear upsample, so full-resolution logits are never
        written."""
        return self.forward(img, return_loss=False, _class_map=True)

    # ---- CUDA graph replay (opt-in) ----
    def enable_cuda_graph(self, flag: bool = True):
        """Capture the inference forward (per input shape) into a CUDA graph and replay it: removes the host-side launch
        cost of the ~190 kernels of one step.  Outputs are then STATIC buffers, overwritten by the next call."""
        self._use_graph = bool(flag)
        self._graphs = {}
        self._max_graphs = 4
        return self

    def _graph_forward(self, img, class_map: bool):
        ver = (_param_versions(self), self.precision, self.training)
        key = (tuple(img.shape), img.device.index, class_map)
        ent = self._graphs.get(key)
        if ent is None or ent["ver"] != ver:
            static_in = img.detach().clone().float().contiguous()
            cur = torch.cuda.current_stream(img.device)
            side = torch.cuda.Stream(device=img.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up outside capture: packs weights, builds plans, sizes the workspace
                for _ in range(2):
                    self._forward_impl(static_in, None, False, {'_class_map': class_map})
            cur.wait_stream(side)
            torch.cuda.synchronize(img.device)
            dev = img.device.index if img.device.index is not None else torch.cuda.current_device()
            n0 = _lib.launch_count(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward_impl(static_in, None, False, {'_class_map': class_map})
            ent = dict(ver=ver, graph=graph, inp=static_in, out=out, launches=_lib.launch_count(dev) - n0,
                       keep=getattr(self, "_last_workspace", None))
            # a few live graphs (each private pool holds a full set of activations): alternating forward()/predict() or two
            # input shapes must not re-capture on every call; the oldest entry goes first
            self._graphs[key] = ent
            while len(self._graphs) > self._max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
        else:
            self._graphs[key] = self._graphs.pop(key)  # mark as most recently used
        ent["inp"].copy_(img, non_blocking=True)
        ent["graph"].replay()
        self.graph_launches_per_step = ent["launches"]
        return ent["out"]

    def forward(self, img, img_metas=None, gt_semantic_seg=None, return_loss=True, **kwargs):
        """Reference denseclip.py:702-916.  Error contract (denseclip.py:738-752, 802-817): the reference catches every
        exception of a stage, logs it and returns ``None`` outputs.  That behaviour is kept for MODEL-LEVEL errors (bad
        shapes, inconsistent configuration: ValueError / TypeError / AttributeError / KeyError / IndexError), so the
        reference's callers see the same dict.  A failure of the NATIVE layer (``DclipError``: missing extension, no sm_100
        device, CPU tensor, kernel launch error) or any CUDA runtime error is never swallowed -- it propagates, because a
        silent ``None`` there would hide a broken accelerator path."""
        try:
            if getattr(self, "_use_graph", False) and not (return_loss and self.training) and img.is_cuda:
                return self._graph_forward(img, bool(kwargs.get('_class_map', False)))
            return self._forward_impl(img, gt_semantic_seg, return_loss, kwargs)
        except (ValueError, TypeError, AttributeError, KeyError, IndexError) as e:
            logger.error("Error during DenseCLIP forward: %s", e, exc_info=True)
            if return_loss and self.training:
                return {'main_output': None, 'depth_output': None, 'aux_losses': {}}
            return {'seg': None, 'depth': None}

    def _forward_impl(self, img, gt_semantic_seg, return_loss, kwargs):
        """Reference denseclip.py:702-916.  Inference returns {'seg': [B,K,H,W], 'depth': [B,1,H,W]} fp32; the training
        branch returns {'main_output','depth_output','aux_losses'} resized to the ground-truth size.  Forward-only:
        the native path carries no autograd graph (backward of the trainable tail is out of scope)."""
        if self.align_corners:
            raise NotImplementedError("align_corners=True resize is not implemented natively (reference default is False)")
        precise = self.precision == "fp32"
        use_tokens = isinstance(self.neck, ViTFeatureFusionNeck) and not precise
        enc = self.backbone.forward_native(img, taps_nchw=not use_tokens, taps_tokens_bf16=use_tokens, last_tokens=True)
        gh, gw = enc["grid"]
        B = img.shape[0]
        self._last_workspace = enc["workspace"]   # (a captured graph holds on to it, see _graph_forward)
        if self.backbone.out_indices[-1] != self.backbone.layers - 1:
            # x[-1] is then an intermediate (un-normalised) tap: rebuild its token view from the NCHW tap
            last_nchw = enc["nchw"][-1] if enc["nchw"] else ops.tap_nchw(enc["tokens_bf16"][-1].float(), gh, gw)
            tokens, _ = ops.nchw_to_tokens(last_nchw, row_off=1, rows=1 + gh * gw)
        else:
            tokens = enc["last_tokens"]
        # 2. score map / context decoder: computed as the reference does, though forward() never returns it.
        #    The branch (vis/global projection, ContextDecoder, score map: ~50 mostly latency-bound launches) shares nothing
        #    with the neck / heads branch below, so it runs on a side stream and the two overlap (also inside a captured
        #    graph, where the fork/join become graph dependencies).  DCLIP_OVERLAP_TAIL=0 runs them back to back.
        #    The branch ALWAYS runs on the model's own side stream (the few-query attention keeps its key-split scratch per
        #    stream and may not grow it during graph capture: warm-up and capture must see the same stream); with
        #    DCLIP_OVERLAP_TAIL=0 the main stream simply joins before the neck instead of after the heads.
        side = self._tail_stream(img.device)
        main = torch.cuda.current_stream(img.device)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            text, score, _ = self._tail_native(tokens, gh, gw)
            join = torch.cuda.Event()
            join.record(side)
        if not _OVERLAP_TAIL:
            main.wait_event(join)
            side = None
        self.last_text_embeddings, self.last_score_map = text, score
        if self.training:
            # training mode (train_denseclip.py:1226): BatchNorm on batch statistics, Dropout, and a tape through neck / heads /
            # resize (the only part loss.backward() reaches: backbone and text tower are frozen, the score-map branch above is
            # computed and dropped, denseclip.py:755-812) -- train_tail.py
            out = self._forward_train(enc, tokens, gh, gw, B, img, gt_semantic_seg, return_loss, kwargs)
            if side is not None:
                main.wait_event(join)
            return out
        # 3. neck
        if self.neck is not None:
            if use_tokens:
                _, feat_b = self.neck.forward_tokens(enc["tokens_bf16"], 1, gh, gw)
            else:
                toks = []
                for f in enc["nchw"]:
                    tf, _ = ops.nchw_to_tokens(f)
                    toks.append(ops.split_bf16(tf.view(-1, tf.shape[2])).view(B, gh * gw, -1) if precise else ops.cast_bf16(tf.view(-1, tf.shape[2])).view(B, gh * gw, -1))
                _, feat_b = self.neck.forward_tokens(toks, 0, gh, gw)
            feat_b = feat_b.view(B, gh * gw, -1)
        else:
            t2 = tokens[:, 1:].reshape(B * gh * gw, -1)
            feat_b = (ops.split_bf16(t2) if precise else ops.cast_bf16(t2)).view(B, gh * gw, -1)
        # 5./6. heads + resize
        if return_loss and self.training:
            gt = gt_semantic_seg if gt_semantic_seg is not None else kwargs.get('gt_depth', kwargs.get('depth_targets', kwargs.get('seg_targets')))
            out_hw = tuple(gt.shape[-2:]) if gt is not None else (gh, gw)
            seg, depth = self._heads_native(feat_b, gh, gw, out_hw)
            if side is not None:
                main.wait_event(join)
            return {'main_output': seg, 'depth_output': depth, 'aux_losses': {}}
        seg, depth = self._heads_native(feat_b, gh, gw, tuple(img.shape[2:]), class_map=bool(kwargs.get('_class_map', False)))
        if side is not None:
            main.wait_event(join)  # the score map / text embeddings are complete when forward() returns, as before
        return {'seg': seg, 'depth': depth}

    def _forward_train(self, enc, tokens, gh, gw, B, img, gt_semantic_seg, return_loss, kwargs):
        """Training-mode neck / heads / resize with a tape (train_tail.py).  The tail's GEMM precision follows the model's:
        precision="fp32" -> three-pass split products (the reference trains in fp32), precision="bf16" -> one bf16 pass."""
        from . import train_tail as T
        precise = self.precision == "fp32"
        for hd in (self.decode_head, self.depth_head):
            if hd is not None and isinstance(hd, IdentityHead):
                raise NotImplementedError("IdentityHead has no native token-major path")
        with torch.enable_grad():
            if self.neck is not None:
                if not isinstance(self.neck, ViTFeatureFusionNeck):
                    raise NotImplementedError("training mode supports the ViTFeatureFusionNeck only")
                if precise:
                    taps = [ops.nchw_to_tokens(f)[0].view(B * gh * gw, -1) for f in enc["nchw"]]
                    fused = T.neck_forward_train(self.neck, taps, 0, gh, gw, split=True)
                else:
                    fused = T.neck_forward_train(self.neck, enc["tokens_bf16"], 1, gh, gw, split=False)
            else:
                fused = tokens[:, 1:].reshape(B * gh * gw, -1).contiguous()
            if return_loss:
                gt = gt_semantic_seg if gt_semantic_seg is not None else kwargs.get('gt_depth', kwargs.get('depth_targets', kwargs.get('seg_targets')))
                out_hw = tuple(gt.shape[-2:]) if gt is not None else (gh, gw)   # denseclip.py:822-836: no GT shape -> unresized
            else:
                out_hw = tuple(img.shape[2:])
            res = []
            for hd in (self.decode_head if self.with_decode_head else None, self.depth_head if self.with_depth_head else None):
                if hd is None:
                    res.append(None)
                    continue
                y, n = T.head_forward_train(hd, fused, B, gh, gw, split=precise)
                res.append(T.upsample_train(y, B, gh, gw, n, out_hw))
        if return_loss:
            return {'main_output': res[0], 'depth_output': res[1], 'aux_losses': {}}
        return {'seg': res[0], 'depth': res[1]}

    def _tail_stream(self, device):
        st = getattr(self, "_tail_side_stream", None)
        if st is None or st.device != torch.device(device):
            st = torch.cuda.Stream(device=device)
            self._tail_side_stream = st
        return st

    # ---- inference helpers (reference denseclip.py:938-1041) ----
    def inference(self, img, img_meta, rescale):
        outputs = self.forward(img, img_metas=img_meta, return_loss=False)
        seg, depth = outputs.get('seg'), outputs.get('depth')
        if rescale and img_meta is not None and len(img_meta) > 0 and 'ori_shape' in img_meta[0]:
            ori = tuple(img_meta[0]['ori_shape'][:2])
            if seg is not None and tuple(seg.shape[-2:]) != ori:
                seg = ops.upsample_bilinear(seg, ori)
            if depth is not None and tuple(depth.shape[-2:]) != ori:
                depth = ops.upsample_bilinear(depth, ori)
        return {'seg': seg, 'depth': depth}

    def simple_test(self, img, img_meta, rescale=True):
        out = self.inference(img, img_meta, rescale)
        seg, depth = out.get('seg'), out.get('depth')
        seg_map = seg.argmax(dim=1).cpu().numpy()[0] if seg is not None else None
        depth_map = depth.squeeze(1).cpu().numpy()[0] if depth is not None else None
        return {'seg': seg_map, 'depth': depth_map}

    def aug_test(self, imgs, img_metas, rescale=True):
        segs, depths = [], []
        for img, meta in zip(imgs, img_metas):
            out = self.inference(img.unsqueeze(0), [meta], rescale)
            if out.get('seg') is not None:
                segs.append(out['seg'])
            if out.get('depth') is not None:
                depths.append(out['depth'])
        seg_map = torch.stack(segs).mean(dim=0).argmax(dim=1).squeeze(0).cpu().numpy() if segs else None
        depth_map = torch.stack(depths).mean(dim=0).squeeze().cpu().numpy() if depths else None
        return {'seg': seg_map, 'depth': depth_map}

    def forward_dummy(self, img):
        out = self.forward(img, return_loss=False)
        return out['seg'] if out['seg'] is not None else torch.zeros(img.shape[0], self.num_classes, *img.shape[2:], device=img.device)

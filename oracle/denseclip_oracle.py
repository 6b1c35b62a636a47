"""ORACLE -- CPU fp32 restatement of the reference's DenseCLIP forward path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module, and only as the checker / CPU baseline -- never as part of the shipped path.  It is written as plain
functions over a flat ``state_dict`` (the reference's own key names), so the same weights drive the reference, this
oracle and the B200-native implementation.  Every function cites the reference lines it restates (paths relative to
``/root/reference/segmentation/denseclip/``).

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against outputs of the
UNMODIFIED reference imported in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this module against those fixtures, and ``tests/test_oracle_vs_reference.py``
checks it against the live reference whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------
# configs
# ---------------------------------------------------------------------------------------------------------------
CITYSCAPES_CLASSES = ['road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'traffic light', 'traffic sign',
                      'vegetation', 'terrain', 'sky', 'person', 'rider', 'car', 'truck', 'bus', 'train', 'motorcycle',
                      'bicycle']  # datasets/cityscapes_depth_seg.py:47-52

# token ids produced by the reference tokenizer (utils.py:295-314) for the names above, context_length 6
CITYSCAPES_TOKEN_IDS = [
    [49406, 1759, 49407, 0, 0, 0], [49406, 23278, 49407, 0, 0, 0], [49406, 2307, 49407, 0, 0, 0],
    [49406, 2569, 49407, 0, 0, 0], [49406, 12679, 49407, 0, 0, 0], [49406, 8170, 49407, 0, 0, 0],
    [49406, 3399, 1395, 49407, 0, 0], [49406, 3399, 2292, 49407, 0, 0], [49406, 33947, 49407, 0, 0, 0],
    [49406, 20184, 49407, 0, 0, 0], [49406, 2390, 49407, 0, 0, 0], [49406, 2533, 49407, 0, 0, 0],
    [49406, 9707, 49407, 0, 0, 0], [49406, 1615, 49407, 0, 0, 0], [49406, 4629, 49407, 0, 0, 0],
    [49406, 2840, 49407, 0, 0, 0], [49406, 3231, 49407, 0, 0, 0], [49406, 10297, 49407, 0, 0, 0],
    [49406, 11652, 49407, 0, 0, 0]]


def model_config(name: str = "vit_b16", context_decoder_layers: int = 3) -> dict:
    """Constructor kwargs for ``DenseCLIP`` (reference and native share them).  ``vit_b16`` = the live yaml
    (configs/denseclip_cityscapes.yaml:18-72) plus the canonical 3-layer ContextDecoder; ``tiny`` = a
    structurally identical miniature used by the golden fixtures."""
    if name == "vit_b16":
        cfg = dict(
            backbone=dict(type='CLIPVisionTransformer', patch_size=16, width=768, layers=12, heads=12, input_resolution=224,
                          output_dim=768, out_indices=list(range(12))),
            text_encoder=dict(type='CLIPTextContextEncoder', context_length=22, vocab_size=49408, transformer_width=512,
                              transformer_heads=8, transformer_layers=12, embed_dim=512),
            neck=dict(type='ViTFeatureFusionNeck', inter_channels=128, out_channels=256),
            decode_head=dict(type='FPNHead', in_channels=256, channels=256, num_classes=19, align_corners=False, dropout_ratio=0.1),
            depth_head=dict(type='FCNHeadDepth', in_channels=256, channels=128, align_corners=False),
            class_names=CITYSCAPES_CLASSES, context_length=6, token_embed_dim=512, text_dim=512, context_feature='attention',
            score_concat_index=-1, text_head=False, tau=0.05)
        if context_decoder_layers:
            cfg['context_decoder'] = dict(type='ContextDecoder', transformer_width=256, transformer_heads=4,
                                          transformer_layers=context_decoder_layers, visual_dim=512, dropout=0.1)
        return cfg
    if name == "vit_l14":
        cfg = model_config("vit_b16", context_decoder_layers)
        cfg['backbone'] = dict(type='CLIPVisionTransformer', patch_size=14, width=1024, layers=24, heads=16,
                               input_resolution=224, output_dim=1024, out_indices=[23])
        cfg['neck'] = dict(type='ViTFeatureFusionNeck', inter_channels=128, out_channels=256)
        return cfg
    if name == "tiny":
        cfg = dict(
            backbone=dict(type='CLIPVisionTransformer', patch_size=16, width=256, layers=2, heads=4, input_resolution=32,
                          output_dim=256, out_indices=[0, 1]),
            text_encoder=dict(type='CLIPTextContextEncoder', context_length=22, vocab_size=49408, transformer_width=128,
                              transformer_heads=2, transformer_layers=2, embed_dim=128),
            neck=dict(type='ViTFeatureFusionNeck', inter_channels=64, out_channels=128),
            decode_head=dict(type='FPNHead', in_channels=128, channels=64, num_classes=19, align_corners=False, dropout_ratio=0.1),
            depth_head=dict(type='FCNHeadDepth', in_channels=128, channels=32, align_corners=False),
            class_names=CITYSCAPES_CLASSES, context_length=6, token_embed_dim=128, text_dim=128, context_feature='attention',
            score_concat_index=-1, text_head=False, tau=0.05)
        if context_decoder_layers:
            cfg['context_decoder'] = dict(type='ContextDecoder', transformer_width=128, transformer_heads=2,
                                          transformer_layers=min(context_decoder_layers, 2), visual_dim=128, dropout=0.1)
        return cfg
    raise KeyError(name)


# ---------------------------------------------------------------------------------------------------------------
# deterministic weights: every tensor of the state_dict is overwritten from a seeded numpy generator (SURVEY N2:
# the reference leaves text positional_embedding / text_projection uninitialised -> NaN without this)
# ---------------------------------------------------------------------------------------------------------------
def seeded_state_dict(shapes: "dict[str, tuple]", seed: int = 0) -> "dict[str, torch.Tensor]":
    """shapes: name -> shape (from ``model.state_dict()``). Values depend only on (seed, sorted key order, shape)."""
    out = {}
    for i, k in enumerate(sorted(shapes)):
        shape = tuple(shapes[k])
        rng = np.random.default_rng([seed, i])
        n = int(np.prod(shape)) if shape else 1
        leaf = k.split('.')[-1]
        if leaf == 'num_batches_tracked':
            out[k] = torch.tensor(1, dtype=torch.long)
            continue
        if leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, n)
        elif leaf == 'running_mean':
            v = rng.normal(0, 0.1, n)
        elif k.endswith(('ln_1.weight', 'ln_2.weight', 'ln_pre.weight', 'ln_post.weight', 'ln_final.weight')) or \
                ('norm' in k and leaf == 'weight') or (leaf == 'weight' and len(shape) == 1):
            v = 1.0 + rng.normal(0, 0.05, n)      # LayerNorm / BatchNorm scales
        elif leaf in ('bias', 'in_proj_bias') or (len(shape) == 1 and leaf != 'gamma' and 'embedding' not in leaf):
            v = rng.normal(0, 0.02, n)
        elif leaf == 'gamma':
            v = np.full(n, 0.5) + rng.normal(0, 0.05, n)   # large enough that the ContextDecoder matters
        elif leaf == 'token_embedding.weight' or k.endswith('token_embedding.weight'):
            v = rng.normal(0, 0.02, n)
        elif 'positional_embedding' in k or leaf in ('class_embedding', 'contexts'):
            v = rng.normal(0, 0.05, n)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            if leaf in ('text_projection', 'proj') and len(shape) == 2:
                fan_in = shape[0]
            v = rng.normal(0, 1.0 / math.sqrt(max(fan_in, 1)), n)
        out[k] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return out


# ---------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------
def layer_norm(x, sd, prefix, eps=1e-5):
    """models.py:243-249 (nn.LayerNorm in fp32)."""
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], eps)


def quick_gelu(x):
    """models.py:252-254."""
    return x * torch.sigmoid(1.702 * x)


def multihead_attention(x, sd, prefix, heads, attn_mask=None):
    """nn.MultiheadAttention(x, x, x) as used at models.py:287-289; x is [L, N, D] (LND)."""
    L, N, D = x.shape
    hd = D // heads
    qkv = x @ sd[prefix + '.in_proj_weight'].t() + sd[prefix + '.in_proj_bias']
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(L, N * heads, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.reshape(L, N * heads, hd).transpose(0, 1)
    v = v.reshape(L, N * heads, hd).transpose(0, 1)
    s = q @ k.transpose(1, 2)
    if attn_mask is not None:
        s = s + attn_mask
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(0, 1).reshape(L, N, D)
    return o @ sd[prefix + '.out_proj.weight'].t() + sd[prefix + '.out_proj.bias']


def residual_attention_block(x, sd, prefix, heads, attn_mask=None):
    """models.py:291-294."""
    x = x + multihead_attention(layer_norm(x, sd, prefix + '.ln_1'), sd, prefix + '.attn', heads, attn_mask)
    h = layer_norm(x, sd, prefix + '.ln_2')
    h = quick_gelu(h @ sd[prefix + '.mlp.c_fc.weight'].t() + sd[prefix + '.mlp.c_fc.bias'])
    return x + (h @ sd[prefix + '.mlp.c_proj.weight'].t() + sd[prefix + '.mlp.c_proj.bias'])


def interpolate_pos_encoding(pos, n_tokens, H, W):
    """models.py:514-540."""
    n_loaded = pos.shape[0] - 1
    if n_tokens - 1 == n_loaded:
        return pos
    g0 = int(np.sqrt(n_loaded))
    if g0 * g0 != n_loaded:
        return pos
    dim = pos.shape[-1]
    patch = F.interpolate(pos[1:].reshape(1, g0, g0, dim).permute(0, 3, 1, 2), size=(H, W), mode='bilinear', align_corners=False)
    patch = patch.permute(0, 2, 3, 1).reshape(-1, dim)
    return torch.cat([pos[0:1], patch], dim=0)


def vit_forward(sd, bcfg, img, prefix='backbone.', return_stream=False):
    """CLIPVisionTransformer.forward, models.py:543-597. Returns list of [B, D, gh, gw], one per sorted out_index."""
    ps, layers, heads = bcfg['patch_size'], bcfg['layers'], bcfg['heads']
    out_indices = sorted(set(bcfg.get('out_indices') or [layers - 1]))
    x = F.conv2d(img, sd[prefix + 'conv1.weight'], stride=ps)
    B, D, gh, gw = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd[prefix + 'class_embedding'].expand(B, 1, -1), x], dim=1)
    x = x + interpolate_pos_encoding(sd[prefix + 'positional_embedding'], x.shape[1], gh, gw)
    x = layer_norm(x, sd, prefix + 'ln_pre').permute(1, 0, 2)
    feats, stream = [], []
    for i in range(layers):
        x = residual_attention_block(x, sd, f'{prefix}transformer.resblocks.{i}', heads)
        if return_stream:
            stream.append(x.permute(1, 0, 2).clone())
        if i in out_indices:
            seq = x.permute(1, 0, 2)
            if i == layers - 1:
                seq = layer_norm(seq, sd, prefix + 'ln_post')
            feats.append(seq[:, 1:, :].permute(0, 2, 1).reshape(B, D, gh, gw))
    return (feats, stream) if return_stream else feats


def text_context_encode(sd, tcfg, texts, contexts, prefix='text_encoder.'):
    """CLIPTextContextEncoder.forward, models.py:844-864, including Transformer.forward's double pass (:305-307)."""
    heads, layers, L = tcfg['transformer_heads'], tcfg['transformer_layers'], tcfg['context_length']
    x_text = sd[prefix + 'token_embedding.weight'][texts]
    K, N1, C = x_text.shape
    B, N2, _ = contexts.shape
    eos = (texts.argmax(dim=-1) + N2).reshape(1, K).expand(B, K).reshape(-1)
    x_text = x_text.reshape(1, K, N1, C).expand(B, K, N1, C)
    ctx = contexts.reshape(B, 1, N2, C).expand(B, K, N2, C)
    x = torch.cat([x_text[:, :, 0:1], ctx, x_text[:, :, 1:]], dim=2).reshape(B * K, N1 + N2, C)
    x = (x + sd[prefix + 'positional_embedding']).permute(1, 0, 2)
    mask = torch.full((L, L), float('-inf')).triu_(1)
    for _ in range(2):  # `for resblock in self.resblocks: x = resblock(x)` and then `return self.resblocks(x)`
        for i in range(layers):
            x = residual_attention_block(x, sd, f'{prefix}transformer.resblocks.{i}', heads, mask)
    x = layer_norm(x.permute(1, 0, 2), sd, prefix + 'ln_final')
    x = x[torch.arange(x.shape[0]), eos] @ sd[prefix + 'text_projection']
    return x.reshape(B, K, -1)


def decoder_attention(q, k, v, sd, prefix, heads):
    """Attention.forward, models.py:328-344 (q/k/v projections without bias when qkv_bias=False)."""
    B, N, C = q.shape
    M = k.shape[1]
    lin = lambda t, n: t @ sd[f'{prefix}.{n}.weight'].t() + (sd[f'{prefix}.{n}.bias'] if f'{prefix}.{n}.bias' in sd else 0)  # noqa: E731
    qh = lin(q, 'q_proj').reshape(B, N, heads, C // heads)
    kh = lin(k, 'k_proj').reshape(B, M, heads, C // heads)
    vh = lin(v, 'v_proj').reshape(B, M, heads, C // heads)
    attn = torch.einsum('bnkc,bmkc->bknm', qh, kh) * ((C // heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = torch.einsum('bknm,bmkc->bnkc', attn, vh).reshape(B, N, C)
    return lin(x, 'proj')


def context_decoder(sd, ccfg, text, visual, prefix='context_decoder.'):
    """ContextDecoder.forward, models.py:909-916 with TransformerDecoderLayer.forward :369-375 (eval: dropout off)."""
    heads, layers = ccfg['transformer_heads'], ccfg['transformer_layers']
    ln = lambda t, p: F.layer_norm(t, (t.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)  # noqa: E731
    lin = lambda t, p: t @ sd[p + '.weight'].t() + sd[p + '.bias']  # noqa: E731
    mem = ln(lin(ln(visual, prefix + 'memory_proj.0'), prefix + 'memory_proj.1'), prefix + 'memory_proj.2')
    x = lin(ln(text, prefix + 'text_proj.0'), prefix + 'text_proj.1')
    for i in range(layers):
        p = f'{prefix}decoder.{i}'
        q = ln(x, p + '.norm1')
        x = x + decoder_attention(q, q, q, sd, p + '.self_attn', heads)
        q = ln(x, p + '.norm2')
        x = x + decoder_attention(q, mem, mem, sd, p + '.cross_attn', heads)
        h = F.gelu(lin(ln(x, p + '.norm3'), p + '.mlp.0'))
        x = x + lin(h, p + '.mlp.3')
    return lin(ln(x, prefix + 'out_proj.0'), prefix + 'out_proj.1')


def process_features(sd, cfg, feats, texts=None):
    """DenseCLIP._process_features, denseclip.py:570-698. Returns (text_embeddings [B,K,C], score_map [B,K,h,w])."""
    vis = feats[-1]
    B = vis.shape[0]
    glob = F.adaptive_avg_pool2d(vis, (1, 1)).flatten(1)
    if 'global_proj.weight' in sd:
        glob = glob @ sd['global_proj.weight'].t() + sd['global_proj.bias']
        vis = F.conv2d(vis, sd['vis_proj.weight'], sd['vis_proj.bias'])
    texts = torch.tensor(CITYSCAPES_TOKEN_IDS) if texts is None else texts
    text = text_context_encode(sd, cfg['text_encoder'], texts, sd['contexts']).expand(B, -1, -1)
    if cfg.get('context_decoder'):
        if cfg.get('context_feature', 'attention') == 'attention':
            ctx = torch.cat([glob.unsqueeze(1), vis.flatten(2).permute(0, 2, 1)], dim=1)
        else:
            ctx = vis.flatten(2).permute(0, 2, 1)
        text = text + sd['gamma'] * context_decoder(sd, cfg['context_decoder'], text, ctx)
    score = torch.einsum('bchw,bkc->bkhw', F.normalize(vis, dim=1, p=2), F.normalize(text, dim=2, p=2))
    return text, score


def conv_bn_relu(x, sd, prefix, padding):
    """ConvBNReLU, models.py:13-20, eval-mode BatchNorm (running statistics, eps 1e-5)."""
    y = F.conv2d(x, sd[prefix + '.0.weight'], None, padding=padding)
    y = F.batch_norm(y, sd[prefix + '.1.running_mean'], sd[prefix + '.1.running_var'], sd[prefix + '.1.weight'],
                     sd[prefix + '.1.bias'], False, 0.0, 1e-5)
    return F.relu(y)


def neck_forward(sd, feats, prefix='neck.'):
    """ViTFeatureFusionNeck.forward, models.py:761-782."""
    proc = [conv_bn_relu(f, sd, f'{prefix}process_layers.{i}', 1) for i, f in enumerate(feats)]
    return conv_bn_relu(torch.cat(proc, dim=1), sd, prefix + 'fusion_layer', 0)


def fcn_head(sd, x, prefix):
    """torchvision FCNHead (conv3x3 no-bias, BN, ReLU, Dropout, conv1x1) + the appended ``classifier`` 1x1 conv
    (denseclip.py:305-309, 343-349); eval mode."""
    y = conv_bn_relu(x, sd, prefix, 1)
    y = F.conv2d(y, sd[prefix + '.4.weight'], sd[prefix + '.4.bias'])
    return F.conv2d(y, sd[prefix + '.classifier.weight'], sd[prefix + '.classifier.bias'])


def denseclip_forward(sd, cfg, img, return_intermediates=False):
    """DenseCLIP.forward inference branch, denseclip.py:702-916 (eval, return_loss=False)."""
    feats = vit_forward(sd, cfg['backbone'], img)
    text, score = process_features(sd, cfg, feats)
    x = neck_forward(sd, feats) if cfg.get('neck') else feats[-1]
    seg_lr = fcn_head(sd, x, 'decode_head') if cfg.get('decode_head') else None
    depth_lr = fcn_head(sd, x, 'depth_head') if cfg.get('depth_head') else None
    size = img.shape[2:]
    out = {'seg': F.interpolate(seg_lr, size=size, mode='bilinear', align_corners=False) if seg_lr is not None else None,
           'depth': F.interpolate(depth_lr, size=size, mode='bilinear', align_corners=False) if depth_lr is not None else None}
    if return_intermediates:
        out.update(feats=feats, text=text, score=score, neck=x, seg_lr=seg_lr, depth_lr=depth_lr)
    return out


# ---------------------------------------------------------------------------------------------------------------
# training step of the trainable tail: restated with torch CPU autograd as the differentiator.
# Pinned by tests/golden/make_golden_train.py (the UNMODIFIED reference in .train(), its own SILogLoss, .backward()).
# ---------------------------------------------------------------------------------------------------------------
def conv_bn_relu_train(x, sd, prefix, padding, running):
    """ConvBNReLU (models.py:13-20) with nn.BatchNorm2d in TRAINING mode: batch statistics, running stats updated with
    momentum 0.1 (torch default) into ``running`` (a dict of clones, so the caller's state_dict is not touched)."""
    y = F.conv2d(x, sd[prefix + '.0.weight'], None, padding=padding)
    rm = running.setdefault(prefix + '.1.running_mean', sd[prefix + '.1.running_mean'].clone())
    rv = running.setdefault(prefix + '.1.running_var', sd[prefix + '.1.running_var'].clone())
    y = F.batch_norm(y, rm, rv, sd[prefix + '.1.weight'], sd[prefix + '.1.bias'], True, 0.1, 1e-5)
    return F.relu(y)


def silog_loss(prediction, target, mask=None, lambd=0.5, eps=1e-6):
    """SILogLoss.forward, losses.py:21-79 (variance form, no square root; 0 when no pixel is valid)."""
    d = torch.log(torch.clamp(prediction, min=eps)) - torch.log(torch.clamp(target, min=eps))
    if mask is not None:
        d = torch.where(mask, d, torch.zeros_like(d))
        T = int(mask.sum())
        if T == 0:
            return prediction.sum() * 0.0
    else:
        T = d.numel()
    return (d ** 2).sum() / T - lambd * (d.sum() ** 2) / (T ** 2)


TRAINABLE_PREFIXES = ('neck.', 'decode_head.', 'depth_head.')   # what receives a gradient (see train_forward)


def train_forward(sd, cfg, img, out_hw, drop_p=0.0, running=None):
    """DenseCLIP.forward(return_loss=True) in .train() (denseclip.py:702-891): the heads consume the neck output of the ORIGINAL
    backbone features (:755-812); outputs are resized to the ground-truth size (:838, :849).  The frozen backbone
    (train_denseclip.py:1040-1044) runs without a tape.  Dropout(0.1) of the FCNHead is parameterised (drop_p) so a test can
    switch it off.  Returns (main_output, depth_output, running-stat dict)."""
    running = {} if running is None else running
    with torch.no_grad():
        feats = vit_forward(sd, cfg['backbone'], img)
    proc = [conv_bn_relu_train(f, sd, f'neck.process_layers.{i}', 1, running) for i, f in enumerate(feats)]
    x = conv_bn_relu_train(torch.cat(proc, dim=1), sd, 'neck.fusion_layer', 0, running)
    outs = []
    for prefix in ('decode_head', 'depth_head'):
        if not cfg.get(prefix):
            outs.append(None)
            continue
        y = conv_bn_relu_train(x, sd, prefix, 1, running)
        y = F.dropout(y, drop_p, training=True)
        y = F.conv2d(y, sd[prefix + '.4.weight'], sd[prefix + '.4.bias'])
        y = F.conv2d(y, sd[prefix + '.classifier.weight'], sd[prefix + '.classifier.bias'])
        outs.append(F.interpolate(y, size=out_hw, mode='bilinear', align_corners=False))
    return outs[0], outs[1], running


def train_step(sd, cfg, img, seg_target, depth_target, depth_mask, ignore_index=255, w_seg=1.0, w_silog=0.1, drop_p=0.0):
    """One loss evaluation + backward of the reference's training step (train_denseclip.py:1226-1330):
    loss = w_seg * CE(main_output, seg_target, ignore_index) + w_silog * SILog(depth_output, depth_target, depth_mask).
    Returns dict(main_output, depth_output, loss_seg, loss_silog, loss, grads {key: tensor}, running {key: tensor})."""
    sdg = {k: (v.clone().requires_grad_(True) if k.startswith(TRAINABLE_PREFIXES) and v.is_floating_point()
               and 'running_' not in k else v) for k, v in sd.items()}
    main, depth, running = train_forward(sdg, cfg, img, tuple(seg_target.shape[-2:]), drop_p)
    loss_seg = F.cross_entropy(main, seg_target, ignore_index=ignore_index)
    loss_silog = silog_loss(depth, depth_target, depth_mask)
    loss = w_seg * loss_seg + w_silog * loss_silog
    loss.backward()
    grads = {k: v.grad for k, v in sdg.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    return dict(main_output=main.detach(), depth_output=depth.detach(), loss_seg=loss_seg.detach(), loss_silog=loss_silog.detach(),
                loss=loss.detach(), grads=grads, running=running)


def synthetic_targets(B, H, W, seed=0, num_classes=19, ignore_index=255):
    """Deterministic training targets: int64 class map with ~10% ignore pixels, positive depth, boolean validity mask (~80%)."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.randint(0, num_classes, (B, H, W), generator=g)
    seg[torch.rand(B, H, W, generator=g) < 0.1] = ignore_index
    depth = 0.5 + 20.0 * torch.rand(B, 1, H, W, generator=g)
    mask = torch.rand(B, 1, H, W, generator=g) < 0.8
    return seg, depth, mask


def synthetic_images(B, H, W, seed=0):
    """CLIP-normalised Cityscapes crops are ~zero-mean/unit-variance per channel."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, H, W, generator=g)

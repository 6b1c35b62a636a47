#!/usr/bin/env python
"""Golden fixtures for the TRAINING step of the trainable tail: the UNMODIFIED reference (read from
/root/reference) in .train() on CPU, its own SILogLoss and torch's CrossEntropyLoss(ignore_index=255), loss = CE + 0.1 * SILog
(train_denseclip.py:1086-1096, 1311-1314), backbone and text encoder frozen (:1040-1044), one loss.backward().

Run in the build container only:  python tests/golden/make_golden_train.py
The FCNHead's Dropout(0.1) is set to p = 0 (its mask is the one thing a different implementation cannot reproduce); everything
else is the stock training forward.  Stored: strided outputs, the losses, the gradient of every parameter that received one, the
list of parameters that did NOT (they are not reached by the loss), and the BatchNorm running statistics after the step.
"""
import json
import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import denseclip_oracle as O  # noqa: E402
from oracle.reference_loader import load_reference_denseclip  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DEPTH_BIAS_SHIFT = 2.0


def min_bn_margin(cfg, B, H, W, seed):
    """Smallest |BatchNorm output| (= distance of a ReLU input from 0) over every training-mode BN of the tail, oracle forward."""
    import torch.nn.functional as F
    model_shapes = min_bn_margin.shapes
    sd = O.seeded_state_dict(model_shapes, seed)
    img = O.synthetic_images(B, H, W, seed=seed + 100)
    margin = float("inf")

    def bn(x, prefix, padding):
        nonlocal margin
        y = F.conv2d(x, sd[prefix + '.0.weight'], None, padding=padding)
        y = F.batch_norm(y, None, None, sd[prefix + '.1.weight'], sd[prefix + '.1.bias'], True, 0.1, 1e-5)
        margin = min(margin, float(y.abs().min()))
        return F.relu(y)
    with torch.no_grad():
        feats = O.vit_forward(sd, cfg['backbone'], img)
        x = bn(torch.cat([bn(f, f'neck.process_layers.{i}', 1) for i, f in enumerate(feats)], 1), 'neck.fusion_layer', 0)
        bn(x, 'decode_head', 1)
        bn(x, 'depth_head', 1)
    return margin


def pick_seed(cfg, B, H, W, first, threshold=1e-4):
    """First seed >= `first` whose fixture keeps every ReLU input at least `threshold` away from 0: a mask flip under 1e-5-level
    arithmetic differences would change one channel's gradients by O(1 / rows) and says nothing about the code under test."""
    for seed in range(first, first + 5000):
        m = min_bn_margin(cfg, B, H, W, seed)
        if m > threshold:
            print(f"seed {seed}: min |BN output| {m:.2e}")
            return seed
    raise RuntimeError("no well-conditioned seed found")


def run_case(name, B, H, W, seed, out_stride):
    cfg = O.model_config("tiny", 2)
    if not hasattr(min_bn_margin, "shapes"):
        _, min_bn_margin.shapes = load_reference_denseclip(cfg, 0)
    seed = pick_seed(cfg, B, H, W, seed)
    model, shapes = load_reference_denseclip(cfg, seed)
    sys.path.insert(0, "/root/reference/segmentation")
    from denseclip.losses import SILogLoss
    # keep the depth predictions away from SILog's clamp at eps: with random weights some pixels land at |pred| ~ 1e-5, where
    # d loss / d pred = O(1 / pred) makes every gradient of the depth head depend on fp32 summation order (ill-conditioned
    # fixture, not a property of the code under test).  The clamp branch itself is covered by the op-level loss tests.
    with torch.no_grad():
        model.depth_head.classifier.bias += DEPTH_BIAS_SHIFT
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    for n, p in model.named_parameters():   # train_denseclip.py:1040-1044
        p.requires_grad = not (n.startswith('backbone.') or n.startswith('text_encoder.'))
    img = O.synthetic_images(B, H, W, seed=seed + 100)
    seg_t, depth_t, mask = O.synthetic_targets(B, H, W, seed=seed + 200)
    out = model(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)
    main, depth = out['main_output'], out['depth_output']
    loss_seg = torch.nn.CrossEntropyLoss(ignore_index=255)(main, seg_t)
    loss_silog = SILogLoss(lambd=0.5, eps=1e-6)(depth, depth_t, mask)
    loss = 1.0 * loss_seg + 0.1 * loss_silog
    loss.backward()
    arrays = dict(main_output=main.detach()[..., ::out_stride, ::out_stride].numpy(),
                  depth_output=depth.detach()[..., ::out_stride, ::out_stride].numpy(),
                  losses=np.array([float(loss_seg.detach()), float(loss_silog.detach()), float(loss.detach())], dtype=np.float64))
    no_grad = []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.grad is None:
            no_grad.append(n)
        else:
            arrays["grad:" + n] = p.grad.numpy()
    for n, b in model.named_buffers():
        if n.startswith(O.TRAINABLE_PREFIXES) and ("running_" in n):
            arrays["buf:" + n] = b.detach().numpy()
    meta = dict(cfg_name="tiny", decoder_layers=2, B=B, H=H, W=W, seed=seed, out_stride=out_stride, torch=torch.__version__,
                trainable_without_grad=no_grad, depth_bias_shift=DEPTH_BIAS_SHIFT, shapes={k: list(v) for k, v in shapes.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)
    print(name, "losses", arrays["losses"], "grads", sum(k.startswith("grad:") for k in arrays), "params without grad", len(no_grad))


if __name__ == "__main__":
    logging.disable(logging.CRITICAL)
    torch.set_num_threads(8)
    run_case("tiny_train_32x64_b2", 2, 32, 64, seed=5, out_stride=1)       # 2x4 grid: gather-conv fallback, 16 rows per BatchNorm
    run_case("tiny_train_128x256_b1", 1, 128, 256, seed=6, out_stride=4)   # 8x16 grid: implicit-conv TMA path

#!/usr/bin/env python
"""Training step of the trainable tail at the BASELINE shape (ViT-B/16, 512x1024): forward in .train() + CE/SILog + backward,
CUDA-event timed.  usage: time_train_step.py [batch] [precision] [iters]"""
import copy
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import denseclip_vit_multimodal_b200 as D  # noqa: E402
from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss  # noqa: E402
from oracle import denseclip_oracle as O  # noqa: E402  (synthetic targets only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
m = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()), precision=precision)
bench.init_uninitialised(m)
m = m.cuda().train()
for n, p in m.named_parameters():
    p.requires_grad = not (n.startswith('backbone.') or n.startswith('text_encoder.'))
img = torch.randn(B, 3, 512, 1024, device='cuda')
seg_t, depth_t, mask = (t.cuda() for t in O.synthetic_targets(B, 512, 1024, seed=1))
ce, sl = CrossEntropyLoss(ignore_index=255), SILogLoss()


def step():
    m.zero_grad(set_to_none=True)
    out = m(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)
    loss = ce(out['main_output'], seg_t) + 0.1 * sl(out['depth_output'], depth_t, mask)
    loss.backward()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
with torch.no_grad():
    m.eval()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m(img, return_loss=False)
    f0.record()
    for _ in range(iters):
        m(img, return_loss=False)
    f1.record()
    torch.cuda.synchronize()
print(f"train step B={B} precision={precision}: {ms:.2f} ms/step ({B / ms * 1e3:.1f} img/s), loss {float(loss):.4f}, "
      f"grad norm neck.fusion {float(m.neck.fusion_layer[0].weight.grad.norm()):.3e}; eager inference forward {f0.elapsed_time(f1) / iters:.2f} ms; "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

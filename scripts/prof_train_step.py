#!/usr/bin/env python
"""One training step of the trainable tail at the bench shape, bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import denseclip_vit_multimodal_b200 as D  # noqa: E402
from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()), precision="bf16")
bench.init_uninitialised(m)
m = m.cuda().train()
for n, p in m.named_parameters():
    p.requires_grad = not (n.startswith('backbone.') or n.startswith('text_encoder.'))
img = torch.randn(B, 3, 512, 1024, device='cuda')
seg_t = torch.randint(0, 19, (B, 512, 1024), device='cuda')
depth_t = 0.5 + 20 * torch.rand(B, 1, 512, 1024, device='cuda')
ce, sl = CrossEntropyLoss(ignore_index=255), SILogLoss()


def step():
    m.zero_grad(set_to_none=True)
    out = m(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)
    (ce(out['main_output'], seg_t) + 0.1 * sl(out['depth_output'], depth_t)).backward()


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

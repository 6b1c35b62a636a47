#!/usr/bin/env python
"""How much of the step is the score-map / ContextDecoder branch?  Same box, CUDA-graph replay at the bench shape:
(a) default (branch on a side stream), (b) DCLIP_OVERLAP_TAIL=0 (branch serialised on the main stream), (c) branch removed
(MEASUREMENT ONLY: _tail_native replaced by a stub returning cached tensors) = the ceiling a perfect overlap could reach."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import denseclip_vit_multimodal_b200 as D  # noqa: E402
from denseclip_vit_multimodal_b200 import denseclip as dc  # noqa: E402


def timed(model, img, iters=30):
    with torch.no_grad():
        for _ in range(5):
            model(img, return_loss=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            model(img, return_loss=False)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
img = torch.randn(16, 3, 512, 1024, device="cuda")
res = {}
for rep in range(2):
    for mode in ("overlap", "serial", "no_branch"):
        dc._OVERLAP_TAIL = mode != "serial"
        m = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()), precision="bf16")
        bench.init_uninitialised(m)
        m = m.eval().cuda()
        if mode == "no_branch":
            with torch.no_grad():
                m(img, return_loss=False)
            cached = (m.last_text_embeddings, m.last_score_map, None)
            m._tail_native = lambda tokens, gh, gw, _c=cached: _c
        m.enable_cuda_graph(True)
        res.setdefault(mode, []).append(timed(m, img))
        del m
        torch.cuda.empty_cache()
for k, v in res.items():
    print(f"{k:10s} ms/step: " + "  ".join(f"{x:.3f}" for x in v))

#!/usr/bin/env python
"""One eager (no CUDA graph) DenseCLIP forward at the bench shape, bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (every kernel on the path, B = 16 shapes).  usage: prof_forward.py [batch] [precision] [model]"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import denseclip_vit_multimodal_b200 as D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
model_name = sys.argv[3] if len(sys.argv) > 3 else "vit_b16"
torch.manual_seed(0)
m = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs(model=model_name)), precision=precision)
bench.init_uninitialised(m)
m = m.eval().cuda()
img = torch.randn(B, 3, 512, 1024, device='cuda')
with torch.no_grad():
    for _ in range(2):
        m(img, return_loss=False)
    m.predict(img)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(img, return_loss=False)
    m.predict(img) if os.environ.get("PROF_PREDICT") else None
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok")

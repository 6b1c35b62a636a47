"""Times the x16 bilinear upsample kernels (seg logits 19 ch, depth 1 ch) at the bench shape; prints GB/s of output written."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denseclip_vit_multimodal_b200 import ops

def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B, gh, gw, H, W = 16, 32, 64, 512, 1024
for C, ld in ((19, 20), (1, 4), (19, 256)):
    x = torch.randn(B, gh * gw, ld, device="cuda")
    ms = t(lambda: ops.upsample_bilinear(x, (H, W), tokens_hw=(gh, gw), channels=C))
    print(f"upsample tok C={C} ld={ld}: {ms*1e3:.1f} us  {B*C*H*W*4/ms/1e6:.0f} GB/s written")
    ms = t(lambda: ops.upsample_argmax(x, (H, W), tokens_hw=(gh, gw), channels=C)) if C > 1 else 0
    if C > 1:
        print(f"upsample+argmax C={C} ld={ld}: {ms*1e3:.1f} us")

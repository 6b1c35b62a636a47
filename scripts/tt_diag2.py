import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F
from conftest import load_golden, rel_err
from test_gpu_train_tail import _native_model
from oracle import denseclip_oracle as O
from denseclip_vit_multimodal_b200 import train_tail as T, ops
from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss
torch.set_printoptions(precision=3, linewidth=200)
name = "tiny_train_128x256_b1"
meta, g = load_golden(name)
model, cfg, sd = _native_model(meta)
img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100).cuda()
seg_t, depth_t, mask = (t.cuda() for t in O.synthetic_targets(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 200))
out = model(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)
ls = CrossEntropyLoss(ignore_index=255)(out["main_output"], seg_t)
ld = SILogLoss(0.5, 1e-6)(out["depth_output"], depth_t, mask)
(ls + 0.1 * ld).backward()
named = dict(model.named_parameters())
for k in ["decode_head.0.weight", "neck.fusion_layer.0.weight", "neck.process_layers.0.0.weight"]:
    a, b = named[k].grad.float().cpu(), torch.from_numpy(g["grad:" + k])
    d = (a - b).abs()
    print(k, "shape", tuple(a.shape), "max|ref|", float(b.abs().max()), "rms ref", float(b.pow(2).mean().sqrt()), "rms err", float(d.pow(2).mean().sqrt()))
    if a.shape[-1] == 3:
        print("  per-tap max err:\n", d.amax(dim=(0, 1)), "\n  per-tap max ref:\n", b.abs().amax(dim=(0, 1)))
    idx = torch.nonzero(d == d.max())[0].tolist()
    print("  worst at", idx, "got", float(a[tuple(idx)]), "ref", float(b[tuple(idx)]))
    print("  per-filter max err (first 8):", d.flatten(1).amax(1)[:8])
    print("  cos sim", float(F.cosine_similarity(a.flatten(), b.flatten(), dim=0)))

# isolated head block vs torch autograd on the GPU, same input
torch.manual_seed(0)
B, gh, gw = 2, 8, 16
head = model.decode_head
x0 = torch.relu(torch.randn(B * gh * gw, 128, device="cuda"))
xin = x0.clone().requires_grad_(True)
model.zero_grad(set_to_none=True)
y, n = T.head_forward_train(head, xin, B, gh, gw)
gy = torch.randn_like(y)
y.backward(gy)
xr = x0.view(B, gh, gw, 128).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
mods = list(head.children())
w0 = mods[0].weight.detach().clone().requires_grad_(True)
t = F.conv2d(xr, w0, padding=1)
t = F.batch_norm(t, None, None, mods[1].weight, mods[1].bias, True, 0.1, 1e-5)
t = F.relu(t)
t = F.conv2d(t, mods[4].weight, mods[4].bias)
t = F.conv2d(t, head.classifier.weight, head.classifier.bias)
yr = t.permute(0, 2, 3, 1).reshape(B * gh * gw, -1)
yr.backward(gy[:, :n])
print("isolated head: y", rel_err(y[:, :n], yr), "dW0", rel_err(mods[0].weight.grad, w0.grad), "cos", float(F.cosine_similarity(mods[0].weight.grad.flatten(), w0.grad.flatten(), dim=0)),
      "dx", rel_err(xin.grad, xr.grad.permute(0, 2, 3, 1).reshape(B * gh * gw, 128)))

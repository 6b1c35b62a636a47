"""Diagnostic: per-parameter gradient error of the native training step vs the reference golden fixtures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, rel_err
from test_gpu_train_tail import _native_model
from oracle import denseclip_oracle as O
from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss

for name in ["tiny_train_128x256_b1"]:
    meta, g = load_golden(name)
    model, cfg, sd = _native_model(meta, "fp32")
    img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100).cuda()
    seg_t, depth_t, mask = (t.cuda() for t in O.synthetic_targets(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 200))
    for which in ("both",):
        model.zero_grad(set_to_none=True)
        out = model(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)
        ls = CrossEntropyLoss(ignore_index=255)(out["main_output"], seg_t)
        ld = SILogLoss(0.5, 1e-6)(out["depth_output"], depth_t, mask)
        loss = {"both": ls + 0.1 * ld, "seg": ls, "depth": 0.1 * ld}[which]
        loss.backward()
        if which != "both":
            # oracle gradient for the single loss
            sdo = {k: v.clone() for k, v in sd.items()}
            r = O.train_step(sdo, cfg, img.cpu(), seg_t.cpu(), depth_t.cpu(), mask.cpu(), w_seg=1.0 if which == "seg" else 0.0,
                             w_silog=0.1 if which == "depth" else 0.0)
            ref = r["grads"]
        else:
            ref = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad:")}
        named = dict(model.named_parameters())
        print("==", name, which, "losses", float(ls), float(ld))
        for k in sorted(ref):
            gr = named[k].grad
            if gr is None:
                print("   %-40s native grad None, ref norm %.3e" % (k, float(ref[k].norm())))
                continue
            d = (gr.float().cpu() - ref[k]).abs()
            print("   %-40s err %.3e  rms-rel %.3e |ref| %.3e |got| %.3e" % (k, rel_err(gr, ref[k]), float(d.pow(2).mean().sqrt() / ref[k].pow(2).mean().sqrt()), float(ref[k].norm()), float(gr.float().norm())))
            if k.endswith("head.0.weight"):
                print("      per-tap max err", d.amax(dim=(0, 1)).flatten().tolist())
                print("      per-filter max err", [round(v, 6) for v in d.flatten(1).amax(1).tolist()])

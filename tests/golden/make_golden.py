#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (read from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The reference imports `timm` and `ftfy`, which are not installed; oracle/refstubs/ provides the two tiny stand-ins
(SURVEY section 8(c)).  Weights are not stored: they are regenerated from oracle.seeded_state_dict(shapes, seed).
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


def run_case(name, cfg_name, B, H, W, seed, decoder_layers=2, out_stride=1):
    cfg = O.model_config(cfg_name, decoder_layers)
    model, shapes = load_reference_denseclip(cfg, seed)
    img = O.synthetic_images(B, H, W, seed=seed + 100)
    with torch.no_grad():
        feats = model.extract_feat(img)
        text, _, score, _ = model._process_features([f.clone() for f in feats])
        out = model(img, return_loss=False)
        neck = model.neck([f.clone() for f in feats])[0]
        seg_lr = model.decode_head(neck)
        depth_lr = model.depth_head(neck)
    arrays = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
    arrays.update(text=text.numpy(), score=score.numpy(), neck=neck.numpy(), seg_lr=seg_lr.numpy(), depth_lr=depth_lr.numpy(),
                  seg=out['seg'][..., ::out_stride, ::out_stride].numpy(),
                  depth=out['depth'][..., ::out_stride, ::out_stride].numpy())
    meta = dict(cfg_name=cfg_name, decoder_layers=decoder_layers, B=B, H=H, W=W, seed=seed, out_stride=out_stride, torch=torch.__version__,
                shapes={k: list(v) for k, v in shapes.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: v.astype(np.float32) for k, v in arrays.items()})
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)
    print(name, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    logging.disable(logging.CRITICAL)
    torch.set_num_threads(8)
    run_case("tiny_32x64_b2", "tiny", 2, 32, 64, seed=1)       # pos-emb interpolation 2x2 -> 2x4, gather-conv fallback
    run_case("tiny_128x256_b1", "tiny", 1, 128, 256, seed=2, out_stride=4)   # 8x16 grid = one 128-pixel tile: implicit-conv TMA path
    tok = {"classes": O.CITYSCAPES_CLASSES, "context_length": 6}
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
    sys.path.insert(0, "/root/reference/segmentation")
    from denseclip.utils import tokenize
    tok["ids"] = [tokenize(c, context_length=6)[0].tolist() for c in O.CITYSCAPES_CLASSES]
    with open(os.path.join(OUT, "cityscapes_token_ids.json"), "w") as f:
        json.dump(tok, f)

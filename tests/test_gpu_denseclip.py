"""GPU parity tests proper: the B200-native DenseCLIP (through the C ABI) against the oracle and the reference's golden
vectors.  Tolerances (BASELINE.json north_star): fp32 path rel err <= 1e-3; bf16 path max-abs <= 2e-2 on the normalised
score map (argmax agreement is reported, see DESIGN.md 'H1' for why random-init argmax is ill-conditioned)."""
import copy

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import denseclip_oracle as O

pytestmark = pytest.mark.gpu


def build_native(meta, precision):
    import denseclip_vit_multimodal_b200 as D
    cfg = O.model_config(meta["cfg_name"], meta["decoder_layers"])
    model = D.DenseCLIP(**copy.deepcopy(cfg), precision=precision)
    sd = O.seeded_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    model.load_state_dict(sd, strict=True)
    return model.eval().cuda(), cfg, sd


@pytest.mark.parametrize("name", ["tiny_32x64_b2", "tiny_128x256_b1"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_reference_golden(golden_cases, name, precision):
    meta, g = golden_cases[name]
    model, cfg, sd = build_native(meta, precision)
    img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100).cuda()
    with torch.no_grad():
        feats = model.extract_feat(img)
        text, _, score, _ = model._process_features(feats)
        out = model(img, return_loss=False)
    torch.cuda.synchronize()
    st = meta["out_stride"]
    got = dict(feat0=feats[0], feat1=feats[1], text=text, score=score, seg=out["seg"][..., ::st, ::st],
               depth=out["depth"][..., ::st, ::st])
    errs = {k: rel_err(v, g[k]) for k, v in got.items()}
    print(name, precision, {k: f"{e:.2e}" for k, e in errs.items()})
    assert all(np.isfinite(v.float().cpu().numpy()).all() for v in got.values())
    if precision == "fp32":
        for k, e in errs.items():
            assert e <= 1e-3, (k, e)   # north_star: fp32 path within 1e-3 relative error
    else:
        assert float(np.abs(score.cpu().numpy() - g["score"]).max()) <= 2e-2   # north_star: bf16 max-abs on the score map
        for k, e in errs.items():
            assert e <= 2e-2, (k, e)
    # forward() also stashes what the reference computes and drops
    assert rel_err(model.last_score_map, score) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_patch14_ragged_grid_against_oracle(precision):
    """ViT-L/14-style geometry in miniature: patch 14 (K = 588, padded to 592 for TMA), trailing pixels dropped, a 3x5
    grid that cannot be tiled by 128 pixels (gather-conv fallback), last-layer-only tap, 'backbone' context feature."""
    import denseclip_vit_multimodal_b200 as D
    cfg = O.model_config("tiny", 2)
    cfg["backbone"].update(patch_size=14, input_resolution=28, out_indices=[1])
    cfg["context_feature"] = "backbone"
    model = D.DenseCLIP(**copy.deepcopy(cfg), precision=precision)
    sd = O.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 21)
    model.load_state_dict(sd, strict=True)
    model = model.eval().cuda()
    img = O.synthetic_images(2, 50, 76, seed=9)      # 50x76 -> 3x5 patches of 14, 8 / 6 trailing pixels dropped
    with torch.no_grad():
        out = model(img.cuda(), return_loss=False)
        ref = O.denseclip_forward(sd, cfg, img, return_intermediates=True)
    tol = 1e-3 if precision == "fp32" else 2e-2
    assert out["seg"].shape == (2, 19, 50, 76)
    assert rel_err(out["seg"], ref["seg"]) <= tol and rel_err(out["depth"], ref["depth"]) <= tol
    assert float((model.last_score_map.cpu() - ref["score"]).abs().max()) <= (1e-4 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 128, 256)])
def test_layernorm_fold_matches_plain_layernorm_and_oracle(shape):
    """bf16 ViT blocks: ln_1 / ln_2 folded into the QKV / c_fc GEMMs (weights pre-multiplied by gamma, row statistics published
    by the residual epilogues; models.py:291-293 applied algebraically) against (a) the same path with stand-alone LayerNorm
    kernels and (b) the fp32 oracle.  LayerNorm gamma / beta are far from (1, 0) so a wrong fold would show."""
    import denseclip_vit_multimodal_b200 as D
    cfg = O.model_config("tiny", 2)["backbone"]
    cfg = {k: v for k, v in cfg.items() if k != "type"}
    torch.manual_seed(5)
    ref_model = D.CLIPVisionTransformer(**copy.deepcopy(cfg), precision="bf16")
    sd = O.seeded_state_dict({k: tuple(v.shape) for k, v in ref_model.state_dict().items()}, 33)
    g = torch.Generator().manual_seed(1)
    for k in sd:
        if ".ln_1." in k or ".ln_2." in k:
            sd[k] = (0.5 + torch.rand(sd[k].shape, generator=g)) if k.endswith("weight") else 0.3 * torch.randn(sd[k].shape, generator=g)
    B, H, W = shape
    img = O.synthetic_images(B, H, W, seed=4)
    outs = {}
    for fold in (False, True):
        m = D.CLIPVisionTransformer(**copy.deepcopy(cfg), precision="bf16")
        m.load_state_dict(sd, strict=True)
        m.ln_fold = fold
        m = m.eval().cuda()
        with torch.no_grad():
            outs[fold] = [f.float().cpu() for f in m(img.cuda())]
    with torch.no_grad():
        ref = O.vit_forward({"backbone." + k: v for k, v in sd.items()}, dict(cfg), img)
    for a, b, r in zip(outs[False], outs[True], ref):
        assert rel_err(b, a) < 1e-2, rel_err(b, a)       # two bf16 roundings of the same fp32 function
        assert rel_err(b, r) < 2e-2 and rel_err(a, r) < 2e-2, (rel_err(b, r), rel_err(a, r))

"""GPU parity at the BASELINE configurations, native path (through the C ABI) against the ORACLE (CPU fp32 restatement of the
reference, pinned to the unmodified reference by tests/golden + tests/test_oracle_vs_reference.py) on the same weights and
the same synthetic images:

  * ViT-B/16, batch 2, 512x1024 (BASELINE configs[1]/[2]: seg + depth heads, 12-tap neck, 3-layer ContextDecoder)
  * ViT-L/14, batch 1, 512x1024 (BASELINE configs[3])

Weights are the reference's own random init (constructor defaults; the two tensors the reference leaves uninitialised,
SURVEY N2, are filled by bench.init_uninitialised), with gamma at the reference init 1e-4 AND at 0.3 (so the ContextDecoder
visibly moves the text embeddings).  Gates (BASELINE.json north_star):

  precision="fp32" (3-pass split tensor-core products):  rel err <= 1e-3 on score map / seg / depth, and RAW per-pixel
      score-map argmax agreement >= 99.9% -- this is the shipped mode that meets the argmax gate (cost: bench.py --precision fp32)
  precision="bf16":  max-abs <= 2e-2 on the normalised score map, seg / depth within 2e-2 rel; raw argmax agreement is
      measured against the oracle and bounded below (>= 99%); DESIGN.md section 4 shows why no path with even one bf16 block
      can reach 99.9% on random-init weights (top-2 gaps: median 8e-4, 1st percentile 1.3e-5, 0.1th percentile 1.3e-6).
"""
import copy
import json

import pytest
import torch

from conftest import rel_err
from oracle import denseclip_oracle as O

pytestmark = pytest.mark.gpu

GAMMAS = (1e-4, 0.3)   # reference init (denseclip.py:405) and a value at which the ContextDecoder matters


@pytest.fixture(scope="module", params=[("vit_b16", 2), ("vit_l14", 1)], ids=["vit_b16_b2", "vit_l14_b1"])
def case(request):
    import bench
    import denseclip_vit_multimodal_b200 as D
    name, B = request.param
    cfg = O.model_config(name, 3)
    torch.manual_seed(0)
    m16 = D.DenseCLIP(**copy.deepcopy(cfg), precision="bf16")
    bench.init_uninitialised(m16)
    sd = {k: v.detach().float().clone() for k, v in m16.state_dict().items()}
    m32 = D.DenseCLIP(**copy.deepcopy(cfg), precision="fp32")
    m32.load_state_dict(sd)
    img = O.synthetic_images(B, 512, 1024, seed=17)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        feats = O.vit_forward(sd, cfg["backbone"], img)            # gamma-independent part, computed once
        x = O.neck_forward(sd, feats)
        seg = torch.nn.functional.interpolate(O.fcn_head(sd, x, "decode_head"), size=(512, 1024), mode="bilinear", align_corners=False)
        depth = torch.nn.functional.interpolate(O.fcn_head(sd, x, "depth_head"), size=(512, 1024), mode="bilinear", align_corners=False)
        ref = dict(seg=seg, depth=depth, score={}, text={})
        for g in GAMMAS:
            sdg = dict(sd)
            sdg["gamma"] = torch.full_like(sd["gamma"], g)
            ref["text"][g], ref["score"][g] = O.process_features(sdg, cfg, feats)
    return dict(name=name, B=B, img=img, ref=ref, m16=m16.eval().cuda(), m32=m32.eval().cuda())


def _run(model, img, gamma):
    with torch.no_grad():
        model.gamma.fill_(gamma)
        out = model(img.cuda(), return_loss=False)
        score = model.last_score_map.float().cpu()
    torch.cuda.synchronize()
    return out["seg"].cpu(), out["depth"].cpu(), score


def _agreement(a, b):
    return float((a.argmax(1) == b.argmax(1)).float().mean())


def test_fp32_path_meets_the_north_star_gates_against_the_oracle(case):
    ref = case["ref"]
    rows = []
    for g in GAMMAS:
        seg, depth, score = _run(case["m32"], case["img"], g)
        r = dict(model=case["name"], B=case["B"], precision="fp32", gamma=g,
                 score_rel=rel_err(score, ref["score"][g]), score_max_abs=float((score - ref["score"][g]).abs().max()),
                 score_argmax_agree=_agreement(score, ref["score"][g]),
                 seg_rel=rel_err(seg, ref["seg"]), depth_rel=rel_err(depth, ref["depth"]),
                 seg_argmax_agree=_agreement(seg, ref["seg"]))
        rows.append(r)
        print("PARITY " + json.dumps(r))
    for r in rows:
        assert r["score_rel"] <= 1e-3 and r["seg_rel"] <= 1e-3 and r["depth_rel"] <= 1e-3, r     # north star: fp32 path
        assert r["score_argmax_agree"] >= 0.999, r                                                 # north star: argmax gate, raw, all pixels


def test_bf16_path_against_the_oracle(case):
    ref = case["ref"]
    rows = []
    for g in GAMMAS:
        seg, depth, score = _run(case["m16"], case["img"], g)
        top2 = ref["score"][g].topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1]).flatten()
        r = dict(model=case["name"], B=case["B"], precision="bf16", gamma=g,
                 score_max_abs=float((score - ref["score"][g]).abs().max()), score_argmax_agree=_agreement(score, ref["score"][g]),
                 seg_rel=rel_err(seg, ref["seg"]), depth_rel=rel_err(depth, ref["depth"]), seg_argmax_agree=_agreement(seg, ref["seg"]),
                 oracle_top2_gap_median=float(gap.median()), oracle_top2_gap_p1=float(gap.quantile(0.01)),
                 oracle_top2_gap_p01=float(gap.quantile(0.001)))
        rows.append(r)
        print("PARITY " + json.dumps(r))
    for r in rows:
        assert r["score_max_abs"] <= 2e-2, r                            # north star: bf16 max-abs on the normalised score map
        assert r["seg_rel"] <= 2e-2 and r["depth_rel"] <= 2e-2, r
        # raw agreement with the oracle; the 99.9% gate is met by precision="fp32" above
        assert r["score_argmax_agree"] >= 0.99 and r["seg_argmax_agree"] >= 0.99, r

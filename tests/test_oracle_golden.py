"""CPU: the oracle (oracle/denseclip_oracle.py) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import denseclip_oracle as O


@pytest.mark.parametrize("name", ["tiny_32x64_b2", "tiny_128x256_b1"])
def test_oracle_matches_reference_golden(golden_cases, name):
    meta, g = golden_cases[name]
    cfg = O.model_config(meta["cfg_name"], meta["decoder_layers"])
    sd = O.seeded_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100)
    with torch.no_grad():
        out = O.denseclip_forward(sd, cfg, img, return_intermediates=True)
    st = meta["out_stride"]
    got = dict(feat0=out["feats"][0], feat1=out["feats"][1], text=out["text"], score=out["score"], neck=out["neck"],
               seg_lr=out["seg_lr"], depth_lr=out["depth_lr"], seg=out["seg"][..., ::st, ::st], depth=out["depth"][..., ::st, ::st])
    for k, v in got.items():
        assert rel_err(v, g[k]) < 2e-6, k  # both are fp32 CPU evaluations of the same algorithm


def test_token_ids_known_answer():
    import json, os
    from conftest import GOLDEN
    from denseclip_vit_multimodal_b200.utils import tokenize
    tok = json.load(open(os.path.join(GOLDEN, "cityscapes_token_ids.json")))
    assert tok["ids"] == O.CITYSCAPES_TOKEN_IDS
    got = torch.cat([tokenize(c, context_length=6) for c in tok["classes"]]).tolist()
    assert got == tok["ids"]
    with pytest.raises(KeyError):
        tokenize("zebra crossing", context_length=6)


def test_seeded_state_dict_is_deterministic():
    a = O.seeded_state_dict({"x.weight": (4, 3), "y.bias": (5,)}, 3)
    b = O.seeded_state_dict({"y.bias": (5,), "x.weight": (4, 3)}, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = O.seeded_state_dict({"x.weight": (4, 3), "y.bias": (5,)}, 4)
    assert not torch.equal(a["x.weight"], c["x.weight"])

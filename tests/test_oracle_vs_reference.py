"""CPU, build container only (skipped where /root/reference is absent): the oracle against the LIVE unmodified reference
on configurations the committed golden vectors do not cover."""
import copy

import pytest
import torch

from conftest import rel_err
from oracle import denseclip_oracle as O
from oracle.reference_loader import load_reference_denseclip, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("variant", ["no_decoder", "backbone_ctx", "last_only"])
def test_oracle_tracks_reference(variant):
    cfg = O.model_config("tiny", 2)
    if variant == "no_decoder":
        cfg.pop("context_decoder")
    elif variant == "backbone_ctx":
        cfg["context_feature"] = "backbone"
    elif variant == "last_only":
        cfg["backbone"]["out_indices"] = [1]
    model, shapes = load_reference_denseclip(cfg, seed=11)
    sd = O.seeded_state_dict(shapes, 11)
    img = O.synthetic_images(2, 48, 80, seed=5)   # 3x5 grid: non-square, pos-emb interpolation, odd sizes
    with torch.no_grad():
        feats = model.extract_feat(img)
        text, _, score, _ = model._process_features([f.clone() for f in feats])
        out = model(img, return_loss=False)
        got = O.denseclip_forward(sd, cfg, img, return_intermediates=True)
    assert rel_err(got["feats"][-1], feats[-1]) < 2e-6
    assert rel_err(got["text"], text) < 2e-6 and rel_err(got["score"], score) < 2e-6
    assert rel_err(got["seg"], out["seg"]) < 2e-6 and rel_err(got["depth"], out["depth"]) < 2e-6


def test_text_tower_double_pass_is_reproduced():
    """Transformer.forward applies the stack twice (models.py:305-307); a single pass must NOT match."""
    cfg = O.model_config("tiny", 2)
    model, shapes = load_reference_denseclip(cfg, seed=3)
    sd = O.seeded_state_dict(shapes, 3)
    texts = torch.tensor(O.CITYSCAPES_TOKEN_IDS)
    with torch.no_grad():
        ref = model.text_encoder(texts, sd["contexts"])
        two = O.text_context_encode(sd, cfg["text_encoder"], texts, sd["contexts"])
    assert rel_err(two, ref) < 2e-6

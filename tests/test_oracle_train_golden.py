"""CPU: the oracle's training step (oracle.train_step: forward in .train() semantics, CE(ignore 255) + 0.1 * SILog, backward) against
the golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden_train.py).  Also pins WHICH parameters the
reference's loss reaches: neck + heads only (SURVEY section 8(f)-4 scope)."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import denseclip_oracle as O

CASES = ["tiny_train_32x64_b2", "tiny_train_128x256_b1"]


def oracle_step(meta):
    cfg = O.model_config(meta["cfg_name"], meta["decoder_layers"])
    sd = O.seeded_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    sd["depth_head.classifier.bias"] = sd["depth_head.classifier.bias"] + meta["depth_bias_shift"]   # see make_golden_train.py
    img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100)
    seg_t, depth_t, mask = O.synthetic_targets(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 200)
    return cfg, sd, img, (seg_t, depth_t, mask), O.train_step(sd, cfg, img, seg_t, depth_t, mask)


@pytest.mark.parametrize("name", CASES)
def test_oracle_train_step_matches_reference_golden(name):
    meta, g = load_golden(name)
    _, _, _, _, r = oracle_step(meta)
    st = meta["out_stride"]
    assert rel_err(r["main_output"][..., ::st, ::st], g["main_output"]) < 1e-5
    assert rel_err(r["depth_output"][..., ::st, ::st], g["depth_output"]) < 1e-5
    assert abs(float(r["loss_seg"]) - g["losses"][0]) < 1e-5 * abs(g["losses"][0])
    assert abs(float(r["loss_silog"]) - g["losses"][1]) < 1e-4 * abs(g["losses"][1])
    gkeys = sorted(k[5:] for k in g if k.startswith("grad:"))
    assert sorted(r["grads"]) == gkeys                       # the same parameters receive a gradient ...
    assert all(k.startswith(O.TRAINABLE_PREFIXES) for k in gkeys)
    assert all(not k.startswith(O.TRAINABLE_PREFIXES) for k in meta["trainable_without_grad"])   # ... and nothing else does
    for k in gkeys:
        assert rel_err(r["grads"][k], g["grad:" + k]) < 2e-4, k
    for k in (k for k in g if k.startswith("buf:")):
        assert rel_err(r["running"][k[4:]], g[k]) < 1e-5, k


def test_silog_edge_cases():
    """losses.py:47-53: an all-False mask gives 0; predictions below eps are clamped (zero gradient there)."""
    pred = torch.tensor([[[[0.5, -1.0], [2.0, 1e-9]]]], requires_grad=True)
    tgt = torch.tensor([[[[1.0, 1.0], [4.0, 1.0]]]])
    assert float(O.silog_loss(pred, tgt, torch.zeros_like(tgt, dtype=torch.bool))) == 0.0
    loss = O.silog_loss(pred, tgt, None)
    loss.backward()
    assert float(pred.grad[0, 0, 0, 1]) == 0.0 and float(pred.grad[0, 0, 1, 1]) == 0.0 and float(pred.grad[0, 0, 0, 0]) != 0.0

"""CPU: host-side mirror of the reference interface -- registries, config handling, state_dict layout, parameter
preprocessing (BN folding, conv weight re-ordering, checkpoint pos-emb resize), sharding helpers, no-CPU-fallback."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import denseclip_vit_multimodal_b200 as D
from denseclip_vit_multimodal_b200 import distributed as dd
from denseclip_vit_multimodal_b200 import models as M
from oracle import denseclip_oracle as O


def test_package_surface_matches_reference_init():
    for name in ["DenseCLIP", "CLIPResNet", "CLIPTextEncoder", "CLIPVisionTransformer", "CLIPResNetWithAttention",
                 "CLIPTextContextEncoder", "ContextDecoder", "IdentityHead", "BACKBONES", "HEADS"]:
        assert hasattr(D, name), name
    assert D.HEADS.get("IdentityHead") is D.IdentityHead
    vit = D.BACKBONES.build(dict(type="CLIPVisionTransformer", width=128, layers=1, heads=2, input_resolution=32))
    assert isinstance(vit, D.CLIPVisionTransformer) and vit.out_indices == [0]


def test_state_dict_layout_matches_reference(golden_cases):
    meta, _ = golden_cases["tiny_32x64_b2"]
    m = D.DenseCLIP(**copy.deepcopy(O.model_config("tiny", 2)))
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == meta["shapes"]          # keys and shapes recorded from the reference's own state_dict
    m.load_state_dict(O.seeded_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, 0), strict=True)


def test_config_errors_follow_reference():
    cfg = O.model_config("tiny", 2)
    bad = copy.deepcopy(cfg); bad["backbone"]["type"] = "Nope"
    with pytest.raises(ValueError):
        D.DenseCLIP(**bad)
    bad = copy.deepcopy(cfg); bad["backbone"].pop("out_indices")
    with pytest.raises(ValueError):
        D.DenseCLIP(**bad)
    bad = copy.deepcopy(cfg); bad["backbone"]["type"] = "CLIPResNet"
    with pytest.raises(NotImplementedError):      # out of the hot-path scope, fails loudly
        D.DenseCLIP(**bad)
    with pytest.raises(ValueError):
        D.CLIPVisionTransformer(layers=2, out_indices=[5])
    m = D.DenseCLIP(**copy.deepcopy(cfg), unknown_extra_key=1)   # unknown kwargs are absorbed (denseclip.py:88)
    assert m.tau == 0.05 and m.gamma.shape == (128,) and m.contexts.shape == (1, 16, 128)


def test_no_cpu_fallback():
    m = D.DenseCLIP(**copy.deepcopy(O.model_config("tiny", 2))).eval()
    with pytest.raises(D.DclipError):
        m(torch.zeros(1, 3, 32, 64), return_loss=False)
    with pytest.raises(D.DclipError):
        m.backbone(torch.zeros(1, 3, 32, 64))


def test_fold_bn_and_conv_weight_order_against_torch():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 6, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(6).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.data.normal_(); bn.bias.data.normal_()
    x = torch.randn(2, 8, 5, 7)
    ref = bn(conv(x))
    wf, bf = M.fold_bn(conv.weight, bn)
    got = F.conv2d(x, wf, bf, padding=1)
    assert torch.allclose(ref, got, atol=1e-5)
    # GEMM K order (ky, kx, c) against an explicit unfold
    wg = M.conv3x3_weight_to_gemm(wf)                                   # [N, 9*C]
    cols = F.unfold(x, 3, padding=1).view(2, 8, 9, 35).permute(0, 3, 2, 1).reshape(70, 72)  # [(b,pix), (tap, c)]
    got2 = (cols @ wg.t() + bf).view(2, 35, 6).permute(0, 2, 1).reshape(2, 6, 5, 7)
    assert torch.allclose(ref, got2, atol=1e-4)


def test_checkpoint_posemb_resize_matches_torch_interpolate():
    rng = np.random.default_rng(0)
    pe = rng.normal(size=(1 + 14 * 14, 32)).astype(np.float32)
    got = M._resize_pos_embed_np(pe, 14, 9)
    t = torch.from_numpy(pe[1:]).reshape(1, 14, 14, 32).permute(0, 3, 1, 2)
    ref = F.interpolate(t, size=(9, 9), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).reshape(-1, 32)
    assert np.allclose(got[1:], ref.numpy(), atol=1e-6) and np.array_equal(got[0], pe[0])


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 16, 129):
        for w in (1, 2, 3, 8):
            spans = [dd.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dd.shard_range(4, 2, 2)


def test_confusion_matrix_and_stats():
    pred = torch.tensor([0, 1, 1, 2, 2, 2]); tgt = torch.tensor([0, 1, 2, 2, 255, 2])
    cm = dd.confusion_matrix(pred, tgt, 3)
    assert cm.tolist() == [[1, 0, 0], [0, 1, 0], [0, 1, 2]]
    _, miou, acc, rmse = dd.reduce_eval_stats(cm, torch.tensor(8.0), torch.tensor(2.0))
    assert abs(acc - 0.8) < 1e-9 and abs(rmse - 2.0) < 1e-9 and abs(miou - (1 + 0.5 + 2 / 3) / 3) < 1e-9


def test_training_tail_fails_loudly_without_a_gpu():
    """The training-mode tail and the native losses have no CPU path: CPU tensors raise DclipError instead of silently
    falling back to torch (the product path must fail loudly when the CUDA side is unavailable)."""
    import pytest
    import torch
    from denseclip_vit_multimodal_b200 import _lib, train_tail as T
    from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss
    x = torch.randn(8, 16)
    with pytest.raises(_lib.DclipError):
        T.col_stats(x, 1e-5)
    with pytest.raises(_lib.DclipError):
        CrossEntropyLoss(ignore_index=255)(torch.randn(1, 3, 4, 4, requires_grad=True), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(_lib.DclipError):
        SILogLoss()(torch.rand(1, 1, 4, 4, requires_grad=True) + 0.1, torch.rand(1, 1, 4, 4) + 0.1)
    with pytest.raises(ValueError):
        SILogLoss(reduction="none")          # same constructor contract as denseclip/losses.py:15-19

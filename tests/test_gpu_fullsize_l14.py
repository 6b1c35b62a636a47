"""GPU, SURVEY 8(d) config C4: ViT-L/14 @512x1024 (grid 36x73 = 2628 patches of 14, the trailing 8 x 2 pixels are dropped;
width 1024, 16 heads, 24 layers, one tap).  The reference class is parametric (models.py:384-396); this checks that the
native path is too, at full size: shapes, batch independence (bit-exact), bf16 path vs fp32 path on the normalised score map
(north_star tolerance 2e-2).  The 2628-pixel grid is not tileable by 128, so the neck/head 3x3 convs take the gather path."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    import bench
    import denseclip_vit_multimodal_b200 as D
    from oracle import denseclip_oracle as O
    torch.manual_seed(0)
    cfg = O.model_config("vit_l14", 3)
    m = D.DenseCLIP(**copy.deepcopy(cfg), precision="bf16")
    bench.init_uninitialised(m)
    with torch.no_grad():
        m.gamma.fill_(0.3)
    m = m.eval().cuda()
    m32 = D.DenseCLIP(**copy.deepcopy(cfg), precision="fp32")
    m32.load_state_dict(m.state_dict())
    return m, m32.eval().cuda()


def test_vit_l14_fullsize_properties(models):
    m, m32 = models
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn(2, 3, 512, 1024, device="cuda", generator=g)
    with torch.no_grad():
        full = m(img, return_loss=False)
        s16 = m.last_score_map.clone()
        singles = [m(img[i:i + 1], return_loss=False) for i in range(2)]
        o32 = m32(img[:1], return_loss=False)
        s32 = m32.last_score_map.clone()
    assert full["seg"].shape == (2, 19, 512, 1024) and full["depth"].shape == (2, 1, 512, 1024)
    assert s16.shape == (2, 19, 36, 73) and torch.isfinite(full["seg"]).all() and torch.isfinite(full["depth"]).all()
    assert torch.equal(torch.cat([s["seg"] for s in singles]), full["seg"])        # no cross-image op anywhere
    assert torch.equal(torch.cat([s["depth"] for s in singles]), full["depth"])
    max_abs = float((s16[:1] - s32).abs().max())
    rel_seg = float((full["seg"][:1] - o32["seg"]).abs().max() / o32["seg"].abs().max())
    print(f"ViT-L/14 bf16 vs fp32-path: score-map max_abs={max_abs:.2e} seg rel={rel_seg:.2e}")
    assert float(s32.abs().max()) <= 1.0 + 1e-5
    assert max_abs <= 2e-2 and rel_seg <= 2e-2   # (parity against the ORACLE at this config: tests/test_gpu_parity_baseline.py)

"""GPU, BASELINE-size shapes (ViT-B/16, 512x1024): size-independent properties of the native forward --
batch independence / shard equality (bit-exact), determinism, CUDA-graph replay equality (also with two live graphs),
pipelined predictor equality.  Parity against the oracle at this size lives in tests/test_gpu_parity_baseline.py."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    import bench
    import denseclip_vit_multimodal_b200 as D
    torch.manual_seed(0)
    m = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()), precision="bf16")
    bench.init_uninitialised(m)
    with torch.no_grad():
        m.gamma.fill_(0.3)   # so that the ContextDecoder visibly contributes to the score map
    m = m.eval().cuda()
    m32 = D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()), precision="fp32")
    m32.load_state_dict(m.state_dict())
    return m, m32.eval().cuda()


def test_batch_independence_determinism_and_shard_equality(models):
    from denseclip_vit_multimodal_b200 import distributed as dd
    m, _ = models
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.randn(3, 3, 512, 1024, device="cuda", generator=g)
    with torch.no_grad():
        full = m(img, return_loss=False)
        again = m(img, return_loss=False)
        assert torch.equal(full["seg"], again["seg"]) and torch.equal(full["depth"], again["depth"])   # deterministic
        parts = [m(dd.shard_batch(img, 2, r), return_loss=False) for r in range(2)]                    # 2-way image shard
    seg = torch.cat([p["seg"] for p in parts]); depth = torch.cat([p["depth"] for p in parts])
    assert torch.equal(seg, full["seg"]) and torch.equal(depth, full["depth"])   # bit-exact: no cross-image op anywhere
    assert full["seg"].shape == (3, 19, 512, 1024) and torch.isfinite(full["seg"]).all()


def test_cuda_graph_replay_matches_eager(models):
    m, _ = models
    g = torch.Generator(device="cuda").manual_seed(3)
    imgs = [torch.randn(2, 3, 512, 1024, device="cuda", generator=g) for _ in range(2)]
    with torch.no_grad():
        eager = [m(i, return_loss=False)["seg"].clone() for i in imgs]
        m.enable_cuda_graph(True)
        try:
            for i, e in zip(imgs, eager):
                assert torch.equal(m(i, return_loss=False)["seg"], e)
            pm = m.predict(imgs[0])
            assert torch.equal(pm["seg"].long(), eager[0].argmax(1))
            # forward() and predict() alternate without re-capturing: both graphs stay live
            n_graphs = len(m._graphs)
            ids = {k: id(v["graph"]) for k, v in m._graphs.items()}
            assert torch.equal(m(imgs[1], return_loss=False)["seg"], eager[1])
            assert torch.equal(m.predict(imgs[1])["seg"].long(), eager[1].argmax(1))
            assert len(m._graphs) == n_graphs == 2 and ids == {k: id(v["graph"]) for k, v in m._graphs.items()}
        finally:
            m.enable_cuda_graph(False)


def test_pipelined_predictor_matches_direct_predict(models):
    from denseclip_vit_multimodal_b200.pipeline import PipelinedPredictor
    m, _ = models
    g = torch.Generator().manual_seed(4)
    hosts = [torch.randn(2, 3, 512, 1024, generator=g).pin_memory() for _ in range(5)]
    with torch.no_grad():
        direct = [{k: v.clone().cpu() for k, v in m.predict(h.cuda()).items()} for h in hosts]
        pipe = PipelinedPredictor(m, (2, 3, 512, 1024), "cuda")
        got = [{k: v.clone() for k, v in r.items()} for r in pipe.run(hosts)]
    assert len(got) == len(hosts)
    for d, r in zip(direct, got):
        assert torch.equal(d["seg"], r["seg"]) and torch.equal(d["depth"], r["depth"])

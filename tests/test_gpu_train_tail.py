"""GPU parity of the TRAINING step of the trainable tail, native path through the C ABI against
(a) the golden vectors of the unmodified reference (tests/golden/tiny_train_*.npz: outputs, losses, gradients, running stats) and
(b) torch fp32 references of the individual ops.  GEMM operands are bf16 (fp32 accumulation), so model-level gradients are
compared at 3e-2 relative error (observed ~3e-3); the elementwise / reduction kernels are compared at fp32 tolerances."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import denseclip_oracle as O

pytestmark = pytest.mark.gpu


def _native_model(meta, precision="bf16"):
    import denseclip_vit_multimodal_b200 as D
    cfg = O.model_config(meta["cfg_name"], meta["decoder_layers"])
    model = D.DenseCLIP(**copy.deepcopy(cfg), precision=precision)
    sd = O.seeded_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    sd["depth_head.classifier.bias"] = sd["depth_head.classifier.bias"] + meta["depth_bias_shift"]
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                   # as in make_golden_train.py
    for n, p in model.named_parameters():               # train_denseclip.py:1040-1044
        p.requires_grad = not (n.startswith('backbone.') or n.startswith('text_encoder.'))
    return model, cfg, sd


def _cos(a, b):
    return float(F.cosine_similarity(a.detach().float().flatten().cpu(), torch.as_tensor(b).float().flatten(), dim=0))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["tiny_train_32x64_b2", "tiny_train_128x256_b1"])
def test_training_step_matches_reference_golden(name, precision):
    """precision="fp32" (three-pass split products everywhere, the reference trains in fp32): outputs, losses and EVERY gradient
    within 1e-3 / 5e-3 of the unmodified reference.  precision="bf16" (one bf16 pass, bf16 encoder): outputs within 2e-2, gradients
    within the noise a bf16 forward leaves after three BatchNorm backward passes (cosine >= 0.97 per parameter; observed >= 0.982)."""
    from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss, SILogLoss
    meta, g = load_golden(name)
    model, cfg, sd = _native_model(meta, precision)
    img = O.synthetic_images(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 100).cuda()
    seg_t, depth_t, mask = (t.cuda() for t in O.synthetic_targets(meta["B"], meta["H"], meta["W"], seed=meta["seed"] + 200))
    out = model(img, gt_semantic_seg=seg_t, gt_depth=depth_t, return_loss=True)      # the reference's call (train_denseclip.py:1226)
    main, depth = out["main_output"], out["depth_output"]
    assert out["aux_losses"] == {} and main.requires_grad and depth.requires_grad
    loss_seg = CrossEntropyLoss(ignore_index=255)(main, seg_t)
    loss_silog = SILogLoss(lambd=0.5, eps=1e-6)(depth, depth_t, mask)
    loss = 1.0 * loss_seg + 0.1 * loss_silog
    loss.backward()
    torch.cuda.synchronize()
    st = meta["out_stride"]
    tol_out, tol_loss = (1e-3, 1e-4) if precision == "fp32" else (2e-2, 2e-2)
    errs = dict(main=rel_err(main.detach()[..., ::st, ::st], g["main_output"]), depth=rel_err(depth.detach()[..., ::st, ::st], g["depth_output"]))
    assert errs["main"] < tol_out and errs["depth"] < tol_out, errs
    assert abs(float(loss_seg) - g["losses"][0]) < tol_loss * abs(g["losses"][0])
    assert abs(float(loss_silog) - g["losses"][1]) < tol_loss * abs(g["losses"][1])
    named = dict(model.named_parameters())
    gkeys = sorted(k[5:] for k in g if k.startswith("grad:"))
    with_grad = sorted(n for n, p in named.items() if p.grad is not None)
    assert with_grad == gkeys                                   # exactly the parameters the reference's loss reaches
    worst = {k: rel_err(named[k].grad, g["grad:" + k]) for k in gkeys}
    cos = {k: _cos(named[k].grad, g["grad:" + k]) for k in gkeys}
    print(name, precision, errs, "worst grad", max(worst.items(), key=lambda kv: kv[1]), "min cos", min(cos.items(), key=lambda kv: kv[1]))
    for k in gkeys:
        if precision == "fp32":
            assert worst[k] < 5e-3, (k, worst[k])
        else:
            assert cos[k] > 0.97, (k, cos[k])
    bufs = dict(model.named_buffers())
    for k in (k for k in g if k.startswith("buf:")):
        assert rel_err(bufs[k[4:]], g[k]) < tol_out, k


def test_training_forward_with_dropout_and_torch_losses():
    """Dropout(0.1) active (FCNHead [3]): the output still backpropagates, torch's own criteria work on the returned tensors
    (drop-in for train_denseclip.py:1086-1096), and eval() afterwards uses the updated running statistics."""
    meta, _ = load_golden("tiny_train_32x64_b2")
    model, cfg, sd = _native_model(meta)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.1
    img = O.synthetic_images(2, 32, 64, seed=3).cuda()
    seg_t, depth_t, mask = (t.cuda() for t in O.synthetic_targets(2, 32, 64, seed=4))
    torch.manual_seed(0)
    out = model(img, gt_semantic_seg=seg_t, return_loss=True)
    loss = F.cross_entropy(out["main_output"], seg_t, ignore_index=255) + 0.1 * out["depth_output"].abs().mean()
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    assert float(model.neck.fusion_layer[1].num_batches_tracked) == float(sd["neck.fusion_layer.1.num_batches_tracked"]) + 1
    model.eval()
    with torch.no_grad():
        ev = model(img, return_loss=False)
    assert ev["seg"].shape == (2, 19, 32, 64) and torch.isfinite(ev["seg"]).all()


@pytest.mark.parametrize("B,K,H,W", [(2, 19, 32, 64), (1, 5, 17, 23)])
def test_cross_entropy_loss_forward_backward(B, K, H, W):
    from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss
    g = torch.Generator().manual_seed(1)
    logits = (3 * torch.randn(B, K, H, W, generator=g)).cuda().requires_grad_(True)
    tgt = torch.randint(0, K, (B, H, W), generator=g)
    tgt[torch.rand(B, H, W, generator=g) < 0.2] = 255
    tgt = tgt.cuda()
    loss = CrossEntropyLoss(ignore_index=255)(logits, tgt)
    (2.5 * loss).backward()
    ref_in = logits.detach().clone().requires_grad_(True)
    ref = F.cross_entropy(ref_in, tgt, ignore_index=255)
    (2.5 * ref).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(logits.grad, ref_in.grad) < 1e-5
    # every pixel ignored: torch returns nan and a zero gradient
    all_ign = torch.full_like(tgt, 255)
    l2 = CrossEntropyLoss(ignore_index=255)(logits, all_ign)
    assert torch.isnan(l2)


def test_silog_loss_forward_backward_and_edge_cases():
    from denseclip_vit_multimodal_b200.losses import SILogLoss
    g = torch.Generator().manual_seed(2)
    pred = (torch.randn(2, 1, 33, 47, generator=g) + 0.7).cuda()      # a third of the predictions below eps: clamp branch
    pred[(pred.abs() < 1e-2)] = 0.5                                     # keep 1/pred well conditioned
    pred.requires_grad_(True)
    tgt = (0.5 + 10 * torch.rand(2, 1, 33, 47, generator=g)).cuda()
    mask = (torch.rand(2, 1, 33, 47, generator=g) < 0.7).cuda()
    for m in (mask, None, mask[:, 0]):
        pred.grad = None
        loss = SILogLoss(0.5, 1e-6)(pred, tgt, m)
        loss.backward()
        ref_in = pred.detach().clone().requires_grad_(True)
        mm = None if m is None else (m if m.dim() == 4 else m.unsqueeze(1))
        ref = O.silog_loss(ref_in, tgt, mm)
        ref.backward()
        assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
        assert rel_err(pred.grad, ref_in.grad) < 1e-4
    zero = SILogLoss()(pred, tgt, torch.zeros_like(mask))
    assert float(zero) == 0.0                                            # losses.py:47-53


@pytest.mark.parametrize("B,gh,gw,H,W,K", [(2, 2, 4, 32, 64, 19), (1, 8, 16, 128, 256, 1), (2, 3, 5, 50, 76, 4), (1, 4, 4, 4, 4, 3)])
def test_upsample_bilinear_backward_is_the_adjoint(B, gh, gw, H, W, K):
    from denseclip_vit_multimodal_b200 import train_tail as T
    g = torch.Generator().manual_seed(3)
    ld = (K + 3) // 4 * 4
    y = torch.randn(B * gh * gw, ld, generator=g).cuda().requires_grad_(True)
    out = T.upsample_train(y, B, gh, gw, K, (H, W))
    gout = torch.randn(B, K, H, W, generator=g).cuda()
    out.backward(gout)
    ref_in = y.detach()[:, :K].reshape(B, gh, gw, K).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.interpolate(ref_in, size=(H, W), mode="bilinear", align_corners=False)
    ref.backward(gout)
    assert rel_err(out, ref) < 1e-6
    want = ref_in.grad.permute(0, 2, 3, 1).reshape(B * gh * gw, K)
    assert rel_err(y.grad[:, :K], want) < 1e-5
    assert float(y.grad[:, K:].abs().max()) == 0.0 if ld > K else True


@pytest.mark.parametrize("M,N", [(16, 64), (4099, 1536), (257, 20)])
def test_batchnorm_training_kernels(M, N):
    """col_stats / bn_apply / col_grad_sums against torch's batch_norm (training=True) + relu + dropout mask autograd."""
    from denseclip_vit_multimodal_b200 import train_tail as T
    g = torch.Generator().manual_seed(4)
    x = (2 * torch.randn(M, N, generator=g) + 0.5).cuda()
    gamma = (0.5 + torch.rand(N, generator=g)).cuda()
    beta = (0.3 * torch.randn(N, generator=g)).cuda()
    mask = (torch.rand(M, N, generator=g) < 0.9).cuda()
    gy = torch.randn(M, N, generator=g).cuda()
    rm, rv = torch.zeros(N).cuda(), torch.ones(N).cuda()
    mean, var, rstd = T.col_stats(x, 1e-5, rm, rv, 0.1)
    y, yb = T.bn_apply(0, M, N, x=x, mean=mean, rstd=rstd, gamma=gamma, beta=beta, relu=True, mask=mask.to(torch.uint8), mask_scale=1 / 0.9,
                       want_f32=True, want_bf16=True)
    sg, sgx = T.col_grad_sums(gy, x, mean, rstd, gamma, beta, relu=True, mask=mask.to(torch.uint8), mask_scale=1 / 0.9)
    dx, _ = T.bn_apply(1, M, N, x=x, g=gy, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sum_g=sg, sum_gx=sgx, relu=True,
                       mask=mask.to(torch.uint8), mask_scale=1 / 0.9, want_f32=True)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(N).cuda(), torch.ones(N).cuda()
    ref = F.relu(F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)) * mask / 0.9
    ref.backward(gy)
    assert rel_err(mean, x.mean(0)) < 1e-5 and rel_err(var, x.var(0, unbiased=False)) < 1e-5
    assert rel_err(rm, rm2) < 1e-5 and rel_err(rv, rv2) < 1e-5
    assert rel_err(y, ref) < 1e-5 and rel_err(yb[:, :N].float(), ref) < 1e-2
    assert rel_err(sg, br.grad) < 1e-4 and rel_err(sgx, gr.grad) < 1e-4
    assert rel_err(dx, xr.grad) < 1e-4


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("B,gh,gw,C,F_", [(2, 8, 16, 128, 32), (1, 3, 5, 64, 20), (3, 16, 16, 256, 128)])
def test_conv3x3_weight_and_input_gradients(B, gh, gw, C, F_, split):
    """The weight-gradient GEMM over the padded pixel axis (dclip_gemm_args.wg_* + dclip_transpose_pad operands) and the input
    gradient as an implicit conv with the transposed / flipped filter, against torch's conv2d autograd: one bf16 pass against the
    same bf16-rounded operands, the three-pass split against the fp32 operands themselves."""
    from denseclip_vit_multimodal_b200 import train_tail as T
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B * gh * gw, C, generator=g).cuda()
    w = (0.1 * torch.randn(F_, C, 3, 3, generator=g)).cuda()
    dy = torch.randn(B * gh * gw, F_, generator=g).cuda()
    geo = T._Geom(B, gh, gw, split)
    dw, dx = T._conv_backward(dy, T._tok_parts_of(x, geo), 0, geo, w, need_dx=True)
    r = (lambda t: t.double().cpu()) if split else (lambda t: t.bfloat16().double().cpu())   # fp64 CPU reference (no TF32)
    xr = r(x).view(B, gh, gw, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = r(w).requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    yr.backward(r(dy).view(B, gh, gw, F_).permute(0, 3, 1, 2).contiguous())
    tol = 2e-5 if split else 1e-4
    assert rel_err(dw, wr.grad) < tol, rel_err(dw, wr.grad)
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1).reshape(B * gh * gw, C)) < tol


def test_conv3x3_weight_gradient_production_shape():
    """One neck tap at the BASELINE shape (ViT-B/16 @512x1024: 32x64 grid, 768 -> 128 channels, batch 2): the weight-gradient GEMM
    runs 256-wide tiles over K = 2 * 33 * 72 padded pixels (75 K blocks) with 27 output blocks; bf16 operands vs an fp64 CPU
    conv2d backward on the same rounded operands."""
    from denseclip_vit_multimodal_b200 import train_tail as T
    B, gh, gw, C, F_ = 2, 32, 64, 768, 128
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B * gh * gw, C, generator=g).cuda()
    dy = torch.randn(B * gh * gw, F_, generator=g).cuda()
    w = torch.zeros(F_, C, 3, 3).cuda()
    geo = T._Geom(B, gh, gw, False)
    dw, _ = T._conv_backward(dy, T._tok_parts_of(x, geo), 0, geo, w, need_dx=False)
    r = lambda t: t.bfloat16().double().cpu()   # noqa: E731
    xr = r(x).view(B, gh, gw, C).permute(0, 3, 1, 2).contiguous()
    gr = r(dy).view(B, gh, gw, F_).permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.grad.conv2d_weight(xr, (F_, C, 3, 3), gr, padding=1)
    assert rel_err(dw, ref) < 1e-4, rel_err(dw, ref)


@pytest.mark.parametrize("split", [False, True])
def test_grouped_weight_gradient_equals_per_tap(split):
    """The grouped weight-gradient launch (G convs with 128 filters each in one GEMM per operand pair, the neck's 12 taps) against
    the per-tap launches and an fp64 CPU reference; taps as bf16 tokens behind a CLS row (row0 = 1), as the encoder delivers them."""
    from denseclip_vit_multimodal_b200 import train_tail as T
    G, B, gh, gw, C, F_ = 3, 2, 8, 16, 64, 128
    g = torch.Generator().manual_seed(8)
    geo = T._Geom(B, gh, gw, split)
    xs = [torch.randn(B, gh * gw, C, generator=g) for _ in range(G)]
    toks = [torch.cat([torch.zeros(B, 1, C), x], 1).cuda().bfloat16().contiguous() for x in xs]
    dy = torch.randn(B * gh * gw, G * F_, generator=g).cuda()
    w = torch.zeros(F_, C, 3, 3).cuda()
    dw_all = T._grouped_wgrad(dy, toks, 1, 1, geo, G, F_, C)
    r = (lambda t: t.double().cpu()) if split else (lambda t: t.bfloat16().double().cpu())
    for i in range(G):
        per_tap, _ = T._conv_backward(dy[:, i * F_:(i + 1) * F_].contiguous(), [toks[i]], 1, geo, w, need_dx=False)
        got = dw_all[i * F_:(i + 1) * F_].view(F_, 3, 3, C).permute(0, 3, 1, 2)
        assert rel_err(got, per_tap) < 1e-6
        xr = toks[i][:, 1:].double().cpu().view(B, gh, gw, C).permute(0, 3, 1, 2).contiguous()
        gr = r(dy[:, i * F_:(i + 1) * F_]).view(B, gh, gw, F_).permute(0, 3, 1, 2).contiguous()
        ref = torch.nn.grad.conv2d_weight(xr, (F_, C, 3, 3), gr, padding=1)
        assert rel_err(got, ref) < (2e-5 if split else 1e-4), (i, rel_err(got, ref))


def test_training_step_seg_only_config_against_oracle():
    """A seg-only configuration (no depth head, denseclip.py:343 skipped) in fp32-class precision: main_output and the gradients of
    every neck / decode-head parameter against torch autograd over the oracle's training forward; depth_output is None."""
    import denseclip_vit_multimodal_b200 as D
    from denseclip_vit_multimodal_b200.losses import CrossEntropyLoss
    cfg = O.model_config("tiny", 2)
    cfg.pop("depth_head")
    model = D.DenseCLIP(**copy.deepcopy(cfg), precision="fp32")
    sd = O.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 5)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    for n, p in model.named_parameters():
        p.requires_grad = not (n.startswith('backbone.') or n.startswith('text_encoder.'))
    img = O.synthetic_images(2, 32, 64, seed=105)
    seg_t, _, _ = O.synthetic_targets(2, 32, 64, seed=205)
    out = model(img.cuda(), gt_semantic_seg=seg_t.cuda(), return_loss=True)
    assert out["depth_output"] is None
    CrossEntropyLoss(ignore_index=255)(out["main_output"], seg_t.cuda()).backward()
    sdg = {k: (v.clone().requires_grad_(True) if k.startswith(O.TRAINABLE_PREFIXES) and v.is_floating_point() and 'running_' not in k else v)
           for k, v in sd.items()}
    main, depth, _ = O.train_forward(sdg, cfg, img, (32, 64))
    assert depth is None
    F.cross_entropy(main, seg_t, ignore_index=255).backward()
    assert rel_err(out["main_output"], main) < 1e-3
    named = dict(model.named_parameters())
    for k, v in sdg.items():
        if isinstance(v, torch.Tensor) and v.requires_grad:
            assert named[k].grad is not None and rel_err(named[k].grad, v.grad) < 5e-3, (k, rel_err(named[k].grad, v.grad))

#!/usr/bin/env python
"""bench.py -- DenseCLIP ViT-B/16 forward @512x1024 on N B200s (one process per GPU, batch sharded by image).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps 3 --warmup 1     # the reference algorithm's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM (CUDA-event timed, max over ranks);
`e2e` = the same metric through the public API (`DenseCLIP.predict`) with pinned HOST buffers, H2D and D2H inside the
timed region; `roofline` = the dominant kernel (tcgen05 flash attention) timed live with CUDA events against the
measured bf16 peak; `cpu_baseline` = the oracle port of the reference timed on the host cores (N=1, rank 0, bounded sample).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec DenseCLIP ViT-B/16 fwd @512x1024"
ENCODER_FLOPS_PER_IMAGE = 505.25e9          # patch 2.416 + QKV 87.016 + out 29.005 + MLP 232.041 + attn 154.770 GF
ATTN_FLOPS_PER_IMAGE_LAYER = 154.770e9 / 12  # QK^T + PV, 12 heads x 2049^2 x 64 x 2 x 2
# measured on this pool's B200s by the driver (MEASURED_PEAKS.json at the time of writing); re-read from the file if present
RECORDED_PEAKS = {"hbm_gbs": 6541.8, "bf16_tflops": 1674.0, "bf16_tflops_sustained": 1403.8}
ATTN_DRAM_TRAFFIC_BYTES = 236.0e6           # profiles/r01_attn_ncu_persistent.txt (ncu --set full): dram read 203.9 MB + write 32.1 MB per launch (B=16)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {k: float(d[k]) for k in RECORDED_PEAKS}, "MEASURED_PEAKS.json"
    return dict(RECORDED_PEAKS), "MEASURED_PEAKS.json values recorded in bench.py (file not shipped to this box)"


def model_kwargs(decoder_layers=3):
    """The live yaml model (configs/denseclip_cityscapes.yaml:18-72) + the canonical 3-layer ContextDecoder."""
    classes = ['road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'traffic light', 'traffic sign', 'vegetation', 'terrain',
               'sky', 'person', 'rider', 'car', 'truck', 'bus', 'train', 'motorcycle', 'bicycle']
    return dict(
        backbone=dict(type='CLIPVisionTransformer', patch_size=16, width=768, layers=12, heads=12, input_resolution=224,
                      output_dim=768, out_indices=list(range(12))),
        text_encoder=dict(type='CLIPTextContextEncoder', context_length=22, vocab_size=49408, transformer_width=512,
                          transformer_heads=8, transformer_layers=12, embed_dim=512),
        context_decoder=dict(type='ContextDecoder', transformer_width=256, transformer_heads=4, transformer_layers=decoder_layers,
                             visual_dim=512, dropout=0.1),
        neck=dict(type='ViTFeatureFusionNeck', inter_channels=128, out_channels=256),
        decode_head=dict(type='FPNHead', in_channels=256, channels=256, num_classes=19, align_corners=False, dropout_ratio=0.1),
        depth_head=dict(type='FCNHeadDepth', in_channels=256, channels=128, align_corners=False),
        class_names=classes, context_length=6, token_embed_dim=512, text_dim=512, context_feature='attention',
        score_concat_index=-1, text_head=False, tau=0.05)


def init_uninitialised(model):
    """The reference leaves text positional_embedding / text_projection as torch.empty; give them values."""
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        te = model.text_encoder
        te.positional_embedding.copy_(torch.randn(te.positional_embedding.shape, generator=g) * 0.01)
        te.text_projection.copy_(torch.randn(te.text_projection.shape, generator=g) * te.text_projection.shape[0] ** -0.5)
        for m in model.modules():  # non-trivial BatchNorm statistics so the folded-BN path is exercised
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        rows, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                clk, pw = float(r[0]), float(r[2])
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            rows.append((clk, pw))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = those drawing at least 60% of the highest power seen (the sampler also sees the idle edges
        # of the region, where the clock sits at its maximum)
        pmax = max((pw for _, pw in rows), default=0.0)
        load = sorted(clk for clk, pw in rows if pw >= 0.6 * pmax) or sorted(clk for clk, _ in rows)
        pw_load = sorted(pw for _, pw in rows if pw >= 0.6 * pmax)
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_min_mhz": load[0] if load else None, "sm_max_mhz": mx or None,
                "power_w": pw_load[len(pw_load) // 2] if pw_load else None, "reasons": sorted(reasons), "samples": len(rows),
                "samples_under_load": len(load)}


def run_reference(args, rank):
    """--impl reference: the reference algorithm's CPU forward (oracle port: the Python reference cannot travel to the GPU
    box) on all host cores.  A step = one forward of ONE synthetic 512x1024 image (bounded sample of the workload)."""
    if rank != 0:
        return
    from oracle import denseclip_oracle as O  # the only place besides cpu_baseline where bench.py executes oracle/
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = O.model_config("vit_b16", 3)
    import denseclip_vit_multimodal_b200 as D
    model = D.DenseCLIP(**copy.deepcopy(model_kwargs()))
    init_uninitialised(model)
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    img = O.synthetic_images(1, args.height, args.width, seed=0)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            O.denseclip_forward(sd, cfg, img)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.denseclip_forward(sd, cfg, img)
        dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = f"{args.steps} forwards of 1 synthetic {args.height}x{args.width} image, fp32, torch CPU ({threads} threads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_name(args):
    return (f"DenseCLIP ViT-B/16 forward, batch {args.batch}/GPU @{args.height}x{args.width}, 19 classes, seg+depth heads, "
            f"12-tap fusion neck, 3-layer ContextDecoder + score map (superset of BASELINE configs[1]); random init")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch kernels eagerly instead of replaying a captured graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import denseclip_vit_multimodal_b200 as D
    from denseclip_vit_multimodal_b200 import _lib, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: the native path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = D.DenseCLIP(**copy.deepcopy(model_kwargs()), precision=args.precision)
    init_uninitialised(model)
    model = model.eval().to(dev)
    if not args.no_cuda_graph:
        model.enable_cuda_graph(True)
    B, H, W = args.batch, args.height, args.width
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_imgs = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    dev_imgs = [h.to(dev) for h in host_imgs]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    with torch.no_grad():
        for i in range(args.warmup):
            model(dev_imgs[i % 2], return_loss=False)
        barrier()
        _lib.reset_launch_count(local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for i in range(args.steps):
                out = model(dev_imgs[i % 2], return_loss=False)
            e1.record()
            barrier()
        launches = _lib.launch_count(local)
        if not args.no_cuda_graph:  # replays do not pass through the C ABI: count = kernels captured per step x steps
            launches = int(getattr(model, "graph_launches_per_step", 0)) * args.steps
        ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        assert out["seg"].shape == (B, 19, H, W) and out["depth"].shape == (B, 1, H, W)
        del out

        # ---------------- encoder-only (the 60%-of-TC-peak target is quoted on the ViT encoder) ----------------
        for _ in range(2):
            model.backbone.forward_native(dev_imgs[0], taps_nchw=False, taps_tokens_bf16=True, last_tokens=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            model.backbone.forward_native(dev_imgs[i % 2], taps_nchw=False, taps_tokens_bf16=True, last_tokens=True)
        e1.record()
        barrier()
        enc_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)

        # ---------------- e2e: public API, pinned host buffers, H2D + D2H inside the timed region ----------------
        # PipelinedPredictor: every step copies that step's batch from pinned host memory and reads that step's result
        # (uint8 class map + fp32 depth) back to the host; copies of step i+1 / i-1 overlap the compute of step i.
        from denseclip_vit_multimodal_b200.pipeline import PipelinedPredictor
        pipe = PipelinedPredictor(model, (B, 3, H, W), dev)

        def e2e_run(n):
            checksum = 0
            for i in range(n):
                if i >= pipe.depth:
                    r = pipe.collect()
                    checksum += int(r["seg"][0, 0, 0])
                pipe.submit(host_imgs[i % 2])
                if world > 1 and i % 8 == 7:  # gather the class maps of all shards (the only collective on the path)
                    gathered = torch.empty(world * B, H, W, dtype=torch.uint8, device=dev)
                    dist.all_gather_into_tensor(gathered, pipe.seg_dev[i % pipe.depth])
            while pipe.n_collected < pipe.n_submitted:
                r = pipe.collect()
                checksum += int(r["seg"][0, 0, 0])
            return checksum

        e2e_run(4)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        barrier()
        h2d = host_imgs[0].numel() * 4
        d2h = pipe.seg_host[0].numel() + pipe.depth_host[0].numel() * 4

        # ---------------- roofline: dominant kernel (flash attention) timed live, alone ----------------
        Dm, Hh, Nt = 768, 12, (H // 16) * (W // 16) + 1
        qkv = (torch.randn(B, Nt, 3 * Dm, device=dev) * 2).to(torch.bfloat16)
        att = torch.empty(B, Nt, Dm, dtype=torch.bfloat16, device=dev)
        q_start = 0
        run_attn = lambda: ops.attention(qkv, qkv, qkv, B=B, H=Hh, Nq=Nt, Nk=Nt, q_col0=0, k_col0=Dm, v_col0=2 * Dm, scale=0.125,  # noqa: E731
                                         out=att, q_start=q_start)
        for _ in range(10):  # the clock needs a few ms to settle after the copy-bound e2e phase
            run_attn()
        torch.cuda.synchronize()
        reps = []
        for _ in range(3):   # three back-to-back averages over `steps` launches; the median is reported
            e0.record()
            for _ in range(args.steps):
                run_attn()
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1) / args.steps)
        attn_ms = sorted(reps)[1]
        del qkv, att

    pk, pk_src = peaks()
    attn_tflops = B * ATTN_FLOPS_PER_IMAGE_LAYER / (attn_ms * 1e-3) / 1e12
    enc_tflops = B * ENCODER_FLOPS_PER_IMAGE / (enc_ms * 1e-3) / 1e12
    clocks = clk.summary()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import denseclip_oracle as O  # checker / CPU baseline leg only
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        cfg = O.model_config("vit_b16", 3)
        img1 = host_imgs[0][:1].clone()
        with torch.no_grad():
            ref = O.denseclip_forward(sd, cfg, img1, return_intermediates=True)   # warm-up, also a live parity check
            t0 = time.perf_counter()
            n_cpu = 3
            for _ in range(n_cpu):
                O.denseclip_forward(sd, cfg, img1)
            dt = (time.perf_counter() - t0) / n_cpu
            got = model(dev_imgs[0][:1], return_loss=False)
        err = float((got["seg"].cpu() - ref["seg"]).abs().max() / ref["seg"].abs().max())
        cpu_baseline = {"value": 1.0 / dt, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": f"{n_cpu} forwards of 1 image {H}x{W} (same model/weights), fp32 torch CPU after 1 warm-up",
                        "seg_rel_err_native_vs_port": err}

    if rank == 0:
        total_imgs_per_s = world * B / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": total_imgs_per_s, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3-split (fp32-class)", "data": "synthetic",
            "config": {"workload": workload_name(args), "cuda_graph": not args.no_cuda_graph, "global_batch": world * B, "image": [H, W], "parallelism": f"dp{world} (batch sharded by image)",
                       "l2": "no explicit flush: each step streams ~4 GB of activations per GPU (>> 126 MB L2) and alternates 2 input batches"},
            "clocks": clocks,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "PipelinedPredictor(DenseCLIP.predict): pinned host images in, uint8 class map + fp32 depth back on the host, every step; copies overlap compute (wall-clock timed)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "attn_fwd_persistent_kernel (tcgen05 flash attention, one CTA per SM, 12 launches/step)", "bound": "tensor",
                         "achieved": attn_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": attn_tflops / pk["bf16_tflops"],
                         "traffic": ATTN_DRAM_TRAFFIC_BYTES * B / 16, "ms_per_launch": attn_ms, "peak_source": pk_src + " (burst: kernel timed alone)"},
            "encoder": {"ms_per_step": enc_ms, "tflops": enc_tflops, "flops_per_image": ENCODER_FLOPS_PER_IMAGE,
                        "frac_of_burst_peak": enc_tflops / pk["bf16_tflops"], "frac_of_sustained_peak": enc_tflops / pk["bf16_tflops_sustained"],
                        "images_per_s": world * B / (enc_ms * 1e-3)},
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

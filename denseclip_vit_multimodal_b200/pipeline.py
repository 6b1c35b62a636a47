"""Host <-> device pipelining for inference: while batch i runs on the compute stream, batch i+1 is copied in and the
results of batch i-1 are copied out on a second stream (double-buffered device staging + pinned host buffers).
PyTorch streams / events only (plumbing); the compute is ``DenseCLIP.predict`` (native kernels, CUDA-graph replay).

The reference has no counterpart: its validation loop does a blocking ``.to(device)`` per batch and reads metrics
back with ``.item()`` (segmentation/train_denseclip.py:428-435)."""
from __future__ import annotations

import torch


class PipelinedPredictor:
    """``submit(host_batch)`` enqueues one pinned-host batch [B,3,H,W] fp32; ``collect()`` returns the oldest finished
    result as pinned host tensors ``{'seg': uint8 [B,H,W], 'depth': fp32 [B,1,H,W]}`` (valid until two more submits)."""

    def __init__(self, model, batch_shape, device=None, depth: int = 2):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        B, _, H, W = batch_shape
        self.depth = depth
        # separate in / out copy streams: an H2D queued behind a D2H that waits for the running step would stall the next step
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.out_stream = torch.cuda.Stream(device=self.device)
        self.compute_stream = torch.cuda.Stream(device=self.device)
        self.in_dev = [torch.empty(batch_shape, dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.seg_dev = [torch.empty(B, H, W, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self.depth_dev = [torch.empty(B, 1, H, W, dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.seg_host = [torch.empty(B, H, W, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.depth_host = [torch.empty(B, 1, H, W, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]     # compute has consumed in_dev[s]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]     # results of slot s are on the host
        self.n_submitted = 0
        self.n_collected = 0
        self.has_depth = True
        # optional consumer of the DEVICE results of a slot (e.g. the multi-GPU gather / eval-statistics reduce of
        # bench.py): called as post(slot) right after the slot's results are complete on the compute stream (it must
        # wait on ev_out[slot] on its own stream) and returns an event that the next overwrite of the slot waits for
        self.post = None
        self.ev_post = [None] * depth

    @torch.no_grad()
    def submit(self, host_batch: torch.Tensor):
        if self.n_submitted - self.n_collected >= self.depth:
            raise RuntimeError("pipeline full: collect() before submitting more")
        s = self.n_submitted % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self.n_submitted >= self.depth:
                self.copy_stream.wait_event(self.ev_free[s])
            self.in_dev[s].copy_(host_batch, non_blocking=True)
            self.ev_in[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.ev_in[s])
            if self.ev_post[s] is not None:      # the previous results of this slot are still being consumed on the device
                self.compute_stream.wait_event(self.ev_post[s])
            out = self.model.predict(self.in_dev[s])
            self.ev_free[s].record(self.compute_stream)
            self.seg_dev[s].copy_(out["seg"], non_blocking=True)       # results may live in CUDA-graph static buffers
            if out.get("depth") is not None:
                self.depth_dev[s].copy_(out["depth"], non_blocking=True)
            else:
                self.has_depth = False
            self.ev_out[s].record(self.compute_stream)
        if self.post is not None:
            self.ev_post[s] = self.post(s)
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(self.ev_out[s])
            self.seg_host[s].copy_(self.seg_dev[s], non_blocking=True)
            if self.has_depth:
                self.depth_host[s].copy_(self.depth_dev[s], non_blocking=True)
            self.ev_done[s].record(self.out_stream)
        self.n_submitted += 1

    def collect(self):
        if self.n_collected >= self.n_submitted:
            raise RuntimeError("nothing to collect")
        s = self.n_collected % self.depth
        self.ev_done[s].synchronize()
        self.n_collected += 1
        return {"seg": self.seg_host[s], "depth": self.depth_host[s] if self.has_depth else None}

    def run(self, host_batches):
        """Generator over results for an iterable of pinned host batches (keeps `depth` batches in flight)."""
        for hb in host_batches:
            if self.n_submitted - self.n_collected >= self.depth:
                yield self.collect()
            self.submit(hb)
        while self.n_collected < self.n_submitted:
            yield self.collect()

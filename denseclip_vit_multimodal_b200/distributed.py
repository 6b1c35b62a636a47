"""Multi-GPU plumbing for the forward path: one process per GPU, the batch sharded BY IMAGE (every image is independent
through the whole forward -- SURVEY 8(e)), weights replicated.  The data path has no collective; NCCL (or gloo in the
CPU tests) is used only to gather the outputs and to reduce evaluation statistics.

Reference behaviour for comparison: DDP for training only, validation on rank 0 alone
(segmentation/train_denseclip.py:321-323, 1050-1054); eval statistics = torchmetrics mIoU/Acc/RMSE (:351-355).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _NoReduce:
    """sentinel: statistics are already reduced"""


_NO_REDUCE = _NoReduce()


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous [start, stop) slice of `n_items` images for `rank`; sizes differ by at most one, order is preserved."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(images: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    a, b = shard_range(images.shape[0], world, rank)
    return images[a:b]


def gather_shards(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank outputs [n_local, ...] (uint8 class maps, depth maps) back into batch order [n_total, ...].
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    max_n = max(b - a for a, b in sizes)
    pad = local
    if local.shape[0] < max_n:
        pad = torch.cat([local, local.new_zeros((max_n - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((world * max_n,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * max_n: r * max_n + (b - a)] for r, (a, b) in enumerate(sizes)])


def confusion_matrix(pred: torch.Tensor, target: torch.Tensor, num_classes: int, ignore_index: int = 255) -> torch.Tensor:
    """int64 [K, K] (rows = target, cols = prediction) over the local shard; pixels with `ignore_index` are skipped."""
    t = target.reshape(-1).to(torch.int64)
    p = pred.reshape(-1).to(torch.int64)
    keep = t != ignore_index
    idx = t[keep] * num_classes + p[keep]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def shard_eval_stats(pred: torch.Tensor, target: torch.Tensor, num_classes: int, ignore_index: int = 255, depth_pred=None,
                     depth_gt=None, depth_mask=None):
    """Evaluation statistics of the local shard, computed by the native kernel (`ops.eval_stats`, one pass over the uint8
    class map / depth map, nothing leaves the GPU): (conf int64 [K, K], depth_sq_err_sum float64 [], depth_count float64 []),
    ready for `reduce_eval_stats`.  Reference: torchmetrics JaccardIndex / Accuracy(ignore_index) and masked RMSE
    (train_denseclip.py:351-355, 582-593)."""
    from . import ops
    conf, ds = ops.eval_stats(pred, target, num_classes, ignore_index, depth_pred, depth_gt, depth_mask)
    return conf, ds[0].clone(), ds[1].clone()


def reduce_eval_stats(conf: torch.Tensor, depth_sq_err_sum: torch.Tensor, depth_count: torch.Tensor, group=None):
    """All-reduce (sum) the evaluation statistics of all shards: 19x19 confusion matrix + depth squared-error sum/count.
    Returns (conf, mIoU, pixel_acc, rmse)."""
    if not isinstance(group, _NoReduce) and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in (conf, depth_sq_err_sum, depth_count):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tp = conf.diag().double()
    union = conf.sum(0).double() + conf.sum(1).double() - tp
    valid = union > 0
    miou = float((tp[valid] / union[valid]).mean()) if valid.any() else float("nan")
    acc = float(tp.sum() / conf.sum().clamp(min=1))
    rmse = float(torch.sqrt(depth_sq_err_sum.double() / depth_count.double().clamp(min=1)))
    return conf, miou, acc, rmse


def reduce_eval_stats_packed(conf: torch.Tensor, depth_stats: torch.Tensor, group=None) -> torch.Tensor:
    """ONE collective per step: the int64 [K, K] confusion matrix and the float64 [2] depth (sum sq err, count) travel as a
    single float64 vector (counts are exact in float64 up to 2^53).  Returns the reduced vector [K*K + 2]; unpack with
    ``unpack_eval_stats``."""
    packed = torch.cat([conf.reshape(-1).to(torch.float64), depth_stats.reshape(-1).to(torch.float64)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def unpack_eval_stats(packed: torch.Tensor, num_classes: int):
    """-> (conf int64 [K, K], mIoU, pixel_acc, rmse) from the vector of ``reduce_eval_stats_packed``."""
    kk = num_classes * num_classes
    conf = packed[:kk].round().to(torch.int64).reshape(num_classes, num_classes)
    return reduce_eval_stats(conf, packed[kk].clone(), packed[kk + 1].clone(), group=_NO_REDUCE)


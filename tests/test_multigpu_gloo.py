"""CPU, world_size 2 over gloo: the N>1 path of the forward is 'shard the batch by image, gather outputs, reduce eval
statistics' -- no data-path collective.  This checks the sharding / gather / reduce plumbing end to end."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    from denseclip_vit_multimodal_b200 import distributed as dd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        batch = torch.randint(0, 19, (n_total, 8, 16), generator=g, dtype=torch.uint8)     # stand-in class maps
        target = torch.randint(0, 19, (n_total, 8, 16), generator=g, dtype=torch.int64)
        a, b = dd.shard_range(n_total, world, rank)
        local = batch[a:b].clone()                       # what this rank's forward would have produced
        full = dd.gather_shards(local, n_total)
        conf = dd.confusion_matrix(local, target[a:b], 19)
        err = (local.float() - target[a:b].float()) ** 2
        conf, miou, acc, rmse = dd.reduce_eval_stats(conf, err.sum(), torch.tensor(float(err.numel())))
        if rank == 0:
            ref_conf = dd.confusion_matrix(batch, target, 19)
            ref_rmse = float(torch.sqrt(((batch.float() - target.float()) ** 2).mean()))
            q.put((bool(torch.equal(full, batch)), bool(torch.equal(conf, ref_conf)), abs(rmse - ref_rmse) < 1e-5, acc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])   # even and ragged shards
def test_shard_gather_reduce_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500 + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    same, conf_ok, rmse_ok, acc = q.get(timeout=5)
    assert same and conf_ok and rmse_ok and 0 <= acc <= 1

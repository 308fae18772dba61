"""world_size-2 gloo test of the data-parallel exchange step: shard the batch, compute
per-replica gradients (oracle, per-replica BN), sum-all-reduce the flat arena, scale by 1/world."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from nvae_tf_b200 import parallel
    from oracle import nvae_oracle as O
    assert parallel.init_from_env("gloo") == world
    cfg = H.oracle_cfg(n_groups_per_scale=(1, 1), n_preprocess_cells=2, n_post_process_cells=1)
    params, trainable, bnl, s = O.build_params(cfg, seed=2 + rank, jitter=0.05)  # deliberately different per rank
    names = sorted(params)
    flat = torch.cat([torch.as_tensor(params[n]).reshape(-1) for n in names])
    parallel.broadcast_parameters([flat], src=0)
    off = 0
    for n in names:
        params[n] = flat[off:off + params[n].size].reshape(params[n].shape).numpy().copy()
        off += params[n].size
    x = O.make_images(cfg, 4, seed=2).numpy()
    eps = [e.numpy() for e in O.make_eps(s, 4, seed=2)]
    sl = parallel.shard_batch(4, rank, world)
    _, g, _, _ = H.run_oracle_step(cfg, params, trainable, bnl, s, x[sl], [e[sl] for e in eps], 50, True)
    gflat = torch.cat([torch.as_tensor(g[n]).reshape(-1) for n in trainable])
    local = gflat.clone()
    parallel.all_reduce_gradients(gflat, bucket_elems=1000)
    gflat /= world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    q.put((rank, float((gflat - sum(gathered) / world).abs().max()), float(gflat.abs().max()),
           float((gathered[0] - gathered[1]).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, mag, diff in res:
        assert err < 1e-12 and mag > 0  # bucketed all-reduce == mean of the per-rank gradients
        assert diff > 0                 # the shards really saw different samples

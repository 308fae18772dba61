"""Data-parallel correctness of the PRODUCT path on a GPU: two ranks, each running
the real kernels on half of the global batch with per-replica BatchNorm, exchange through NVAE.apply_gradients (the
all-reduce the product calls), and must (a) hold the average of the two per-replica gradients the float64 oracle
computes for the same shards, (b) end up with bit-identical parameters after k optimizer steps, (c) draw different
epsilons per rank.  Both ranks share cuda:0 (the test box has one GPU) and talk through gloo, which accepts CUDA
tensors; NCCL needs one device per rank and is exercised by bench.py --gpus N, which asserts (b) after its timed loop."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import helpers as H

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK="0")
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import torch.distributed as dist
        from nvae_tf_b200 import parallel
        from nvae_tf_b200.models import NVAE, Adamax, CosineDecay
        from oracle import nvae_oracle as O
        torch.cuda.set_device(0)
        assert parallel.init_from_env("gloo") == world
        cfg = H.oracle_cfg()
        GB = 8
        params, trainable, bnl, s = O.build_params(cfg, seed=4, jitter=0.1)
        params = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in params.items()}
        x = O.make_images(cfg, GB, seed=4).numpy()
        eps = [np.asarray(e.numpy(), np.float32).astype(np.float64) for e in O.make_eps(s, GB, seed=4)]
        sl = parallel.shard_batch(GB, rank, world)
        m = NVAE(**H.mirror_kwargs(cfg, GB // world), training=True, seed=1)
        m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)))
        seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(seeds, torch.tensor([m.rt.philox_seed % (1 << 62)], dtype=torch.int64))
        m.rt.load_named(params)
        m.rt.inject_eps([e[sl] for e in eps])
        m.steps = 10
        # one step WITHOUT the optimizer: reduce exactly as apply_gradients does, then compare with the oracle's emulation
        m.train_step(x[sl], apply_gradients=False)
        dist.all_reduce(m.rt.grads)
        got = {k: v / world for k, v in m.rt.named_grads().items()}
        want = None
        for r in range(world):  # per-replica BN statistics, KL-balance coefficients and batch mean, then the average
            rs = parallel.shard_batch(GB, r, world)
            _, g, _, _ = H.run_oracle_step(cfg, params, trainable, bnl, s, x[rs], [e[rs] for e in eps], 10, True)
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        worst = H.compare_grads(got, want, 1e-3)
        # k real steps through the product path (Philox epsilons, rank-folded key): replicas must stay identical
        m.rt.inject_eps(None)
        m.rt.load_named(params)
        for _ in range(3):
            m.train_step(x[sl])
        torch.cuda.synchronize()
        # the same 3 steps through the CAPTURED path bench.py times at N > 1: two graphs, the postprocess bucket all-reduced
        # while the second graph runs -- must land on bit-identical parameters (Philox is counter-keyed, sums of two ranks
        # commute)
        g = NVAE(**H.mirror_kwargs(cfg, GB // world), training=True, seed=1)
        g.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)))
        g.rt.load_named(params)
        g.steps = m.steps - 3        # the eager model took its 3 optimizer steps from here ...
        g._counters[1] = 1           # ... and from optimizer iteration 1 (the gradient-only step above advanced it)
        static_in, replay = g.capture_train_step((GB // world, 32, 32, 1))
        assert g._graph2 is not None, "overlapped two-graph step expected with more than one rank"
        static_in.copy_(torch.as_tensor(np.asarray(x[sl], np.float32)))
        for _ in range(3):
            replay()
        torch.cuda.synchronize()
        graph_equals_eager = bool(torch.equal(g.rt.params, m.rt.params))
        p = m.rt.params.double()
        mine = torch.stack([p.sum(), (p * p).sum(), p.abs().max()]).cpu()
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        q.put((rank, worst, [int(t.item()) for t in seeds], [t.tolist() for t in allv], graph_equals_eager))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of a queue timeout
        q.put((rank, ("exception: " + repr(e), 1e9), [], [], False))
        raise


@pytest.mark.timeout(600)
def test_two_replicas_average_to_the_oracle_and_stay_identical(lib_built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, worst, seeds, sums, graph_ok in res:
        assert worst[1] <= 1e-3, worst                 # averaged gradient == oracle's 2-replica emulation
        assert graph_ok                                # overlapped two-graph replay == eager steps, bit for bit
        assert seeds[0] != seeds[1]                    # the ranks draw different epsilons
        assert sums[0] == sums[1]                      # identical parameters on both ranks after 3 optimizer steps

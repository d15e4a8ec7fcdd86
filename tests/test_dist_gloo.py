""" Data-parallel sharding of the fit step (N>1 host logic) on CPU: two gloo ranks each run the
(emulated) device step on their shard with inv_n = 1/B_global, all-reduce [grads | loss], and must
reproduce the single-rank result; the Philox stream must not depend on the world size. """
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import emul_harness as E
    from helpers import load_golden
    from oracle import philox as ph
    from pydens_b200.engine import shard_batch
    g = load_golden('burgers')
    spec = E.spec_for('burgers')
    B = 101                                                   # ragged on purpose
    pts_all = ph.sample([(0, -1.0, 2.0), (0, 0.0, 3.0)], 2, seed=5, step=3, point_offset=0, n=B)
    n_loc, off = shard_batch(B, world, rank)
    pts = ph.sample([(0, -1.0, 2.0), (0, 0.0, 3.0)], 2, seed=5, step=3, point_offset=off, n=n_loc)
    assert np.array_equal(pts, pts_all[off:off + n_loc])
    # emul_step scales by 1/n_local: rescale to the global mean like the kernel's inv_global_n
    loss, _, grads = E.emul_step(spec, g['params'], pts)
    buf = torch.from_numpy(np.concatenate([grads, [loss]]).astype(np.float32) * (n_loc / B))
    dist.all_reduce(buf)
    if rank == 0:
        l1, _, g1 = E.emul_step(spec, g['params'], pts_all)
        ret['loss_err'] = abs(float(buf[-1]) - l1) / abs(l1)
        ret['grad_err'] = float(np.linalg.norm(buf[:-1].numpy() - g1) / np.linalg.norm(g1))
        ret['shards'] = [shard_batch(B, world, r) for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_rank():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret['shards'] == [(51, 0), (50, 51)]
    assert ret['loss_err'] <= 1e-6
    assert ret['grad_err'] <= 1e-5


def test_shard_batch_covers_everything():
    from pydens_b200.engine import shard_batch
    for B in (1, 7, 100, 100000, 4000000):
        for W in (1, 2, 4, 8):
            if B < W:
                continue
            parts = [shard_batch(B, W, r) for r in range(W)]
            assert sum(n for n, _ in parts) == B
            assert all(parts[r][1] == sum(n for n, _ in parts[:r]) for r in range(W))
            assert max(n for n, _ in parts) - min(n for n, _ in parts) <= 1

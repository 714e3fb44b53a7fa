"""N>1 host logic on CPU (gloo, world_size 2): frame sharding + all-reduce of the per-shard global
blocks reproduces the whole problem's reduced normal equations.  The CUDA path cannot run here, so
the oracle builds each shard's blocks; what is under test is vicalib_b200.synth.shard /
shard_frames and the reduction layout the device path uses (C, gc, cost summed; frame blocks local).
"""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Oracle
    from vicalib_b200 import synth

    p = synth.make_problem(models=("poly3", "fov"), n_frames=11, seed=3)
    ps = synth.shard(p, rank, world)
    ne = Oracle(ps).normal_equations()
    G = ne["C"].shape[0]
    buf = torch.from_numpy(np.concatenate([ne["C"].ravel(), ne["gc"], [ne["cost"]]]))
    dist.all_reduce(buf)
    f0, f1 = synth.shard_frames(p.n_frames, rank, world)
    if rank == 0:
        full = Oracle(p).normal_equations()
        red = buf.numpy()
        ok = (np.allclose(red[: G * G].reshape(G, G), full["C"], rtol=1e-12, atol=1e-9)
              and np.allclose(red[G * G: G * G + G], full["gc"], rtol=1e-12, atol=1e-9)
              and abs(red[-1] - full["cost"]) <= 1e-12 * full["cost"]
              and np.allclose(ne["B"], full["B"][f0:f1], rtol=1e-13) and np.allclose(ne["E"], full["E"][f0:f1], rtol=1e-13))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frames_partition():
    sys.path.insert(0, ROOT)
    from vicalib_b200 import synth

    for n in (1, 7, 64, 2000):
        for w in (1, 2, 3, 8):
            edges = [synth.shard_frames(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


def test_sharded_blocks_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert q.get(timeout=5) is True

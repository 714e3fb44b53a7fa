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


def _worker_inertial(rank, world, port, q):
    """Inertial shards: every rank but the last carries the next rank's first frame as a ghost and owns the IMU factor
    that reaches it.  Summed over the ranks, the global blocks, the cost and — ghost block of rank p added to the first
    block of rank p+1 — the separator frames' blocks are those of the whole problem."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Oracle
    from vicalib_b200 import synth

    flags = dict(inertial=1, bias_active=1, scale_active=1, optimize_ts=1)
    p = synth.make_problem(models=("poly3",), n_frames=13, inertial=True, seed=4)
    ps = synth.shard(p, rank, world)
    f0, f1 = synth.shard_frames(p.n_frames, rank, world)
    assert ps.n_frames == (f1 - f0) + (1 if rank < world - 1 else 0)
    ne = Oracle(ps, **flags).normal_equations()
    G = ne["C"].shape[0]
    fd = ne["B"].shape[1]
    # ghost block of this rank (zeros on the last rank) travels to the next rank's separator
    ghost = np.concatenate([ne["B"][-1].ravel(), ne["gf"][-1]]) if rank < world - 1 else np.zeros(fd * fd + fd)
    buf = torch.from_numpy(np.concatenate([ne["C"].ravel(), ne["gc"], [ne["cost"]], ghost]))
    dist.all_reduce(buf)
    if rank == world - 1:
        full = Oracle(p, **flags).normal_equations()
        red = buf.numpy()
        scale = np.abs(full["C"]).max()
        ok = (np.abs(red[: G * G].reshape(G, G) - full["C"]).max() <= 1e-11 * scale
              and np.abs(red[G * G: G * G + G] - full["gc"]).max() <= 1e-11 * max(np.abs(full["gc"]).max(), 1.0)
              and abs(red[G * G + G] - full["cost"]) <= 1e-12 * full["cost"])
        # this rank's first frame is a separator: its block = local part + the previous rank's ghost part
        gB = red[G * G + G + 1: G * G + G + 1 + fd * fd].reshape(fd, fd)
        gg = red[G * G + G + 1 + fd * fd:]
        ok = ok and np.abs(ne["B"][0] + gB - full["B"][f0]).max() <= 1e-11 * np.abs(full["B"][f0]).max()
        ok = ok and np.abs(ne["gf"][0] + gg - full["gf"][f0]).max() <= 1e-9 * max(np.abs(full["gf"][f0]).max(), 1.0)
        # interior frames and the coupling blocks U = H[f-1, f] are local
        ok = ok and np.allclose(ne["B"][1:], full["B"][f0 + 1:f1], rtol=1e-12, atol=1e-6)
        ok = ok and np.allclose(ne["U"][1:], full["U"][f0 + 1:f1], rtol=1e-12, atol=1e-6)
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_inertial_blocks_with_ghost_frames_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_inertial, args=(r, 2, 29733, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(180)
        assert pr.exitcode == 0
    assert q.get(timeout=5) is True


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

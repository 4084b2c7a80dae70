"""CPU: the N>1 host logic (work partition + final gather) on the gloo backend, world_size 2.
The per-block evaluator here is the CPU oracle standing in for a GPU rank's engine call."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from madaiemulator_b200 import datasets as ds
    from madaiemulator_b200 import sharding
    from oracle.pyoracle import PortOracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    X = ds.synthetic_design(48, 2)
    y = ds.synthetic_response(X)
    o = PortOracle(X, y, 1, 0)
    rng = np.random.default_rng(5)
    thetas = np.stack([np.array([rng.uniform(-5, -2), rng.uniform(0, 1), rng.uniform(0, 1)]) for _ in range(7)])

    def evaluate(block):
        out = []
        for th in block:
            r = o.loglik_grad(th)
            out.append(np.concatenate([[r["negL"]], r["grad"]]))
        return np.array(out).reshape(len(block), 4)

    full = sharding.shard_map_rows(evaluate, thetas)
    t = sharding.max_over_ranks(1.0 + rank)
    # the strong-scaling jobs of bench.py: components round-robin over ranks, query points in contiguous blocks
    ncomp = 5
    mine = sharding.round_robin(ncomp, world, rank)
    blocks = sharding.gather_row_blocks(np.array([[10.0 * c, c + 0.5] for c in mine]).reshape(len(mine), 2),
                                        [len(sharding.round_robin(ncomp, world, r)) for r in range(world)])
    comps = sharding.scatter_round_robin(blocks, ncomp)
    lo, hi = sharding.block_range(11, world, rank)
    qblocks = sharding.gather_row_blocks(np.arange(lo, hi, dtype=np.float64)[:, None] ** 2,
                                         [b - a for a, b in (sharding.block_range(11, world, r) for r in range(world))])
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put((full, t, np.array([evaluate(thetas[i:i + 1])[0] for i in range(7)]), comps, np.concatenate(qblocks)[:, 0]))


def test_block_ranges_cover_everything():
    from madaiemulator_b200 import sharding
    for n in (0, 1, 7, 8, 10 ** 7):
        for w in (1, 2, 3, 8):
            ends = [sharding.block_range(n, w, r) for r in range(w)]
            assert ends[0][0] == 0 and ends[-1][1] == n
            assert all(ends[i][1] == ends[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in ends) - min(h - l for l, h in ends) <= 1
    assert sharding.round_robin(8, 8, 3) == [3] and sharding.round_robin(10, 4, 1) == [1, 5, 9]


def test_shard_map_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, t, serial, comps, qall = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (7, 4)
    assert np.array_equal(full, serial)
    assert t == 2.0
    assert np.array_equal(comps, np.array([[10.0 * c, c + 0.5] for c in range(5)]))
    assert np.array_equal(qall, np.arange(11.0) ** 2)

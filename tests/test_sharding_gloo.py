"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a batch by problem index, each rank solves its block, rank 0
gathers in global order. The per-rank solver here is the CPU oracle (tests may use it); on the GPU box the same sharding code
feeds one BatchedLqSolver per rank (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

from ocs2_b200.sharding import all_shard_bounds, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("total,world", [(16384, 8), (16384, 1), (10, 4), (3, 8), (0, 2), (65536, 3)])
def test_shards_are_contiguous_disjoint_and_cover(total, world):
    b = all_shard_bounds(total, world)
    assert b[0][0] == 0 and sum(c for _, c in b) == total
    for (b0, c0), (b1, _) in zip(b, b[1:]):
        assert b0 + c0 == b1
    assert max(c for _, c in b) - min(c for _, c in b) <= 1


def test_shard_bounds_rejects_bad_ranks():
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from ocs2_b200.sharding import gather_arrays, shard_bounds
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, m, nc, N, dt = 6, 3, 0, 12, 0.01
    st = orc.make_settings(algorithm=0, reduced_form=True, hessian_multiple=1e-5, time_step=dt)
    begin, count = shard_bounds(total, world, rank)
    K, x = [], []
    for i in range(begin, begin + count):  # global problem index -> the same seeded problem on every layout of ranks
        pb, x0 = orc.generate_problem(3, i, 0, n, m, nc, N, dt)
        ref = orc.backward(st, pb)
        xs, _, _, _ = orc.rollout(st, pb, ref, x0)
        K.append(ref.K)
        x.append(xs)
    local = {"K": np.stack(K) if K else np.zeros((0, N + 1, m, n)), "x": np.stack(x) if x else np.zeros((0, N + 1, n))}
    full = gather_arrays(local, total, dist, dst=0)
    if rank == 0:
        np.savez(out_path, **full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo_shard_solve_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    total = 7  # ragged: ranks own 4 and 3 problems
    port = _free_port()
    out2 = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, port, total, out2), nprocs=2, join=True)
    port = _free_port()
    out1 = str(tmp_path / "w1.npz")
    mp.spawn(_worker, args=(1, port, total, out1), nprocs=1, join=True)
    a, b = np.load(out2), np.load(out1)
    assert a["K"].shape[0] == total
    assert np.array_equal(a["K"], b["K"]) and np.array_equal(a["x"], b["x"])  # bit-exact: sharding moves no arithmetic

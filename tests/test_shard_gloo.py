"""CPU, world_size 2 over gloo: the N>1 host logic — instance sharding and the whole-mix histogram all-reduce.
The per-rank "banks" here are CPU oracles (no GPU in this container); the GPU equivalent is
tests/test_ebu_gpu.py::test_sharded_banks_equal_single_bank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _oracle as O
    import _signals as S
    from meters_lv2_b200 import shard
    lo, cnt = shard.shard_range(n_total, rank, world)
    x = S.white(2 * n_total, 1024 * 260, seed=55)[2 * lo:2 * (lo + cnt)]      # identical global stream, local rows
    e = O.Ebu(cnt, 2); e.integr("start")
    for b in range(260):
        e.process(np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024]))
    mix = np.zeros(shard.MIX_WORDS, np.int64)
    for i in range(cnt):
        hm, hs, c = e.hist(i)
        mix[:751] += hm; mix[752:752 + 751] += hs; mix[1504:1508] += c
    t = torch.from_numpy(mix.astype(np.int32))
    shard.allreduce_mix(t)
    parts = shard.gather_results(e.read())
    if rank == 0:
        q.put((t.numpy().copy(), np.concatenate(parts)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_process():
    sys.path.insert(0, HERE)
    import _oracle as O
    import _signals as S
    from meters_lv2_b200 import shard
    n_total = 7                                   # odd: unbalanced shards 3 + 4
    assert [shard.shard_range(n_total, r, 2) for r in range(2)] == [(0, 3), (3, 4)]
    assert shard.shard_rows(8192, 2, 3, 8) == (6144, 2048)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    mix, res = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process truth
    x = S.white(2 * n_total, 1024 * 260, seed=55)
    e = O.Ebu(n_total, 2); e.integr("start")
    for b in range(260):
        e.process(np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024]))
    ref = np.zeros(shard.MIX_WORDS, np.int64)
    for i in range(n_total):
        hm, hs, c = e.hist(i)
        ref[:751] += hm; ref[752:752 + 751] += hs; ref[1504:1508] += c
    assert np.array_equal(mix, ref.astype(np.int32)), "integer all-reduce must be bit-exact"
    assert np.array_equal(res.view(np.uint32), e.read().view(np.uint32)), "sharded results == single-process results"

"""Host logic of the N>1 path on CPU: world_size 2, gloo.  The walk function injected here is the CPU
oracle (test infrastructure) -- what is under test is the row sharding into batches, the global-row RNG keying and
the batch-wise all-gather layout of pecanpy_b200/dist.py, which are device independent.  The CUDA kernels on two
ranks are covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_rows, batches, q, gather="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from pecanpy_b200.dist import shard_rows, sharded_walks
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hub400_sparseotf_n2v.npz"))
    start = z["start"][:total_rows]
    L = 12

    def walk_block(lo, hi, out_block):
        w = orc.walk_csr("SparseOTF", z["indptr"], z["indices"], z["data"], 4, 0.25, start[lo:hi], L,
                         rng=orc.RNG_PHILOX, seed=11, row0=lo, nthreads=1)
        out_block[: hi - lo].copy_(torch.from_numpy(w.view(np.int32)))

    full = sharded_walks(walk_block, total_rows, L + 2, "cpu", batches=batches, gather=gather)
    blocks, B = shard_rows(total_rows, world, rank, batches)
    q.put((rank, blocks, B, full.numpy().view(np.uint32).copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_rows,batches,gather", [(800, 1, "nccl"), (777, 1, "nccl"), (777, 3, "nccl"), (50, 4, "nccl"),
                                                        (777, 2, "mirror"), (777, 2, "push")])
def test_sharded_walks_match_single_process(total_rows, batches, gather):
    # ("mirror" / "push" need peer-mapped device matrices: on host tensors they must fall back to the all-gather)
    from oracle import oracle as orc
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hub400_sparseotf_n2v.npz"))
    want = orc.walk_csr("SparseOTF", z["indptr"], z["indices"], z["data"], 4, 0.25, z["start"][:total_rows], 12,
                        rng=orc.RNG_PHILOX, seed=11, nthreads=1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_rows, batches, q, gather)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = sorted(res, key=lambda t: t[0])
    covered = np.zeros(total_rows, dtype=np.int32)
    for _, blocks, B, full in res:
        assert len(blocks) == batches
        for lo, hi in blocks:
            covered[lo:hi] += 1
        assert np.array_equal(full, want)      # every rank ends with the whole matrix
    assert (covered == 1).all()                # the ranks' blocks tile the rows exactly once


def test_shard_rows_cover_everything():
    from pecanpy_b200.dist import shard_rows
    for tot in [0, 1, 7, 8, 9, 1000, 10_000_000]:
        for world in [1, 2, 3, 4, 8]:
            for batches in [1, 2, 8]:
                covered = 0
                spans = []
                for r in range(world):
                    blocks, B = shard_rows(tot, world, r, batches)
                    assert B * world * batches >= tot
                    for b, (lo, hi) in enumerate(blocks):
                        assert hi - lo <= B and (hi == lo or lo == (b * world + r) * B)
                        spans.append((lo, hi))
                        covered += hi - lo
                assert covered == tot
                spans = sorted(s for s in spans if s[1] > s[0])
                for a, b2 in zip(spans, spans[1:]):
                    assert a[1] == b2[0]


def _sleep_worker(rank, world, port, fail, q):
    import sys
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    done_at = [None]

    def work():
        time.sleep(0.5)
        done_at[0] = time.time()
        if fail:
            raise RuntimeError("rank 0 failed")

    raised = False
    try:
        bench.on_rank0_while_others_sleep(dist, rank, work)
    except RuntimeError:
        raised = True
    q.put((rank, done_at[0], time.time(), raised))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail", [False, True])
def test_bench_other_ranks_sleep_until_rank0_is_done(fail):
    """bench.py's e2e leg at N > 1: rank 0 works, the others sleep on the CPU and leave only after it has finished;
    a failure on rank 0 is raised there and does not strand the others."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sleep_worker, args=(r, 3, port, fail, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t_done = res[0][1]
    assert t_done is not None
    for rank, _, left_at, raised in res:
        assert left_at >= t_done
        assert raised == (fail and rank == 0)

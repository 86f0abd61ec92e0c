"""Host logic of the N>1 path on CPU: world_size 2, gloo.  The walk function injected here is the CPU
oracle (test infrastructure) -- what is under test is the row sharding, the global-row RNG keying and
the all-gather layout of pecanpy_b200/dist.py, which are device independent."""
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


def _worker(rank, world, port, total_rows, q):
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

    full = sharded_walks(walk_block, total_rows, L + 2, "cpu")
    lo, hi, R = shard_rows(total_rows, world, rank)
    q.put((rank, lo, hi, R, full.numpy().view(np.uint32).copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_rows", [800, 777])
def test_sharded_walks_match_single_process(total_rows):
    from oracle import oracle as orc
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hub400_sparseotf_n2v.npz"))
    want = orc.walk_csr("SparseOTF", z["indptr"], z["indices"], z["data"], 4, 0.25, z["start"][:total_rows], 12,
                        rng=orc.RNG_PHILOX, seed=11, nthreads=1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blocks = sorted(res)
    assert blocks[0][1] == 0 and blocks[0][2] == blocks[1][1] and blocks[1][2] == total_rows
    for _, _, _, _, full in blocks:           # every rank ends with the whole matrix
        assert np.array_equal(full, want)


def test_shard_rows_cover_everything():
    from pecanpy_b200.dist import shard_rows
    for tot in [0, 1, 7, 8, 9, 1000, 10_000_000]:
        for world in [1, 2, 3, 4, 8]:
            spans = [shard_rows(tot, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == tot
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[2] * world >= tot for s in spans)

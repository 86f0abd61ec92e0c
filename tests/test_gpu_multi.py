"""More than one GPU worker on the CUDA path: (a) b2w_walk_multi -- several graph replicas driven from one process,
one host matrix; (b) two torch.distributed ranks running the CUDA kernels on their row blocks and all-gathering
batch by batch.  Both must reproduce the single-worker matrix bit for bit (Philox is keyed by the global row).
On a one-GPU box the replicas / ranks share cuda:0 (the gather of (b) is then staged through gloo); with two or more
GPUs they spread out and (b) uses NCCL."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _graph():
    from pecanpy_b200.synth import power_law_csr
    return power_law_csr(20000, 400000, seed=7)


@pytest.mark.parametrize("mode", ["SparseOTF", "PreComp", "SparseOTF+extend"])
@pytest.mark.parametrize("replicas", [2, 3])
def test_walk_multi_equals_single_device(mode, replicas):
    import torch
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.multi import walk_host_engines
    from pecanpy_b200.synth import power_law_csr
    extend = mode.endswith("+extend")
    mode = mode.split("+")[0]
    indptr, indices, data = power_law_csr(20000, 400000, seed=7, weighted=(mode == "PreComp" or extend))
    ndev = torch.cuda.device_count()
    engines = []
    for k in range(replicas):
        e = WalkEngine.from_csr(indptr, indices, data, device=f"cuda:{k % ndev}")
        if extend:
            e.compute_thresholds(0.25)
        if mode == "PreComp":
            e.build_alias(indptr, 0.5, 2.0)
        engines.append(e)
    start = orc.shuffled_start(20000, 3, 1)[:50001]                  # not a multiple of the replica count
    L = 25
    got = walk_host_engines(engines, mode, 0.5, 2.0, extend, start, L, seed=9)
    want = engines[0].walk_host(mode, 0.5, 2.0, start, L, seed=9, extend=extend)
    assert np.array_equal(got, want)
    assert walk_host_engines.last_stats["steps"] == int((want[:, -1].astype(np.int64) - 1).sum())
    k = 4000
    alias = orc.alias_build(indptr, indices, data, 0.5, 2.0) if mode == "PreComp" else None
    thr = orc.noise_thresholds_csr(indptr, data, 0.25) if extend else None
    ref = orc.walk_csr(mode, indptr, indices, data, 0.5, 2.0, start[:k], L, extend=extend, thr=thr, alias=alias,
                       rng=orc.RNG_PHILOX, seed=9)
    assert np.array_equal(got[:k], ref)
    for e in engines:
        e.close()


def test_dropin_devices_attribute_uses_every_listed_gpu():
    """model.devices = [...]: simulate_walks_array shards over them from this process and returns the same matrix."""
    import torch
    from pecanpy_b200 import pecanpy as b2
    indptr, indices, data = _graph()
    ndev = torch.cuda.device_count()
    one = b2.SparseOTF(p=4, q=0.25, random_state=3)
    one.indptr, one.indices, one.data = indptr, indices, data
    one.set_node_ids(None, implicit_ids=True, num_nodes=indptr.size - 1)
    want = one.simulate_walks_array(2, 30)
    many = b2.SparseOTF(p=4, q=0.25, random_state=3)
    many.indptr, many.indices, many.data = indptr, indices, data
    many.set_node_ids(None, implicit_ids=True, num_nodes=indptr.size - 1)
    many.devices = [f"cuda:{k % ndev}" for k in range(max(2, ndev))] if ndev > 1 else ["cuda:0", "cuda:0"]
    if ndev == 1:
        # two replicas on the one GPU: the second one is a separate handle (the model's own engine serves the first)
        from pecanpy_b200.multi import walk_host_engines
        e2 = many._make_engine(device="cuda:0")
        got = walk_host_engines([many.engine, e2], "SparseOTF", 4, 0.25, False, many._start_nodes(2), 30, 3)
        e2.close()
    else:
        got = many.simulate_walks_array(2, 30)
        assert len(many.last_multi_stats["devices"]) == len(many.devices)
    assert np.array_equal(got, want)
    one.release(); many.release()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, batches, gather, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    backend = "nccl" if ndev >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    from oracle import oracle as orc
    from pecanpy_b200.dist import simulate_walks_distributed
    from pecanpy_b200.engine import WalkEngine
    indptr, indices, data = _graph()
    eng = WalkEngine.from_csr(indptr, indices, data, device=dev)
    start = orc.shuffled_start(20000, 2, 5)[:33333]
    if gather in ("push", "mirror") and ndev < world:
        gather = "nccl"                                     # the peers' matrices are on the same device: nothing to map
    full = simulate_walks_distributed(eng, "SparseOTF", 4.0, 0.25, start, 30, seed=21, batches=batches, gather=gather)
    torch.cuda.synchronize()
    q.put((rank, backend, eng.kernel_name("SparseOTF", 4.0, 0.25), full.cpu().numpy().view(np.uint32).copy()))
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.parametrize("batches,gather", [(1, "nccl"), (4, "nccl"), (3, "push"), (1, "mirror"), (2, "mirror")])
def test_two_ranks_cuda_walks_equal_single_rank(batches, gather):
    """CUDA kernels on two ranks (row0 offsets, batch-interleaved blocks, all-gather): every rank's matrix equals
    the one-rank matrix and the oracle."""
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    indptr, indices, data = _graph()
    start = orc.shuffled_start(20000, 2, 5)[:33333]
    eng = WalkEngine.from_csr(indptr, indices, data)
    want = eng.walk("SparseOTF", 4.0, 0.25, start, 30, seed=21).cpu().numpy().view(np.uint32)
    eng.close()
    ref = orc.walk_csr("SparseOTF", indptr, indices, data, 4.0, 0.25, start[:3000], 30, rng=orc.RNG_PHILOX, seed=21)
    assert np.array_equal(want[:3000], ref)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, batches, gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, backend, kernel, full in res:
        assert kernel == "walk_uw_edge_kernel"
        assert np.array_equal(full, want), (rank, backend)


@pytest.mark.parametrize("n_mirrors,row_off,L,pad", [(1, 0, 31, 0), (3, 5, 31, 0), (7, 2, 80, 0), (2, 1, 5, 0), (2, 3, 47, 3)])
def test_mirrored_walk_writes_every_mirror(n_mirrors, row_off, L, pad):
    """b2w_walk_mirrored (the all-gather fused into the kernel): the rows land, identical, in the local matrix and in
    every mirror -- here other buffers of the same GPU stand in for the peers' matrices -- and nothing else of the
    mirrors is touched.  Odd row lengths put the rows at all eight sector phases; `pad` widens the leading dimension."""
    import torch
    from pecanpy_b200.engine import WalkEngine
    indptr, indices, data = _graph()
    from oracle import oracle as orc
    start = orc.shuffled_start(20000, 1, 9)[:7777]
    eng = WalkEngine.from_csr(indptr, indices, data)
    want = eng.walk("SparseOTF", 4.0, 0.25, start, L, seed=3, row0=row_off).cpu().numpy()
    rows = start.size + row_off + 3
    local = torch.full((rows, L + 2 + pad), -7, dtype=torch.int32, device=eng.device)
    mirrors = [torch.full((rows, L + 2 + pad), -7, dtype=torch.int32, device=eng.device) for _ in range(n_mirrors)]
    out = local[row_off:row_off + start.size, :L + 2]
    ptrs = [m[row_off:].data_ptr() for m in mirrors]
    eng.walk("SparseOTF", 4.0, 0.25, start, L, seed=3, row0=row_off, out=out, mirrors=ptrs)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    for m in mirrors:
        assert torch.equal(m, local)
    assert int((local[:row_off] != -7).sum()) == 0 and int((local[row_off + start.size:] != -7).sum()) == 0
    assert int((local[:, L + 2:] != -7).sum()) == 0
    # kernels that do not mirror refuse instead of silently skipping the peers
    with pytest.raises(Exception):
        eng.walk("SparseOTF", 4.0, 0.25, start, L, seed=3, out=out, mirrors=ptrs, flags=0x40)   # B2W_FLAG_NO_EDGE_INDEX
    with pytest.raises(Exception):
        eng.walk("SparseOTF", 4.0, 0.25, start, L, seed=3, out=out, mirrors=[ptrs[0] + 4])       # not sector-congruent
    eng.close()


def test_allgather_rows_entry_point():
    """b2w_allgather_rows (SURVEY 8b's name for the trailing all-gather; here copy engines over mapped matrices): each
    of three `ranks` -- three matrices of one GPU stand in for the mapped peers -- pushes its own row block; afterwards
    all matrices are equal and hold every block."""
    import ctypes as C
    import torch
    from pecanpy_b200 import _capi as capi
    lib = capi.lib()
    world, R, ld = 3, 1000, 82
    mats = [torch.zeros((world * R, ld), dtype=torch.int32, device="cuda:0") for _ in range(world)]
    for r, m in enumerate(mats):
        m[r * R:(r + 1) * R] = torch.randint(1, 1 << 30, (R, ld), dtype=torch.int32, device="cuda:0")
    want = torch.cat([mats[r][r * R:(r + 1) * R] for r in range(world)])
    peers = (C.c_void_p * world)(*[m.data_ptr() for m in mats])
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        capi.check(lib.b2w_allgather_rows(0, peers, world, r, R, ld, C.c_void_p(st)), "b2w_allgather_rows")
    torch.cuda.synchronize()
    for m in mats:
        assert torch.equal(m, want)
    assert lib.b2w_allgather_rows(0, peers, world, world, R, ld, C.c_void_p(st)) != capi.OK     # self out of range


def test_push_rows_streams_entry_point():
    """b2w_push_rows_streams (one stream per peer, what PeerMatrix.push uses): arbitrary row blocks of each `rank` land
    in every other matrix; again three matrices of one GPU stand in for the mapped peers."""
    import ctypes as C
    import torch
    from pecanpy_b200 import _capi as capi
    lib = capi.lib()
    world, R, ld = 3, 777, 33
    mats = [torch.zeros((world * R, ld), dtype=torch.int32, device="cuda:0") for _ in range(world)]
    for r, m in enumerate(mats):
        m[r * R:(r + 1) * R] = torch.randint(1, 1 << 30, (R, ld), dtype=torch.int32, device="cuda:0")
    want = torch.cat([mats[r][r * R:(r + 1) * R] for r in range(world)])
    torch.cuda.synchronize()                                # the side streams below do not wait for the default stream
    peers = (C.c_void_p * world)(*[m.data_ptr() for m in mats])
    streams = [torch.cuda.Stream(device="cuda:0") for _ in range(world)]
    sptr = (C.c_void_p * world)(*[s.cuda_stream for s in streams])
    for r in range(world):
        half = R // 2                                       # two pushes per rank: [r R, r R + half) and the rest
        capi.check(lib.b2w_push_rows_streams(0, peers, world, r, r * R, half, 4 * ld, sptr), "b2w_push_rows_streams")
        capi.check(lib.b2w_push_rows_streams(0, peers, world, r, r * R + half, R - half, 4 * ld, sptr),
                   "b2w_push_rows_streams")
    torch.cuda.synchronize()
    for m in mats:
        assert torch.equal(m, want)

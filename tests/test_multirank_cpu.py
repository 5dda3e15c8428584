"""World-size-2 gloo test (CPU) of the N>1 host logic of corpus retrieval: the shard plan
(contiguous video ranges, global moment-id bases), the all-gather layout [P, Q, k] and the merge
rule (ascending (score, id)) reproduce the single-bank top-k.  The per-shard top-k and the merge are
played by the CPU oracle here - the CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vfr_b200  # noqa: F401
from vfr_b200 import synth
from vfr_b200.retrieval import shard_range
from oracle import cal_oracle as orc

V, S, D, Q, K = 40, 6, 16, 9, 12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_topk(clips, qs, v0, v1):
    full = orc.score_matrix(clips[v0 * S:v1 * S], np.arange(v1 - v0 + 1) * S, qs).numpy()
    ids = np.arange(full.shape[1], dtype=np.int64) + v0 * 21
    order = np.argsort(full, axis=1, kind="stable")[:, :K]
    return np.take_along_axis(full, order, axis=1), ids[order]


def _merge(scores, ids):
    P, Qn, k = scores.shape
    out_s, out_i = np.empty((Qn, k), np.float32), np.empty((Qn, k), np.int64)
    for q in range(Qn):
        s, i = scores[:, q].reshape(-1), ids[:, q].reshape(-1)
        order = np.lexsort((i, s))[:k]
        out_s[q], out_i[q] = s[order], i[order]
    return out_s, out_i


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clips = synth.make_bank(7, V, S, D)
    qs = synth.make_query_embeddings(7, Q, D)
    v0, v1 = shard_range(V, rank, world)
    s, i = _local_topk(clips, qs, v0, v1)
    gs = [torch.empty(Q, K) for _ in range(world)]
    gi = [torch.empty(Q, K, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gs, torch.from_numpy(s))
    dist.all_gather(gi, torch.from_numpy(i))
    ms, mi = _merge(torch.stack(gs).numpy(), torch.stack(gi).numpy())
    ws, wi = _local_topk(clips, qs, 0, V)
    ok = np.array_equal(ms, ws) and np.array_equal(mi, wi)
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_two_rank_sharded_topk_merge_equals_global():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_ranges_partition_the_bank():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (1094, 4)):
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _pooled_worker(rank, world, port, ret):
    """The pooled-sample threshold protocol of the sharded filter + refine search (retrieval._sel_score_sharded), its
    host logic played with numpy: strided sample per shard -> 32 smallest per query -> all-gather -> j-th smallest of
    the union (j from the library's own vfr_sel_sample_rank) -> per-shard counts -> all-reduce(sum) certificate."""
    from vfr_b200 import _lib
    from vfr_b200.retrieval import _DistComm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = _DistComm()
    Vb, Qn, k = 6000, 7, 5
    clips = synth.make_bank(11, Vb, S, D)
    qs = synth.make_query_embeddings(11, Qn, D)
    v0, v1 = shard_range(Vb, rank, world)
    mine = clips[v0 * S:v1 * S]
    d2 = ((mine[None, :, :] - qs[:, None, :] + 1e-6) ** 2).sum(-1).astype(np.float32)        # [Q, shard clips]
    sample = d2[:, ::16]                                                                      # strided sample
    smallest = torch.from_numpy(np.sort(sample, axis=1)[:, :32].copy())
    sizes = comm.all_reduce_sum(torch.tensor([sample.shape[1], d2.shape[1]], dtype=torch.int64)).tolist()
    j = _lib.load().vfr_sel_sample_rank(k, sizes[0], sizes[1])
    pooled = comm.all_gather(smallest)                                                        # [P, Q, 32]
    ok = pooled.shape == (world, Qn, 32) and 1 <= j <= 32
    bound = torch.kthvalue(pooled.permute(1, 0, 2).reshape(Qn, -1), j, dim=1).values
    count = comm.all_reduce_sum(torch.from_numpy((d2 <= bound.numpy()[:, None]).sum(1).astype(np.int32)))
    ok = ok and bool((count >= k).all())                       # the certificate holds on this (unordered) bank
    # and then the bound covers the global k closest clips: every shard may drop what lies above it
    full = ((clips[None, :, :] - qs[:, None, :] + 1e-6) ** 2).sum(-1).astype(np.float32)
    kth = np.sort(full, axis=1)[:, k - 1]
    ok = ok and bool((kth <= bound.numpy()).all())
    mn = comm.all_reduce_min(torch.tensor([float(rank + 3)]))
    ok = ok and float(mn.item()) == 3.0
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_two_rank_pooled_sample_bound_protocol():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pooled_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_sample_rank_is_the_smallest_safe_rank():
    """vfr_sel_sample_rank(k, sampled, total): the smallest j <= 32 with P(Poisson(k * sampled / total) >= j) < 1e-10
    (0 if there is none) - checked against scipy's survival function."""
    from scipy.stats import poisson
    from vfr_b200 import _lib
    lib = _lib.load()
    for k, n_s, n in ((100, 16384 * 8, 6_000_000), (1, 16384, 6_000_000), (100, 131072, 6_000_000), (128, 4096, 48000),
                      (10, 2048, 48128), (100, 500_000, 1_000_000)):
        j = lib.vfr_sel_sample_rank(k, n_s, n)
        x = k * n_s / n
        if j == 0:
            assert poisson.sf(31, x) >= 1e-10                  # not even rank 32 is safe
        else:
            assert 1 <= j <= 32 and poisson.sf(j - 1, x) < 1.01e-10
            assert j == 1 or poisson.sf(j - 2, x) >= 0.99e-10
    assert lib.vfr_sel_sample_rank(0, 10, 10) == 0 and lib.vfr_sel_sample_rank(5, 0, 10) == 0
    assert lib.vfr_sel_tiles(257) == 2 and lib.vfr_sel_tiles(0) == 0


def _a2a_worker(rank, world, port, ret):
    """The query-slice exchange: rank r sends record j to rank j and ends up with the P records of slice r."""
    from vfr_b200.retrieval import _DistComm, slice_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = _DistComm()
    per, k = slice_rows(7, world), 3
    blk = per * (k * 12 + 4)
    send = torch.empty((world, blk), dtype=torch.uint8)
    for j in range(world):
        send[j] = (rank * 16 + j)                       # record (src = rank, dst = j)
    recv = comm.all_to_all(send)
    ok = all(bool((recv[j] == (j * 16 + rank)).all()) for j in range(world))
    gathered = comm.all_gather(torch.tensor([rank, rank + 10], dtype=torch.int32), out=torch.empty(2 * world, dtype=torch.int32))
    ok = ok and gathered.tolist() == [[r, r + 10] for r in range(world)]
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_two_rank_query_slice_all_to_all():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_query_slices_partition_the_batch():
    from vfr_b200.retrieval import slice_rows
    for q, w in ((37888, 8), (37888, 2), (151, 4), (3, 4), (1, 8), (100000, 8)):
        per = slice_rows(q, w)
        assert per % 2 == 0 and per * w >= q and (per - 2) * w < q
        rows = [(min(r * per, q), min((r + 1) * per, q)) for r in range(w)]
        assert rows[0][0] == 0 and rows[-1][1] == q and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))

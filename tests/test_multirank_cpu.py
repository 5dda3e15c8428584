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

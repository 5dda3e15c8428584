"""Run under torchrun on N GPUs: sharded retrieval (K3 by query slice -> all-gather -> K4 per shard with the threshold
protocol -> all-to-all of the query-slice records -> K7; every rank returns ITS slice of the batch) must equal the rows of
the single-bank search computed on every rank from the full bank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import vfr_b200
from vfr_b200 import models, synth
from vfr_b200.retrieval import MomentRetriever, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, S, Q, k = 60000, 6, 1000, 100
sd = synth.make_state_dict(3, 8, 500, spread=4.0)
model = models.CALModel(visual_input_dim=18, pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]))
model.load_state_dict({n: torch.from_numpy(v) for n, v in sd.items()})
model = model.to(dev).eval()
clips = torch.from_numpy(synth.make_bank(3, V, S, 100)).to(dev)
tokens = synth.make_queries(3, synth.make_videos(3, 4, 8), Q, 500)["tokens"]
ok = True
for engine in ("sel", "tc", "exact"):
    v0, v1 = shard_range(V, rank, world)
    shard = MomentRetriever(model, clips[v0 * S:v1 * S], np.arange(v1 - v0 + 1) * S, id_base=v0 * 21, max_queries=Q, k=k,
                            engine=engine, text_engine="tc" if engine == "sel" else engine)
    s, i = shard.search(tokens)                      # host in, host out, all-gather + merge inside
    s, i = s.clone(), i.clone()
    full = MomentRetriever(model, clips, np.arange(V + 1) * S, max_queries=Q, k=k, engine=engine,
                           text_engine="tc" if engine == "sel" else engine, world=1, rank=0)   # single-bank reference on every rank
    fs, fi = full.search(tokens)
    q0, q1 = shard.owned_range(Q)
    same = torch.equal(s, fs[q0:q1]) and torch.equal(i, fi[q0:q1])
    t = torch.tensor([int(same)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"engine={engine}: sharded({world}) == single bank: {bool(t.item())}")
    ok = ok and bool(t.item())
# a bank large enough for the pooled-sample protocol (one list per query: n_split = 1)
V2 = 400000
clips2 = torch.from_numpy(synth.make_bank(5, V2, S, 100)).to(dev)
v0, v1 = shard_range(V2, rank, world)
shard = MomentRetriever(model, clips2[v0 * S:v1 * S], np.arange(v1 - v0 + 1) * S, id_base=v0 * 21, max_queries=Q, k=k,
                        engine="sel", text_engine="tc", n_split=1)
s, i = shard.search(tokens)
s, i = s.clone(), i.clone()
pooled_ran = shard._sel_rank_cache.get(Q, (0, 0))[0] > 0
full = MomentRetriever(model, clips2, np.arange(V2 + 1) * S, max_queries=Q, k=k, engine="exact", text_engine="tc", world=1, rank=0)
fs, fi = full.search(tokens)
q0, q1 = shard.owned_range(Q)
same = torch.equal(s, fs[q0:q1]) and torch.equal(i, fi[q0:q1]) and pooled_ran
t = torch.tensor([int(same)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"engine=sel, pooled-sample protocol (ran: {pooled_ran}, fix-ups: {getattr(shard, 'n_fixups', 0)}): sharded({world}) == exact single bank: {bool(t.item())}")
ok = ok and bool(t.item())
del clips2, shard, full
# exact rank of the first positive over the sharded bank (tau all-reduce(min) + count all-reduce(sum), NCCL)
from vfr_b200 import ops, evaluate as vev
from vfr_b200.retrieval import sharded_rank_first_positive
videos = synth.make_videos(3, 2000, 8)
queries = synth.make_queries(3, videos, 400, 500)
nseg = np.array([v["num_segments"] for v in videos])
vo = np.concatenate([[0], np.cumsum(nseg)])
g = torch.Generator(device=dev).manual_seed(7)
bank_clips = torch.randn(int(vo[-1]), 100, device=dev, generator=g) * 0.3
q_emb = torch.randn(400, 100, device=dev, generator=g) * 0.3
full_bank = ops.Bank(bank_clips, vo)
want = vev.rank_first_positive(full_bank, q_emb, queries["video_idx"], queries["times"], [0.5, 0.7])["rank"]
a, b = shard_range(2000, rank, world)
shard_bank = ops.Bank(bank_clips[int(vo[a]):int(vo[b])], vo[a:b + 1] - vo[a])
got, _ = sharded_rank_first_positive(shard_bank, a, q_emb, queries["video_idx"], nseg[queries["video_idx"]], queries["times"], [0.5, 0.7])
same = bool(np.array_equal(got.cpu().numpy(), want))
t = torch.tensor([int(same)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded exact rank of the first positive ({world} shards, NCCL all-reduce) == single bank: {bool(t.item())}")
ok = ok and bool(t.item())
dist.destroy_process_group()
sys.exit(0 if ok else 1)

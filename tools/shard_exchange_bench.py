"""One GPU plays ONE of P shards of a V-video bank: K4 of the shard with and without the mid-scan bound exchange
(the bound it would receive from the all-reduce(min) is computed beforehand from all P shards).  Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops, _lib
V, P, Q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
S, D, k = 6, 100, int(os.environ.get("K", "100"))
FRAC = int(os.environ.get("FIRST_DIV", "8"))
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
clips = ((torch.randn(V, 1, D, device="cuda", generator=g) + 0.6 * torch.randn(V, S, D, device="cuda", generator=g)) * 0.05).reshape(-1, D)
q = torch.randn(Q, D, device="cuda", generator=g) * 0.06
stream = torch.cuda.current_stream().cuda_stream
per = V // P
banks = [ops.Bank(clips[r * per * S:(r + 1) * per * S], np.arange(per + 1) * S) for r in range(P)]
nc = per * S
tiles = lib.vfr_sel_tiles(nc)
first = min(tiles, max(32, tiles // FRAC))
qp = torch.empty(lib.vfr_sel_query_bytes(Q, 100), dtype=torch.uint8, device="cuda")
ws = torch.empty(lib.vfr_sel_topk_bytes(Q, nc, 0), dtype=torch.uint8, device="cuda")
out_s = torch.empty((Q, k), dtype=torch.float32, device="cuda")
out_i = torch.empty((Q, k), dtype=torch.int64, device="cuda")
bound = torch.empty(Q, dtype=torch.float32, device="cuda")

def phase1(b):
    _lib.call("vfr_sel_query_pack", q.data_ptr(), Q, D, b.sel().data_ptr(), nc, qp.data_ptr(), stream)
    _lib.call("vfr_sel_filter", b.sel().data_ptr(), nc, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, 0, first, 0, stream)
    _lib.call("vfr_sel_bound_get", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, bound.data_ptr(), stream)

def rest(b, gb):
    _lib.call("vfr_sel_bound_put", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, gb.data_ptr(), stream)
    _lib.call("vfr_sel_filter", b.sel().data_ptr(), nc, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, first, tiles, 1, stream)
    _lib.call("vfr_sel_refine", b.clips.data_ptr(), b.vid_off.data_ptr(), b.mom_off.data_ptr(), b.n_videos, nc, b.n_max, D,
              qp.data_ptr(), q.data_ptr(), Q, k, 0, out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), 0, stream)

bounds = []
for b in banks:
    phase1(b)
    bounds.append(bound.clone())
gb = torch.stack(bounds).min(dim=0).values
own = bounds[0]

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b0 = banks[0]
t_plain = timeit(lambda: ops.score_topk_sel(b0, q, k))
t_p1 = timeit(lambda: phase1(b0))
t_x = timeit(lambda: (phase1(b0), rest(b0, gb)))
t_own = timeit(lambda: (phase1(b0), rest(b0, own)))
print(f"V={V} P={P} Q={Q} k={k} shard tiles={tiles} first={first}: one-call {t_plain:.2f} ms | phase1 {t_p1:.2f} ms | "
      f"two-phase own bound {t_own:.2f} ms | two-phase exchanged bound {t_x:.2f} ms   (ideal 1/P of the 1-GPU K4)", flush=True)

def pack(b): _lib.call("vfr_sel_query_pack", q.data_ptr(), Q, D, b.sel().data_ptr(), nc, qp.data_ptr(), stream)
def filt(b): _lib.call("vfr_sel_filter", b.sel().data_ptr(), nc, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, 0, tiles, 0, stream)
def refine(b): _lib.call("vfr_sel_refine", b.clips.data_ptr(), b.vid_off.data_ptr(), b.mom_off.data_ptr(), b.n_videos, nc, b.n_max, D,
              qp.data_ptr(), q.data_ptr(), Q, k, 0, out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), 0, stream)
pack(b0); filt(b0); refine(b0)
print(f"  breakdown of the one-call path: query pack {timeit(lambda: pack(b0)):.2f} ms | filter {timeit(lambda: filt(b0)):.2f} ms | refine {timeit(lambda: refine(b0)):.2f} ms")
os.environ["VFR_SEL_SAMPLE"] = "0"
print(f"  without the sample pass: filter {timeit(lambda: filt(b0)):.2f} ms")

# pooled-sample protocol: T from the samples of ALL shards
os.environ.pop("VFR_SEL_SAMPLE", None)
import ctypes as C
lists = lib.vfr_sel_sample_lists(Q, nc, 0, 100)
samp = torch.empty(Q * lists * 32, dtype=torch.float32, device="cuda")
count = torch.empty(Q, dtype=torch.int32, device="cuda")
n_s = C.c_int64(0)
def sample(b):
    _lib.call("vfr_sel_query_pack", q.data_ptr(), Q, D, b.sel().data_ptr(), nc, qp.data_ptr(), stream)
    _lib.call("vfr_sel_sample", b.sel().data_ptr(), nc, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, samp.data_ptr(), C.byref(n_s), stream)
pool = []
for b in banks:
    sample(b)
    m = samp.view(Q, lists * 32)
    pool.append((torch.topk(m, 32, dim=1, largest=False).values if lists > 1 else m).clone())
j = lib.vfr_sel_sample_rank(k, n_s.value * P, nc * P)
T = torch.kthvalue(torch.stack(pool).permute(1, 0, 2).reshape(Q, -1), j, dim=1).values.contiguous()
ranks = sorted({max(1, -(-j // d)) for d in (1, 2, 4, 8)}, reverse=True)
levels = torch.sort(torch.stack(pool).permute(1, 0, 2).reshape(Q, -1), dim=1).values[:, [r - 1 for r in ranks]].t().contiguous()
counts = torch.empty((len(ranks), Q), dtype=torch.int32, device="cuda")
def scan(b):
    sample(b)
    _lib.call("vfr_sel_bound_put", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, T.data_ptr(), stream)
    _lib.call("vfr_sel_filter", b.sel().data_ptr(), nc, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, 0, tiles, 1, stream)
    for i in range(len(ranks)):
        _lib.call("vfr_sel_count_under", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, levels[i].data_ptr(), counts[i].data_ptr(), stream)
tot4 = torch.zeros((len(ranks), Q), dtype=torch.int64, device="cuda")
for b in banks:
    scan(b); tot4 += counts
ok = tot4 >= k
best = ok.to(torch.int32).sum(dim=0).clamp_(min=1) - 1
T2 = levels.gather(0, best.view(1, Q).to(torch.int64)).view(Q).contiguous()
print("  tightest certified level per query (0 = rank j ... 3 = rank j/8):", torch.bincount(best, minlength=len(ranks)).tolist())
def pooled(b):
    scan(b)
    _lib.call("vfr_sel_bound_put", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, T2.data_ptr(), stream)
    _lib.call("vfr_sel_count_under", qp.data_ptr(), Q, nc, D, k, ws.data_ptr(), 0, T2.data_ptr(), count.data_ptr(), stream)
    refine(b)
t_scan = timeit(lambda: scan(b0))
t_pool = timeit(lambda: pooled(b0))
print(f"  scan part {t_scan:.2f} ms")
tot = torch.zeros(Q, dtype=torch.int64, device="cuda")
for b in banks:
    pooled(b); tot += count
print(f"  pooled-sample protocol (j = {j}, {n_s.value} sampled clips per shard): {t_pool:.2f} ms per shard; "
      f"queries whose global count check fails: {int((tot < k).sum())}; candidates kept by shard 0: {float(count.float().mean()):.1f} per query")

"""K4 (filter + refine) phase by phase on the bench workload (CUDA events): query pack, sample pass + filter, refine."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import _lib, ops

V, Q, k, D = int(os.environ.get("V", "1000000")), int(os.environ.get("Q", "37888")), int(os.environ.get("K", "100")), int(os.environ.get("D", "100"))
g = torch.Generator(device="cuda").manual_seed(0)
clips = ((torch.randn(V, 1, D, device="cuda", generator=g) + 0.6 * torch.randn(V, 6, D, device="cuda", generator=g)) * 0.05).reshape(-1, D)
q = (torch.randn(Q, D, device="cuda", generator=g) * 0.06).contiguous()
bank = ops.Bank(clips, np.arange(V + 1) * 6)
lib = _lib.load()
n_clips = V * 6
qp = torch.empty(lib.vfr_sel_query_bytes(Q, D), dtype=torch.uint8, device="cuda")
ws = torch.empty(lib.vfr_sel_topk_bytes(Q, n_clips, 0), dtype=torch.uint8, device="cuda")
out_s = torch.empty((Q, k), dtype=torch.float32, device="cuda")
out_i = torch.empty((Q, k), dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tiles = lib.vfr_sel_tiles(n_clips)

def phases():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    _lib.call("vfr_sel_query_pack", q.data_ptr(), Q, D, bank.sel().data_ptr(), n_clips, qp.data_ptr(), st)
    ev[1].record()
    _lib.call("vfr_sel_filter", bank.sel().data_ptr(), n_clips, D, qp.data_ptr(), Q, k, ws.data_ptr(), 0, 0, tiles, 0, st)
    ev[2].record()
    _lib.call("vfr_sel_refine", bank.clips.data_ptr(), bank.vid_off.data_ptr(), bank.mom_off.data_ptr(), V, n_clips, 6, D, qp.data_ptr(),
              q.data_ptr(), Q, k, 0, out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), 0, st)
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]

for fast in (os.environ.get("VFR_RF_FAST", "512"), "0"):
  os.environ["VFR_RF_FAST"] = fast
  for _ in range(3):
    phases()
  t = np.mean([phases() for _ in range(8)], axis=0)
  pairs = Q * V * 21
  print(json.dumps(dict(refine_fast_max=fast, V=V, Q=Q, k=k, D=D, ms_query_pack=t[0], ms_sample_and_filter=t[1], ms_refine=t[2], ms_total=float(t.sum()),
                      filter_tflops=Q * n_clips * 2.0 * D / t[1] / 1e9, pairs_per_s=pairs / t.sum() * 1e3)))
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
os.environ["VFR_RF_DBG"] = hex(dbg.data_ptr())
os.environ["VFR_RF_FAST"] = "512"
phases()
d = dbg.cpu().numpy().astype(np.float64) / Q
print("refine, cycles per query CTA: gather+kth %.0f | vids+unique %.0f | distances %.0f | moments %.0f | final select %.0f ; unique videos %.1f, moments kept %.1f" % (d[0], d[1], d[2], d[3], d[4], d[6], d[7]))
os.environ.pop("VFR_RF_DBG")
stats = torch.zeros(8, dtype=torch.int64, device="cuda")
_lib.call("vfr_sel_stats", qp.data_ptr(), Q, n_clips, D, k, ws.data_ptr(), 0, stats.data_ptr(), st)
print("stats", stats.tolist())

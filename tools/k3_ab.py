"""K3 (text embedding, tcgen05 path) timed alone (B = queries, VFR_GEMM2=0 selects the one-CTA GEMM of round 1) and the
tile-phase counters of its step GEMMs (VFR_GEMM_DBG)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import models
import bench

dev = "cuda"
model = bench.make_model(dev)
model.engine = "tc"
B = int(os.environ.get("B", "37888"))
tok = torch.from_numpy(bench.make_tokens(B, 1000)).to(dev)
def run(n=5):
    with torch.no_grad():
        for _ in range(2): model(tok, False, dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): model(tok, False, dev)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
for rep in range(4):
    print(json.dumps(dict(B=B, gemm2=os.environ.get("VFR_GEMM2", "1"), ms=run())), flush=True)

# where the cycles go (VFR_GEMM_DBG: sums over every GEMM launched while it is set)
dbg = torch.zeros(8, dtype=torch.int64, device=dev)
os.environ["VFR_GEMM_PRE"] = "0"
with torch.no_grad():
    model(tok, False, dev)
    torch.cuda.synchronize()
    os.environ["VFR_GEMM_DBG"] = hex(dbg.data_ptr())
    model(tok, False, dev)
    torch.cuda.synchronize()
os.environ.pop("VFR_GEMM_DBG")
d = dbg.cpu().numpy().astype(np.float64)
if os.environ.get("VFR_GEMM2", "1") != "0":
    # CTA-pair kernel: sums over the leader CTAs of all launches (d[7] = pairs x launches, d[6] = tiles)
    n = d[6]
    print("K3 of %d queries, pair kernel: %d tiles over %d pair-launches; cycles per tile: MMA issuer loop %.0f (waits: free accumulator %.0f, operands %.0f) | "
          "epilogue warp: wait for accumulator %.0f, work %.0f | producer waits for free stages %.0f" % (
              B, n, d[7], d[2] / n, d[0] / n, d[1] / n, d[3] / n, d[4] / n, d[5] / n))
else:
    n = d[7]
    print("K3 of %d queries: %d working tiles; cycles per tile: set-up %.0f | wait for first operands %.0f | main loop %.0f | epilogue %.0f | whole tile %.0f" % (
        B, n, d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[4] / n))

"""K2 split-weight form once warmed up, then 3 calls: run under `ncu --metrics gpu__time_duration.sum` for the per-kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import models
dev = "cuda"
torch.manual_seed(123)
model = models.CALModel(visual_input_dim=8194, pretrained_emb=torch.randn(50, 100) * 0.4).to(dev).eval()
model.visual_engine = "tc"
V, n = int(os.environ.get("V", "4096")), 6
g = torch.Generator(device=dev).manual_seed(0)
seg = torch.rand((V * n, 4096), device=dev, generator=g)
ctx = torch.rand((V, 4096), device=dev, generator=g)
vid_off = np.arange(V + 1) * n
with torch.no_grad():
    for _ in range(4):
        model.embed_clips(seg, ctx, vid_off)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        model.embed_clips(seg, ctx, vid_off)
    b.record(); torch.cuda.synchronize()
print("V", V, "ms", a.elapsed_time(b) / 5)
if os.environ.get("K2_DBG"):
    # tile phases of the CTA-pair GEMMs of one call (sums over the three GEMMs; the clip GEMM dominates)
    dbg = torch.zeros(8, dtype=torch.int64, device=dev)
    os.environ["VFR_GEMM_DBG"] = hex(dbg.data_ptr())
    with torch.no_grad():
        model.embed_clips(seg, ctx, vid_off)
    torch.cuda.synchronize()
    os.environ.pop("VFR_GEMM_DBG")
    d = dbg.cpu().numpy().astype(np.float64)
    n = d[6]
    print("work items %d over %d pair-launches; cycles per item: MMA issuer loop %.0f (waits: free accumulator %.0f, operands %.0f) | "
          "epilogue warp: wait for accumulator %.0f, work %.0f | producer waits for free stages %.0f" % (
              n, d[7], d[2] / n, d[0] / n, d[1] / n, d[3] / n, d[4] / n, d[5] / n))

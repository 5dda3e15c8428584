"""Single launch of the filter + refine top-k for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops
V = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
g = torch.Generator(device="cuda").manual_seed(0)
clips = ((torch.randn(V, 1, 100, device="cuda", generator=g) + 0.6 * torch.randn(V, 6, 100, device="cuda", generator=g)) * 0.05).reshape(-1, 100)
q = torch.randn(Q, 100, device="cuda", generator=g) * 0.06
bank = ops.Bank(clips, np.arange(V + 1) * 6)
for _ in range(2):
    s, i = ops.score_topk_sel(bank, q, int(os.environ.get("K", "100")))
torch.cuda.synchronize()
print("ok", float(s[:, 0].sum()))

"""Single launch of the tensor-core top-k for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops
V = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
g = torch.Generator(device="cuda").manual_seed(0)
clips = torch.randn(V * 6, 100, device="cuda", generator=g) * 0.25
q = torch.randn(Q, 100, device="cuda", generator=g) * 0.25
bank = ops.Bank(clips, np.arange(V + 1) * 6)
for _ in range(2):
    s, i = ops.score_topk_tc(bank, q, 100)
torch.cuda.synchronize()
print("ok", float(s[:, 0].sum()))

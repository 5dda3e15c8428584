"""Device-side timing of the tensor-core top-k at several batch sizes (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
S, D, k = 6, 100, 100
g = torch.Generator(device="cuda").manual_seed(0)
clips = torch.randn(V * S, D, device="cuda", generator=g) * 0.25
bank = ops.Bank(clips, np.arange(V + 1) * S)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for Q in [int(x) for x in (sys.argv[2:] or ["4096", "18944"])]:
    q = torch.randn(Q, D, device="cuda", generator=g) * 0.25
    for nt in (3, 1):
        t = timeit(lambda: ops.score_topk_tc(bank, q, k, n_terms=nt))
        print(f"tc topk n_terms={nt} V={V} Q={Q}: {t:.2f} ms  {Q*V*21/t/1e6:.1f} Gpairs/s  clip-pairs {Q*V*S/t/1e6:.1f} G/s  per SM-clk@1.9GHz {Q*V*S/t/1e6/148/1.9:.2f}")

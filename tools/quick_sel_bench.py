"""Device-side timing of the filter + refine top-k at several batch sizes (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops, _lib
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
S, D, k = 6, int(os.environ.get("D", "100")), int(os.environ.get("K", "100"))
g = torch.Generator(device="cuda").manual_seed(0)
clips = ((torch.randn(V, 1, D, device="cuda", generator=g) + 0.6 * torch.randn(V, S, D, device="cuda", generator=g)) * 0.05).reshape(-1, D)
bank = ops.Bank(clips, np.arange(V + 1) * S)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for Q in [int(x) for x in (sys.argv[2:] or ["4096", "18944"])]:
    q = torch.randn(Q, D, device="cuda", generator=g) * 0.06
    t = timeit(lambda: ops.score_topk_sel(bank, q, k))
    s, i, flags, (qp, ws) = ops.score_topk_sel(bank, q, k, return_flags=True)
    print(f"sel topk k={k} V={V} Q={Q} R={os.environ.get('VFR_SEL_R','auto')}: {t:.2f} ms  {Q*V*21/t/1e6:.1f} Gpairs/s  clip-pairs {Q*V*S/t/1e6:.1f} G/s  "
          f"per SM-clk@1.9GHz {Q*V*S/t/1e6/148/1.9:.2f}  flags={int(flags.abs().sum())} ws={ws.numel()/1e6:.0f}MB", flush=True)

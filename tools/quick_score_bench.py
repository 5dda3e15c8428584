"""Quick device-side timing of the K4 modes (development aid, not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops, synth

V = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S, D, k = 6, 100, 100
g = torch.Generator(device="cuda").manual_seed(0)
clips = torch.randn(V * S, D, device="cuda", generator=g) * 0.25
q = torch.randn(Q, D, device="cuda", generator=g) * 0.25
bank = ops.Bank(clips, np.arange(V + 1) * S)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pairs = Q * V * 21
t = timeit(lambda: ops.score_topk(bank, q, k))
print(f"topk  V={V} Q={Q}: {t:.2f} ms  {pairs / t / 1e6:.1f} Gpairs/s  clip-pairs/s {Q*V*S/t/1e6:.1f} G")
tau = torch.full((Q, 2), 1.0, device="cuda")
qv = torch.zeros(Q, dtype=torch.int32, device="cuda")
t = timeit(lambda: ops.score_count(bank, q, tau, qv))
print(f"count V={V} Q={Q}: {t:.2f} ms  {pairs / t / 1e6:.1f} Gpairs/s")
t = timeit(lambda: ops.score_topk_tc(bank, q, k))
print(f"tc topk V={V} Q={Q}: {t:.2f} ms  {pairs / t / 1e6:.1f} Gpairs/s  clip-pairs/s {Q*V*S/t/1e6:.1f} G")
t = timeit(lambda: ops.score_topk_tc(bank, q, k, n_terms=1))
print(f"tc topk bf16 V={V} Q={Q}: {t:.2f} ms  {pairs / t / 1e6:.1f} Gpairs/s")

import sys, os
sys.path.insert(0, "/root/repo")
import torch, vfr_b200
from vfr_b200 import ops
Q, k = 37888, 100
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for P in (2, 8):
    s = torch.rand(P, Q, k, device="cuda").sort(dim=2).values
    i = torch.randint(0, 21000000, (P, Q, k), device="cuda")
    print(P, "topk_merge ms", timeit(lambda: ops.topk_merge(s, i)))
    x = torch.rand(P, Q, 32, device="cuda")
    print(P, "sort pooled ms", timeit(lambda: torch.sort(x.permute(1, 0, 2).reshape(Q, -1), dim=1)))
    print(P, "contiguous copies ms", timeit(lambda: (s[:, :Q].contiguous(), i[:, :Q].contiguous())))

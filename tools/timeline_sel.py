"""Timeline of CTA 0 of the filter kernel (development aid): VFR_SEL_DBG points the kernel at a debug buffer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import ops
V, Q, k = int(os.environ.get("V", "1000000")), int(os.environ.get("Q", "18944")), int(os.environ.get("K", "1"))
D = int(os.environ.get("D", "100"))
g = torch.Generator(device="cuda").manual_seed(0)
clips = ((torch.randn(V, 1, D, device="cuda", generator=g) + 0.6 * torch.randn(V, 6, D, device="cuda", generator=g)) * 0.05).reshape(-1, D)
q = torch.randn(Q, D, device="cuda", generator=g) * 0.06
bank = ops.Bank(clips, np.arange(V + 1) * 6)
ops.score_topk_sel(bank, q, k)
dbg = torch.zeros(9 * 256 * 4 + 8, dtype=torch.int64, device="cuda")
os.environ["VFR_SEL_DBG"] = hex(dbg.data_ptr())
ops.score_topk_sel(bank, q, k)
torch.cuda.synchronize()
cnt = dbg.cpu().numpy()[9 * 256 * 4:]
d = dbg.cpu().numpy()[:9 * 256 * 4].reshape(9, 256, 4)
t0 = d[0, 0, 0]
print("MMA thread (job = tile*2 + r): wait_start wait_end issued(commit) | wait issue")
for j in range(40, 48):
    print("  job %3d buf %d  %8d %8d %8d | wait %5d (probes + pending commit %5d, bank tile there after %5d) issue %5d" % (j, d[0, j, 3] & 255, d[0, j, 0] - t0, d[0, j, 1] - t0, d[0, j, 2] - t0, d[0, j, 1] - d[0, j, 0], (d[0, j, 3] >> 8) & 0xffffff, d[0, j, 3] >> 32, d[0, j, 2] - d[0, j, 1]))
for s in (0, 1):
    print("epilogue set %d: per warp (ew) full_seen / released / done, visits 20..22" % s)
    for v in range(20, 23):
        print("  visit %d: " % v + "  ".join("ew%d %d/%d/%d" % (w, d[1 + w, v, 1] - t0, d[1 + w, v, 2] - t0, d[1 + w, v, 3] - t0) for w in range(4 * s, 4 * s + 4)))
print("steady state: %.0f cycles per job" % ((d[0, 200, 0] - d[0, 40, 0]) / 160.0))
hold = (d[1:9, :128, 2] - d[1:9, :128, 1])
print("hold (full_seen -> release) per epilogue warp over the last 128 visits: median / p90 / max")
for w in range(8):
    h = np.sort(hold[w])
    print("  ew%d: %5d %5d %5d   stragglers(>1200): %3d  at visits mod 4: %s" % (w, h[64], h[115], h[-1], int((hold[w] > 1200).sum()),
          np.bincount(np.nonzero(hold[w] > 1200)[0] % 4, minlength=4).tolist()))
tail = (d[1:9, :128, 3] - d[1:9, :128, 2])
print("tail (release -> done): median %d p90 %d max %d" % (np.median(tail), np.percentile(tail, 90), tail.max()))
wait = (d[1:9, :128, 1] - d[1:9, :128, 0])
print("wait for tmem_full: median %d p90 %d" % (np.median(wait), np.percentile(wait, 90)))
n_cta = 148
print("compaction, all CTAs: %d warp events, %d lists, %.0f clk per list, %.0f clk per event; per CTA %.2f ms if serialised (1.9 GHz)" % (
    cnt[1], cnt[2], cnt[0] / max(cnt[2], 1), cnt[0] / max(cnt[1], 1), cnt[0] / n_cta / 1.9e6))
print("  per list: load %.0f clk, select %.0f clk, write-back %.0f clk" % tuple(cnt[3 + i] / max(cnt[2], 1) for i in range(3)))
if os.environ.get("VFR_SEL_PDBG"):
    # producer of CTA 0: when it asked for the two chunks of each of the last 128 bank tiles (role 8 slots), against the
    # issuer's first job of that tile (jobs 2 * tile of the last 256 jobs)
    print("bank tile: requested chunk0 / chunk1 (producer) | issuer starts looking / has it  -> request-to-arrival of chunk1")
    for tl in range(20, 28):
        j = 2 * tl
        looking = d[0, j, 0] + ((d[0, j, 3] >> 8) & 0xffffff)
        has = d[0, j, 0] + (d[0, j, 3] >> 32)
        print("  tile %3d: %8d %8d | %8d %8d -> %6d" % (tl, d[8, tl, 0] - t0, d[8, tl, 1] - t0, looking - t0, has - t0, has - d[8, tl, 1]))

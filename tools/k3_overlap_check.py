"""K3 with overlapped step GEMMs (VFR_K3_OVERLAP=1, programmatic dependent launch + row-block dependency counters) against
the serial launches: bitwise equality over random batches of several sizes, then timing of both."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
import bench

dev = "cuda"
model = bench.make_model(dev)
model.engine = "tc"
rng = np.random.default_rng(0)
bad = 0
for it, B in enumerate([1, 7, 255, 256, 257, 700, 4736, 4737, 9000, 18944, 37888] * int(os.environ.get("REPS", "3"))):
    if os.environ.get("SKIP_EQ"): break
    tok = np.zeros((B, 20), dtype=np.int64)
    lens = np.clip(rng.poisson(6.5, size=B) + 1, 0, 20) if it % 3 else rng.integers(0, 21, size=B)
    for r in range(B):
        tok[r, :lens[r]] = rng.integers(1, 1000, size=lens[r])
    t = torch.from_numpy(tok).to(dev)
    with torch.no_grad():
        os.environ["VFR_K3_OVERLAP"] = "0"
        ref = model(t, False, dev).clone()
        os.environ["VFR_K3_OVERLAP"] = "1"
        for rep in range(3):
            got = model(t, False, dev)
            if not torch.equal(got.view(torch.int32), ref.view(torch.int32)):
                bad += 1
                print("MISMATCH B=%d rep=%d max abs diff %.3e rows %d" % (B, rep, (got - ref).abs().max().item(), int((got != ref).any(1).sum())), flush=True)
print("mismatching runs:", bad, flush=True)

def run(tok, n=5):
    with torch.no_grad():
        for _ in range(2): model(tok, False, dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): model(tok, False, dev)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
for B in [int(x) for x in os.environ.get("SIZES", "4736,37888").split(",")]:
    tok = torch.from_numpy(bench.make_tokens(B, 1000)).to(dev)
    for rep in range(3):
        for mode in ("0", "1"):
            os.environ["VFR_K3_OVERLAP"] = mode
            print(json.dumps(dict(B=B, overlap=mode, ms=run(tok))), flush=True)
sys.exit(1 if bad else 0)

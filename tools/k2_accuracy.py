"""K2 (visual MLP) on tensor cores: error against float64 and time per clip as a function of the K-segment length of the
accumulation (VFR_VIS_FLUSH; 0 = one uninterrupted accumulation), for the general and the split-weight form."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vfr_b200
from vfr_b200 import models

dev = "cuda"
torch.manual_seed(123)
model = models.CALModel(visual_input_dim=8194, pretrained_emb=torch.randn(50, 100) * 0.4).to(dev).eval()
rng = np.random.default_rng(0)
V, n = 4096, 6
seg = rng.random((V * n, 4096), dtype=np.float32)
ctx = rng.random((V, 4096), dtype=np.float32)
seg /= np.linalg.norm(seg, axis=1, keepdims=True) + 1e-5
ctx /= np.linalg.norm(ctx, axis=1, keepdims=True) + 1e-5
tef = np.tile(np.stack([np.arange(n) / np.float32(n), (np.arange(n) + 1) / np.float32(n)], 1).astype(np.float32), (V, 1))
x = torch.from_numpy(np.concatenate([seg, np.repeat(ctx, n, axis=0), tef], axis=1)).to(dev)
seg_t, ctx_t, vid_off = torch.from_numpy(seg).to(dev), torch.from_numpy(ctx).to(dev), np.arange(V + 1) * n
lin1, lin2 = model.visual_fc[0], model.visual_fc[2]
with torch.no_grad():
    want = torch.relu(x[:6000].double() @ lin1.weight.double().t() + lin1.bias.double()) @ lin2.weight.double().t() + lin2.bias.double()
scale = want.abs().max().item()

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

with torch.no_grad():
    model.visual_engine = "exact"
    e = model(x[:6000])
    print(json.dumps(dict(engine="exact", err=(e.double() - want).abs().max().item() / scale, ms=timeit(lambda: model(x)), rows=x.shape[0])))
    model.visual_engine = "tc_bf16x3"
    e = model(x[:6000])
    print(json.dumps(dict(engine="tc_bf16x3", err=(e.double() - want).abs().max().item() / scale, ms=timeit(lambda: model(x)))))
    model.visual_engine = "tc"
    for flush in (0, 2048, 1024, 512, 256, 128, 64):
        os.environ["VFR_VIS_FLUSH"] = str(flush)
        g = model(x[:6000])
        s = model.embed_clips(seg_t[:6000], ctx_t[:1000], vid_off[:1001])
        print(json.dumps(dict(engine="tc split-fp16", flush=flush, err_general=(g.double() - want).abs().max().item() / scale,
                              err_split=(s.double() - want).abs().max().item() / scale, ms_general=timeit(lambda: model(x)),
                              ms_split=timeit(lambda: model.embed_clips(seg_t, ctx_t, vid_off)), rows=x.shape[0])))

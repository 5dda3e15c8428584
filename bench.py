"""Contract benchmark: query-moment pairs scored per second on the corpus-retrieval workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4]): a bank of 1,000,000 six-clip videos
(6 M clip embeddings, D = 100, 21 M candidate moments) resident in HBM; the 100,000 queries arrive in
batches of --batch queries (default 37,888 = 2 x 148 query tiles: one CTA per SM walks the whole bank shard for
two query tiles, so a query keeps a single candidate list).  One STEP = one batch of tokenised queries through the
retrieval hot path: K3 query embedding (GloVe gather -> BiLSTM -> Linear) -> K4 fused distance / moment-mean /
top-100 over the whole bank (-> all-to-all by query slice + K7 merge when the bank is sharded over N GPUs).
Strong scaling: the bank is fixed and split by contiguous video ranges over the ranks.

value      = batch * 21 M pairs / step time, inputs (token ids) already resident in HBM.
e2e        = the same through the public host-buffer call (MomentRetriever.search): pinned host token ids in, host
             top-k lists out, copies inside the timed region (on N ranks every rank moves its own query slice).
roofline   = the dominant kernel (K4) timed with CUDA events INSIDE the timed steps (an event between K3 and K4 of every
             step); roofline_k3 = the same for K3.
parity_check = OUTSIDE the timed regions: a subsample of the last step's results re-scored by the CPU oracle.
filter_stats = candidates kept per query / compactions / flagged queries / exact-engine reruns of the last step.
cpu_baseline / --impl reference = the UNMODIFIED reference (oracle/_ref, a verbatim copy of /root/reference/model made by
             oracle/build_ref.py): its own evaluate.evaluate + evaluate_single.evaluate on the host cores, on a bounded
             sample of the same workload.
library_baseline = stock PyTorch on the same B200: the reference's loop with device='cuda' (bounded sample) and a
             batched torch formulation (cuDNN LSTM, cdist, cumsum, topk) of the same step.
"""
import argparse
import io
import json
import os
import sys
import threading
import time
from contextlib import redirect_stderr, redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VIDEOS = 1_000_000
N_SEG = 6
VOCAB = 10_000
TOPK = 100
MOMENTS_PER_VIDEO = N_SEG * (N_SEG + 1) // 2
BANK_BLOCKS = 64          # the synthetic bank is generated in fixed blocks so any sharding sees the same data
SEED = 123


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=37888, help="queries per step (296 query tiles of 128: one CTA of two tiles per SM, no bank split)")
    ap.add_argument("--engine", type=str, default="sel", choices=["sel", "tc", "tc_bf16", "exact"])
    ap.add_argument("--videos", type=int, default=N_VIDEOS)
    ap.add_argument("--dim", type=int, default=100, help="joint embedding dimension (BASELINE configs[2] names 1024 as a variant)")
    ap.add_argument("--bank", type=str, default="gaussian", choices=["gaussian", "clustered"],
                    help="clustered: 1,000 centroids, sigma 1e-3 - the near-duplicate worst case of the filter + refine engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# synthetic data (seeded; identical for every N)
# ------------------------------------------------------------------------------------------------
def make_state(dim):
    """CALModel weights of the reference's shapes (visual branch unused by the retrieval step: a token 18-d input keeps
    the 500 x 8194 matrix out of memory-irrelevant setup time)."""
    torch.manual_seed(SEED)
    table = torch.randn(VOCAB, 100) * 0.4
    table[0] = 0
    return table


def make_model(device, dim=100):
    import vfr_b200  # noqa: F401
    from vfr_b200 import models
    table = make_state(dim)
    torch.manual_seed(SEED)
    model = models.CALModel(visual_input_dim=2 * 4096 + 2, pretrained_emb=table, emb_dim=dim)
    return model.to(device).eval()


def make_tokens(n, seed):
    rng = np.random.default_rng(seed)
    tok = np.zeros((n, 20), dtype=np.int64)
    lens = np.clip(rng.poisson(6.5, size=n) + 1, 1, 20)
    for i in range(n):
        tok[i, :lens[i]] = rng.integers(1, VOCAB, size=lens[i])
    return tok


def bank_block(block, n_videos, device, dim=100, kind="gaussian"):
    """Clip embeddings of one fixed block of videos (shared per-video component + per-clip part)."""
    v0, v1 = (n_videos * block) // BANK_BLOCKS, (n_videos * (block + 1)) // BANK_BLOCKS
    g = torch.Generator(device=device).manual_seed(SEED * 1000 + block)
    if kind == "clustered":
        # 1,000 centroids shared by the whole bank, every clip = a centroid + N(0, 1e-3): masses of near-duplicates
        gc = torch.Generator(device=device).manual_seed(SEED * 7)
        cent = torch.randn((1000, dim), device=device, generator=gc) * 0.05
        pick = torch.randint(0, 1000, ((v1 - v0) * N_SEG,), device=device, generator=g)
        return cent[pick] + 1e-3 * torch.randn(((v1 - v0) * N_SEG, dim), device=device, generator=g), v0, v1
    base = torch.randn((v1 - v0, 1, dim), device=device, generator=g)
    clip = torch.randn((v1 - v0, N_SEG, dim), device=device, generator=g)
    return ((base + 0.6 * clip) * 0.05).reshape(-1, dim), v0, v1


def make_shard(n_videos, rank, world, device, dim=100, kind="gaussian"):
    from vfr_b200.retrieval import shard_range
    v0, v1 = shard_range(n_videos, rank, world)
    parts = []
    for b in range(BANK_BLOCKS):
        b0, b1 = (n_videos * b) // BANK_BLOCKS, (n_videos * (b + 1)) // BANK_BLOCKS
        lo, hi = max(b0, v0), min(b1, v1)
        if lo >= hi:
            continue
        emb, _, _ = bank_block(b, n_videos, device, dim, kind)
        parts.append(emb[(lo - b0) * N_SEG:(hi - b0) * N_SEG])
    clips = torch.cat(parts, dim=0)
    vid_off = np.arange(v1 - v0 + 1, dtype=np.int64) * N_SEG
    return clips, vid_off, v0 * MOMENTS_PER_VIDEO


def bank_rows(video_ids, n_videos, device, dim, kind):
    """Clip embeddings [len(video_ids), 6, dim] of arbitrary videos of the bench bank (regenerated block by block)."""
    video_ids = np.asarray(video_ids, dtype=np.int64)
    out = torch.empty((len(video_ids), N_SEG, dim), dtype=torch.float32)
    blocks = np.searchsorted([(n_videos * (b + 1)) // BANK_BLOCKS for b in range(BANK_BLOCKS)], video_ids, side="right")
    for b in np.unique(blocks):
        emb, v0, _ = bank_block(int(b), n_videos, device, dim, kind)
        sel = np.nonzero(blocks == b)[0]
        out[sel] = emb.view(-1, N_SEG, dim)[torch.as_tensor(video_ids[sel] - v0, device=device)].cpu()
    return out


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference (oracle/_ref) on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
class ReferenceArm:
    """One step = the reference's own ``evaluate.evaluate`` (model/evaluate.py:28-90: batch-1 text embedding, one
    ``index_select().mean().item()`` per moment, ``np.argsort`` over all distances) on ``n_q`` queries x ``n_v`` videos of
    block 0 of the bench bank, followed by its ``evaluate_single.evaluate`` on the same queries.  Only feature assembly
    is adapted (the bank is generated at embedding level: the dataset hands the clip embeddings through and the model's
    ``visual_fc`` is ``nn.Identity()``); ``CALModel``'s text branch, the samplers, collates and both ``evaluate``
    functions are the reference's files, sha256-checked against oracle/_ref/MANIFEST.json."""

    def __init__(self, n_videos, n_q=48, n_v=384, device="cpu", dim=100, kind="gaussian"):
        from oracle import ref_harness
        self.h = ref_harness
        self.ref = ref_harness.load()
        self.kind = "reference" if os.path.abspath(self.ref.root) == os.path.abspath(ref_harness.REF_COPY) or \
            self.ref.root == ref_harness.REF_SOURCE else "port"
        self.n_q, self.n_v, self.device, self.dim = n_q, min(n_v, n_videos // BANK_BLOCKS), device, dim
        ours = make_model("cpu", dim)
        sd = {k: v.detach() for k, v in ours.state_dict().items()}
        self.model = ref_harness.ref_model(self.ref, sd, 4096)
        self.model.visual_fc = torch.nn.Identity()
        self.model = self.model.to(device).eval()
        emb, _, _ = bank_block(0, n_videos, "cpu", dim, kind)
        self.clips = emb[:self.n_v * N_SEG].contiguous()
        self.step_idx = 0
        self.pairs_per_step = self.n_q * self.n_v * MOMENTS_PER_VIDEO
        from vfr_b200 import synth
        self.prior = synth.make_prior((N_SEG,))

    def step(self):
        from torch.utils.data import DataLoader
        rng = np.random.default_rng(20_000 + self.step_idx)
        tok = make_tokens(self.n_q, 10_000 + self.step_idx)
        self.step_idx += 1
        q_video = rng.integers(0, self.n_v, size=self.n_q)
        times = []
        for _ in range(self.n_q):
            s = int(rng.integers(0, N_SEG))
            times.append([[s, s], [s, s], [s, min(s + 1, N_SEG - 1)], [max(s - 1, 0), s]])
        ds, annotations, names = self.h.bank_dataset(self.ref, self.clips, N_SEG, tok, q_video, times)
        vit = DataLoader(ds, shuffle=False, collate_fn=self.ref.data.validate_collate,
                         batch_sampler=self.ref.data.VideoBatchSampler(names, ds.num_segments_info))
        lit = DataLoader(ds, shuffle=False, collate_fn=self.ref.data.validate_collate,
                         batch_sampler=self.ref.data.LanguageBatchSampler(annotations, ds.num_segments_info))
        sink = io.StringIO()
        with redirect_stdout(sink), redirect_stderr(sink):
            corpus = self.ref.evaluate.evaluate(self.model, vit, lit, annotations, self.device, preliminary=10 ** 9)
            single = self.ref.evaluate_single.evaluate(self.model, vit, lit, annotations, self.device, ["model"], self.prior)
        return corpus, single

    def describe(self):
        return (f"{self.n_q} queries x {self.n_v} videos x {MOMENTS_PER_VIDEO} moments per step (block 0 of the bench bank) "
                f"through the unmodified reference evaluate.evaluate (model/evaluate.py:28-90) + evaluate_single.evaluate "
                f"(model/evaluate_single.py:28-87) on {self.device}")


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm uses the cores the box really has."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    n = max(1, min(n, 64))
    torch.set_num_threads(n)
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    arm = ReferenceArm(args.videos, dim=args.dim, kind=args.bank)
    for _ in range(min(args.warmup, 1)):          # (a warm-up step is 5-10 s of CPU time: one is enough)
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = time.perf_counter() - t0
    value = arm.pairs_per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "query-moment pairs scored/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world=args.gpus),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": arm.kind, "sample": arm.describe(),
                         "host_cpus": os.cpu_count(), "reference_root": os.path.relpath(arm.ref.root, ROOT)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(args, world):
    return {"workload": "corpus retrieval: 100k queries x %d videos x 21 moments (6 clips, D=%d), top-%d, "
                        "query batches of %d (BASELINE configs[4])" % (args.videos, args.dim, TOPK, args.batch),
            "query_batch": args.batch, "n_videos": args.videos, "n_moments": args.videos * MOMENTS_PER_VIDEO,
            "dim": args.dim, "topk": TOPK, "bank": getattr(args, "bank", "gaussian"),
            "parallelism": f"bank sharded by video range over {world} GPU(s); queries embedded by slice, all-gathered; "
                           f"top-k lists exchanged by query slice (all-to-all)", "engine": getattr(args, "engine", "sel"),
            "l2": "inputs larger than L2: the packed bank shard (%.2f GB) is streamed every step"
                  % (args.videos / world * N_SEG * 128 * 2 / 1e9)}


# ------------------------------------------------------------------------------------------------
# library bar: stock PyTorch on the same GPU
# ------------------------------------------------------------------------------------------------
def library_baseline(args, device, model, clips, tokens_dev, budget_s=20.0):
    """(a) the reference's loop itself with device='cuda' on a bounded sample; (b) the same STEP written with stock
    batched torch ops: cuDNN BiLSTM + Linear for K3, torch.cdist (cuBLAS SGEMM, TF32 off) + cumsum + gather for the
    moment means, torch.topk - chunks of 64 queries against this rank's whole bank."""
    out = {}
    try:
        arm = ReferenceArm(args.videos, n_q=8, n_v=256, device=str(device), dim=args.dim, kind=args.bank)
        arm.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 0.4 * budget_s:
            arm.step()
            n += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["reference_cuda"] = {"value": arm.pairs_per_step * n / dt, "unit": "pairs/s", "sample": arm.describe(),
                                 "kind": arm.kind}
    except Exception as e:  # noqa: BLE001 - the bar is informative; the bench line must still be printed
        out["reference_cuda"] = {"unavailable": repr(e)[:200]}
    # (b) batched stock torch
    import torch.nn as nn
    H, E, D = model.hidden_size, model.word_embedding.weight.shape[1], model.lang_fc.weight.shape[0]
    lstm = nn.LSTM(E, H, num_layers=1, batch_first=True, bidirectional=True).to(device)
    lstm.load_state_dict(model.lstm.state_dict())
    table, fc_w, fc_b = model.word_embedding.weight, model.lang_fc.weight, model.lang_fc.bias
    V = clips.shape[0] // N_SEG
    moments = [(j, j) for j in range(N_SEG)] + [(s, e) for s in range(N_SEG) for e in range(s + 1, N_SEG)]
    s_idx = torch.tensor([m[0] for m in moments], device=device)
    e_idx = torch.tensor([m[1] + 1 for m in moments], device=device)
    inv_len = (1.0 / (e_idx - s_idx).float()).view(1, 1, -1)
    n_lib, chunk = 1024, 64
    tok = tokens_dev[:n_lib]

    def step():
        with torch.no_grad():
            _, (h, _) = lstm(table[tok])
            q = torch.nn.functional.linear(h.transpose(0, 1).reshape(n_lib, 2 * H), fc_w, fc_b)
            best = []
            for c0 in range(0, n_lib, chunk):
                d = torch.cdist(q[c0:c0 + chunk] - 1e-6, clips).view(-1, V, N_SEG)          # |v - q + eps|
                cs = torch.nn.functional.pad(torch.cumsum(d, dim=2), (1, 0))
                sc = (cs[:, :, e_idx] - cs[:, :, s_idx]) * inv_len
                best.append(torch.topk(sc.view(sc.shape[0], -1), TOPK, dim=1, largest=False))
            return best
    try:
        step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        t0 = time.perf_counter()
        a.record()
        while time.perf_counter() - t0 < 0.4 * budget_s and n < 8:
            step()
            torch.cuda.synchronize()
            n += 1
        b.record()
        torch.cuda.synchronize()
        out["torch_batched"] = {"value": n_lib * V * MOMENTS_PER_VIDEO * n / (a.elapsed_time(b) * 1e-3), "unit": "pairs/s",
                                "sample": f"{n} steps of {n_lib} queries x {V} videos: nn.LSTM (cuDNN) + F.linear + torch.cdist "
                                          f"(SGEMM) + cumsum + gather + torch.topk, {chunk}-query chunks"}
    except Exception as e:  # noqa: BLE001
        out["torch_batched"] = {"unavailable": repr(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
# parity of the benchmark configuration itself (outside the timed regions)
# ------------------------------------------------------------------------------------------------
def parity_check(args, retr, model, tokens_host_last, res_s, res_i, device, n_check=64, block_videos=16384):
    """Re-score a subsample of the LAST e2e step with the CPU oracle (oracle/cal_oracle.py, pinned to the reference by
    tests/golden): (1) the query embeddings, (2) every returned (query, moment id) pair - the videos are regenerated from
    the bank's seeds, wherever they live, (3) completeness: all moments of one whole 16 k-video block are scored by the
    oracle and every one that beats the returned k-th score must be in the returned list.  Raises on a mismatch."""
    from oracle import cal_oracle as orc
    q0, q1 = retr.owned_range(tokens_host_last.shape[0])
    n = min(n_check, q1 - q0)
    pick = np.linspace(0, q1 - q0 - 1, n).astype(np.int64)
    tok = tokens_host_last[q0:q1][pick].numpy()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        q_oracle = orc.text_embed(sd, tok)
    q_oracle = np.asarray(q_oracle, dtype=np.float32)
    q_ours = retr.q_emb[q0:q1][torch.as_tensor(pick, device=device)].cpu().numpy()
    emb_err = float(np.abs(q_ours - q_oracle).max() / np.abs(q_oracle).max())
    s = res_s[pick].numpy()
    ids = res_i[pick].numpy()
    assert (ids >= 0).all(), "fewer than k results"
    vids, moms = ids // MOMENTS_PER_VIDEO, ids % MOMENTS_PER_VIDEO
    uniq, inv = np.unique(vids, return_inverse=True)
    rows = bank_rows(uniq, args.videos, device, args.dim, args.bank).numpy()             # [U, 6, D]
    moments = orc.generate_moments(N_SEG)
    worst_score = 0.0
    inv = inv.reshape(vids.shape)
    for qi in range(n):
        v = rows[inv[qi]]                                                                  # [k, 6, D]
        d = np.sqrt((((v - q_ours[qi][None, None, :]) + np.float32(1e-6)) ** 2).sum(-1, dtype=np.float32))
        want = np.array([d[j, moments[m][0]:moments[m][1] + 1].mean(dtype=np.float32) for j, m in enumerate(moms[qi])],
                        dtype=np.float32)
        worst_score = max(worst_score, float(np.abs(want - s[qi]).max() / np.abs(want).max()))
        assert (np.diff(s[qi]) >= 0).all(), "returned list not sorted"
    assert emb_err < 2e-5, f"query embeddings differ from the oracle: {emb_err:.2e}"
    assert worst_score < 1e-5, f"returned scores differ from the oracle: {worst_score:.2e}"
    # completeness against one whole block (video range of bank block 0 lies in rank 0's shard)
    nb = min(block_videos, args.videos // BANK_BLOCKS)
    emb, v0, _ = bank_block(0, args.videos, device, args.dim, args.bank)
    full = orc.score_matrix(emb[:nb * N_SEG].cpu().numpy(), np.arange(nb + 1) * N_SEG, q_ours).numpy()   # [n, nb * 21]
    kth = s[:, -1:]
    tol = 4e-6 * np.abs(kth)
    must = full < (kth - tol)                                  # certainly better than the returned k-th score
    block_ids = (v0 * MOMENTS_PER_VIDEO + np.arange(full.shape[1]))[None, :].repeat(n, 0)
    missing, n_must, id_mismatch = 0, int(must.sum()), 0
    for qi in range(n):
        have = set(ids[qi].tolist())
        for mid, sc in zip(block_ids[qi][must[qi]], full[qi][must[qi]]):
            if int(mid) not in have:
                missing += 1
        # ids of this block inside the returned list: their oracle score must match the returned score (id <-> score)
        inb = np.nonzero((ids[qi] >= block_ids[qi][0]) & (ids[qi] <= block_ids[qi][-1]))[0]
        for j in inb:
            if abs(full[qi][ids[qi][j] - block_ids[qi][0]] - s[qi][j]) > 1e-5 * abs(s[qi][j]):
                id_mismatch += 1
    assert missing == 0, f"{missing} moments of block 0 beat the returned k-th score but are not in the lists"
    assert id_mismatch == 0, f"{id_mismatch} returned ids of block 0 carry a score that is not theirs"
    return {"queries": int(n), "pairs_rescored": int(ids.size), "max_rel_err_scores": worst_score,
            "max_rel_err_query_emb": emb_err, "block_videos": int(nb), "block_moments_scored": int(full.size),
            "block_moments_that_must_be_returned": n_must, "missing": missing, "id_score_mismatch": id_mismatch,
            "tolerance": 1e-5, "checker": "oracle.cal_oracle (score_matrix / text_embed), outside the timed regions",
            "ok": True}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import vfr_b200  # noqa: F401
    from vfr_b200 import _lib
    from vfr_b200.retrieval import MomentRetriever

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    model = make_model(device, args.dim)
    clips, vid_off, id_base = make_shard(args.videos, rank, world, device, args.dim, args.bank)
    retr = MomentRetriever(model, clips, vid_off, id_base=id_base, max_queries=args.batch, k=TOPK, engine=args.engine)
    n_batches = args.warmup + args.steps
    tokens_host = [torch.from_numpy(make_tokens(args.batch, 1000 + i)).pin_memory() for i in range(n_batches)]
    tokens_dev = [t.to(device) for t in tokens_host]
    pairs_per_step = args.batch * args.videos * MOMENTS_PER_VIDEO

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident steps (value) ----
    for i in range(args.warmup):
        retr.search_device(tokens_dev[i])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.vfr_launch_count()
    # (per step three event records on the stream: step start, K3 done, step end - the two dominant stages are timed
    #  INSIDE the timed steps, in the clock / cache state the step really has)
    names = ("start", "k3_done", "scan_start", "scan_end", "end")
    marks = [{n: torch.cuda.Event(enable_timing=True) for n in names} for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        marks[i]["start"].record()
        retr.search_device(tokens_dev[args.warmup + i], marks=marks[i])
        marks[i]["end"].record()
    ev1.record()
    barrier()
    launches = lib.vfr_launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    fixups_value = retr.n_fixups
    k3 = max_over_ranks(float(np.mean([m["start"].elapsed_time(m["k3_done"]) for m in marks])))
    k4 = max_over_ranks(float(np.mean([m["k3_done"].elapsed_time(m["end"]) for m in marks])))
    scan = None
    if args.engine == "sel":
        try:
            scan = max_over_ranks(float(np.mean([m["scan_start"].elapsed_time(m["scan_end"]) for m in marks])))
        except Exception:          # noqa: BLE001 - (the small-shard protocol does not pass the marked call)
            scan = None
    stats = retr.filter_stats(args.batch)
    k4_parts = None
    if world == 1 and args.engine == "sel":
        # K4 stage by stage through the shard-level entry points (same kernels, same query embeddings): outside the timed region
        import ctypes as C_
        p_, b_ = retr.plan, retr.bank
        st_ = torch.cuda.current_stream().cuda_stream
        tiles_ = lib.vfr_sel_tiles(retr.n_clips)
        parts = []
        for _ in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            _lib.call("vfr_sel_query_pack", p_.q_emb, args.batch, b_.dim, p_.bank_tc, retr.n_clips, p_.q_tc, st_)
            ev[1].record()
            _lib.call("vfr_sel_filter", p_.bank_tc, retr.n_clips, b_.dim, p_.q_tc, args.batch, TOPK, p_.topk_ws, p_.n_split, 0, tiles_, 0, st_)
            ev[2].record()
            _lib.call("vfr_sel_refine", p_.bank_clips, p_.vid_off, p_.mom_off, b_.n_videos, retr.n_clips, b_.n_max, b_.dim, p_.q_tc,
                      p_.q_emb, args.batch, TOPK, p_.id_base, p_.out_scores_dev, p_.out_ids_dev, p_.topk_ws, p_.n_split, st_)
            ev[3].record()
            torch.cuda.synchronize()
            parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
        pm = np.mean(parts[1:], axis=0)
        k4_parts = {"query_pack_ms": float(pm[0]), "sample_pass_and_filter_ms": float(pm[1]), "refine_ms": float(pm[2]),
                    "filter_tflops_algorithmic": args.batch * retr.n_clips * 2.0 * args.dim / pm[1] / 1e9,
                    "filter_frac_of_sustained_peak": args.batch * retr.n_clips * 2.0 * args.dim / pm[1] / 1e9 / peaks()["bf16_tflops_sustained"]}
    stage_ms = None
    if world > 1:
        retr.profile = True
        for i in range(args.steps):
            retr.search_device(tokens_dev[args.warmup + i])
        stage_ms = retr.stage_ms()
        retr.profile = False

    # ---- end to end through the host-buffer API ----
    for i in range(args.warmup):
        retr.search(tokens_host[i])
    barrier()
    if world > 1:
        retr.trace_search = {}
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        s, ids = retr.search(tokens_host[args.warmup + i])
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_host = None
    if world > 1 and retr.trace_search:
        n_tr = max(retr.trace_search.pop("steps", 1), 1)
        e2e_host = {k: v / n_tr * 1e3 for k, v in retr.trace_search.items()}       # rank 0's host-side ms per step
        retr.trace_search = None
    checksum = float(s[:, 0].double().sum().item())
    if world > 1:
        t = torch.tensor([checksum], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        checksum = float(t.item())

    pk = peaks()
    local_pairs = args.batch * retr.bank.m_total                 # pairs this rank's K4 launch scores
    flop_per_pair = 4.0 * args.dim / (N_SEG + 1)                 # 2*D*S / (S(S+1)/2)
    # the dominant KERNEL is the filter's scan (sl_filter_kernel); the whole of K4 (query pack, scan, exact re-scoring and,
    # on N ranks, the threshold protocol + list exchange + merge) is reported next to it as k4_*
    dom_ms = scan if scan else k4
    achieved_tflops = local_pairs * flop_per_pair / (dom_ms * 1e-3) / 1e12
    kernel_names = {
        "sel": ("sl_filter_kernel (fp16 tcgen05 GEMM + min / threshold epilogue: the scan of vfr_sel_topk" + (", sample pass included)" if world == 1 else " over this rank's shard)")
                if scan else "K4 = vfr_sel_query_pack + vfr_sel_topk (sl_filter_kernel + sl_refine_kernel)"),
        "tc": "vfr_score_topk_tc (score_tc_kernel<TOPK> + threshold init + topk_finish_kernel)",
        "tc_bf16": "vfr_score_topk_tc (score_tc_kernel<TOPK>, plain bf16)",
        "exact": "vfr_score_topk (score_kernel<TOPK> + topk_finish_kernel)"}
    notes = {
        "sel": "one fp16 tcgen05 pass (K padded to a multiple of 16: %d executed FLOP per (query, clip) vs %d algorithmic), the k best "
               "moments are provably inside the videos of the ~k closest clips, which are re-scored exactly in fp32; results "
               "bit-identical to the exact engine" % (2 * ((args.dim + 3 + 15) // 16 * 16), 2 * args.dim),
        "tc": "tcgen05 split-bf16 GEMM (3 MMA passes) + fused sqrt / moment-mean / top-k epilogue; algorithmic FLOPs "
              "count ONE fp32 pass (4D/(S+1) per pair), so frac understates tensor-pipe use 3.4x",
        "tc_bf16": "tcgen05 plain-bf16 GEMM + fused epilogue (1e-2 tolerance)",
        "exact": "exact-fp32 CUDA-core path (FADD+FFMA direct-difference form): 2 fp32 instr per (clip, dim); fp32 FFMA peak "
                 "~72 TFLOP/s is the real ceiling of this path"}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        key = f"{args.engine}:{world}:{args.batch}:{args.videos}:{args.dim}:{args.bank}"
        traffic = tj.get(key, {}).get("bytes_per_launch")
    roofline = {
        "kernel": kernel_names[args.engine],
        "bound": "tensor", "achieved": achieved_tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved_tflops / pk["bf16_tflops_sustained"], "traffic": traffic,
        "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
        "ms_per_launch": dom_ms, "algorithmic_flop_per_pair": flop_per_pair,
        "note": notes[args.engine], "share_of_step": dom_ms * args.steps / ms_total,
        "timing": "CUDA events recorded inside every timed step (start, K3 done, scan start / end, end); max over ranks of the means",
        "k4_ms": k4, "k4_frac": local_pairs * flop_per_pair / (k4 * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
        "k4_share_of_step": k4 * args.steps / ms_total,
        "k4_is": "everything after K3: query pack, scan, exact re-scoring" + (", threshold protocol, list exchange, merge, flags" if world > 1 else ""),
        "stages": k4_parts,
    }
    # K3: 352 MFLOP per query in the reference's form (every padded step of both directions)
    H, E = model.hidden_size, model.word_embedding.weight.shape[1]
    k3_flop = 2.0 * 20 * 2 * 4 * H * (E + H) + 2.0 * 2 * H * args.dim
    k3_tflops = (args.batch / world) * k3_flop / (k3 * 1e-3) / 1e12
    roofline_k3 = {
        "kernel": "K3 = vfr_text_embed_tc (gather + 20 x (join + gemm_tc_kernel<EpiLstmTc>) + fc)" + (" + all-gather of the query slices" if world > 1 else ""),
        "bound": "tensor", "achieved": k3_tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": k3_tflops / pk["bf16_tflops_sustained"], "ms_per_launch": k3, "algorithmic_flop_per_query": k3_flop,
        "note": "reference-form FLOPs (padding fed through both directions); the kernels run 3 bf16 passes per fp32 product "
                "and skip the backward direction's padding, so executed tensor work is ~2.2x the algorithmic count",
        "share_of_step": k3 * args.steps / ms_total, "queries_per_rank": args.batch // world,
    }

    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(args, retr, model, tokens_host[-1], s, ids, device)
        except AssertionError as e:
            parity = {"ok": False, "error": str(e)}
    if world > 1:
        ok = torch.tensor([1 if (parity is None or parity.get("ok")) else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if parity is not None:
            parity["all_ranks_ok"] = bool(ok.item())

    cpu_baseline = library = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        arm = ReferenceArm(args.videos, dim=args.dim, kind=args.bank)
        t0 = time.perf_counter()
        n = 0
        while n < 2 or time.perf_counter() - t0 < 12.0:
            arm.step()
            n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": arm.pairs_per_step * n / dt, "unit": "pairs/s", "cores": cores, "kind": arm.kind,
                        "sample": arm.describe() + f", {n} steps in {dt:.1f} s", "host_cpus": os.cpu_count()}
    if rank == 0 and world == 1 and not args.no_library_baseline:
        library = library_baseline(args, device, model, retr.bank.clips, tokens_dev[0])

    if rank == 0:
        line = {
            "metric": "query-moment pairs scored/sec", "value": pairs_per_step * args.steps / (ms_total * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"sel": "f16 tensor-core filter (fp32 accumulate, rigorous error band) + f32 exact re-scoring: scores bit-identical to the fp32 path",
                      "tc": "bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32 scores within 1e-5)", "tc_bf16": "bf16", "exact": "f32"}[args.engine],
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": pairs_per_step * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": retr.h2d_bytes(args.batch), "d2h_bytes_per_step": retr.d2h_bytes(args.batch),
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "MomentRetriever.search" + (" -> vfr_search_host" if world == 1 else " (every rank moves and owns its query slice)"),
                    "host_ms_per_step": e2e_host},
            "gpu_launches": int(launches), "gpu_launches_source": "vfr_launch_count() around the timed region (rank 0)",
            "roofline": roofline, "roofline_k3": roofline_k3, "comm": stage_ms,
            "filter_stats": dict(stats or {}, n_fixups_value_steps=fixups_value, n_fixups_total=retr.n_fixups),
            "parity_check": parity, "cpu_baseline": cpu_baseline, "library_baseline": library, "checksum_top1": checksum,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity.get("ok"):
        sys.exit(3)


_JSON_FD = 1


def _emit(line):
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # stdout carries the ONE JSON line and nothing else: whatever libraries print there (NCCL's version banner ...)
    # goes to stderr
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

"""bench.py -- adapter-train embeddings/sec on CelebA-shaped synthetic RN50 embeddings (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one stage-1 training epoch of the reference (train_one_epoch, final_main.py:426-496) over the
GPU-resident 162,770 x 1024 fp32 embedding matrix at batch size 1024 (159 SGD steps) -- 667 MB per pass, larger
than the 126 MB L2, so consecutive timed steps cannot reuse cached inputs.  N > 1 (launched by torchrun, one
process per GPU): data parallel, weak scaling -- every rank holds its own 162,770-row shard and contributes 1024
rows to each global batch of N x 1024; BatchNorm statistics and gradients are all-reduced over NCCL.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline` (dominant kernel, live CUDA-event
timing), `cpu_baseline` (oracle port of the reference step on the host cores), `e2e` (host buffers, H2D/D2H
inside the timed region), `eval` (validate()-style forward throughput, the HBM-bound half of the path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, D, H, C, G = 162770, 1024, 128, 2, 4
BATCH = 1024
TRAIN_GROUPS = (71629, 66874, 22880, 1387)
ALG_BYTES_PER_EMB = 4096                       # one fp32 row of X
ALG_FLOP_TRAIN = 1.319e6                       # fwd 528,384 + bwd 786,432 + logits grad
ALG_FLOP_EVAL = 528384


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


def traffic_from_profiles(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full captures (profiles/r2_traffic.json, then the
    round-1 file), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            v = json.load(open(path)).get(kernel)
            if v is not None:
                return v
    return None


def synth_rows(n, seed, dim=D):
    """CelebA-shaped rows: x = base + k*mu_group + eps, rounded through fp16."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal(dim).astype(np.float32)
    mu = rng.standard_normal((4, dim)).astype(np.float32)
    pr = np.array(TRAIN_GROUPS, np.float64) / sum(TRAIN_GROUPS)
    g = rng.choice(4, size=n, p=pr).astype(np.int32)
    x = np.empty((n, dim), np.float32)
    for s in range(0, n, 16384):
        e = min(n, s + 16384)
        x[s:e] = base + 0.25 * mu[g[s:e]] + rng.standard_normal((e - s, dim), dtype=np.float32)
    x = x.astype(np.float16).astype(np.float32)
    T = np.stack([base + mu[[0, 1]].mean(0), base + mu[[2, 3]].mean(0)], 1).astype(np.float32)
    return x, (g // 2).astype(np.int32), g, T


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].startswith("Active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def extra_legs(torch, ops, dev, P, e0, e1):
    """Zero-shot head (config 4 shape per GPU slice) and contrastive regulariser (config 3 shape), device-resident."""
    res = {}
    rng = np.random.default_rng(5)
    # head: 262,144 rows x 1024-d against 1,000 prompt columns; 2 * D * C flop per row, single pass over X
    n, c = 262144, 1000
    U = torch.randn(n, D, device=dev).half().float()
    yh = torch.randint(0, c, (n,), device=dev, dtype=torch.int32)
    gh = torch.randint(0, G, (n,), device=dev, dtype=torch.int32)
    Th = ops.normalize_text(torch.randn(D, c, device=dev))
    st = ops.BatchStatsBuffers((n + 1023) // 1024, G, device=dev)
    U16 = U.half()                              # the resident format of CLIP embeddings (lossless: the rows are fp16-valued)

    def time_head(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3 * 1e-3
    t32 = time_head(lambda: ops.logits_ce(U, yh, gh, Th, 100.0, st, 1024, G=G))
    t = time_head(lambda: ops.logits_ce_f16(U16, yh, gh, Th, 100.0, st, 1024, G=G))
    fl = 2.0 * D * c * n
    res["head_c1000"] = {"rows": n, "emb_per_s": n / t, "ms": t * 1e3, "algorithmic_tflops": fl / t / 1e12,
                         "frac_of_bf16_sustained_peak": fl / t / 1e12 / P["tc_sustained"],
                         "resident_dtype": "f16 (dbmm_logits_ce_f16: kind::f16 x 2 terms, X as stored * (That_hi, 2^11 That_lo) fp16 pair)",
                         "fp32_resident": {"ms": t32 * 1e3, "algorithmic_tflops": fl / t32 / 1e12,
                                           "frac_of_bf16_sustained_peak": fl / t32 / 1e12 / P["tc_sustained"],
                                           "note": "dbmm_logits_ce: kind::tf32 x 2 terms, 4x the bf16 cost per algorithmic flop"}}
    del U, U16
    # contrastive: B = 8192, d = 768, forward + backward; 6 * B * d flop per row
    B, d = 8192, 768
    Z = torch.nn.functional.normalize(torch.randn(B, d, device=dev), dim=1).contiguous()
    lab = torch.randint(0, 4, (B,), device=dev, dtype=torch.int32)
    sc = ops.SupconState(device=dev)
    for _ in range(2):
        sc.zero_(); ops.supcon_fwd(Z, lab, sc); ops.supcon_bwd(Z, sc)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        sc.zero_(); ops.supcon_fwd(Z, lab, sc); ops.supcon_bwd(Z, sc)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    fl = 6.0 * B * d * B
    res["supcon_b8192_d768"] = {"rows": B, "emb_per_s": B / t, "ms": t * 1e3, "algorithmic_tflops": fl / t / 1e12,
                                "frac_of_bf16_sustained_peak": fl / t / 1e12 / P["tc_sustained"], "loss": sc.loss(),
                                "note": "kind::f16 on fp16 pairs of the power-of-two-scaled operands (three products, one TMEM accumulator: 3x the bf16 cost per "
                                        "algorithmic flop); whole-batch backward as ONE GEMM dZ = (G + G^T) Z; similarity matrix staged in HBM once"}
    return res


def multi_gpu_legs(torch, dist, ops, parallel, dev, P, e0, e1, world, rank, barrier):
    """BASELINE configs 4 and 3 at N > 1.  Config 4: the zero-shot head over 10 M x 1024-d rows x 1,000 prompts sharded by rows
    (1.25 M rows = 5.1 GB per GPU, the 8-GPU share; rows independent, counters and loss all-reduced ONCE at the end).  Config 3: the
    contrastive regulariser at global B = 8192, d = 768 with all-gathered negatives (parallel.supcon_distributed: NCCL all-gather
    of the normalised rows, tcgen05 similarity GEMMs on each rank's anchors, reduce-scatter of the contrast-role gradient)."""
    res = {}
    n, c = 1250000, 1000
    U = torch.empty(n, D, device=dev, dtype=torch.float16)          # fp16-resident, as the packed store keeps CLIP embeddings
    for s0 in range(0, n, 250000):
        U[s0:s0 + 250000] = torch.randn(min(250000, n - s0), D, device=dev).half()
    yh = torch.randint(0, c, (n,), device=dev, dtype=torch.int32)
    gh = torch.randint(0, G, (n,), device=dev, dtype=torch.int32)
    Th = ops.normalize_text(torch.randn(D, c, device=dev, generator=torch.Generator(device=dev).manual_seed(9)))
    st = ops.BatchStatsBuffers((n + 1023) // 1024, G, device=dev)

    def head_pass():
        st.zero_()
        ops.logits_ce_f16(U, yh, gh, Th, 100.0, st, 1024, G=G)
        tot = torch.cat([st.loss_sum.sum().reshape(1), st.counts.sum(0).reshape(-1).to(torch.float64)])
        dist.all_reduce(tot)                       # the one exchange of the sharded head: loss sum + 2 G counters
        return tot
    head_pass()
    barrier()
    e0.record()
    for _ in range(2):
        tot = head_pass()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts = float(t.item()) * 1e-3
    fl = 2.0 * D * c * n * world
    res["head_c1000_sharded"] = {"rows_total": n * world, "rows_per_gpu": n, "prompts": c, "emb_per_s": n * world / ts, "ms": ts * 1e3,
                                 "algorithmic_tflops_per_gpu": fl / ts / 1e12 / world,
                                 "frac_of_bf16_sustained_peak": fl / ts / 1e12 / world / P["tc_sustained"],
                                 "rows_counted": int(tot[1 + G:].sum().item()), "exchange": "one all-reduce of loss sum + 2G counters",
                                 "resident_dtype": "f16 (dbmm_logits_ce_f16)"}
    del U
    B, d = 8192, 768
    Bl = B // world
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    Z = torch.nn.functional.normalize(torch.randn(Bl, d, device=dev, generator=gen), dim=1).contiguous()
    lab = torch.randint(0, 4, (Bl,), device=dev, dtype=torch.int32, generator=gen)
    for _ in range(2):
        loss, dZ = parallel.supcon_distributed(Z, lab)
    barrier()
    e0.record()
    for _ in range(3):
        loss, dZ = parallel.supcon_distributed(Z, lab)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts = float(t.item()) * 1e-3
    fl = 6.0 * B * d * B
    res["supcon_b8192_d768_global_negatives"] = {"global_batch": B, "rows_per_gpu": Bl, "emb_per_s": B / ts, "ms": ts * 1e3,
                                                 "algorithmic_tflops_per_gpu": fl / ts / 1e12 / world, "loss": loss,
                                                 "dZ_finite": bool(torch.isfinite(dZ).all().item()),
                                                 "collectives": "all-gather Z + labels, all-reduce (loss, n_valid), reduce-scatter dZ (NCCL)"}
    # the same regulariser TRAINING an adapter under data parallelism (parallel.contrastive_step_distributed): forward_ca of the
    # rank's rows, global negatives, D-wide backward, gradient all-reduce, SGD; the replicas must stay identical
    from dbmm.modules import Adapter
    torch.manual_seed(11)
    ad_c = Adapter(d, 128).to(dev).tensors()
    buf_c = ops.TrainBuffers(d, 128, device=dev)
    Xc = torch.randn(16 * Bl, d, device=dev, generator=gen).half().float()
    yc = torch.randint(0, 2, (16 * Bl,), device=dev, dtype=torch.int32, generator=gen)
    Xc += 0.5 * (yc.float()[:, None] * 2 - 1) * torch.randn(1, d, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    Xc = Xc.half().float()
    losses = []
    for s in range(2):
        parallel.contrastive_step_distributed(Xc, yc, ad_c, buf_c, 0.05, torch.arange(s * Bl, (s + 1) * Bl, device=dev, dtype=torch.int32))
    barrier()
    e0.record()
    for s in range(2, 12):
        losses.append(parallel.contrastive_step_distributed(Xc, yc, ad_c, buf_c, 0.05,
                                                            torch.arange(s * Bl, (s + 1) * Bl, device=dev, dtype=torch.int32)))
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    chk = torch.stack([ad_c.W1.double().sum(), ad_c.W2.double().abs().sum()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    res["contrastive_adapter_train_step_dp"] = {"global_batch": B, "rows_per_gpu": Bl, "ms_per_step": float(t.item()),
                                                "emb_per_s": B / (float(t.item()) * 1e-3), "loss_first": losses[0], "loss_last": losses[-1],
                                                "replicas_identical": bool(all(torch.equal(c, allc[0]) for c in allc)),
                                                "what": "forward_ca + all-gathered negatives + D-wide backward + gradient all-reduce + SGD per step "
                                                        "(host-driven: includes the loss read-back of every step)"}
    return res


def cpu_reference_step(rows_x, rows_y, rows_g, T, n_sgd_steps):
    """`n_sgd_steps` batches through the reference's train_one_epoch on the host cores, tensors pre-loaded (bounded sample).
    kind "reference": the UNMODIFIED reference modules (final_main.Adapter / CustomCLIP / train_one_epoch, demo.util
    set_optimizer) staged by oracle/stage_reference.py into oracle/_ref; kind "port" (only if nothing was staged): the
    torch-CPU restatement oracle/ref_port.py."""
    import torch
    from oracle import ref_run
    torch.set_num_threads(os.cpu_count() or 1)
    if ref_run.available():
        r = ref_run.time_train_epochs(rows_x, rows_y, rows_g, T, H, BATCH, n_sgd_steps, device="cpu", lr=0.1)
        r["kind"] = "reference"
        return r
    from oracle import ref_port
    r = ref_port.time_train_steps(rows_x, rows_y.astype(np.int64), rows_g.astype(np.int64), T, H, BATCH, n_sgd_steps)
    r["kind"] = "port"
    return r


def reference_dataloader_leg():
    """the reference's own Dataset / DataLoader path (pandas JSON parse, per-item lookups) feeding
    its train_one_epoch on a Waterbirds-shaped file (config 0), host cores.  None when the reference is not staged."""
    import contextlib
    import io
    from oracle import ref_run
    if not ref_run.available():
        return None
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref_run.time_dataloader_epoch(n_rows=4795, dim=D, H=H, batch_size=BATCH)
    return {"value": r["emb_per_s_epoch"], "unit": "embeddings/s", "rows": r["rows"], "epoch_seconds": r["epoch_seconds"],
            "dataset_build_seconds": r["build_seconds"], "value_with_build": r["emb_per_s_with_build"],
            "num_workers": r["num_workers"], "cores": r["threads"], "note": r["note"]}


def accuracy_leg():
    """The other half of BASELINE.json's metric: worst-group accuracy delta against the reference on the FULL config 0 / 1
    hyper-parameters (run_final_main.sh:1-31: 4,795 train rows, bs 1024 / bsr 256, 100 epochs / 40 feature-learning, lr 1.0).
    The reference's numbers for these synthetic files and this seed are the committed golden (tests/golden/e2e_cases.json,
    generated by oracle/make_golden.py from the unmodified reference); the drop-in runs the same CLI here."""
    import contextlib
    import io
    import tempfile
    from dbmm import cli, synth
    case = json.load(open(os.path.join(ROOT, "tests", "golden", "e2e_cases.json")))["waterbirds_full"]
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        ds = synth.make_dataset(**case["synth"])
        paths = synth.write_reference_files(ds, tmp)
        argv = [a for a in case["argv"] if a != "--save_results"]
        for k, v in paths.items():
            argv += [f"--{k}", v]
        t1 = time.perf_counter()
        (tr, va, te), _ = cli.train_all_epochs(cli.parse_option(argv))
        run = cli.train_all_epochs.last_run
    t2 = time.perf_counter()
    ref_test = case["final"][2]
    keys = ("acc_0_0", "acc_0_1", "acc_1_0", "acc_1_1", "mean_acc", "worst_acc")
    return {"case": "waterbirds-shaped synthetic, adapter_reg_seq_alter --add_adapter --warm_reg, bs 1024 / bsr 256, lr 1.0, 100 epochs / 40 FL, seed 42",
            "worst_group_acc": float(te["worst_acc"]), "reference_worst_group_acc": float(ref_test["worst_acc"]),
            "worst_group_acc_delta": float(te["worst_acc"]) - float(ref_test["worst_acc"]),
            "selected_epoch": int(run["best_epoch"]), "reference_selected_epoch": int(case["best_epoch"]),
            "test_dict_max_abs_delta": max(abs(float(te[k]) - float(ref_test[k])) for k in keys),
            "train_seconds": t2 - t1, "file_write_seconds": t1 - t0,
            "reference_noise_floor": _noise_floor()}


def _noise_floor():
    """The unmodified reference against ITSELF on this config with another CPU thread count (oracle/config1_noise_floor.py)."""
    try:
        nf = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_noise_floor.json")))
        return {"what": "unmodified reference re-run with 1 / 3 CPU threads vs its 8-thread golden (same seed, same data)",
                "runs": [{"threads": r["threads"], "epochs_with_identical_test_dict": r["epochs_exact"], "of": r["epochs"],
                          "first_differing_epoch": r["first_differing_epoch"], "worst_group_acc_delta": r["worst_acc_delta_vs_golden"],
                          "selected_epoch": r["best_epoch"]} for r in nf["runs"] if r["threads"] != nf["golden_threads"]]}
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sgd = 159                                 # one epoch's worth of SGD steps of 1024 rows per bench step (bounded sample:
    x, y, g, T = synth_rows(48 * BATCH, seed=1234)      # cycling over the first 49,152 rows of the workload)
    cpu_reference_step(x, y, g, T, 2)           # page-in / thread pool warm-up
    for _ in range(args.warmup):
        cpu_reference_step(x, y, g, T, n_sgd)
    t0 = time.perf_counter()
    rows = 0
    for _ in range(args.steps):
        r = cpu_reference_step(x, y, g, T, n_sgd)
        rows += r["rows"]
    dt = time.perf_counter() - t0
    val = rows / dt
    sample = (f"{n_sgd} SGD steps of {BATCH} rows per bench step through the reference's train_one_epoch "
              f"(kind {r['kind']}: {'unmodified reference modules from oracle/_ref' if r['kind'] == 'reference' else 'oracle/ref_port.py'}), "
              "tensors pre-loaded, torch CPU, all host threads")
    dl = reference_dataloader_leg()
    print(json.dumps({
        "impl": "reference", "metric": "adapter-train embeddings/sec", "value": val, "unit": "embeddings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "embeddings/s", "cores": r["threads"], "kind": r["kind"], "sample": sample,
                         "dataloader_inclusive": dl},
        "e2e": {"value": val, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus):
    return {"workload": "celeba-shaped synthetic CLIP-RN50 embeddings, stage-1 adapter training epoch "
                        "(train_one_epoch: Linear-BN-ReLU-Linear adapter, L2-norm, cosine logits/0.01 vs 2 class prompts, "
                        "CE, SGD momentum 0.9 wd 5e-5, per-group counters)",
            "rows_per_gpu": N_TRAIN, "dim": D, "adapter_feat_dim": H, "classes": C, "groups": G,
            "batch_size_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "sgd_steps_per_step": (N_TRAIN + BATCH - 1) // BATCH,
            "parallelism": f"dp{n_gpus}", "l2_policy": "inputs (667 MB per step) larger than L2 (126 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dbmm", choices=["dbmm", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-accuracy", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import dbmm
    from dbmm import ops, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = parallel.bind_to_gpu_numa_node(local_rank) if world > 1 else None      # host staging next to the GPU (e2e leg)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"            # the version banner goes to stdout, where ONE JSON line is expected
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    P = peaks()

    # ---- data: this rank's shard, resident in HBM; pinned host copy for the e2e leg
    x_np, y_np, g_np, T_np = synth_rows(N_TRAIN, seed=1234 + rank)
    if world > 1:
        T_np = synth_rows(8, seed=1234)[3] if rank else T_np        # same prompts everywhere
    X = torch.from_numpy(x_np).to(dev)
    y = torch.from_numpy(y_np).to(dev)
    g = torch.from_numpy(g_np).to(dev)
    That = ops.normalize_text(torch.from_numpy(T_np).to(dev))
    steps_per_epoch = (N_TRAIN + BATCH - 1) // BATCH

    def fresh_model():
        torch.manual_seed(42)
        from dbmm.modules import Adapter
        a = Adapter(D, H).to(dev)
        return a, a.tensors()

    mod, ad = fresh_model()
    buf = ops.TrainBuffers(D, H, device=dev)
    stats = ops.BatchStatsBuffers(steps_per_epoch, G, device=dev)
    lrs = np.full(steps_per_epoch, 0.1, np.float32)      # CelebA setting of the reference: lr 0.1
    gen = torch.Generator().manual_seed(7)
    orders = [torch.randperm(N_TRAIN, generator=gen).to(torch.int32).to(dev) for _ in range(4)]
    order_buf = torch.empty_like(orders[0])      # fixed address: the epoch's CUDA graph is cached by argument addresses
    dp = parallel.DataParallelTrainer(local_batches=True) if world > 1 else None

    def train_epoch(i):
        stats.zero_()
        if world == 1:
            order_buf.copy_(orders[i % 4])
            ops.train_epoch(X, order_buf, BATCH, y, g, ad, That, 100.0, buf, lrs, stats, G=G)
        else:
            # weak scaling: every rank holds its own 162,770-row shard and contributes its own 1024 rows to each global
            # batch of world x 1024; BatchNorm statistics / CE mean / gradients are those of the global batch
            order_buf.copy_(orders[i % 4])
            dp.train_epoch(X, order_buf, BATCH, y, g, ad, That, 100.0, buf, lrs, stats, G=G)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up epochs, then exactly K timed epochs
    for i in range(args.warmup):
        train_epoch(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        train_epoch(args.warmup + i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss_sum, counts = stats.host()
    final_loss = float(loss_sum.sum() / N_TRAIN)
    value = world * N_TRAIN * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps

    out = None
    if rank == 0:
        out = {"metric": "adapter-train embeddings/sec", "value": value, "unit": "embeddings/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(world), "clocks": clocks, "numa_binding": numa,
               "us_per_sgd_step": 1e3 * ms_per_step / steps_per_epoch, "final_epoch_mean_loss": final_loss,
               # per SGD step: k_gemm1_tc, k_reduce_stats, k_rows_train, k_wgrad_tc, k_tail_w1 + (second graph branch) k_sum_spart_g, k_hs_w2,
               # k_sum_gpart -- the same eight on one GPU and under data parallelism (the exchanges ride inside them); + 3 per epoch prologue
               "gpu_launches": args.steps * (steps_per_epoch * 8 + 3)}

    # ---- per-kernel timing inside the running step (CUDA events between the kernels, stream launches) -> roofline of the
    #      dominant kernel.  Algorithmic bytes / flops per launch: GEMM-1 and dW1 each stream the
    #      batch's X once (B * 4096 B) and do 2*D*H flop per row; the row kernel moves A + dahat (2 * B * H * 4 B).
    if world == 1:
        order_buf.copy_(orders[0])
        kus = ops.train_epoch_profile(X, order_buf, BATCH, y, g, ad, That, 100.0, buf, lrs, stats, G=G)
    else:
        kus = None
        if rank == 0:
            # N > 1: the per-kernel intervals are measured at N = 1 only (dbmm_train_epoch_profile is single-GPU); the line
            # carries the whole step of ONE rank against the roofs, like roofline.step of the N = 1 line
            t_step = ms_per_step * 1e-3 / steps_per_epoch
            gbs = ALG_BYTES_PER_EMB * BATCH / t_step / 1e9
            out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": P["hbm"], "unit": "GB/s", "frac": gbs / P["hbm"],
                               "traffic": None, "kernel": "whole data-parallel step of one rank (per-kernel figures: N = 1 line)",
                               "peak_source": P["src"], "algorithmic_bytes_per_launch": ALG_BYTES_PER_EMB * BATCH,
                               "step": {"bound": "tensor", "achieved_tflops": ALG_FLOP_TRAIN * BATCH / t_step / 1e12,
                                        "frac": ALG_FLOP_TRAIN * BATCH / t_step / 1e12 / P["tc_sustained"]}}
    if rank == 0 and kus is not None:
        np_ = 2 * D * H + 3 * H + D
        alg = {"gemm1_tc": (ALG_BYTES_PER_EMB * BATCH, 2.0 * D * H * BATCH), "wgrad_tc": (ALG_BYTES_PER_EMB * BATCH, 2.0 * D * H * BATCH),
               "reduce_stats": (2 * 4 * H * BATCH, 2.0 * H * BATCH), "rows_train": (2 * 4 * H * BATCH, 4.0 * H * (H + C) * BATCH),
               "finalize_grads": (4 * np_, 2.0 * D * (H + 1) * (H + 1 + C)), "update": (5 * 4 * np_, 4.0 * np_),
               "tail_w1": (4 * (16 + 5) * D * H, 20.0 * D * H), "tail_w2": (4 * 4 * D * H, 4.0 * D * (H + 1) * (H + 1 + C))}
        # the event records between the kernels cost device time themselves (the intervals sum to ~1.5x the step time
        # measured inside the epoch graph): scale them so that the kernels on the step's critical path add up to the
        # graph-mode step.  With the fused tail (mode 2) k_tail_w2 runs on a second branch of the graph, overlapped with
        # k_wgrad_tc / k_tail_w1 / the next step's k_gemm1_tc / k_reduce_stats: it is timed but not on the critical path.
        tmode = ops.train_tail_mode(BATCH, N_TRAIN - (steps_per_epoch - 1) * BATCH, 1, D, H, C)
        off_path = {"tail_w2"} if tmode == 2 else set()
        us_step = 1e3 * ms_per_step / steps_per_epoch
        raw_sum = sum(v for k, v in kus.items() if k not in off_path)
        kus_raw = dict(kus)
        kus = {k: v * us_step / raw_sum for k, v in kus.items()}
        dom = max((k for k in kus if k not in off_path), key=kus.get)
        dom_bytes, dom_flop = alg[dom]
        t_dom = kus[dom] * 1e-6
        t_hbm, t_tc = dom_bytes / (P["hbm"] * 1e9), dom_flop / (P["tc_sustained"] * 1e12)
        if t_tc >= t_hbm:
            roof = {"bound": "tensor", "achieved": dom_flop / t_dom / 1e12, "peak": P["tc_sustained"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": dom_bytes / t_dom / 1e9, "peak": P["hbm"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel"] = "k_" + dom
        roof["traffic"] = traffic_from_profiles("k_" + dom)
        roof["peak_source"] = P["src"] + (" (sustained bf16 figure: kernel timed inside a long step)" if roof["bound"] == "tensor" else "")
        roof["kernel_us"] = kus
        roof["kernel_us_event_timed"] = kus_raw
        roof["tail_mode"] = tmode
        roof["overlapped_kernels"] = sorted(off_path)
        roof["algorithmic_bytes_per_launch"] = dom_bytes
        # whole training step against its binding roof (the tensor roof binds the reference
        # formulation of the step); the dependent-phase latency floor is discussed in DESIGN.md section 5
        t_step = ms_per_step * 1e-3 / steps_per_epoch
        roof["step"] = {"bound": "tensor", "achieved_tflops": ALG_FLOP_TRAIN * BATCH / t_step / 1e12,
                        "frac": ALG_FLOP_TRAIN * BATCH / t_step / 1e12 / P["tc_sustained"],
                        "hbm_gbs": ALG_BYTES_PER_EMB * BATCH / t_step / 1e9,
                        "hbm_frac": ALG_BYTES_PER_EMB * BATCH / t_step / 1e9 / P["hbm"]}
        out["roofline"] = roof

    # ---- eval leg: validate()-style forward over the resident matrix (HBM-bound half of the path).  The resident format is
    #      what the packed store holds: fp16 (lossless for CLIP embeddings) -> dbmm_eval_fwd_f16; the fp32-resident kernel
    #      (dbmm_eval_fwd, tf32 hi + lo) is timed beside it.  Roofline: ALGORITHMIC bytes of SURVEY.md section 8d (one fp32 row,
    #      4,096 B per embedding) over the measured copy bandwidth, whatever the kernel really moves (`traffic`).
    st_e = ops.BatchStatsBuffers((N_TRAIN + 511) // 512, G, device=dev)
    X16 = X.half()
    f16_exact = bool(torch.equal(X16.float(), X)) and ops.eval_f16_supported(D, H, C)

    def time_eval(fn, reps=5):
        for _ in range(2):
            fn()
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    ms_e32 = time_eval(lambda: ops.eval_fwd(X, y, g, ad, That, 100.0, st_e, 512, G=G))
    ms_e = time_eval(lambda: ops.eval_fwd_f16(X16, y, g, ad, That, 100.0, st_e, 512, G=G)) if f16_exact else ms_e32
    del X16
    if rank == 0:
        ev = N_TRAIN / (ms_e * 1e-3)
        out["eval"] = {"value": ev * world, "unit": "embeddings/s", "ms_per_pass": ms_e,
                       "resident_dtype": "f16 (dbmm_eval_fwd_f16, kind::f16 tcgen05)" if f16_exact else "f32",
                       "roofline": {"bound": "hbm", "achieved": ev * ALG_BYTES_PER_EMB / 1e9, "peak": P["hbm"], "unit": "GB/s",
                                    "frac": ev * ALG_BYTES_PER_EMB / 1e9 / P["hbm"],
                                    "traffic": traffic_from_profiles("eval_fwd_f16_per_row" if f16_exact else "eval_fwd_per_row"),
                                    "traffic_unit": "DRAM bytes per row (algorithmic 4096)"},
                       "fp32_resident": {"value": N_TRAIN / (ms_e32 * 1e-3) * world, "ms_per_pass": ms_e32,
                                         "frac": N_TRAIN / (ms_e32 * 1e-3) * ALG_BYTES_PER_EMB / 1e9 / P["hbm"]}}

    # ---- batched-adapter sweep (BASELINE config 5): 64 members in lock step over one resident matrix, members sharded over
    #      the ranks with no communication.  Rows bounded to 40,960 per member-epoch to keep the leg short.
    M_total, n_b = 64, 40 * BATCH
    M_local = M_total // world
    from dbmm.modules import Adapter
    members = []
    for m in range(M_local):
        torch.manual_seed(1000 + rank * M_local + m)
        members.append(ops.SweepMember(order=torch.randperm(n_b, device=dev).to(torch.int32), ad=Adapter(D, H).to(dev).tensors(),
                                       buf=ops.TrainBuffers(D, H, device=dev), stats=ops.BatchStatsBuffers(n_b // BATCH, G, device=dev),
                                       lrs=np.full(n_b // BATCH, 0.01, np.float32)))
    Xb = X[:n_b]
    for _ in range(2):          # warm-up: the first call carries first_step, the second captures the steady-state epoch graph
        ops.train_epoch_batched(Xb, members, BATCH, y, g, That, 100.0)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e0.record()
    for i in range(3):
        evs[i].record()
        ops.train_epoch_batched(Xb, members, BATCH, y, g, That, 100.0)
    evs[3].record()
    e1.record()
    barrier()
    per_epoch_b = [evs[i].elapsed_time(evs[i + 1]) for i in range(3)]
    print("batched sweep epochs (ms):", per_epoch_b, file=sys.stderr)
    # median of the three timed epochs (all three are reported as `epochs_ms`): the first replay after a graph instantiation
    # occasionally carries a one-off host-side upload stall of tens of ms that is not part of the steady state
    ms_b = sorted(per_epoch_b)[1]
    if world > 1:
        t = torch.tensor([ms_b], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_b = float(t.item())
    del members
    if rank == 0:
        rate = M_total * n_b / (ms_b * 1e-3)
        out["batched_sweep"] = {"members": M_total, "members_per_gpu": M_local, "rows_per_member_epoch": n_b, "value": rate,
                                "unit": "embeddings/s (all members, all GPUs)", "ms_per_epoch": ms_b, "epochs_ms": per_epoch_b,
                                "timing": "median of 3 timed epochs after 2 warm-up epochs (max over ranks)",
                                "us_per_member_step": 1e3 * ms_b / (n_b // BATCH) / M_local,
                                "algorithmic_tflops_per_gpu": ALG_FLOP_TRAIN * rate / world / 1e12,
                                "frac_of_bf16_sustained_peak": ALG_FLOP_TRAIN * rate / world / 1e12 / P["tc_sustained"],
                                "hbm_frac": ALG_BYTES_PER_EMB * rate / world / 1e9 / P["hbm"],
                                "scaling": "members sharded over the ranks, zero communication",
                                "api": "dbmm_train_epoch_batched (k_gemm1_tc / k_hs_rows / k_wgrad_tc / k_tail_w1 / k_hs_w2, one launch per step for all members)"}

    # ---- data-parallel parity (N > 1, outside the timed region): one epoch over shards of ONE global batch order against
    #      the single-rank run of the same order from the same initial weights
    if world > 1:
        xs, ys, gs, _ = synth_rows(12 * BATCH + 333, seed=777)
        Xs, ys_d, gs_d = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), torch.from_numpy(gs).to(dev)
        n_s = Xs.shape[0]
        st_s = (n_s + BATCH - 1) // BATCH
        order_s = torch.randperm(n_s, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(dev)
        lr_s = np.full(st_s, 0.05, np.float32)
        _, ad_dp = fresh_model()
        buf_dp, stats_dp = ops.TrainBuffers(D, H, device=dev), ops.BatchStatsBuffers(st_s, G, device=dev)
        dp_g = parallel.DataParallelTrainer(local_batches=False)
        dp_g.train_epoch(Xs, order_s, BATCH, ys_d, gs_d, ad_dp, That, 100.0, buf_dp, lr_s, stats_dp, G=G, reduce_stats=True)
        barrier()
        if rank == 0:
            _, ad_1 = fresh_model()
            buf_1, stats_1 = ops.TrainBuffers(D, H, device=dev), ops.BatchStatsBuffers(st_s, G, device=dev)
            ops.train_epoch(Xs, order_s, BATCH, ys_d, gs_d, ad_1, That, 100.0, buf_1, lr_s, stats_1, G=G)
            torch.cuda.synchronize()
            a, b = ad_dp.to_numpy(), ad_1.to_numpy()
            dev_max = max(float(np.abs(a[k].astype(np.float64) - b[k]).max() / max(np.abs(b[k]).max(), 1e-30))
                          for k in ("W1", "gamma", "beta", "W2", "b2", "running_mean", "running_var"))
            c_dp, c_1 = stats_dp.host()[1], stats_1.host()[1]
            out["dp_parity"] = {"max_rel_dev": dev_max, "counters_identical": bool(np.array_equal(c_dp, c_1)),
                                "counter_max_abs_diff": int(np.abs(c_dp - c_1).max()), "rows": int(n_s), "steps": int(st_s),
                                "what": f"one epoch, global batch {BATCH} sharded over {world} ranks vs the same order on one rank"}
        dp_g.close()
        barrier()

    # ---- kernels of BASELINE configs 3 / 4 (tcgen05 + TMA GEMMs), short legs, N = 1 only
    if rank == 0 and world == 1:
        out["extra"] = extra_legs(torch, ops, dev, P, e0, e1)
    if world > 1:
        ex = multi_gpu_legs(torch, dist, ops, parallel, dev, P, e0, e1, world, rank, barrier)
        if rank == 0:
            out["extra"] = ex

    # ---- e2e leg: host buffers in, statistics out, copies inside the timed region.  Two device buffer sets: the
    #      pinned-host -> device copy of epoch i+1's inputs runs on a copy stream while epoch i trains.  Host embeddings
    #      are held the way the packed store (dbmm/pack.py) holds them: fp16 when that is lossless (CLIP emits fp16; the
    #      synthetic rows are fp16-valued like the real ones), widened exactly on the device by dbmm_widen_f16 on the copy
    #      stream -- half the PCIe bytes of an fp32 host matrix.  The fp32-host variant is measured beside it.
    yh, gh = torch.from_numpy(y_np).pin_memory(), torch.from_numpy(g_np).pin_memory()
    oh = [o.cpu().pin_memory() for o in orders]
    sets = [dict(X=torch.empty_like(X), y=torch.empty_like(y), g=torch.empty_like(g), o=torch.empty_like(orders[0]),
                 st=ops.BatchStatsBuffers(steps_per_epoch, G, device=dev), ready=torch.cuda.Event(), free=torch.cuda.Event())
            for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    loss_h = torch.empty(steps_per_epoch, dtype=torch.float64).pin_memory()
    cnt_h = torch.empty(steps_per_epoch, 2, G, dtype=torch.int64).pin_memory()
    main_stream = torch.cuda.current_stream()
    x16_np = x_np.astype(np.float16)
    fp16_lossless = bool(np.array_equal(x16_np.astype(np.float32), x_np))

    def run_e2e(half_host):
        xh = torch.from_numpy(x16_np if half_host else x_np).pin_memory()
        stage = torch.empty(xh.shape, dtype=torch.float16, device=dev) if half_host else None

        def prefetch(i):
            b = sets[i % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(b["free"])             # the epoch that last used this buffer set has finished
                if half_host:
                    stage.copy_(xh, non_blocking=True)
                    ops.widen_f16(stage, out=b["X"])          # on the copy stream
                else:
                    b["X"].copy_(xh, non_blocking=True)
                b["y"].copy_(yh, non_blocking=True); b["g"].copy_(gh, non_blocking=True)
                b["o"].copy_(oh[i % 4], non_blocking=True)
                b["ready"].record(copy_stream)

        def e2e_epoch(i):
            b = sets[i % 2]
            main_stream.wait_event(b["ready"])
            prefetch(i + 1)                                   # overlaps with this epoch's kernels
            b["st"].zero_()
            if world == 1:
                ops.train_epoch(b["X"], b["o"], BATCH, b["y"], b["g"], ad, That, 100.0, buf, lrs, b["st"], G=G)
            else:
                dp.train_epoch(b["X"], b["o"], BATCH, b["y"], b["g"], ad, That, 100.0, buf, lrs, b["st"], G=G)
            loss_h.copy_(b["st"].loss_sum, non_blocking=True); cnt_h.copy_(b["st"].counts, non_blocking=True)
            b["free"].record(main_stream)
            main_stream.synchronize()                         # the step's result (loss / counters) is read on the host
            return float(loss_h.sum())

        n_e2e = max(3, min(args.steps, 5))
        for b in sets:
            b["free"].record(main_stream)
        prefetch(0)
        e2e_epoch(0); e2e_epoch(1)                            # warm-up: both buffer sets' graphs exist
        barrier()
        e0.record()
        for i in range(2, 2 + n_e2e):
            e2e_epoch(i)
        e1.record()
        barrier()
        copy_stream.synchronize()
        ms2 = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        h2d = xh.numel() * xh.element_size() + y_np.nbytes + g_np.nbytes + 4 * N_TRAIN
        return world * N_TRAIN * n_e2e / (ms2 * 1e-3), ms2 / n_e2e, int(h2d)

    v32, ms32, h2d32 = run_e2e(False)
    v16, ms16, h2d16 = run_e2e(True) if fp16_lossless else (v32, ms32, h2d32)
    if rank == 0:
        d2h = loss_h.numel() * 8 + cnt_h.numel() * 8
        out["e2e"] = {"value": v16, "unit": "embeddings/s",
                      "h2d_bytes_per_step": h2d16, "d2h_bytes_per_step": int(d2h), "ms_per_step": ms16,
                      "host_dtype": "f16 (lossless store, widened on the device)" if fp16_lossless else "f32",
                      "fp32_host": {"value": v32, "ms_per_step": ms32, "h2d_bytes_per_step": h2d32},
                      "api": "dbmm_widen_f16 + dbmm_train_epoch (C ABI): pinned host buffers copied in every step on a copy stream "
                             "(double-buffered against the previous step's kernels), per-batch statistics read back"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port of the reference step on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_rows_s = 48 * BATCH
        r = cpu_reference_step(x_np[:n_rows_s], y_np[:n_rows_s], g_np[:n_rows_s], T_np, 24)       # warm-up + rate estimate
        n_sgd = int(min(max(12.0 * r["emb_per_s"] / BATCH, 48), 6000))                             # ~12 s of CPU work
        r = cpu_reference_step(x_np[:n_rows_s], y_np[:n_rows_s], g_np[:n_rows_s], T_np, n_sgd)
        out["cpu_baseline"] = {"value": r["emb_per_s"], "unit": "embeddings/s", "cores": r["threads"], "kind": r["kind"],
                               "sample": f"{n_sgd} SGD steps of {BATCH} rows cycling over the first {n_rows_s} rows of the same "
                                         f"workload ({r['seconds']:.1f} s) through the reference's train_one_epoch, tensors pre-loaded "
                                         + ("(unmodified reference modules staged in oracle/_ref)" if r["kind"] == "reference"
                                            else "(oracle/ref_port.py restatement: the reference was not staged)"),
                               "dataloader_inclusive": reference_dataloader_leg()}
        # the same reference code on this B200 under stock PyTorch eager: the incumbent the kernels replace
        from oracle import ref_run
        if ref_run.available():
            ref_run.time_train_epochs(x_np[:n_rows_s], y_np[:n_rows_s], g_np[:n_rows_s], T_np, H, BATCH, 8, device="cuda")
            rg = ref_run.time_train_epochs(x_np[:n_rows_s], y_np[:n_rows_s], g_np[:n_rows_s], T_np, H, BATCH, 159, device="cuda", epochs=2)
            out["reference_on_b200"] = {"value": rg["emb_per_s"], "unit": "embeddings/s",
                                        "what": "unmodified reference train_one_epoch under stock PyTorch eager on cuda:0, batches resident",
                                        "seconds": rg["seconds"], "rows": rg["rows"], "speedup_of_value": value / rg["emb_per_s"]}
    if rank == 0 and world == 1 and not args.no_accuracy:
        try:
            out["accuracy"] = accuracy_leg()
        except Exception as exc:                      # the throughput line must not die with the accuracy leg
            out["accuracy"] = {"error": repr(exc)[:300]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

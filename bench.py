#!/usr/bin/env python
"""bench.py — positive links/sec of the neg_shared training loop on the C3 workload (synthetic 1M users x 1M items,
100M power-law links, dim 128, skip-gram, batch_size_p 512), plus whole@k users/sec as an extra.

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path through the C-ABI)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (NumPy transcription)

One "step" = one pass of the hot path over one device batch = R replicas x batch_size_p positive links
(R independent neg_shared batches computed against one table snapshot: synchronous data-parallel virtual workers;
R = 1 is the reference's strictly sequential loop and is reported alongside as `sequential`).
Prints ONE JSON line on rank 0.
"""
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

WORKLOAD = "C3: synthetic 1M users x 1M items, 100M power-law links, dim 128, neg_shared skip-gram, batch_size_p 512"
N_USERS = N_ITEMS = 1_000_000
N_LINKS = 100_000_000
DIM = 128
BATCH = 512
LAMBDA = 128.0
LR = 0.01


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------------------------------------------
# synthetic data
# ----------------------------------------------------------------------------------------------------------------
def synth_links_device(n_links, n_users, n_items, seed, torch):
    """int32 [n_links, 3] rows (user, item, 1) on the device: items ~ (rank+10)^-1.0, users ~ (rank+10)^-0.8 over
    seeded permutations of the id spaces; duplicates kept (the reference never dedupes)."""
    g = torch.Generator(device="cuda").manual_seed(seed)

    def draw(n_ids, expo, perm_seed, n):
        p = torch.pow(torch.arange(n_ids, device="cuda", dtype=torch.float64) + 10.0, -expo)
        cdf = torch.cumsum(p, 0)
        cdf = cdf / cdf[-1]
        r = torch.searchsorted(cdf, torch.rand(n, device="cuda", generator=g, dtype=torch.float64), right=True)
        r.clamp_(max=n_ids - 1)
        perm = torch.randperm(n_ids, device="cuda", generator=torch.Generator(device="cuda").manual_seed(perm_seed))
        return perm[r].to(torch.int32)

    out = torch.empty((n_links, 3), dtype=torch.int32, device="cuda")
    chunk = 20_000_000
    for s in range(0, n_links, chunk):
        n = min(chunk, n_links - s)
        out[s:s + n, 0] = draw(n_users, 0.8, 124, n)
        out[s:s + n, 1] = draw(n_items, 1.0, 123, n)
    out[:, 2] = 1
    return out


def synth_block_device(n_links, n_users_local, n_items_local, world, seed, torch):
    """links of ONE stratified block with LOCAL ids: the rank's shard of a power-law id space is itself (about) a power
    law over its local rows, p(j) ∝ (j + 10 / N)^-a (the ids of shard s are every N-th rank of a random permutation)"""
    g = torch.Generator(device="cuda").manual_seed(seed)

    def draw(n_ids, expo, perm_seed, n):
        p = torch.pow(torch.arange(n_ids, device="cuda", dtype=torch.float64) + 10.0 / world, -expo)
        cdf = torch.cumsum(p, 0)
        cdf = cdf / cdf[-1]
        r = torch.searchsorted(cdf, torch.rand(n, device="cuda", generator=g, dtype=torch.float64), right=True)
        r.clamp_(max=n_ids - 1)
        perm = torch.randperm(n_ids, device="cuda", generator=torch.Generator(device="cuda").manual_seed(perm_seed))
        return perm[r].to(torch.int32)

    uid = torch.empty(n_links, dtype=torch.int32, device="cuda")
    cid = torch.empty(n_links, dtype=torch.int32, device="cuda")
    chunk = 20_000_000
    for s in range(0, n_links, chunk):
        n = min(chunk, n_links - s)
        uid[s:s + n] = draw(n_users_local, 0.8, 124, n)
        cid[s:s + n] = draw(n_items_local, 1.0, 123, n)
    return uid, cid


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        # NVML directly (nvidia_ml_py): a sample costs ~0.1 ms, so even a 50 ms timed region gets dozens of samples;
        # rows have the layout of the nvidia-smi query below, which is the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                    ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                    ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                    ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))]
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = reasons_fn(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), "%.1f" % pw] + ["Active" if (r & b) else "Not Active" for _, b in bits])
                time.sleep(0.002)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in self.rows if len(r) >= 7)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's path, NumPy transcription (oracle/), all host threads BLAS can use
# ----------------------------------------------------------------------------------------------------------------
def cpu_links_per_sec(n_steps, warmup, n_rows=N_USERS):
    from oracle import nncf_oracle as O
    rng = np.random.RandomState(7)
    EU = rng.uniform(-0.05, 0.05, size=(n_rows, DIM)).astype(np.float32)
    EV = rng.uniform(-0.05, 0.05, size=(n_rows, DIM)).astype(np.float32)
    from nncf_b200.data_utils import powerlaw_ids
    n = (n_steps + warmup) * BATCH
    uid = powerlaw_ids(n_rows, n, 0.8, 124, rng).astype(np.int64)
    cid = powerlaw_ids(n_rows, n, 1.0, 123, rng).astype(np.int64)
    for s in range(warmup):
        O.baseline_neg_shared_sgd_step(EU, EV, uid[s * BATCH:(s + 1) * BATCH], cid[s * BATCH:(s + 1) * BATCH], LAMBDA, LR)
    t0 = time.perf_counter()
    for s in range(warmup, warmup + n_steps):
        O.baseline_neg_shared_sgd_step(EU, EV, uid[s * BATCH:(s + 1) * BATCH], cid[s * BATCH:(s + 1) * BATCH], LAMBDA, LR)
    dt = time.perf_counter() - t0
    return n_steps * BATCH / dt, dt


_OUT = sys.stdout


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20000))
    v, dt = cpu_links_per_sec(steps, max(3, min(args.warmup, 50)))
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "positive links/sec train (neg_shared)", "value": v, "unit": "links/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_size_p": BATCH, "dim": DIM, "optimizer": "sparse SGD",
                   "note": "reference CPU path: Keras 1.2.2 / TF 1.0 / py2 cannot be installed offline -> NumPy "
                           "transcription of the same step (oracle/nncf_oracle.py baseline_neg_shared_sgd_step), one "
                           "batch of 512 links per step, strictly sequential"},
        "cpu_baseline": {"value": v, "unit": "links/s", "cores": cores, "kind": "port",
                         "sample": "%d sequential neg_shared steps of 512 links on 1M x 128 fp32 tables" % steps},
        "e2e": {"value": v, "unit": "links/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nncf_b200 import ops
    from nncf_b200.ops import FusedStep, StepSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = _peaks()
    R, B, d = args.replicas, BATCH, DIM
    links_per_step = R * B

    # ---- resident state.  N = 1: the two 1M x 128 tables live on the GPU.  N > 1 (weak scaling: per-GPU work fixed, every
    #      rank trains on its own links of the SAME global 1M x 1M problem):
    #        stratified (default): DSGD schedule, tables sharded by row, rank r owns user shard r and the item shards rotate
    #          round the ring between sub-epochs; every step touches local rows only (nncf_b200/parallel.py)
    #        peer: rows read and updated in the owners' shards over NVLink peer memory inside the step kernels
    #      Links are shuffled once on the device; np.random.shuffle semantics are exercised by the parity tests, the
    #      order itself is not part of the timed hot path.
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd",
                    learn_rate=LR, replicas=R, neg_loss_weight=LAMBDA)
    sharded = strat = None
    n_links = args.links
    loss_buf = torch.empty(max(args.steps, args.warmup, 1) * R, dtype=torch.float32, device="cuda")
    pos = {"step": 0, "rot": 0}
    if world > 1 and args.parallelism == "stratified":
        from nncf_b200.parallel import StratifiedTrainer, shard_rows
        strat = StratifiedTrainer(spec, N_USERS, N_ITEMS, rank, world, seed=7)
        step = strat.step
        links_per_block = n_links // world
        blocks = [synth_block_device(links_per_block, strat.rows_u, shard_rows(N_ITEMS, v, world), world,
                                     2017 + rank * world + v, torch) for v in range(world)]
        steps_per_pass = links_per_block // links_per_step          # steps until the item shards rotate
        assert steps_per_pass >= 1
        tables = lambda: (strat.users, strat.items)                 # noqa: E731  (the item tensor changes at every rotation)
        ids_now = lambda: blocks[strat.held]                        # noqa: E731

        def end_of_pass():
            strat.rotate()
            pos["rot"] += 1
    else:
        if world > 1:
            from nncf_b200.parallel import ShardedTrainer
            sharded = ShardedTrainer(spec, N_USERS, N_ITEMS, rank, world, seed=7)
            step, EU, EV = sharded.step, sharded.users.local, sharded.items.local
        else:
            EU = (torch.rand((N_USERS, d), device="cuda", generator=g) - 0.5) * 0.1
            EV = (torch.rand((N_ITEMS, d), device="cuda", generator=g) - 0.5) * 0.1
            step = FusedStep(spec)
        train = synth_links_device(n_links, N_USERS, N_ITEMS, 2017 + rank, torch)
        perm = torch.randperm(n_links, device="cuda", generator=g)
        train = ops.permute_rows(train, perm)
        uid_all = train[:, 0].contiguous()
        cid_all = train[:, 1].contiguous()
        del train, perm
        steps_per_pass = n_links // links_per_step
        assert steps_per_pass >= 1
        tables = lambda: (EU, EV)                                   # noqa: E731
        ids_now = lambda: (uid_all, cid_all)                        # noqa: E731

        def end_of_pass():
            pass

    def run_steps(k, start_step=None):
        """k consecutive steps from the current position (wraps around the link array / moves to the next stratified
        block, rotating the item shards, when a pass is exhausted)"""
        done = 0
        while done < k:
            n = min(k - done, steps_per_pass - pos["step"])
            off = pos["step"] * links_per_step
            u, c = ids_now()
            tu, tv = tables()
            step.run(tu, tv, u[off:], c[off:], n, loss_out=loss_buf[done * R:])
            done += n
            pos["step"] += n
            if pos["step"] == steps_per_pass:
                pos["step"] = 0
                end_of_pass()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- value: device-resident throughput ------------------------------------------------------------------
    run_steps(args.warmup, 0)
    if strat is not None:
        end_of_pass()            # untimed: the first send/recv sets up the NCCL peer-to-peer channels (~0.3 s)
        pos["step"] = 0
        run_steps(min(20, steps_per_pass))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rot0 = pos["rot"]
    e0.record()
    run_steps(args.steps, args.warmup)
    if strat is not None and pos["rot"] - rot0 < -(-args.steps // steps_per_pass):
        # the timed window must carry its share of rotations even when it ends inside a block (rounded UP)
        pos["step"] = 0
        end_of_pass()
    e1.record()
    barrier()
    rotations = pos["rot"] - rot0
    launches = ops.launch_count() - launches0
    sampler.stop_flag = True
    ms = max_over_ranks(e0.elapsed_time(e1))
    sampler.join(timeout=2)
    final_loss = float(loss_buf[(args.steps - 1) * R:(args.steps) * R].mean().item())
    assert np.isfinite(final_loss), "training diverged"
    value = world * args.steps * links_per_step / (ms * 1e-3)

    # ---- roofline of the dominant kernel: CUDA events around the score+gradient kernel, on its stream (every rank
    #      runs the same number of profiled steps so that the device barriers of the sharded mode pair up)
    step.set_profile(True)
    nprof = min(200, steps_per_pass)
    pos["step"] = 0
    run_steps(nprof)
    torch.cuda.synchronize()
    phase_ms, psteps = step.get_profile()
    step.set_profile(False)
    gather_ms, score_ms, final_ms = [x / max(psteps, 1) for x in phase_ms]

    # ---- e2e: the public host-fed call (FusedStep.run_host -> nncf_train_steps_host): link ids in pinned HOST memory,
    #      every step copies its own ids H2D (overlapping the previous step's kernels) and its R losses D2H; wall clock
    #      around the call, which returns only when every step and copy has completed.  `per_call` is the same work issued
    #      as one blocking train_on_batch-style call per step (H2D, step, loss.cpu()) from Python.
    e2e_steps = max(1, min(args.steps, 1000, steps_per_pass - 5))
    h_uid = torch.empty((e2e_steps + 5, links_per_step), dtype=torch.int32).pin_memory()
    h_cid = torch.empty((e2e_steps + 5, links_per_step), dtype=torch.int32).pin_memory()
    uid_all, cid_all = ids_now()
    EU, EV = tables()
    h_uid.copy_(uid_all[:h_uid.numel()].view(h_uid.shape).cpu())
    h_cid.copy_(cid_all[:h_cid.numel()].view(h_cid.shape).cpu())
    h_loss = torch.empty((e2e_steps + 5) * R, dtype=torch.float32).pin_memory()
    step.run_host(EU, EV, h_uid, h_cid, 5, h_loss)
    barrier()
    t0 = time.perf_counter()
    step.run_host(EU, EV, h_uid[5:], h_cid[5:], e2e_steps, h_loss)
    if strat is not None:
        # e2e carries the rotations of its window too (rounded up to one), synchronised like the step calls
        end_of_pass()
        torch.cuda.synchronize()
        EU, EV = tables()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * e2e_steps * links_per_step / e2e_s
    assert np.isfinite(float(h_loss[:e2e_steps * R].mean())), "e2e training diverged"
    barrier()
    d_uid = torch.empty(links_per_step, dtype=torch.int32, device="cuda")
    d_cid = torch.empty(links_per_step, dtype=torch.int32, device="cuda")

    def e2e_step(i):
        d_uid.copy_(h_uid[i], non_blocking=True)
        d_cid.copy_(h_cid[i], non_blocking=True)
        out = step.run(EU, EV, d_uid, d_cid, 1)
        return out["loss"].cpu()            # D2H + sync: the python-float loss Keras' train_on_batch returns

    per_call_steps = min(e2e_steps, 300)
    for i in range(5):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(5, 5 + per_call_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    per_call_s = max_over_ranks(time.perf_counter() - t0)
    per_call_value = world * per_call_steps * links_per_step / per_call_s
    barrier()

    # ---- extra: whole@k users/sec, users sharded over the ranks (C4 shape: every rank scores its own users against all
    #      2M items, k = 50; no data-path collective, the metric sums would be all-reduced).  Time = max over ranks.
    extra = {}
    if not args.no_eval:
        n_eval_users, n_eval_items, k = args.eval_users, 2_000_000, 50
        ge = torch.Generator(device="cuda").manual_seed(99)
        Ve = torch.randn((n_eval_items, d), device="cuda", generator=ge) / d ** 0.5            # replicated item table
        Ue = torch.randn((n_eval_users, d), device="cuda", generator=g) / d ** 0.5             # this rank's user shard
        ops.eval_topk(Ue[:1024], Ve, k, "bf16")
        barrier()
        e0.record()
        ids, _ = ops.eval_topk(Ue, Ve, k, "bf16")
        e1.record()
        barrier()
        ems = max_over_ranks(e0.elapsed_time(e1))
        etf = 2.0 * world * n_eval_users * n_eval_items * d / (ems * 1e-3) / 1e12
        extra = {"whole_at_k": {"users_per_sec": world * n_eval_users / (ems * 1e-3), "k": k, "users": world * n_eval_users,
                                "items": n_eval_items, "dim": d, "ms": ems, "sharding": "users over %d GPU(s)" % world,
                                "roofline": {"bound": "tensor", "achieved": etf, "peak": world * peaks["bf16_burst"],
                                             "unit": "TFLOP/s", "frac": etf / (world * peaks["bf16_burst"])}}}
        # the other two k of the C4 configuration (same users, same 2M items), one timed call each
        by_k = {}
        for kk in (10, 100):
            ops.eval_topk(Ue[:1024], Ve, kk, "bf16")
            barrier()
            e0.record()
            ops.eval_topk(Ue, Ve, kk, "bf16")
            e1.record()
            barrier()
            kms = max_over_ranks(e0.elapsed_time(e1))
            by_k[str(kk)] = {"users_per_sec": world * n_eval_users / (kms * 1e-3), "ms": kms,
                             "tflops": 2.0 * world * n_eval_users * n_eval_items * d / (kms * 1e-3) / 1e12}
        extra["whole_at_k"]["other_k"] = by_k
        del Ue, Ve, ids

    if rank != 0:
        if sharded is not None:
            dist.barrier()
            sharded.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (same R, B, d)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01b_traffic.json")))
        if R == 37:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    flops = 6.0 * B * B * d * R
    achieved_tf = flops / (score_ms * 1e-3) / 1e12
    step_bytes = links_per_step * (8 + 16 * d)
    step_s = ms / args.steps * 1e-3
    roofline = {"bound": "tensor", "kernel": "score_grad_tc_kernel<2,skip-gram,neg_shared>", "achieved": achieved_tf,
                "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["bf16_burst"], "traffic": traffic,
                "peak_source": peaks["src"], "ms_per_launch": score_ms, "flops_per_launch": flops,
                "note": "algorithmic 6*B^2*d*R flops per launch; the one-sided kernel executes 8*B^2*dp*R (S is recomputed "
                        "by the item side); its epilogue is MUFU-bound (DESIGN.md 3.1)",
                "phases_ms": {"gather_prepare": gather_ms, "score_grad": score_ms, "finalize_update": final_ms},
                "step_hbm": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / step_s / 1e9,
                             "peak_gbs": peaks["hbm"], "frac": step_bytes / step_s / 1e9 / peaks["hbm"]}}

    seq_info, cpu = None, None
    if world == 1:
        # ---- sequential reference semantics (R = 1) ----------------------------------------------------------
        seq = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d,
                                 optimizer="sgd", learn_rate=LR, replicas=1, neg_loss_weight=LAMBDA))
        nseq = min(2000, n_links // B)
        seq.run(EU, EV, uid_all, cid_all, 50)
        torch.cuda.synchronize()
        e0.record()
        seq.run(EU, EV, uid_all, cid_all, nseq)
        e1.record()
        torch.cuda.synchronize()
        seq_info = {"value": nseq * B / (e0.elapsed_time(e1) * 1e-3), "unit": "links/s", "replicas_per_gpu": 1, "steps": nseq}
        # ---- the reference's optimizer family: sparse (lazy) Adam on the gathered rows, same R / B / d ------------
        try:
            adam = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d,
                                      optimizer="lazy_adam", learn_rate=0.001, replicas=R, neg_loss_weight=LAMBDA))
            state = [torch.zeros_like(EU), torch.zeros_like(EU), torch.zeros_like(EV), torch.zeros_like(EV)]
            nad = min(600, n_links // (R * B))
            adam.run(EU, EV, uid_all, cid_all, 30, adam_state=state)
            torch.cuda.synchronize()
            e0.record()
            adam.run(EU, EV, uid_all, cid_all, nad, adam_state=state)
            e1.record()
            torch.cuda.synchronize()
            ams = e0.elapsed_time(e1) / nad
            seq_info["lazy_adam"] = {"value": R * B / (ams * 1e-3), "unit": "links/s", "replicas_per_gpu": R, "ms_per_step": ams,
                                     "hbm_frac": R * B * (8 + 48 * d) / (ams * 1e-3) / 1e9 / peaks["hbm"],
                                     "note": "lazy Adam (utils/optimizer.py _apply_sparse rule), algorithmic 8 + 48 d bytes per link"}
            del state, adam
        except Exception as ex:      # an extra must never cost the headline line
            seq_info["lazy_adam"] = {"error": str(ex)[:200]}
        # ---- the other batch sizes SURVEY.md 8 names for C3 (same tables, same links; fewer replicas so that a step still
        #      fills the 148 SMs once) and the tensor-bound C5 shape are reported beside the headline -----------------------
        sweep = {}
        for Bx, Rx in ((4096, 5), (8192, 2)):
            try:
                sx = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=Bx, dim=d,
                                        optimizer="sgd", learn_rate=LR, replicas=Rx, neg_loss_weight=LAMBDA))
                nx = min(300, n_links // (Rx * Bx) - 12)
                sx.run(EU, EV, uid_all, cid_all, 10)
                torch.cuda.synchronize()
                e0.record()
                sx.run(EU, EV, uid_all, cid_all, nx)
                e1.record()
                torch.cuda.synchronize()
                xms = e0.elapsed_time(e1) / nx
                sweep[str(Bx)] = {"value": Rx * Bx / (xms * 1e-3), "unit": "links/s", "replicas_per_gpu": Rx, "ms_per_step": xms,
                                  "tflops_algorithmic": 6.0 * Bx * Bx * d * Rx / (xms * 1e-3) / 1e12}
                del sx
            except Exception as ex:
                sweep[str(Bx)] = {"error": str(ex)[:200]}
        seq_info["batch_size_sweep"] = sweep
        # ---- BASELINE config 5: neg_shared max-margin, batch 16,384, dim 256, l2-normalised rows (tensor-bound contraction)
        try:
            g5 = torch.Generator(device="cuda").manual_seed(55)
            U5 = (torch.rand((1_000_000, 256), device="cuda", generator=g5) - 0.5) * 0.1
            V5 = (torch.rand((1_000_000, 256), device="cuda", generator=g5) - 0.5) * 0.1
            n5 = 40
            u5 = torch.randint(0, 1_000_000, ((n5 + 5) * 16384,), device="cuda", generator=g5, dtype=torch.int32)
            c5 = torch.randint(0, 1_000_000, ((n5 + 5) * 16384,), device="cuda", generator=g5, dtype=torch.int32)
            s5 = FusedStep(StepSpec(scheme="neg_shared", loss="max-margin", precision="bf16", batch_size_p=16384, dim=256, norm_u=True,
                                    norm_v=True, optimizer="sgd", learn_rate=LR, replicas=1, neg_loss_weight=LAMBDA, loss_gamma=0.1))
            s5.run(U5, V5, u5, c5, 5)
            torch.cuda.synchronize()
            e0.record()
            s5.run(U5, V5, u5[5 * 16384:], c5[5 * 16384:], n5)
            e1.record()
            torch.cuda.synchronize()
            ms5 = e0.elapsed_time(e1) / n5
            tf5 = 6.0 * 16384 * 16384 * 256 / (ms5 * 1e-3) / 1e12
            seq_info["c5_max_margin_b16384_d256"] = {"value": 16384 / (ms5 * 1e-3), "unit": "links/s", "ms_per_step": ms5,
                                                     "roofline": {"bound": "tensor", "achieved": tf5, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                                                                  "frac": tf5 / peaks["bf16_burst"],
                                                                  "note": "algorithmic 6 B^2 d; the one-sided kernel executes 8 B^2 d"}}
            del U5, V5, u5, c5, s5
        except Exception as ex:
            seq_info["c5_max_margin_b16384_d256"] = {"error": str(ex)[:200]}
        # ---- cpu baseline: bounded sample on the box's host cores (rank 0, N = 1 only) ------------------------
        cpu_v, cpu_dt = cpu_links_per_sec(args.cpu_steps, 5)
        cpu = {"value": cpu_v, "unit": "links/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d sequential neg_shared steps of 512 links (%.1f s) on 1M x 128 fp32 tables, NumPy/BLAS" % (args.cpu_steps, cpu_dt)}

    line = {
        "metric": "positive links/sec train (neg_shared)", "value": value, "unit": "links/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_size_p": B, "dim": d, "replicas_per_gpu": R, "links_per_step_per_gpu": links_per_step,
                   "links_resident_per_gpu": n_links, "optimizer": "sparse SGD (atomic scatter-add)",
                   "precision": "bf16 operands, fp32 accumulate (tcgen05)",
                   "parallelism": "1 GPU" if world == 1 else (
                       "stratified SGD (DSGD) over %d GPUs: tables row-sharded (owner = id mod N), rank r owns user shard r, item "
                       "shards rotate round the ring (NCCL send/recv) between sub-epochs of %d steps, every step touches local "
                       "rows only; %d rotation(s) inside the timed region" % (world, steps_per_pass, rotations)
                       if strat is not None else
                       "tables row-sharded over %d GPUs (owner = id mod N), rows read and updated over NVLink peer memory "
                       "inside the step kernels, 2 device barriers per step" % world),
                   "semantics": "each step = R independent neg_shared batches per GPU against one table snapshot (synchronous "
                                "data-parallel virtual workers); R=1 (the reference's sequential loop) is reported in `sequential`",
                   "l2": "inputs larger than L2: 1.02 GB of embedding tables, random rows, batches never repeat within a pass"},
        "sequential": seq_info,
        "final_loss": final_loss,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "links/s", "h2d_bytes_per_step": 2 * 4 * links_per_step,
                "d2h_bytes_per_step": 4 * R, "steps": e2e_steps,
                "per_call": {"value": per_call_value, "unit": "links/s", "steps": per_call_steps,
                             "note": "one blocking train_on_batch-style Python call per step (H2D, step, loss.cpu())"},
                "note": "FusedStep.run_host / nncf_train_steps_host: per step and per GPU, that step's ids are copied from "
                        "pinned host memory (overlapping the previous step's kernels) and its R losses are copied back; "
                        "wall clock around the call, which returns after the last copy"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "extra": extra,
    }
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if sharded is not None:
        dist.barrier()
        sharded.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=37)  # 37 x 8 CTAs = 2 full waves of 148 SMs
    ap.add_argument("--links", type=int, default=N_LINKS)
    ap.add_argument("--cpu-steps", type=int, default=3000)
    ap.add_argument("--eval-users", type=int, default=75776)   # 4 full waves of 148 CTAs x 128 users
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--parallelism", default="stratified", choices=["stratified", "peer"])   # N > 1 only
    args = ap.parse_args()
    # stdout carries ONE JSON line and nothing else: libraries that print to the process's stdout (NCCL's "NCCL version ..."
    # banner under torchrun, for one) are sent to stderr for the whole run, the line goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

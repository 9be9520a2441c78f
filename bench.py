#!/usr/bin/env python
"""bench.py -- JGHM trees/sec (sample + full BP posterior) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-eval 65536]

Workload (BASELINE.json configs[1]): the CLIP-risk evaluation shape.
ClipSampler([4,4],[3,3], p_flip=[.2,.2], K=4, q=10), n_eval = 65 536 matched pairs per GPU
-> 5*65 536 text trees + 5*65 536 image trees = 655 360 trees per step per GPU.  One step =
sample both modalities (Philox), materialise the int64 leaves [B, 81], fused root-posterior BP,
symmetric K-way Bayes contrastive reduction -> {sum, sumsq, count}; N > 1: every rank owns its
own 65 536 pairs (weak scaling; Philox tree offsets = rank * B) and the 24-byte sums are
all-reduced over NCCL each step.

`value`  : device-resident trees/s (CUDA events around exactly K steps, max over ranks).  N > 1: the 24-byte
           all-reduce of step k runs on NCCL's stream and overlaps the kernels of the following steps (ring of
           accumulators); every all-reduce has joined the compute stream before the closing event.
`e2e`    : one step = one grid point of the reference's p_flip sweep (figures/eval-clip-ood.py:73-79): new float64
           transition matrices come from the host, derived tables are built in pinned memory and uploaded
           (ghm_model_update), then ClipSampler.get_Bayes(n_eval) returns two host floats (24-byte D2H read).
`e2e_get_batch`: get_Bayes plus what ClipSampler.get_batch returns (int64 leaves and f32 posteriors of both
           modalities) copied into pinned host memory through the host-buffer C entry point (PCIe bound).
`roofline`: dominant kernel = fused sampler+BP k_tree2; algorithmic bytes = int64 leaves + roots +
           f32 posterior per tree (K1 656 B + K2 40 B out), timed with CUDA events per launch.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py)
           called through its own ClipSampler.get_Bayes, one process per host core (its own `&` fan-out); the NumPy
           oracle port (oracle/ghm_oracle.py) is timed next to it (`port_ncore`) and replaces it only when
           baseline/_ref is absent (`kind: "port"`).
`strong_scaling`: BASELINE configs[4]: 1 048 576 paired trees (sample both modalities + text BP_CLS + image BP_DNS with
           the text root message + risk) SPLIT over the N ranks through ConditionalDenoiseSampler.get_Bayes(
           distributed=True) with one NCCL all-reduce per evaluation; q = 10 (FP32 CUDA cores) and q = 256 (tcgen05 TF32).
`configs`: (N = 1) driver-visible numbers of the other BASELINE configs (C1, C3, C4, C5), each next to the reference /
           port on the same host cores.
"""
import os

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import argparse
import json
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-ghm_b200"))

N_LAYERS, N_CHILDS, P_FLIPS, K_CLIP, Q = [4, 4], [3, 3], [0.2, 0.2], 4, 10
METRIC = "JGHM trees/sec (sample + full BP posterior)"
N_SM = 148
# Figures of the dominant kernel taken from its committed `ncu --set full` capture (per launch of 327 680 trees)
KTREE = {
    "kernel": "k_tree_fast<Q=10,S=3,PHILOX,BP> (fused sampler + root-posterior BP: leaf memo, one tree per thread at 7 CTAs/SM, "
              "int64 leaves out through cp.async.bulk)",
    "source": "profiles/r02_ncu_full_k_tree_fast.csv",
    "dram_bytes_per_tree": 518.4,          # 169.62 MB written + 0.23 MB read per launch (below the 696 B/tree algorithmic
                                           # figure: the tail of the leaves is still in the 126 MB L2 when the kernel ends)
    "warp_inst_per_tree": 97016320 / 327680,
    "share": 0.95,                         # of the GPU time in the serialised ncu launch list of this script (profiles/r02_launches_bench_clip.csv)
    "note": "issue slots 0.71 at 7 warps per scheduler (70 registers, 30.8 KB), ALU pipe 0.39, FMA-heavy pipe 0.33 (Philox "
            "IMAD.WIDE 4.3 clk per warp instruction), uniform datapath 0.22, L1 0.77; 31 % of the stall samples sit in the "
            "int64 flush at the end of a tile (bulk copies draining at HBM write speed while every warp of the wave flushes), "
            "10 % in the dependent chain of the last climb; the leaf memo removed 27 of the 40 matvecs of a tree (107 -> 97 M "
            "warp instructions per launch at twice the warps); the HBM fraction is reported, not padded; see roofline.secondary "
            "and DESIGN.md 3.1",
}


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref) or, when it is absent, the oracle port -- all host cores
# --------------------------------------------------------------------------------------
def _cpu_arms():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import cpu_arms
    return cpu_arms


def run_reference_arm(args, rank, world):
    """`--impl reference`: time the reference's CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import multiprocessing as mp
    A = _cpu_arms()
    cores = A.host_cores()
    impl = "reference" if A.reference_available() else "port"
    n_eval_core = args.cpu_n_eval if args.cpu_n_eval else (5000 if impl == "reference" else 10000)
    with mp.get_context("fork").Pool(cores) as pool:
        for w in range(min(args.warmup, 1)):
            A.fan_out(pool, cores, "c2_clip", impl, 400, 1000 + w)
        t_total, trees_total = 0.0, 0
        for k in range(args.steps):
            r = A.fan_out(pool, cores, "c2_clip", impl, n_eval_core, 2000 + 1000 * k)
            t_total += r["seconds"]
            trees_total += r["trees"]
    value = trees_total / t_total
    what = ("the unmodified reference (baseline/_ref: ghmclip.data.data_random_GHM.ClipSampler)" if impl == "reference"
            else "oracle port (oracle/ghm_oracle.py; baseline/_ref absent)")
    sample = "each step: %d cores x ClipSampler.get_Bayes(n_eval=%d) = %d trees; %s, 1 BLAS thread per process" % (
        cores, n_eval_core, cores * n_eval_core * (K_CLIP + 1) * 2, what)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n_eval, 1),
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": impl, "sample": sample},
        "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_eval, world):
    return {"workload": "C2 CLIP-risk eval shape: ClipSampler(L=[4,4],s=[3,3],p=[.2,.2],K=4,q=10), "
                        "n_eval=%d pairs/GPU -> %d trees/GPU/step (sample both modalities + int64 leaves + "
                        "BP root posterior + Bayes contrastive reduction)" % (n_eval, n_eval * (K_CLIP + 1) * 2),
            "n_eval_per_gpu": n_eval, "trees_per_gpu_step": n_eval * (K_CLIP + 1) * 2, "parallelism": "dp%d" % world,
            "rng": "philox4x32-10", "leaf_dtype": "int64",
            "cache": "outputs (%.0f MB/step/GPU) exceed the 126 MB L2; inputs are the 11 KB transition tables (constant "
                     "bank + shared memory), nothing is re-read between steps" % (n_eval * (K_CLIP + 1) * 2 * 81 * 8 / 1e6)}


# --------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ghm_b200 import ops
    from ghm_b200.data_random_GHM import ClipSampler

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's 24-byte all-reduce must not be starved by the compute kernels that already fill every SM: give its
        # stream priority so its single CTA is placed as soon as one of ours retires
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        # stdout carries exactly one JSON line: the communicator banner ("NCCL version ...", written to fd 1 from native
        # code) is sent to stderr by pointing fd 1 there while the communicator is created
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    u = np.ones(Q) / Q
    sampler = ClipSampler(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, K=K_CLIP, variable_type=Q, device=dev, rng="philox",
                          seed=1234)
    tm, im = sampler.t_model, sampler.i_model
    n, K = args.n_eval, K_CLIP
    B = n * (K + 1)
    nLt, nLi = tm.n_leaves, im.n_leaves
    trees_step = 2 * B
    tree_off = rank * B                                     # weak scaling: disjoint global tree ranges
    # Two software pipelines (own streams + own output buffers): consecutive steps are independent risk evaluations,
    # so step k+1 starts sampling while the tail waves of step k drain -- inside a step the text and image launches
    # already run on two streams (the image kernel re-draws the shared roots from the text key).
    class Pipe:
        def __init__(self):
            self.st = torch.cuda.Stream(device=dev)
            self.si = torch.cuda.Stream(device=dev)
            self.t_leaves = torch.empty((B, nLt), dtype=torch.int64, device=dev)
            self.i_leaves = torch.empty((B, nLi), dtype=torch.int64, device=dev)
            self.t_root = torch.empty(B, dtype=torch.int64, device=dev)
            self.t_pp = torch.empty((B, Q), dtype=torch.float32, device=dev)
            self.i_pp = torch.empty((B, Q), dtype=torch.float32, device=dev)
            self.ev = torch.cuda.Event()
    pipes = [Pipe(), Pipe()]                                # deeper pipelines measured no faster (3, 4: same 0.30 ms/step)
    # Risk accumulators: two banks of RING slots; one all-reduce per RING steps (RING x 24 bytes) on NCCL's
    # high-priority stream while the other bank is being filled.
    RING = 8
    banks = [torch.zeros((RING, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    bank_work = [None, None]
    step_no = [0]
    cur = torch.cuda.current_stream(dev)

    def flush_bank(bk):
        for pp_ in pipes:
            cur.wait_stream(pp_.st)
        if world > 1:
            bank_work[bk] = dist.all_reduce(banks[bk], async_op=True)

    def step(seed, reduce=True):
        """sample text / image (+ fused BP, leaves materialised) + contrastive reduction; every RING steps one all-reduce."""
        k = step_no[0]
        step_no[0] += 1
        bk, slot = (k // RING) % 2, k % RING
        pipe = pipes[k % len(pipes)]
        if slot == 0:
            if bank_work[bk] is not None:
                bank_work[bk].wait()                     # the all-reduce that last used this bank (16 steps ago) is done
                bank_work[bk] = None
            for pp_ in pipes:
                pp_.st.wait_stream(cur)
            with torch.cuda.stream(pipes[0].st):
                banks[bk].zero_()
            for pp_ in pipes[1:]:
                pp_.st.wait_stream(pipes[0].st)
        sums = banks[bk][slot]
        pipe.si.wait_stream(pipe.st)
        with torch.cuda.stream(pipe.si):
            ops.sample_paired_into(im, B, 2 * n, seed, seed ^ ops.IMAGE_SEED_XOR, tree_off, None, pipe.i_leaves, pipe.i_pp, None)
        with torch.cuda.stream(pipe.st):
            ops.sample_into(tm, B, ops.ROOT_UNIFORM, None, seed, tree_off, pipe.t_root, pipe.t_leaves, pipe.t_pp, None)
            pipe.st.wait_stream(pipe.si)
            ops.risk_clip(pipe.t_pp, pipe.i_pp, n, K, Q, sums=sums)
        if slot == RING - 1 and reduce:
            flush_bank(bk)
        return sums

    def drain():
        k = step_no[0]
        if k % RING != 0:
            flush_bank((k // RING) % 2)                  # partially filled bank
            step_no[0] = (k // RING + 1) * RING
        for pp_ in pipes:
            cur.wait_stream(pp_.st)
            cur.wait_stream(pp_.si)
        for i, w in enumerate(bank_work):
            if w is not None:
                w.wait()
                bank_work[i] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(100 + w)
    drain()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                                   # (before the barrier: spawning nvidia-smi takes milliseconds and
    barrier()                                            #  would skew rank 0 against the others inside the timed region)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    t_issue = time.perf_counter()
    for k in range(args.steps):
        last = step(1000 + k)
    host_ms_per_step = 1e3 * (time.perf_counter() - t_issue) / args.steps    # host time to ENQUEUE a step (no sync inside)
    drain()                                              # every all-reduce has joined the compute stream before e1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if os.environ.get("GHM_BENCH_DEBUG"):
        sys.stderr.write("rank %d: %.4f ms/step device, %.4f ms/step host issue\n" % (rank, ms / args.steps, host_ms_per_step))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    risk_mean, risk_se = ops.mean_se(last)
    # keep the GPU busy while the clock sampler gets enough samples if the run was very short
    if rank == 0 and ms < 1500:
        t_end = time.time() + 1.5
        while time.time() < t_end:
            step(5000, reduce=False)                     # rank 0 only: no collective in here
        step_no[0] = (step_no[0] // RING + 1) * RING
        for pp_ in pipes:
            cur.wait_stream(pp_.st)
            cur.wait_stream(pp_.si)
        torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    value = world * trees_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (fused sampler + BP) ------------------------------------
    # k_tree2 launches of the two modalities and of consecutive steps overlap on four streams, so a per-launch event
    # interval would count its neighbours: the dominant kernel's throughput is taken over the whole timed region
    # (all 2K launches; k_tree2 is 95 % of the GPU time in the serialised ncu launch list under profiles/).
    launch_bytes = args.steps * (B * (8 * nLt + 4 * Q + 8) + B * (8 * nLi + 4 * Q))
    peak, peak_kind = read_peaks()
    achieved = launch_bytes / (ms * 1e-3) / 1e9
    sm_hz = 1e6 * float((clk or {}).get("sm_mhz") or 1965.0)
    issue_peak = N_SM * 4 * sm_hz                                        # warp-instructions / s the schedulers can issue
    k_trees_s = trees_step * args.steps / (ms * 1e-3)                    # this rank's trees/s
    roofline = {"bound": "hbm", "kernel": KTREE["kernel"] + "; the text and image launches of a step run concurrently on "
                                          "two streams and are timed as one unit",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": KTREE["dram_bytes_per_tree"] * B, "traffic_source": KTREE["source"],
                "bytes_per_tree": 8 * nLt + 4 * Q + 8, "trees_per_launch": B, "launches": 2 * args.steps,
                "avg_launch_ms": ms / (2 * args.steps), "kernel_share_of_gpu_time_ncu": KTREE["share"],
                # what actually binds the kernel: issue slots (warp-instructions per tree from the ncu capture x trees/s
                # over 4 schedulers x 148 SMs x the SM clock sampled during the run)
                "secondary": {"bound": "issue", "achieved": k_trees_s * KTREE["warp_inst_per_tree"], "peak": issue_peak,
                              "unit": "warp-inst/s", "frac": k_trees_s * KTREE["warp_inst_per_tree"] / issue_peak,
                              "warp_inst_per_tree": KTREE["warp_inst_per_tree"], "source": KTREE["source"]},
                "note": KTREE["note"]}

    # ---- end to end through the reference-facing facade call (host in / host out) --------------
    # One e2e step = one grid point of the reference's p_flip sweep (figures/eval-clip-ood.py:73-79): new
    # transition tables for BOTH modalities arrive from the host (float64 matrices -> derived tables in pinned
    # memory -> one H2D copy per modality), then get_Bayes(n_eval) -> two host floats (24-byte D2H read).
    # The NumPy draw of the matrices themselves (GenTransition, a sampler-construction one-off) is done ahead.
    # N > 1: ONE evaluation of world * n pairs sharded on the pair index, a real NCCL all-reduce of the sums per call.
    from ghm_b200.data_random_GHM import GenTransition
    grid = []
    for p in [0.02 * (i + 1) for i in range(20)]:
        np.random.seed(42)
        grid.append((p, GenTransition(N_LAYERS[0], N_CHILDS[0], Q, p, 1.0), GenTransition(N_LAYERS[1], N_CHILDS[1], Q, p, 1.0)))
    n_glob = n * world
    dist_kw = {"distributed": True} if world > 1 else {}

    def e2e_call(k, lazy):
        p, tt, it = grid[k % len(grid)]
        sampler.reparameterize([p, p], transitions=(tt, it))
        sampler.tree_offset = 0
        return sampler.get_Bayes(n_eval=n_glob, keep_batch=True, lazy=lazy, **dist_kw)

    e2e_steps = max(args.steps, 100)                         # a 20-call region is 6 ms: pipeline fill / drain would be 10 % of it

    def e2e_run(lazy):
        for w in range(max(6, args.warmup)):                # same mode as the timed loop (the lazy path allocates its pinned slots once)
            r = e2e_call(w, lazy)
            if lazy:
                r = r.result()
        barrier()
        t0 = time.perf_counter()
        pend = []
        for k in range(e2e_steps):
            h = e2e_call(k, lazy)
            if lazy:                                     # read evaluation k-2 while k-1 and k run: consecutive lazy evaluations
                pend.append(h)                           # share the GPU on two streams, so k-1 completes only shortly before k
                if len(pend) > 2:
                    r = pend.pop(0).result()
            else:
                r = h
        for h in pend:
            r = h.result()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        return world * trees_step * e2e_steps / el, r

    e2e_sync_v, r_sync = e2e_run(False)
    e2e_lazy_v, r_lazy = e2e_run(True)
    call = ("sampler.reparameterize(p_k) [host float64 transition matrices -> pinned derived tables -> H2D] + "
            "ClipSampler.get_Bayes(n_eval=%d%s, keep_batch=True%s) -> (mean, se) host floats; int64 leaves of the batch "
            "materialised on the device like `value`; p_k walks the 20-point p_flip grid"
            % (n_glob, ", distributed=True: pairs sharded over %d ranks, NCCL all-reduce per call" % world if world > 1 else "",
               ", lazy=True: evaluations alternate between two internal streams over double-buffered tables; the 24-byte result of "
               "call k is read after call k+2 has been enqueued"))
    e2e = {"value": e2e_lazy_v, "unit": "trees/s", "h2d_bytes_per_step": tm.table_bytes + im.table_bytes,
           "d2h_bytes_per_step": 24, "steps": e2e_steps, "call": call, "bayes_last": float(r_lazy[0]),
           "synchronous": {"value": e2e_sync_v, "note": "same call with lazy=False: every evaluation blocks on its own "
                                                        "24-byte read before the next one is enqueued"}}
    sampler.reparameterize(P_FLIPS, transitions=(grid[9][1], grid[9][2]))
    # variant that also brings the sampled batch back (what get_batch returns)
    tl = torch.empty((B, nLt), dtype=torch.int64).pin_memory()
    il = torch.empty((B, nLi), dtype=torch.int64).pin_memory()
    tp = torch.empty((B, Q), dtype=torch.float32).pin_memory()
    ip = torch.empty((B, Q), dtype=torch.float32).pin_memory()
    ops.host_clip_bayes(tm, im, n, K, seed=7, tree_offset=tree_off, leaves_out=(tl, il), pp_out=(tp, ip))
    barrier()
    t0 = time.perf_counter()
    reps = max(1, min(args.steps, 5))
    for k in range(reps):
        ops.host_clip_bayes(tm, im, n, K, seed=8 + k, tree_offset=tree_off, leaves_out=(tl, il), pp_out=(tp, ip))
    el = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([el], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    d2h = 24 + tl.numel() * 8 + il.numel() * 8 + tp.numel() * 4 + ip.numel() * 4
    e2e_b = {"value": world * trees_step * reps / el, "unit": "trees/s", "h2d_bytes_per_step": 0,
             "d2h_bytes_per_step": d2h,
             "call": "ghm_host_clip_bayes: as get_Bayes plus int64 leaves + f32 posteriors of both modalities "
                     "copied to pinned host memory (what ClipSampler.get_batch returns); N > 1: independent replicas"}
    del tl, il, tp, ip

    # ---- strong scaling: configs[4], 1 M paired trees split over the ranks ------------------------
    strong = None
    if not args.no_strong:
        strong = strong_scaling(args, rank, world, dev, barrier)

    # ---- the other BASELINE configs + CPU arms (rank 0, N = 1 only) --------------------------------
    cpu = configs = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import multiprocessing as mp
        A = _cpu_arms()
        cores = A.host_cores()
        with mp.get_context("fork").Pool(cores) as pool:
            A.fan_out(pool, cores, "c2_clip", "port", 300, 1)            # warm-up (imports, page-in)
            have_ref = A.reference_available()
            port = A.fan_out(pool, cores, "c2_clip", "port", 10000, 50, reps=3)
            cpu = {"value": port["trees_per_s"], "unit": "trees/s", "cores": cores, "kind": "port",
                   "sample": "%d cores x 3 x ClipSampler.get_Bayes(n_eval=10000) via oracle/ghm_oracle.py (NumPy port, 1 BLAS "
                             "thread per process) = %d trees in %.1f s" % (cores, port["trees"], port["seconds"])}
            if have_ref:
                A.fan_out(pool, cores, "c2_clip", "reference", 200, 2)
                ref_n = A.fan_out(pool, cores, "c2_clip", "reference", 5000, 60, reps=4)
                ref_1 = A.fan_out(pool, 1, "c2_clip", "reference", 5000, 61, reps=2)
                cpu = {"value": ref_n["trees_per_s"], "unit": "trees/s", "cores": cores, "kind": "reference",
                       "sample": "%d cores x 4 x the UNMODIFIED reference's ClipSampler.get_Bayes(n_eval=5000) (baseline/_ref, "
                                 "1 BLAS thread per process) = %d trees in %.1f s" % (cores, ref_n["trees"], ref_n["seconds"]),
                       "ref_ncore": ref_n["trees_per_s"], "ref_1core": ref_1["trees_per_s"],
                       "port_ncore": port["trees_per_s"]}
            if not args.no_configs:
                configs = config_blocks(dev, pool, cores, A, have_ref, peak)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms_per_step,
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n, world),
                "clocks": clk, "e2e": e2e, "e2e_get_batch": e2e_b, "gpu_launches": 3 * args.steps,
                "roofline": roofline, "cpu_baseline": cpu, "strong_scaling": strong, "configs": configs,
                "allreduce": "one NCCL all-reduce of the {sum, sumsq, n} slots per %d steps (bank of %d x 24 bytes) in `value`; "
                             "one per evaluation in `e2e` and `strong_scaling`" % (RING, RING),
                "bayes_clip_risk": {"mean": risk_mean, "se": risk_se}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# strong scaling (BASELINE configs[4]): 1 M paired trees split over the ranks
# --------------------------------------------------------------------------------------
def strong_scaling(args, rank, world, dev, barrier):
    """ConditionalDenoiseSampler([4,4],[3,3],p=.2,sigma=1).get_Bayes(n_eval = 1 048 576, distributed=True): every rank
    samples its slice of the GLOBAL pair index (Philox counter = global index, so any split draws the same trees), runs
    text BP_CLS (fused into the sampling launch at q <= 16) -> root message -> image BP_DNS -> risk sums, then ONE NCCL
    all-reduce of {sum, sumsq, n}.  Every rank then repeats the whole evaluation alone: the reduced sums must equal it
    (count exactly, sums to 1e-12 relative) and its time (max over ranks) gives speedup_vs_1 inside the same run."""
    import torch
    import torch.distributed as dist
    from ghm_b200.data_random_GHM import ConditionalDenoiseSampler
    out = {}
    for tag, q, n_pairs, gemm in (("q10_fp32", 10, args.strong_pairs, 0), ("q256_tf32", 256, args.strong_pairs_wide, 1)):
        if n_pairs <= 0:
            continue
        u = np.ones(q) / q
        s = ConditionalDenoiseSampler(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, sigma=1.0, variable_type=q, device=dev,
                                      rng="philox", seed=4321)
        s.t_model.set_gemm_mode(gemm)
        s.i_model.set_gemm_mode(gemm)
        if q > 16:
            s.bayes_chunk = 32768

        def evaluate(distributed):
            s.tree_offset = 0
            return s.get_Bayes(n_eval=n_pairs, distributed=distributed, lazy=True)

        def timed(distributed, reps):
            for _ in range(2):                                # two warm evaluations: the caching allocator settles its blocks
                evaluate(distributed).result()
            if distributed:
                barrier()
            else:
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                h = evaluate(distributed)
            e1.record()
            h.result()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps, h

        reps = 10 if q <= 16 else 2                          # back-to-back evaluations (lazy handles): steady-state time per evaluation
        if os.environ.get("GHM_BENCH_DEBUG"):
            st0 = torch.cuda.memory_stats(dev)
        ms, h = timed(world > 1, reps)
        if os.environ.get("GHM_BENCH_DEBUG"):
            st1 = torch.cuda.memory_stats(dev)
            sys.stderr.write("strong %s: %.3f ms; device allocs %d -> %d, frees %d -> %d, retries %d, reserved %.1f GB\n" % (
                tag, ms, st0["num_device_alloc"], st1["num_device_alloc"], st0["num_device_free"], st1["num_device_free"],
                st1["num_alloc_retries"], st1["reserved_bytes.all.current"] / 1e9))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        sums = h.sums()
        blk = {"pairs": n_pairs, "trees": 2 * n_pairs, "ms": ms, "trees_per_s": 2 * n_pairs / (ms * 1e-3),
               "risk_mean": sums[0] / sums[2], "count": sums[2],
               "op": "ConditionalDenoiseSampler(L=[4,4],s=[3,3],p=.2,sigma=1,q=%d).get_Bayes(n_eval=%d%s): sample text+image, "
                     "text BP_CLS -> root message -> image BP_DNS, risk sums%s"
                     % (q, n_pairs, ", distributed=True" if world > 1 else "",
                        ", one NCCL all-reduce per evaluation" if world > 1 else ""),
               "arithmetic": "fp32 CUDA cores" if gemm == 0 else "tcgen05 TF32 GEMMs (wide path)"}
        if world > 1:
            # every rank now evaluates ALL pairs alone (no collective; all ranks busy, so nobody spins in a barrier)
            barrier()
            ms1, h1 = timed(False, reps)
            s1 = h1.sums()
            ok = bool(s1[2] == sums[2] and abs(s1[0] - sums[0]) <= 1e-12 * abs(s1[0])
                      and abs(s1[1] - sums[1]) <= 1e-12 * abs(s1[1]))
            t = torch.tensor([ms1, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            blk["single_gpu_ms_same_run"] = float(t[0].item())
            blk["speedup_vs_1"] = float(t[0].item()) / ms
            blk["sums_equal_single_gpu"] = bool(t[1].item() == 0.0)
            blk["sum_rel_diff"] = abs(s1[0] - sums[0]) / abs(s1[0])
        out[tag] = blk
        del s
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------
# the other BASELINE configs (C1, C3, C4, C5) at N = 1, each next to the CPU arms
# --------------------------------------------------------------------------------------
def _gpu_ms(fn, reps, warm=2):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _wall_ms(fn, reps, warm=2):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(warm + i)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def config_blocks(dev, pool, cores, A, have_ref, hbm_peak):
    """Per-config numbers: trees/s on the GPU, the dominant kernel with its HBM and FP32 / MUFU fractions (per-tree
    algorithmic bytes, flops and transcendentals: SURVEY 8(d) / Appendix C), and the CPU arms on the same host."""
    import torch
    from ghm_b200 import ops, sweeps
    from ghm_b200 import data_random_GHM as G
    u = np.ones(Q) / Q
    fp32_peak = N_SM * 128 * 2 * 1.965e9                   # FFMA lanes x 2 flop x max SM clock
    mufu_peak = N_SM * 16 * 1.965e9

    def cpu_arm(task, size_ref, size_port, reps=4):
        r = {"port_ncore": A.fan_out(pool, cores, task, "port", size_port, 7, reps=reps)}
        if have_ref:
            r["reference_ncore"] = A.fan_out(pool, cores, task, "reference", size_ref, 9, reps=reps)
        return r

    def fracs(trees_per_s, bytes_tree, flop_tree, mufu_tree):
        return {"hbm_frac": trees_per_s * bytes_tree / 1e9 / hbm_peak, "fp32_frac": trees_per_s * flop_tree / fp32_peak,
                "mufu_frac": trees_per_s * mufu_tree / mufu_peak,
                "per_tree": {"bytes": bytes_tree, "flop": flop_tree, "mufu": mufu_tree}}

    def vs(block, gpu_tps):
        c = block["cpu"]
        base = c.get("reference_ncore") or c["port_ncore"]
        block["gpu_over_cpu_ncore"] = gpu_tps / base["trees_per_s"]
        block["cpu_kind"] = base["impl"]
        return block

    out = {}
    # ---- C1: the reference's own test default, batch 1024 (tests/test_data_randomghm.py:14-22,41,50) ----
    B1 = 1024
    c1 = {"workload": "ConditionalDenoiseSampler([3,4],[3,3],p=.1,sigma=.1).get_batch(1024, guide=True) and "
                      "DenoiseSampler(3,3,p=.1,sigma=.1).get_batch(1024, guide=True) through the facade (wall time per "
                      "call incl. the float64 posterior D2H the return type implies)", "batch": B1}
    for rng in ("philox", "numpy"):
        cd = G.ConditionalDenoiseSampler([3, 4], [3, 3], [u, u], [.1, .1], sigma=.1, device=dev, rng=rng)
        dn = G.DenoiseSampler(3, 3, u, p_flip=.1, sigma=.1, device=dev, rng=rng)
        ms_cd = _wall_ms(lambda i: cd.get_batch(B1, guide=True, device=dev), 10)
        ms_dn = _wall_ms(lambda i: dn.get_batch(B1, guide=True, device=dev), 10)
        c1["cdm_%s" % rng] = {"ms": ms_cd, "trees_per_s": 2 * B1 / ms_cd * 1e3}
        c1["dns_%s" % rng] = {"ms": ms_dn, "trees_per_s": B1 / ms_dn * 1e3,
                              "note": "returns its 2L+1 guide tensors on the CPU like the reference (:737): 8.9 MB pageable D2H"}
    cdp = G.ConditionalDenoiseSampler([3, 4], [3, 3], [u, u], [.1, .1], sigma=.1, device=dev, rng="philox")
    Bl = 65536
    ms_l = _gpu_ms(lambda i: cdp.get_batch(Bl, guide=True, device=dev, async_=True), 3)
    by = 4 * 81 * Q * (5 * 4 + 2) + 4 * 3 * 27 * Q + 8 * (27 + 81) + 4 * 81 * 2     # image dns guides + text cls guides + leaves + z, mean
    c1["cdm_philox_B65536_async"] = dict({"ms": ms_l, "trees_per_s": 2 * Bl / ms_l * 1e3,
                                          "dominant_kernel": "k_guides_dns_fused_c (HBM-write bound)"},
                                         **fracs(Bl / ms_l * 1e3, by, 61650 + 3190, 6420 + 250))
    c1["cpu"] = {"cdm": cpu_arm("c1_cdm", B1, B1, reps=8), "dns": cpu_arm("c1_dns", B1, B1, reps=16)}
    base = c1["cpu"]["cdm"].get("reference_ncore") or c1["cpu"]["cdm"]["port_ncore"]
    c1["gpu_over_cpu_ncore"] = c1["cdm_philox"]["trees_per_s"] / base["trees_per_s"]
    c1["gpu_B65536_over_cpu_ncore"] = c1["cdm_philox_B65536_async"]["trees_per_s"] / base["trees_per_s"]
    c1["cpu_kind"] = base["impl"]
    out["C1"] = c1
    del cd, dn, cdp

    # ---- C3: CDM sigma sweep, 65 536 pairs x 6 noise levels ---------------------------------------------
    s3 = G.ConditionalDenoiseSampler(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, sigma=1.0, device=dev, rng="philox", seed=77)
    n3 = 65536
    res3 = {}

    def run3(i):
        res3["r"] = sweeps.cdm_sigma_sweep(sigmas=A.SIGMAS_C3, n_eval=n3, sampler=s3)
    ms3 = _wall_ms(run3, 5)
    tps3 = n3 * len(A.SIGMAS_C3) / ms3 * 1e3
    out["C3"] = vs(dict({"workload": "sweeps.cdm_sigma_sweep: one paired sample of 65 536 + text BP_CLS, then per sigma in "
                                     "%s: Philox noise + image BP_DNS(z, sigma, text root message) + risk; one D2H read per sweep; "
                                     "trees = pairs x sigmas (denoiser passes)" % (list(A.SIGMAS_C3),),
                         "ms_per_sweep": ms3, "trees_per_s": tps3, "dominant_kernel": "k_dns2 (FP32 issue bound; LDCU : FFMA2 = 1 : 1)",
                         "bayes_mse_per_sigma": res3["r"]["Bayes"],
                         "cpu": cpu_arm("c3_sigma", 4096, 8192)}, **fracs(tps3, 688, 61650, 6420)), tps3)

    # ---- C4: VLM next-token posterior at every prefix ------------------------------------------------------
    s4 = G.NextWordPredictSampler(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, device=dev, rng="philox", seed=78)
    ms4 = _wall_ms(lambda i: s4.get_Bayes(n_eval=10000), 10)
    tm4 = s4.t_model
    o4 = tm4.sample(65536, seed=5, root_mode=ops.ROOT_UNIFORM, want_root_hd=True, want_post=True)
    lv4, ext4 = o4["leaves"], o4["root_hd"]
    ms4k = _gpu_ms(lambda i: tm4.bp_nwp(lv4, ext4), 5)
    ms4g = _gpu_ms(lambda i: tm4.guides_nwp(lv4, ext4), 3)
    c4 = {"workload": "NextWordPredictSampler([4,4],[3,3],p=.2): image BP_CLS -> root message -> text next-token posterior at "
                      "all 80 prefixes",
          "get_Bayes_10000": {"ms": ms4, "trees_per_s": 2 * 10000 / ms4 * 1e3, "note": "facade call, blocking host read"},
          "bp_nwp_B65536": dict({"ms": ms4k, "trees_per_s": 65536 / ms4k * 1e3,
                                 "dominant_kernel": "k_nwp_full_u + k_nwp_pos_u (dependent LDCU -> FFMA2 chains)"},
                                **fracs(65536 / ms4k * 1e3, 3888, 153600, 13600)),
          "guides_nwp_B65536": dict({"ms": ms4g, "trees_per_s": 65536 / ms4g * 1e3, "dominant_kernel": "k_nwp_pos (HBM write of the 2L+1 guide tensors)"},
                                    **fracs(65536 / ms4g * 1e3, 3888 + 41600, 153600, 13600)),
          "cpu": {"no_guides": cpu_arm("c4_nwp", 4096, 4096), "guides": cpu_arm("c4_nwp_guides", 1024, 1024)}}
    base = c4["cpu"]["no_guides"].get("reference_ncore") or c4["cpu"]["no_guides"]["port_ncore"]
    c4["gpu_over_cpu_ncore"] = c4["get_Bayes_10000"]["trees_per_s"] / base["trees_per_s"]
    baseg = c4["cpu"]["guides"].get("reference_ncore") or c4["cpu"]["guides"]["port_ncore"]
    c4["gpu_guides_over_cpu_ncore"] = 2 * c4["guides_nwp_B65536"]["trees_per_s"] / baseg["trees_per_s"]
    c4["cpu_kind"] = base["impl"]
    out["C4"] = c4
    del s4, o4, lv4, ext4
    torch.cuda.empty_cache()

    # ---- C5: sample + BP_CLS + BP_DNS(sigma = 1, ext), q = 10 and q = 256 (single GPU; the split is `strong_scaling`) --
    c5 = {"workload": "sample + BP_CLS + z = x + N(0,1) + BP_DNS(z, 1, ext = root message), L = 4, s = 3, device-resident"}
    for tag, q, B5, gemm, flop, mufu, byt in (("q10_fp32", 10, 262144, 0, 10210 + 61650, 790 + 6420, 656 + 688 + 80),
                                              ("q256_fp32", 256, 16384, 0, 5.17e6 + 31.8e6, 20224 + 164352, 656 + 2696 + 1672),
                                              ("q256_tf32", 256, 16384, 1, 5.17e6 + 31.8e6, 20224 + 164352, 656 + 2696 + 1672)):
        np.random.seed(42)
        T = G.GenTransition(4, 3, q, 0.2, 1.0)
        m = ops.GhmModel(T, 4, 3, q, p_y=np.ones(q) / q, device=dev)
        m.set_gemm_mode(gemm)

        def run5(i):
            if q <= 16:
                o = m.sample(B5, seed=10 + i, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)
                lv, hd = o["leaves"], o["root_hd"]
            else:
                lv = m.sample(B5, seed=10 + i, root_mode=ops.ROOT_UNIFORM)["leaves"]
                _, hd = m.bp_cls(lv)
            z = m.gauss_noise(lv, 1.0, seed=99 + i)
            return m.bp_dns(z, 1.0, hd)
        ms5 = _gpu_ms(run5, 3 if q <= 16 else 2, warm=1)
        tps = B5 / ms5 * 1e3
        blk = dict({"B": B5, "ms": ms5, "trees_per_s": tps,
                    "dominant_kernel": "k_dns2" if q <= 16 else ("wide path: k_wide_sgemm + row kernels" if gemm == 0 else
                                                                 "wide path: k_wide_gemm_tma (tcgen05 TF32, TMA-fed) + row kernels")},
                   **fracs(tps, byt, flop, mufu))
        if gemm == 1:
            blk["tensor_tflops"] = tps * flop / 1e12
        c5[tag] = blk
        del m
        torch.cuda.empty_cache()
    c5["cpu"] = {"q10": cpu_arm("c5_q10", 16384, 16384), "q256": cpu_arm("c5_q256", 512, 512, reps=2)}
    for tag, key in (("q10_fp32", "q10"), ("q256_fp32", "q256"), ("q256_tf32", "q256")):
        base = c5["cpu"][key].get("reference_ncore") or c5["cpu"][key]["port_ncore"]
        c5[tag]["gpu_over_cpu_ncore"] = c5[tag]["trees_per_s"] / base["trees_per_s"]
    c5["cpu_kind"] = base["impl"]
    out["C5"] = c5

    # ---- training-loop feed: batch 128, guide=True, every iteration (train_CDNS.py:128, train_NWP.py:128) ----
    from ghm_b200.feed import BatchPrefetcher
    feed = {"workload": "sampler.get_batch(batch_size=128, guide=True) per training iteration: the reference calls it synchronously on "
                        "the CPU; here feed.BatchPrefetcher keeps two batches in flight on a side stream (device tensors, no host "
                        "synchronisation) and the consumer waits on an event", "batch": 128}
    for name, cls, task in (("cdm", G.ConditionalDenoiseSampler, "feed_cdm"), ("vlm", G.NextWordPredictSampler, "feed_vlm")):
        smp = cls(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, device=dev, rng="philox", seed=91)
        pf = BatchPrefetcher(smp, batch_size=128, guide=True, depth=2)
        for _ in range(20):
            next(pf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nb = 200
        for _ in range(nb):
            next(pf)
        torch.cuda.synchronize()
        us = 1e6 * (time.perf_counter() - t0) / nb
        sync_ms = _wall_ms(lambda i: smp.get_batch(batch_size=128, guide=True, device=dev), 50)
        cpu = A.fan_out(pool, 1, task, "reference" if have_ref else "port", 128, 5, reps=8)
        feed[name] = {"prefetched_us_per_batch": us, "synchronous_get_batch_us": 1e3 * sync_ms,
                      "cpu_1core_us_per_batch": 1e6 * cpu["seconds"] / cpu["calls_per_core"], "cpu_kind": cpu["impl"]}
        del pf, smp
    out["feed"] = feed
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-eval", type=int, default=65536)
    ap.add_argument("--cpu-n-eval", type=int, default=0,
                    help="--impl reference: pairs per core per step (default 5000 for the real reference, 10000 for the port)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU arms and the config blocks")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4/C5 blocks")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg")
    ap.add_argument("--strong-pairs", type=int, default=1048576)
    ap.add_argument("--strong-pairs-wide", type=int, default=1048576, help="pairs of the q = 256 TF32 strong-scaling leg (0: skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

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
`cpu_baseline` / `--impl reference`: the NumPy oracle port of the reference algorithm
           (oracle/ghm_oracle.py, pinned to the reference's fixtures) fanned out over all host cores.
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


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference, all host cores
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    n_eval, seed = args
    from oracle import ghm_oracle as O
    u = np.ones(Q) / Q
    m = O.PairedModel(N_LAYERS, N_CHILDS, [u, u], P_FLIPS, q=Q)      # seeds the global RNG (reference :654)
    np.random.seed(seed)
    t0 = time.perf_counter()
    val = O.clip_bayes(m, n_eval, K_CLIP)
    return time.perf_counter() - t0, val[0]


def cpu_clip_step(pool, cores, n_eval_per_core, seed0):
    """One bounded CPU sample: every core evaluates clip_bayes(n_eval_per_core).  Returns (trees, seconds)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(n_eval_per_core, seed0 + i) for i in range(cores)])
    wall = time.perf_counter() - t0
    trees = cores * n_eval_per_core * (K_CLIP + 1) * 2
    return trees, wall, res


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    """`--impl reference`: time the CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    n_eval_core = args.cpu_n_eval
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_clip_step(pool, cores, max(200, n_eval_core // 4), 1000 + w)
        t_total, trees_total = 0.0, 0
        for k in range(args.steps):
            trees, wall, _ = cpu_clip_step(pool, cores, n_eval_core, 2000 + 100 * k)
            t_total += wall
            trees_total += trees
    value = trees_total / t_total
    sample = "each step: %d cores x ClipSampler.get_Bayes(n_eval=%d) = %d trees (oracle port, 1 BLAS thread per process)" % (
        cores, n_eval_core, cores * n_eval_core * (K_CLIP + 1) * 2)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n_eval, 1),
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": "port", "sample": sample},
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

    # ---- roofline of the dominant kernel (fused sampler + BP), from the per-launch events ------
    # k_tree2 launches of the two modalities and of consecutive steps overlap on four streams, so a per-launch event
    # interval would count its neighbours: the dominant kernel's throughput is taken over the whole timed region
    # (all 2K launches; k_tree2 is 95 % of the GPU time in the serialised ncu launch list, profiles/r01p_launches_bench_clip.csv).
    launch_bytes = args.steps * (B * (8 * nLt + 4 * Q + 8) + B * (8 * nLi + 4 * Q))
    peak, peak_kind = read_peaks()
    achieved = launch_bytes / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_tree2<Q=10,S=3,TPT=2,PHILOX,BP> (fused sampler + root-posterior BP, int64 leaves out); the text and "
                          "image launches of a step run concurrently on two streams and are timed as one unit",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "unit": "GB/s", "frac": achieved / peak,
                # DRAM bytes per launch from the committed `ncu --set full` capture (profiles/r01f_ncu_full_k_tree2.csv:
                # 171.37 MB written + 0.08 MB read for a 327 680-tree launch = 523.2 B/tree; below the 696 B/tree
                # algorithmic figure because part of the last leaves is still in the 126 MB L2 when the kernel ends)
                "traffic": 523.2 * B, "traffic_source": "profiles/r01f_ncu_full_k_tree2.csv",
                "bytes_per_tree": 8 * nLt + 4 * Q + 8, "trees_per_launch": B, "launches": 2 * args.steps,
                "avg_launch_ms": ms / (2 * args.steps), "kernel_share_of_gpu_time_ncu": 0.955,
                "note": "issue-slot / FP32-pipe bound by design (about 10k thread-instructions per tree: Philox 1.2k, "
                        "alias draws 1k, BP 2.6k FFMA2/FMUL2 + their LDCU/LDS operands); the HBM fraction is reported, not padded"}

    # ---- end to end through the reference-facing facade call (host in / host out) --------------
    e2e = e2e_b = None
    if True:
        # One e2e step = one grid point of the reference's p_flip sweep (figures/eval-clip-ood.py:73-79): new
        # transition tables for BOTH modalities arrive from the host (float64 matrices -> derived tables in pinned
        # memory -> one H2D copy per modality), then get_Bayes(n_eval) -> two host floats (24-byte D2H read).
        # The NumPy draw of the matrices themselves (GenTransition, a sampler-construction one-off) is done ahead.
        from ghm_b200.data_random_GHM import GenTransition
        grid = []
        for p in [0.02 * (i + 1) for i in range(20)]:
            np.random.seed(42)
            grid.append((p, GenTransition(N_LAYERS[0], N_CHILDS[0], Q, p, 1.0), GenTransition(N_LAYERS[1], N_CHILDS[1], Q, p, 1.0)))
        sampler.tree_offset = tree_off

        def e2e_step(k):
            p, tt, it = grid[k % len(grid)]
            sampler.reparameterize([p, p], transitions=(tt, it))
            return sampler.get_Bayes(n_eval=n)

        for w in range(max(1, args.warmup)):
            e2e_step(w)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            r = e2e_step(k)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": world * trees_step * args.steps / el, "unit": "trees/s",
               "h2d_bytes_per_step": tm.table_bytes + im.table_bytes, "d2h_bytes_per_step": 24,
               "call": "sampler.reparameterize(p_k) [host float64 transition matrices -> pinned derived tables -> H2D] + "
                       "ClipSampler.get_Bayes(n_eval=%d) -> (mean, se) host floats; p_k walks the 20-point p_flip grid" % n,
               "bayes_last": r[0]}
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
                         "copied to pinned host memory (what ClipSampler.get_batch returns)"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import multiprocessing as mp
        cores = host_cores()
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_clip_step(pool, cores, 500, 1)                    # warm-up (imports, page-in)
            trees, wall = 0, 0.0
            for rep in range(args.cpu_reps):
                tr, wl, res = cpu_clip_step(pool, cores, args.cpu_n_eval, 50 + 100 * rep)
                trees += tr
                wall += wl
        cpu = {"value": trees / wall, "unit": "trees/s", "cores": cores, "kind": "port",
               "sample": "%d reps x %d cores x ClipSampler.get_Bayes(n_eval=%d) via oracle/ghm_oracle.py (NumPy port of the "
                         "reference, 1 BLAS thread per process) = %d trees in %.1f s wall"
                         % (args.cpu_reps, cores, args.cpu_n_eval, trees, wall)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "trees/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms_per_step,
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n, world),
                "clocks": clk, "e2e": e2e, "e2e_get_batch": e2e_b, "gpu_launches": 3 * args.steps,
                "roofline": roofline, "cpu_baseline": cpu,
                "bayes_clip_risk": {"mean": risk_mean, "se": risk_se}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-eval", type=int, default=65536)
    ap.add_argument("--cpu-n-eval", type=int, default=10000,
                    help="pairs per core in one bounded CPU sample (10000 = the reference's own n_eval)")
    ap.add_argument("--cpu-reps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
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

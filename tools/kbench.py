#!/usr/bin/env python
"""Kernel micro-benchmark: time one C-ABI entry point on device-resident buffers with CUDA events.

    python tools/kbench.py --op sample_bp [--L 4 --s 3 --q 10 --B 327680 --reps 20] [--leaf u8|i64|none]

Development aid (not the bench contract): prints trees/s and the HBM-roofline fraction of the op.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-ghm_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="sample_bp",
                    choices=["sample", "sample_bp", "bp_cls", "bp_dns", "bp_nwp", "guides_cls", "guides_dns", "guides_nwp"])
    ap.add_argument("--L", type=int, default=4)
    ap.add_argument("--s", type=int, default=3)
    ap.add_argument("--q", type=int, default=10)
    ap.add_argument("--B", type=int, default=327680)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--leaf", default="i64", choices=["i64", "u8", "none"])
    ap.add_argument("--non-ti", action="store_true")
    ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--gemm", default="f32", choices=["f32", "tf32", "bf16"], help="wide-q (q > 16) GEMM arithmetic")
    a = ap.parse_args()
    import torch
    from ghm_b200 import ops
    from ghm_b200.data_random_GHM import GenTransition
    np.random.seed(42)
    T = GenTransition(a.L, a.s, a.q, 0.2, 1.0, translation_invariance=not a.non_ti)
    dev = torch.device("cuda", 0)
    m = ops.GhmModel(T, a.L, a.s, a.q, p_y=np.ones(a.q) / a.q, device=dev)
    m.set_gemm_mode({"f32": 0, "tf32": 1, "bf16": 2}[a.gemm])
    B, nL, q = a.B, m.n_leaves, a.q
    ldt = {"i64": torch.int64, "u8": torch.uint8, "none": None}[a.leaf]
    leaves = torch.empty((B, nL), dtype=ldt, device=dev) if ldt is not None else None
    root = torch.empty(B, dtype=torch.int64, device=dev)
    post = torch.empty((B, q), dtype=torch.float32, device=dev)
    lsz = 0 if ldt is None else (8 if ldt == torch.int64 else 1)
    if a.op in ("sample", "sample_bp"):
        bp = a.op == "sample_bp"
        fn = lambda i: ops.sample_into(m, B, ops.ROOT_UNIFORM, None, 100 + i, 0, root, leaves, post if bp else None, None)
        nbytes = B * (lsz * nL + 8 + (4 * q if bp else 0))
    else:
        out = m.sample(B, seed=1, root_mode=ops.ROOT_UNIFORM, leaf_dtype=ldt or torch.int64)
        lv = out["leaves"]
        ext = m.bp_cls(lv)[1]
        z = m.gauss_noise(lv, a.sigma, seed=3)
        lb = lv.element_size()
        if a.op == "bp_cls":
            fn = lambda i: m.bp_cls(lv); nbytes = B * (lb * nL + 8 * q)
        elif a.op == "bp_dns":
            fn = lambda i: m.bp_dns(z, a.sigma, ext); nbytes = B * (8 * nL + 4 * q)
        elif a.op == "bp_nwp":
            fn = lambda i: m.bp_nwp(lv, ext); nbytes = B * (lb * nL + 4 * q + 4 * q * (nL - 1))
        elif a.op == "guides_cls":
            fn = lambda i: m.guides_cls(lv); nbytes = B * (lb * nL + 8 * q + 4 * a.L * nL * q)
        elif a.op == "guides_dns":
            fn = lambda i: m.guides_dns(z, a.sigma, ext); nbytes = B * (8 * nL + 4 * q + 4 * nL * q * (5 * a.L + 2))
        else:
            fn = lambda i: m.guides_nwp(lv, ext); nbytes = B * (lb * nL + 4 * q + 4 * q * (nL - 1) + 4 * (nL - 1) * q * (3 * a.L + 1))
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.reps):
        fn(10 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    E_int = sum(a.s ** l for l in range(1, a.L))
    E = E_int + nL
    flops = {"bp_cls": 2.0 * q * q * E_int, "bp_dns": 4.0 * q * q * E}.get(a.op, 0.0) * B
    print(json.dumps({"op": a.op, "L": a.L, "s": a.s, "q": a.q, "B": B, "leaf": a.leaf, "gemm": a.gemm,
                      "gemm_tflops": round(flops / ms / 1e9, 2),
                      "ms": round(ms, 4), "trees_per_s": round(B / ms * 1e3), "GBps": round(nbytes / ms / 1e6, 1),
                      "hbm_frac": round(nbytes / ms / 1e6 / peak, 4)}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Config-5 scaling sweep on one GPU: sample + BP_CLS + BP_DNS(sigma=1, ext from a paired tree) over a grid of
(L, s, q), device-resident, CUDA events.  q > 16 runs the wide path in FP32 and (q >= 64) tcgen05 TF32 / BF16.

    python tools/sweep.py [--out profiles/r01_sweep_c5.json] [--cpu]

--cpu adds the oracle port's trees/s on ONE host core for a bounded sample (orientation only; the contract CPU
baseline is bench.py's).  Development / evidence tool, not the bench contract.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-ghm_b200"))

GRID_LS = [(3, 3), (4, 3), (6, 2), (3, 4), (3, 8)]
GRID_Q = [4, 10, 16, 32, 64, 128, 256]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from ghm_b200 import ops
    from ghm_b200.data_random_GHM import GenTransition
    dev = torch.device("cuda", 0)
    rows = []
    for (L, s) in GRID_LS:
        nL = s ** L
        n_nodes = sum(s ** l for l in range(L + 1))
        for q in GRID_Q:
            np.random.seed(42)
            T = GenTransition(L, s, q, 0.2, 1.0)
            m = ops.GhmModel(T, L, s, q, p_y=np.ones(q) / q, device=dev)
            per_tree = (3 * n_nodes + 3 * nL) * max(q, 64) * 4 if q > 16 else 4 * n_nodes * q * 4 + 16 * nL
            B = int(max(1024, min(262144, (6 << 30) // per_tree)) // 256 * 256)
            modes = ["f32"] + (["tf32"] if q > 16 else []) + (["bf16"] if q >= 64 else [])
            for mode in modes:
                m.set_gemm_mode({"f32": 0, "tf32": 1, "bf16": 2}[mode])

                def run(i):
                    if q <= 16:
                        out = m.sample(B, seed=10 + i, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)
                        lv, hd = out["leaves"], out["root_hd"]
                    else:
                        out = m.sample(B, seed=10 + i, root_mode=ops.ROOT_UNIFORM)
                        lv = out["leaves"]
                        _, hd = m.bp_cls(lv)
                    z = m.gauss_noise(lv, 1.0, seed=99 + i)
                    return m.bp_dns(z, 1.0, hd)

                run(0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.reps):
                    run(1 + i)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                row = {"L": L, "s": s, "q": q, "n_leaves": nL, "B": B, "gemm": mode, "ms": round(ms, 3),
                       "trees_per_s": round(B / ms * 1e3)}
                if a.cpu and mode == "f32":
                    from oracle import ghm_oracle as O
                    Bc = max(64, min(4096, int(2e7 / (n_nodes * q * q))))
                    rng = np.random.RandomState(0)
                    t0 = time.perf_counter()
                    vals = O.sample_tree(T, L, s, q, Bc, root=rng.randint(0, q, size=Bc), U=rng.rand(O.n_edges(L, s), Bc))
                    post, hdc = O.bp_cls(T, vals[-1], L, s, q, np.ones(q) / q)
                    zc = vals[-1] + rng.randn(nL, Bc)
                    O.bp_dns(T, zc, 1.0, L, s, q, ext=hdc[0][0])
                    row["cpu_1core_trees_per_s"] = round(Bc / (time.perf_counter() - t0))
                rows.append(row)
                print(json.dumps(row), flush=True)
            del m
            torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"what": "C5 sweep: sample + BP_CLS + BP_DNS(sigma=1, ext), device-resident, one B200",
                       "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

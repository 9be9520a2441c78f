#!/usr/bin/env python
"""One-shot measurement report (development aid, not the bench contract): per-kernel micro-benchmarks at the paper
config plus BASELINE configs[0] (the reference's own test default, batch 1024) through the facade, next to the
oracle port on one host core.

    python tools/report.py > profiles/rNN_kernel_table.json

* `kernels`: tools/kbench.py per C-ABI entry point (CUDA events, device-resident buffers), trees/s and HBM fraction.
* `config0`: `ConditionalDenoiseSampler([3,4],[3,3],p=.1,sigma=.1).get_batch(1024, guide=True)` and
  `DenoiseSampler(3,3,p=.1,sigma=.1).get_batch(1024, guide=True)` (tests/test_data_randomghm.py:14-22,41,50 of the
  reference), wall time per call including the host copies the reference's return types imply, in NumPy-parity and
  Philox mode; `oracle_1core` = the same recipes (sampling + BP + guide tensors) in oracle/ghm_oracle.py.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-ghm_b200"))
sys.path.insert(0, ROOT)


def kbench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kbench.py")] + list(args), capture_output=True,
                         text=True, timeout=300)
    return json.loads(out.stdout.strip().splitlines()[-1])


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import torch
    rep = {"kernels": [], "config0": {}}
    for args in (["--op", "sample", "--B", "327680"], ["--op", "sample", "--B", "327680", "--leaf", "u8"],
                 ["--op", "sample_bp", "--B", "327680"], ["--op", "bp_cls", "--B", "327680"],
                 ["--op", "bp_dns", "--B", "262144"], ["--op", "bp_nwp", "--B", "65536"],
                 ["--op", "guides_cls", "--B", "65536"], ["--op", "guides_dns", "--B", "65536"],
                 ["--op", "guides_nwp", "--B", "65536"],
                 ["--op", "bp_cls", "--B", "32768", "--q", "256", "--gemm", "tf32"],
                 ["--op", "bp_dns", "--B", "32768", "--q", "256", "--gemm", "tf32"]):
        rep["kernels"].append(kbench(*args))
    from ghm_b200 import data_random_GHM as G
    from oracle import ghm_oracle as O
    u = np.ones(10) / 10
    B = 1024
    for rng in ("numpy", "philox"):
        cd = G.ConditionalDenoiseSampler([3, 4], [3, 3], [u, u], [.1, .1], sigma=.1, rng=rng)
        dn = G.DenoiseSampler(3, 3, u, p_flip=.1, sigma=.1, rng=rng)

        def f_cd():
            r = cd.get_batch(B, guide=True, device="cuda")
            torch.cuda.synchronize()
            return r

        def f_dn():
            r = dn.get_batch(B, guide=True, device="cuda")
            torch.cuda.synchronize()
            return r
        rep["config0"]["facade_%s" % rng] = {"cdm_get_batch_ms": 1e3 * timed(f_cd, 20), "dns_get_batch_ms": 1e3 * timed(f_dn, 20)}
    pm = O.PairedModel([3, 4], [3, 3], [u, u], [.1, .1])
    sm = O.SingleModel(3, 3, u, .1)

    def o_cd():
        r = O.cdm_get_batch(pm, B, sigma=.1)
        O.guides_cls(r["t_hd"], 3, 3)
        O.guides_dns(r["hd"], r["qd"], r["bu"], 4, 3)

    def o_dn():
        r = O.dns_get_batch(sm, B, .1)
        O.guides_dns(r["hd"], r["qd"], r["bu"], 3, 3)
    rep["config0"]["oracle_1core"] = {"cdm_get_batch_ms": 1e3 * timed(o_cd, 3), "dns_get_batch_ms": 1e3 * timed(o_dn, 3)}
    rep["config0"]["trees_per_call"] = {"cdm": 2 * B, "dns": B}
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()

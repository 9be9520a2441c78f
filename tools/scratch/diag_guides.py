"""diagnostic: guide tensors vs fixtures with NO mask (how do the <= -80 entries compare?)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-ghm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TREE_CASES, load_tree_case
from ghm_b200 import ops

def bq(x): return torch.from_numpy(np.ascontiguousarray(x.T)).float().cuda()

def rep(tag, got, c, key, atol=2e-5, rtol=1e-5):
    for i, g in enumerate(got):
        ref = c[f"{key}_guide{i}"]; g = g.cpu().numpy()
        nf_ref = ~np.isfinite(ref); nf_g = ~np.isfinite(g)
        low = np.isfinite(ref) & (ref <= -80)
        err = np.abs(g - ref); tol = atol + rtol * np.abs(ref)
        bad = np.isfinite(ref) & ~(err <= tol)
        print(f"{tag} {key} g{i} shape{ref.shape} nonfinite ref={nf_ref.sum()} got={nf_g.sum()} low={low.sum()} bad={bad.sum()} "
              f"bad_low={(bad&low).sum()} min_ref={np.nanmin(ref):.1f} worst_ratio={np.nanmax(np.where(np.isfinite(ref), err/tol, 0)):.2f}")
        if bad.sum():
            idx = np.argwhere(bad)[:3]
            for ix in idx:
                ix = tuple(ix); print("    ", ix, "ref", ref[ix], "got", g[ix])

for name in TREE_CASES:
    c = load_tree_case(name)
    m = ops.GhmModel(c["T"], c["L"], c["s"], c["q"], p_y=c["p_y"], device="cuda:0")
    leaves = torch.from_numpy(np.ascontiguousarray(c[f"val{c['L']}"].T)).cuda()
    for tag in ("dns", "dnsx"):
        ext = bq(c["ext"]) if tag == "dnsx" else None
        guides, _ = m.guides_dns(bq(c["z"]), c["sigma"], ext)
        rep(name, guides, c, tag)
    guides, _, _ = m.guides_cls(leaves)
    rep(name, guides, c, "cls")
    for tag in ("nwp", "nwpx"):
        if f"{tag}_guide0" not in c: continue
        ext = bq(c["ext"]) if tag == "nwpx" else None
        guides, _ = m.guides_nwp(leaves, ext)
        rep(name, guides, c, tag)

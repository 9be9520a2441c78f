"""Generate golden fixtures by running the REAL reference (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Imports ``/root/reference/src/ghmclip`` read-only, runs its sampler/BP code on
small seeded configurations and stores inputs + outputs as ``tests/golden/*.npz``;
copies the Bayes / mis-specified-BP columns of the reference's own
``figures/data/ghm-data/*.json`` into ``tests/golden/kat.json``.  The fixtures
are committed; nothing at test time on the GPU box reads /root/reference.
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("GHM_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))

import torch  # noqa: E402
from ghmclip.data import data_random_GHM as R  # noqa: E402


def flat_T(transition):
    return {f"T{l}": np.stack(level) for l, level in enumerate(transition)}


def node_stack(tree, layer, attr):
    return np.stack([np.asarray(getattr(n, attr)) for n in tree.Tree[layer]])


def tree_case(name, L, s, q, ti, B, p_flip, sigma, seed, given_root, p_y=None):
    np.random.seed(seed)
    T = R.GenTransition(L, s, q, p_flip, 1.0, translation_invariance=ti)
    p_y = np.ones(q) / q if p_y is None else np.asarray(p_y)
    root = np.random.choice(q, size=B) if given_root else None
    state = np.random.get_state()
    tree = R.GHMTree(L, s, q, p_y, p_flip, T, B, build_tree=True, root=root)
    after = np.random.get_state()
    # re-draw what gen_values consumed
    np.random.set_state(state)
    if not given_root:
        root_drawn = np.random.choice(q, size=B, p=p_y)
        assert np.array_equal(root_drawn, tree.T_value[0][0])
    E = sum(s ** l for l in range(1, L + 1))
    U = np.random.rand(E, B)
    assert np.random.get_state()[2] == after[2] and np.array_equal(np.random.get_state()[1], after[1])
    out = dict(flat_T(T))
    out.update(L=L, s=s, q=q, ti=int(ti), B=B, p_flip=p_flip, sigma=sigma, p_y=p_y, U=U,
               root=np.asarray(tree.T_value[0][0]), given_root=int(given_root))
    for l in range(1, L + 1):
        out[f"val{l}"] = np.asarray(tree.T_value[l], dtype=np.int64)
    # --- BP_CLS + cls guides
    post = tree.BP_CLS()
    out["cls_post"] = post.copy()
    out["cls_root_hd"] = tree.root_node.hd_message.copy()
    for i, g in enumerate(tree.guided_info()):
        out[f"cls_guide{i}"] = g.numpy()
    ext = np.log(R._softmax_row(np.random.normal(0, 2.0, [B, q])).T)   # a synthetic (q,B) external message
    ext = ext - ext.max(0)
    out["ext"] = ext
    z = np.random.randn(s ** L, B) * sigma + np.asarray(tree.leaves_values)
    out["z"] = z
    # --- BP_DNS without / with external message, + dns guides (need cls_flag off)
    for tag, e in (("dns", None), ("dnsx", ext)):
        tree.build_tree()
        tree.cls_flag = False
        mean = tree.BP_DNS(z, sigma, external_hd_message=None if e is None else e.copy())
        out[f"{tag}_mean"] = mean.copy()
        for i, g in enumerate(tree.guided_info()):
            out[f"{tag}_guide{i}"] = g.numpy()
    # --- NWP without / with external message, with guides
    if s ** L >= 2:
        for tag, e in (("nwp", None), ("nwpx", ext)):
            tree.build_tree()
            pp, guides = tree.BP_NWP_autoregressive(guide_info=True, device="cpu",
                                                    external_hd_message=None if e is None else e.copy())
            out[f"{tag}_pp"] = pp.numpy()
            for i, g in enumerate(guides):
                out[f"{tag}_guide{i}"] = g.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1 and k[0] != "T"})


def sampler_cases():
    """Sampler-level get_batch outputs (return structure, shapes, dtypes, values)."""
    u = np.ones(10) / 10
    out = {}
    s = R.ClipSampler([2, 3], [2, 2], [u, u], [.2, .3], K=4)
    rt, ri = s.get_batch(batch_size=6, guide=True)
    out.update(clip_t_leaves=rt[0].numpy(), clip_t_root=rt[1].numpy(), clip_t_pp=rt[3],
               clip_i_leaves=ri[0].numpy(), clip_i_root=ri[1].numpy(), clip_i_pp=ri[3])
    for i, g in enumerate(rt[2]):
        out[f"clip_t_guide{i}"] = g.numpy()
    for i, g in enumerate(ri[2]):
        out[f"clip_i_guide{i}"] = g.numpy()
    out["clip_loss"] = np.array(R.PPCLIPLoss(rt[3].T, ri[3].T, 6, K=4, variable_type=10))
    s = R.ClipSampler([2, 3], [2, 2], [u, u], [.2, .3], K=4)
    out["clip_bayes_n50"] = np.array(s.get_Bayes(n_eval=50))

    s = R.ConditionalDenoiseSampler([2, 3], [3, 2], [u, u], [.2, .1], sigma=0.7)
    rt, ri = s.get_batch(batch_size=9, guide=True)
    out.update(cdm_t_leaves=rt[0].numpy(), cdm_t_root=rt[1].numpy(), cdm_t_pp=rt[3],
               cdm_z=ri[0].numpy(), cdm_i_leaves=ri[1].numpy(), cdm_mean=ri[3])
    for i, g in enumerate(rt[2]):
        out[f"cdm_t_guide{i}"] = g.numpy()
    for i, g in enumerate(ri[2]):
        out[f"cdm_i_guide{i}"] = g.numpy()
    s = R.ConditionalDenoiseSampler([2, 3], [3, 2], [u, u], [.2, .1], sigma=0.7)
    out["cdm_bayes_n64"] = np.array(s.get_Bayes(n_eval=64))

    s = R.NextWordPredictSampler([3, 2], [2, 3], [u, u], [.15, .25])
    rt, ri = s.get_batch(batch_size=5, guide=True)
    out.update(nwp_in=rt[0].numpy(), nwp_tgt=rt[1].numpy(), nwp_pp=rt[3].numpy(),
               nwp_i_leaves=ri[0].numpy(), nwp_i_root=ri[1].numpy(), nwp_i_pp=ri[3])
    for i, g in enumerate(rt[2]):
        out[f"nwp_t_guide{i}"] = g.numpy()
    for i, g in enumerate(ri[2]):
        out[f"nwp_i_guide{i}"] = g.numpy()
    s = R.NextWordPredictSampler([3, 2], [2, 3], [u, u], [.15, .25])
    b = s.get_Bayes(n_eval=40)
    out["nwp_bayes_n40"] = np.array([b[0].item(), b[1].item()])

    s = R.DoubleSampler([2, 2], [2, 3], [u, u], [.2, .2])
    tl, il, tpp, ipp, root = s.get_zeroshot_batch(batch_size=7)
    out.update(zs_t_leaves=tl, zs_i_leaves=il, zs_t_pp=tpp, zs_i_pp=ipp, zs_root=root)

    py = np.array([.3, .1, .05, .05, .1, .1, .1, .1, .05, .05])
    s = R.ClassificationSampler(3, 2, py, p_flip=.25)
    r = s.get_batch(batch_size=8, guide=True)
    out.update(cls_leaves=r[0].numpy(), cls_root=r[1].numpy(), cls_pp=r[3], cls_py=py)
    for i, g in enumerate(r[2]):
        out[f"cls_guide{i}"] = g.numpy()
    s = R.ClassificationSampler(3, 2, py, p_flip=.25)
    out["cls_bayes_n80"] = np.array(s.get_Bayes(n_eval=80))

    s = R.DenoiseSampler(2, 3, py, p_flip=.2, sigma=0.5)
    r = s.get_batch(batch_size=8, guide=True)
    out.update(dns_z=r[0].numpy(), dns_x=r[1].numpy(), dns_mean=r[3])
    for i, g in enumerate(r[2]):
        out[f"dns_guide{i}"] = g.numpy()
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), **out)
    print("wrote samplers", len(out))


def kat_json():
    D = os.path.join(REF, "figures", "data", "ghm-data")
    kat = {"_source": "reference figures/data/ghm-data/*.json (risk columns that need no checkpoints)",
           "_recipes": "SURVEY.md Appendix D"}
    for fname, cols in (("clip-risk.json", ["Bayes"]), ("cdm-risk.json", ["Bayes"]),
                        ("vlm-risk.json", ["Bayes"]), ("zsc-risk.json", ["Bayes"]),
                        ("vlm-ood.json", ["Bayes", "Mis-spec. BP"]),
                        ("ood-clip.json", ["Bayes", "Mis-spec. BP"])):
        with open(os.path.join(D, fname)) as f:
            d = json.load(f)
        pkey = "p_flip" if "p_flip" in d else ("p" if "p" in d else None)
        kat[fname] = {c: d[c] for c in cols if c in d}
        if pkey:
            kat[fname]["p_flip"] = d[pkey]
        else:
            kat[fname]["_keys"] = list(d.keys())
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote kat.json", {k: list(v.keys()) for k, v in kat.items() if isinstance(v, dict)})


if __name__ == "__main__":
    py = np.array([.4, .35, .25])
    tree_case("tree_L1s2q3", 1, 2, 3, True, 9, 0.3, 0.8, 1, True)
    tree_case("tree_L2s2q3_nonTI_py", 2, 2, 3, False, 11, 0.25, 0.6, 2, False, p_y=py)
    tree_case("tree_L3s3q10", 3, 3, 10, True, 13, 0.1, 0.1, 3, True)
    tree_case("tree_L4s3q10", 4, 3, 10, True, 8, 0.2, 1.0, 4, True)
    tree_case("tree_L3s2q5_nonTI", 3, 2, 5, False, 10, 0.35, 1.5, 5, True)
    tree_case("tree_L2s4q16", 2, 4, 16, True, 6, 0.05, 0.3, 6, True)
    tree_case("tree_L5s2q4", 5, 2, 4, True, 7, 0.4, 2.0, 7, False)
    sampler_cases()
    kat_json()

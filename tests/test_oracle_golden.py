"""Pin the CPU oracle (oracle/ghm_oracle.py) to the real reference.

Fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py from
/root/reference (reference data_random_GHM.py); kat.json holds the reference's
own shipped risk values (figures/data/ghm-data/*.json).  Integer outputs must be
bit-exact; float64 outputs within 1e-12 (only summation order can differ).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_tree_case
from oracle import ghm_oracle as O

RT = dict(rtol=1e-12, atol=1e-13)
u10 = np.ones(10) / 10


def test_sample_bit_exact(tree_case):
    c = tree_case
    vals = O.sample_tree(c["T"], c["L"], c["s"], c["q"], c["B"], root=c["root"], U=c["U"])
    for l in range(1, c["L"] + 1):
        assert np.array_equal(vals[l], c[f"val{l}"]), f"level {l}"


def test_bp_cls(tree_case):
    c = tree_case
    post, hd = O.bp_cls(c["T"], c[f"val{c['L']}"], c["L"], c["s"], c["q"], c["p_y"])
    np.testing.assert_allclose(post, c["cls_post"], **RT)
    np.testing.assert_allclose(hd[0][0], c["cls_root_hd"], **RT)
    for i, g in enumerate(O.guides_cls(hd, c["L"], c["s"])):
        assert g.dtype == np.float32 and g.shape == c[f"cls_guide{i}"].shape
        np.testing.assert_allclose(g, c[f"cls_guide{i}"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["dns", "dnsx"])
def test_bp_dns(tree_case, tag):
    c = tree_case
    ext = c["ext"] if tag == "dnsx" else None
    mean, hd, qd, bu = O.bp_dns(c["T"], c["z"], c["sigma"], c["L"], c["s"], c["q"], ext=ext)
    np.testing.assert_allclose(mean, c[f"{tag}_mean"], rtol=1e-11, atol=1e-12)
    for i, g in enumerate(O.guides_dns(hd, qd, bu, c["L"], c["s"])):
        ref = c[f"{tag}_guide{i}"]
        assert g.shape == ref.shape
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(g))
        np.testing.assert_allclose(g[fin], ref[fin], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["nwp", "nwpx"])
def test_bp_nwp(tree_case, tag):
    c = tree_case
    ext = c["ext"] if tag == "nwpx" else None
    pp, guides = O.bp_nwp(c["T"], c[f"val{c['L']}"], c["L"], c["s"], c["q"], ext=ext, guide=True)
    np.testing.assert_allclose(pp, c[f"{tag}_pp"], rtol=2e-6, atol=1e-7)
    assert len(guides) == 2 * c["L"] + 1
    for i, g in enumerate(guides):
        np.testing.assert_allclose(g, c[f"{tag}_guide{i}"], rtol=1e-6, atol=1e-6)


def test_gen_transition_matches_fixture_tables():
    for name, seed in (("tree_L3s3q10", 3), ("tree_L3s2q5_nonTI", 5)):
        c = load_tree_case(name)
        np.random.seed(seed)
        T = O.gen_transition(c["L"], c["s"], c["q"], c["p_flip"], 1.0, bool(c["ti"]))
        for l in range(c["L"]):
            assert np.array_equal(np.stack(T[l]), c[f"T{l}"])


def test_sampler_recipes_small():
    g = dict(np.load(os.path.join(GOLDEN, "samplers.npz")))
    m = O.PairedModel([2, 3], [2, 2], [u10, u10], [.2, .3])
    r = O.clip_get_batch(m, 6, K=4)
    assert np.array_equal(r["t_leaves"].T, g["clip_t_leaves"]) and np.array_equal(r["i_leaves"].T, g["clip_i_leaves"])
    assert np.array_equal(r["i_root"], g["clip_i_root"])
    np.testing.assert_allclose(r["t_pp"].T, g["clip_t_pp"], **RT)
    np.testing.assert_allclose(r["i_pp"].T, g["clip_i_pp"], **RT)
    np.testing.assert_allclose(O.clip_loss(r["t_pp"], r["i_pp"], 6, 4, 10), g["clip_loss"], rtol=1e-12)
    m = O.PairedModel([2, 3], [2, 2], [u10, u10], [.2, .3])
    np.testing.assert_allclose(O.clip_bayes(m, 50), g["clip_bayes_n50"], rtol=1e-12)

    m = O.PairedModel([2, 3], [3, 2], [u10, u10], [.2, .1])
    r = O.cdm_get_batch(m, 9, sigma=0.7)
    assert np.array_equal(r["i_leaves"].T, g["cdm_i_leaves"])
    np.testing.assert_allclose(r["z"].T.astype(np.float32), g["cdm_z"], rtol=1e-6)
    np.testing.assert_allclose(r["mean"].T, g["cdm_mean"], rtol=1e-11)
    np.testing.assert_allclose(r["t_pp"], g["cdm_t_pp"], **RT)      # (q,B): untransposed in the reference
    m = O.PairedModel([2, 3], [3, 2], [u10, u10], [.2, .1])
    np.testing.assert_allclose(O.cdm_bayes(m, 64, sigma=0.7), g["cdm_bayes_n64"], rtol=1e-11)

    m = O.PairedModel([3, 2], [2, 3], [u10, u10], [.15, .25])
    r = O.nwp_get_batch(m, 5, guide=True)
    assert np.array_equal(r["t_leaves"].T[:, :-1], g["nwp_in"]) and np.array_equal(r["t_leaves"].T[:, 1:], g["nwp_tgt"])
    np.testing.assert_allclose(r["pp"], g["nwp_pp"], rtol=2e-6, atol=1e-7)
    m = O.PairedModel([3, 2], [2, 3], [u10, u10], [.15, .25])
    np.testing.assert_allclose(O.vlm_bayes(m, 40), g["nwp_bayes_n40"], rtol=1e-5)

    m = O.PairedModel([2, 2], [2, 3], [u10, u10], [.2, .2])
    r = O.zeroshot_batch(m, 7)
    assert np.array_equal(r["t_leaves"].T, g["zs_t_leaves"]) and np.array_equal(r["root"], g["zs_root"])
    np.testing.assert_allclose(r["i_pp"].T, g["zs_i_pp"], **RT)

    m = O.SingleModel(3, 2, g["cls_py"], .25)
    r = O.cls_get_batch(m, 8)
    assert np.array_equal(r["leaves"].T, g["cls_leaves"]) and np.array_equal(r["root"], g["cls_root"])
    np.testing.assert_allclose(r["post"].T, g["cls_pp"], **RT)
    m = O.SingleModel(3, 2, g["cls_py"], .25)
    np.testing.assert_allclose(O.cls_bayes(m, 80), g["cls_bayes_n80"], rtol=1e-6)

    m = O.SingleModel(2, 3, g["cls_py"], .2)
    r = O.dns_get_batch(m, 8, 0.5)
    np.testing.assert_allclose(r["z"].T.astype(np.float32), g["dns_z"], rtol=1e-6)
    np.testing.assert_allclose(r["mean"].T, g["dns_mean"], rtol=1e-11)


def _kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def test_kat_cdm_bayes_p02():
    """cdm-risk.json Bayes[0]: ConditionalDenoiseSampler([4,4],[3,3],p=.02).get_Bayes(10000) (train_CDNS.py:66-75)."""
    m = O.PairedModel([4, 4], [3, 3], [u10, u10], [.02, .02])
    val, _ = O.cdm_bayes(m, 10000, sigma=1)
    assert val == pytest.approx(_kat()["cdm-risk.json"]["Bayes"][0], rel=1e-12)


def test_kat_clip_bayes_p20():
    """clip-risk.json Bayes[9]: ClipSampler([4,4],[3,3],p=.2).get_Bayes(10000) (train_CLIP.py:67-76)."""
    m = O.PairedModel([4, 4], [3, 3], [u10, u10], [.2, .2])
    val, _ = O.clip_bayes(m, 10000, K=4)
    assert val == pytest.approx(_kat()["clip-risk.json"]["Bayes"][9], rel=1e-12)


def test_kat_vlm_bayes_p02():
    """vlm-risk.json Bayes[0] (float32): NextWordPredictSampler(p=.02).get_Bayes(10000) (train_NWP.py:65-74)."""
    m = O.PairedModel([4, 4], [3, 3], [u10, u10], [.02, .02])
    val, _ = O.vlm_bayes(m, 10000)
    assert val == pytest.approx(_kat()["vlm-risk.json"]["Bayes"][0], rel=1e-6)


def test_kat_zsc_bayes_p20():
    """zsc-risk.json Bayes[9] (float32): figures/eval-zsc-risk.py:66-83."""
    m = O.PairedModel([4, 4], [3, 3], [u10, u10], [.2, .2])
    assert O.zsc_bayes(m, 7500) == pytest.approx(_kat()["zsc-risk.json"]["Bayes"][9], rel=1e-6)


def test_brute_force_tiny_tree():
    """Exactness of the restated BP against enumeration (L=2, s=2, q=3, non-TI)."""
    np.random.seed(11)
    L, s, q = 2, 2, 3
    T = O.gen_transition(L, s, q, 0.4, 1.0, False)
    py = np.array([.5, .3, .2])
    vals = O.sample_tree(T, L, s, q, 4, p_y=py)
    leaves = vals[-1]
    post, hd = O.bp_cls(T, leaves, L, s, q, py)
    z = np.random.randn(s ** L, 4) * 0.8 + leaves
    ext = np.log(np.array([[.2, .5, .3]] * 4).T)
    mean, *_ = O.bp_dns(T, z, 0.8, L, s, q, ext=ext)
    pp, _ = O.bp_nwp(T, leaves, L, s, q, ext=ext)
    for b in range(4):
        rp, _ = O.brute_force_posteriors(T, L, s, q, py, leaves=leaves[:, b])
        np.testing.assert_allclose(post[:, b], rp, rtol=1e-12)
        _, lm = O.brute_force_posteriors(T, L, s, q, None, z=z[:, b], sigma=0.8, ext=ext[:, b])
        np.testing.assert_allclose(mean[:, b], lm @ np.arange(q), rtol=1e-12)
    # next-token: p(leaf_{t+1} | leaves<=t, ext) by enumeration with a partially observed prefix
    import itertools
    for b in range(2):
        for pos in range(s ** L - 1):
            acc = np.zeros(q)
            for rest in itertools.product(range(q), repeat=s ** L - pos - 1):
                lv = list(leaves[:pos + 1, b]) + list(rest)
                # unnormalised joint of this full leaf assignment, uniform prior + ext
                w = 0.0
                for r in range(q):
                    pr = np.exp(ext[r, b])
                    inner = 1.0
                    for j1 in range(s):
                        t1 = 0.0
                        for v1 in range(q):
                            t = T[0][j1][r, v1]
                            for j2 in range(s):
                                t *= T[1][j1 * s + j2][v1, lv[j1 * s + j2]]
                            t1 += t
                        inner *= t1
                    w += pr * inner
                acc[rest[0]] += w
            np.testing.assert_allclose(pp[b, pos], acc / acc.sum(), rtol=2e-6)

"""GPU tests of the reference-compatible facade (ghm_b200.data_random_GHM) in PARITY mode.

The facade consumes NumPy's global stream exactly like the reference, so the reference's
shipped risk values (tests/golden/kat.json <- figures/data/ghm-data/*.json) and the reference
get_batch fixtures (tests/golden/samplers.npz) must be reproduced: integers bit-exact, float32
BP within 1e-5 relative.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
u10 = np.ones(10) / 10


@pytest.fixture(scope="module")
def G():
    from ghm_b200 import data_random_GHM as mod
    assert torch.cuda.is_available()
    return mod


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "samplers.npz")))


def test_kat_clip_bayes(G, kat):
    """clip-risk.json Bayes[p]: ClipSampler([4,4],[3,3],[u,u],[p,p]).get_Bayes(10000) (reference train_CLIP.py:67-76)."""
    for idx in (9, 0):
        p = 0.02 * (idx + 1)
        s = G.ClipSampler([4, 4], [3, 3], [u10, u10], [p, p])
        val, se = s.get_Bayes(n_eval=10000)
        assert val == pytest.approx(kat["clip-risk.json"]["Bayes"][idx], rel=1e-5)


def test_kat_zsc_bayes(G, kat):
    """zsc-risk.json Bayes[9] via the reference recipe figures/eval-zsc-risk.py:66-83 on the facade."""
    s = G.DoubleSampler(n_layers=[4, 4], n_childs=[3, 3], variable_type=10, p_ys=[u10, u10], p_flips=[.2, .2], seedtree=42)
    tl, il, tpp, ipp, root = s.get_zeroshot_batch(batch_size=7500)
    x = ipp
    for lay in s.t_transition:
        x = x @ lay[0]
    loss = torch.nn.functional.cross_entropy(torch.log(torch.tensor(x, dtype=torch.float)),
                                             torch.tensor(tl, dtype=torch.long)[:, 0]).item()
    assert loss == pytest.approx(kat["zsc-risk.json"]["Bayes"][9], rel=1e-5)
    assert tl.shape == (7500, 81) and tl.dtype == np.int64 and tpp.shape == (7500, 10) and tpp.dtype == np.float64


def test_clip_get_batch_structure_and_values(G, gold):
    s = G.ClipSampler([2, 3], [2, 2], [u10, u10], [.2, .3], K=4)
    rt, ri = s.get_batch(batch_size=6, guide=False)
    assert isinstance(rt, list) and rt[2] is None and rt[3] is None
    assert rt[0].dtype == torch.int64 and rt[0].device.type == "cpu"
    assert np.array_equal(rt[0].numpy(), gold["clip_t_leaves"]) and np.array_equal(ri[0].numpy(), gold["clip_i_leaves"])
    assert np.array_equal(rt[1].numpy(), gold["clip_t_root"]) and np.array_equal(ri[1].numpy(), gold["clip_i_root"])
    s = G.ClipSampler([2, 3], [2, 2], [u10, u10], [.2, .3], K=4)
    val = s.get_Bayes(n_eval=50)
    np.testing.assert_allclose(val, gold["clip_bayes_n50"], rtol=1e-5)
    assert G.PPCLIPLoss(gold["clip_t_pp"].T, gold["clip_i_pp"].T, 6, K=4, variable_type=10)[0] == \
        pytest.approx(gold["clip_loss"][0], rel=1e-6)


def test_mis_specified_clip_bp_recipe(G, kat):
    """ood-clip.json Mis-spec. BP[0]: BP-only part of figures/eval-clip-ood.py:58-92 (B=5000) run on the facade,
    including the caller-side mutation of T_value[-1] and re-build."""
    B = 5000
    ts = G.DoubleSampler([4, 4], [3, 3], [u10, u10], [.2, .2])
    text_tree, image_tree = ts.get_zeroshot_batch(batch_size=B * 5, return_tree=True)
    p = 0.02
    s = G.ClipSampler([4, 4], [3, 3], [u10, u10], [p, p])
    bayes, _ = s.get_Bayes(n_eval=10000)
    assert bayes == pytest.approx(kat["ood-clip.json"]["Bayes"][0], rel=1e-5)
    res_text, res_image = s.get_batch(device="cpu", batch_size=B, guide=False)
    text_tree.T_value[-1] = [res_text[0][:, idx].tolist() for idx in range(81)]
    image_tree.T_value[-1] = [res_image[0][:, idx].tolist() for idx in range(81)]
    text_tree.build_tree()
    image_tree.build_tree()
    text_tree.BP_CLS()
    image_tree.BP_CLS()
    loss, _ = G.PPCLIPLoss(text_tree.posterior_probability_CLS, image_tree.posterior_probability_CLS, B, K=4,
                           variable_type=10)
    assert loss == pytest.approx(kat["ood-clip.json"]["Mis-spec. BP"][0], rel=1e-5)


def test_philox_clip_bayes_statistical_and_sharded(G, kat):
    """Philox mode lands within the fixture's own standard error; pair-sharded evaluation equals the whole."""
    s = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=5)
    val, se = s.get_Bayes(n_eval=40000)
    assert abs(val - kat["clip-risk.json"]["Bayes"][9]) < 5 * (se + 0.0078)
    # sharding invariance: emulate 3 ranks on one GPU by evaluating pair ranges of the same global batch
    from ghm_b200 import ops
    from ghm_b200.sharding import shard_range, mean_se_from_sums
    n = 3001
    s1 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=9)
    whole, _ = s1.get_Bayes(n_eval=n)
    tot = ops.new_sums("cuda:0")
    for r in range(3):
        s2 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=9)
        lo, hi = shard_range(n, r, 3)
        rr = s2._sample_layout(n, want_leaves=False, want_post=True, pair_lo=lo, pair_hi=hi)
        ops.risk_clip(rr["t"]["post"], rr["i"]["post"], hi - lo, 4, 10, sums=tot)
    assert mean_se_from_sums(tot)[0] == pytest.approx(float(whole), rel=1e-12)

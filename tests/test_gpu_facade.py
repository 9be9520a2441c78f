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
    # the same recipe with the projection + CE reduction on the device (ghm_risk_zsc): two more grid points
    for idx, p in ((9, .2), (0, .02)):
        s2 = G.DoubleSampler(n_layers=[4, 4], n_childs=[3, 3], variable_type=10, p_ys=[u10, u10], p_flips=[p, p], seedtree=42)
        val, se = G.zeroshot_bayes(s2, 7500)
        assert val == pytest.approx(kat["zsc-risk.json"]["Bayes"][idx], rel=2e-5) and 0 < se < 0.1


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
    # the shard is ONE launch per modality (ghm_sample_blocked): its rows are bit-identical to the rows of the whole layout
    import torch
    K = 4
    s3 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=9)
    full = s3._sample_layout(n, want_leaves=True, want_post=True)
    lo, hi = shard_range(n, 1, 3)
    s4 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=9)
    part = s4._sample_layout(n, want_leaves=True, want_post=True, pair_lo=lo, pair_hi=hi)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(j * n + lo, j * n + hi) for j in range(K + 1)]).cuda()
    for side in ("t", "i"):
        for key in ("root", "leaves", "post"):
            assert torch.equal(part[side][key], full[side][key][rows]), (side, key)


def test_kat_cdm_bayes(G, kat):
    """cdm-risk.json Bayes[0,9]: ConditionalDenoiseSampler(...).get_Bayes(10000), sigma=1 (reference train_CDNS.py:66-75)."""
    for idx in (0, 9):
        p = 0.02 * (idx + 1)
        s = G.ConditionalDenoiseSampler([4, 4], [3, 3], [u10, u10], [p, p])
        val, se = s.get_Bayes(n_eval=10000)
        assert val == pytest.approx(kat["cdm-risk.json"]["Bayes"][idx], rel=1e-5)


def test_kat_vlm_bayes(G, kat):
    """vlm-risk.json Bayes[0]: NextWordPredictSampler(...).get_Bayes(10000) (reference train_NWP.py:65-74)."""
    s = G.NextWordPredictSampler([4, 4], [3, 3], [u10, u10], [.02, .02])
    val, se = s.get_Bayes(n_eval=10000)
    assert isinstance(val, torch.Tensor) and val.dtype == torch.float32
    assert val.item() == pytest.approx(kat["vlm-risk.json"]["Bayes"][0], rel=1e-5)


def test_mis_specified_vlm_bp_recipe(G, kat):
    """vlm-ood.json Mis-spec. BP[0]: BP-only part of figures/eval-vlm-ood.py:98-132 (B=1000), with the
    caller-side item assignment into T_value[-1] (:117)."""
    B = 1000
    ts = G.DoubleSampler([4, 4], [3, 3], [u10, u10], [.2, .2])
    text_tree, image_tree = ts.get_zeroshot_batch(batch_size=B, return_tree=True)
    s = G.NextWordPredictSampler([4, 4], [3, 3], [u10, u10], [.02, .02])
    s.get_Bayes(n_eval=10000)
    res_text, res_image = s.get_batch(device="cpu", batch_size=B, guide=False)
    for idx in range(80):
        text_tree.T_value[-1][idx] = res_text[0][:, idx].tolist()
    image_tree.T_value[-1] = [res_image[0][:, idx].tolist() for idx in range(81)]
    text_tree.build_tree()
    image_tree.build_tree()
    image_tree.BP_CLS()
    ext = image_tree.root_node.hd_message
    assert ext.shape == (10, B)
    out, _ = text_tree.BP_NWP_autoregressive(external_hd_message=ext, device="cpu", guide_info=False)
    pred = out.reshape(-1, 10)
    target = res_text[1].reshape(-1)
    loss = torch.mean(-torch.log(pred[range(len(target)), target])).item()
    assert loss == pytest.approx(kat["vlm-ood.json"]["Mis-spec. BP"][0], rel=1e-5)


def test_get_batch_structures_vs_reference_fixture(G, gold):
    """Return structure, shapes, dtypes and values of every sampler's get_batch."""
    def cmp_guides(got, prefix, atol=2e-5):
        n = len([k for k in gold if k.startswith(prefix)])
        assert len(got) == n
        for i, g in enumerate(got):
            ref = gold[f"{prefix}{i}"]
            assert g.dtype == torch.float32 and tuple(g.shape) == ref.shape
            assert np.isfinite(ref).all()                   # no mask: every entry is compared (log-domain kernels)
            np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=1e-5, atol=atol)

    s = G.ClipSampler([2, 3], [2, 2], [u10, u10], [.2, .3], K=4)
    rt, ri = s.get_batch(batch_size=6, guide=True)
    assert np.array_equal(rt[0].numpy(), gold["clip_t_leaves"]) and np.array_equal(ri[1].numpy(), gold["clip_i_root"])
    assert rt[3].dtype == np.float64 and rt[3].shape == gold["clip_t_pp"].shape
    np.testing.assert_allclose(rt[3], gold["clip_t_pp"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ri[3], gold["clip_i_pp"], rtol=1e-5, atol=1e-7)
    cmp_guides(rt[2], "clip_t_guide")
    cmp_guides(ri[2], "clip_i_guide")

    s = G.ConditionalDenoiseSampler([2, 3], [3, 2], [u10, u10], [.2, .1], sigma=0.7)
    rt, ri = s.get_batch(batch_size=9, guide=True)
    assert np.array_equal(rt[0].numpy(), gold["cdm_t_leaves"]) and np.array_equal(rt[1].numpy(), gold["cdm_t_root"])
    assert rt[3].shape == gold["cdm_t_pp"].shape            # (q, B): untransposed in the reference (:884)
    np.testing.assert_allclose(rt[3], gold["cdm_t_pp"], rtol=1e-5, atol=1e-7)
    assert ri[0].dtype == torch.float32 and ri[1].dtype == torch.int64
    np.testing.assert_allclose(ri[0].numpy(), gold["cdm_z"], rtol=1e-6)
    assert np.array_equal(ri[1].numpy(), gold["cdm_i_leaves"])
    assert ri[3].dtype == np.float64
    np.testing.assert_allclose(ri[3], gold["cdm_mean"], rtol=1e-5, atol=2e-6)
    cmp_guides(rt[2], "cdm_t_guide")
    cmp_guides(ri[2], "cdm_i_guide")
    s = G.ConditionalDenoiseSampler([2, 3], [3, 2], [u10, u10], [.2, .1], sigma=0.7)
    np.testing.assert_allclose(s.get_Bayes(n_eval=64), gold["cdm_bayes_n64"], rtol=1e-5)

    s = G.NextWordPredictSampler([3, 2], [2, 3], [u10, u10], [.15, .25])
    rt, ri = s.get_batch(batch_size=5, guide=True)
    assert np.array_equal(rt[0].numpy(), gold["nwp_in"]) and np.array_equal(rt[1].numpy(), gold["nwp_tgt"])
    assert rt[3].dtype == torch.float32
    np.testing.assert_allclose(rt[3].numpy(), gold["nwp_pp"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ri[3], gold["nwp_i_pp"], rtol=1e-5, atol=1e-7)
    cmp_guides(rt[2], "nwp_t_guide")
    cmp_guides(ri[2], "nwp_i_guide")
    s = G.NextWordPredictSampler([3, 2], [2, 3], [u10, u10], [.15, .25])
    b = s.get_Bayes(n_eval=40)
    np.testing.assert_allclose([b[0].item(), b[1].item()], gold["nwp_bayes_n40"], rtol=2e-5)

    py = gold["cls_py"]
    s = G.ClassificationSampler(3, 2, py, p_flip=.25)
    r = s.get_batch(batch_size=8, guide=True)
    assert np.array_equal(r[0].numpy(), gold["cls_leaves"]) and np.array_equal(r[1].numpy(), gold["cls_root"])
    np.testing.assert_allclose(r[3], gold["cls_pp"], rtol=1e-5, atol=1e-7)
    cmp_guides(r[2], "cls_guide")
    with pytest.raises(AttributeError):
        G.ClassificationSampler(3, 2, py, p_flip=.25).get_batch(batch_size=4, guide=False)
    s = G.ClassificationSampler(3, 2, py, p_flip=.25)
    np.testing.assert_allclose(s.get_Bayes(n_eval=80), gold["cls_bayes_n80"], rtol=2e-5)

    s = G.DenoiseSampler(2, 3, py, p_flip=.2, sigma=0.5)
    r = s.get_batch(batch_size=8, guide=True)
    np.testing.assert_allclose(r[0].numpy(), gold["dns_z"], rtol=1e-6)
    np.testing.assert_allclose(r[1].numpy(), gold["dns_x"])
    np.testing.assert_allclose(r[3], gold["dns_mean"], rtol=1e-5, atol=2e-6)
    cmp_guides(r[2], "dns_guide")

    s = G.DoubleSampler([2, 2], [2, 3], [u10, u10], [.2, .2])
    tl, il, tpp, ipp, root = s.get_zeroshot_batch(batch_size=7)
    assert np.array_equal(tl, gold["zs_t_leaves"]) and np.array_equal(il, gold["zs_i_leaves"])
    assert np.array_equal(root, gold["zs_root"])
    np.testing.assert_allclose(tpp, gold["zs_t_pp"], rtol=1e-5, atol=1e-7)


def test_reference_unit_tests_run_against_facade(G):
    """The reference's own tests (tests/test_data_randomghm.py:24-54), restated verbatim on the facade:
    |E[m^2] - E[m x]| < 3e-3 for the conditional and the unconditional denoiser, B = 10000, sigma = 0.1."""
    def denoise_test(true_leave_values, pred_leave_values):
        true_leave_values = np.array(true_leave_values)
        mean_power_res = np.mean(np.power(pred_leave_values, 2), 1)
        mean_pred_true_res = np.mean(np.multiply(pred_leave_values, true_leave_values), 1)
        return np.abs(np.mean(mean_power_res) - np.mean(mean_pred_true_res))
    sampler = G.ConditionalDenoiseSampler([3, 4], [3, 3], [u10, u10], [0.1, 0.1], flip_scale=1, sigma=0.1,
                                          translation_invariance=True, variable_type=10)
    _, res_image = sampler.get_batch(batch_size=10000, guide=True)
    assert denoise_test(res_image[1], res_image[-1]) < 3e-3
    sampler = G.DenoiseSampler(3, 3, u10, 0.1, flip_scale=1, sigma=0.1, translation_invariance=True, variable_type=10)
    res = sampler.get_batch(batch_size=10000, guide=True)
    assert denoise_test(res[1], res[-1]) < 3e-3


def test_ood_sweeps_reproduce_reference_columns(kat):
    """ghm_b200.sweeps (device-resident p_flip sweeps, one D2H copy at the end) against the reference's own
    ood-clip.json / vlm-ood.json columns at grid points 2, 4, 20 and 40 % (figures/eval-clip-ood.py:69-94,
    eval-vlm-ood.py:104-132)."""
    from ghm_b200 import sweeps
    pts = [2, 4, 20, 40]
    idx = [p // 2 - 1 for p in pts]
    res = sweeps.clip_ood_sweep(pts, n_eval=10000, batch_size=5000)
    assert res["p_flip"] == pts
    for j, i in enumerate(idx):
        assert res["Bayes"][j] == pytest.approx(kat["ood-clip.json"]["Bayes"][i], rel=1e-5)
        assert res["Mis-spec. BP"][j] == pytest.approx(kat["ood-clip.json"]["Mis-spec. BP"][i], rel=1e-5)
    res = sweeps.vlm_ood_sweep(pts, n_eval=10000, batch_size=1000)
    for j, i in enumerate(idx):
        assert res["Bayes"][j] == pytest.approx(kat["vlm-ood.json"]["Bayes"][i], rel=1e-5)
        assert res["Mis-spec. BP"][j] == pytest.approx(kat["vlm-ood.json"]["Mis-spec. BP"][i], rel=1e-5)


def test_ood_sweeps_philox_consistency(tmp_path):
    """Philox mode: at p == p_model the mis-specified BP *is* the Bayes rule, so the two columns agree within their
    sampling error; away from it the mis-specified risk is larger; the JSON written is the reference's layout."""
    from ghm_b200 import sweeps
    res = sweeps.clip_ood_sweep([10, 20, 30], n_eval=40000, batch_size=40000, rng="philox", seed=11)
    se = res["Bayes SE"]
    assert abs(res["Mis-spec. BP"][1] - res["Bayes"][1]) < 6 * se[1]
    for k in (0, 2):                                       # no rule beats the Bayes rule (two independent estimates)
        assert res["Mis-spec. BP"][k] > res["Bayes"][k] - 6 * se[k]
    assert res["Mis-spec. BP"][2] > res["Bayes"][2]
    res_c = sweeps.cdm_ood_sweep([20, 30], n_eval=20000, batch_size=20000, rng="philox", seed=3)
    assert abs(res_c["Mis-spec. BP"][0] - res_c["Bayes"][0]) < 6 * res_c["Bayes SE"][0]
    assert res_c["Mis-spec. BP"][1] > res_c["Bayes"][1]
    out = sweeps.write_reference_json(res, tmp_path / "clip-ood.json", extra={"Standard TF": [1.0, 2.0, 3.0]})
    back = json.load(open(tmp_path / "clip-ood.json"))
    assert list(back) == ["p_flip", "Bayes", "Mis-spec. BP", "Standard TF"] and back == out
    assert back["p_flip"] == [10, 20, 30]


def test_cdm_sigma_sweep(G):
    """BASELINE config 3: Bayes denoising risk over a sigma grid from one paired sample; monotone in sigma, and the
    sigma = 1 point agrees with ConditionalDenoiseSampler.get_Bayes (independent draw) within sampling error."""
    from ghm_b200 import sweeps
    res = sweeps.cdm_sigma_sweep(sigmas=(0.1, 0.5, 1.0, 2.0, 4.0), n_eval=20000, seed=21)
    b = res["Bayes"]
    assert all(b[i] < b[i + 1] for i in range(len(b) - 1)) and b[0] > 0
    s = G.ConditionalDenoiseSampler([4, 4], [3, 3], [u10, u10], [.2, .2], sigma=1.0, rng="philox", seed=77)
    ref, se = s.get_Bayes(n_eval=20000)
    assert abs(b[2] - ref) < 6 * (se + res["Bayes SE"][2])


# ---------------------------------------------------------------------------------------------------
# round 2: CDM mis-specified-BP recipe (parity mode), BP_DNS root aliasing, async feed, checkpoint key
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kat_regen():
    with open(os.path.join(GOLDEN, "kat_regenerated.json")) as f:
        return json.load(f)


def test_mis_specified_cdm_bp_recipe(G, kat_regen):
    """BP-only part of figures/eval-cdm-ood.py:98-127 (B = 5000) run verbatim on the facade in parity mode, against
    the value the REAL reference produces for that recipe (tests/golden/make_golden_recipes.py; the shipped
    cdm-ood.json column is stale)."""
    gold = kat_regen["cdm-ood.recipe"]
    B = gold["batch_size"]
    ts = G.DoubleSampler([4, 4], [3, 3], [u10, u10], [.2, .2])
    text_tree, image_tree = ts.get_zeroshot_batch(batch_size=B, return_tree=True)
    s = G.ConditionalDenoiseSampler([4, 4], [3, 3], [u10, u10], [.02, .02])
    bayes, _ = s.get_Bayes(n_eval=gold["n_eval"])
    assert bayes == pytest.approx(gold["Bayes"][0], rel=1e-5)
    res_text, res_image = s.get_batch(device="cpu", batch_size=B, guide=False)
    text_tree.T_value[-1] = [res_text[0][:, idx].tolist() for idx in range(81)]
    image_tree.T_value[-1] = [res_image[1][:, idx].tolist() for idx in range(81)]
    text_tree.build_tree()
    image_tree.build_tree()
    text_tree.BP_CLS()
    ext = text_tree.root_node.hd_message
    image_tree.BP_DNS(res_image[0].T.numpy(), 1, external_hd_message=ext)
    pred = image_tree.posterior_mean_DNS.T
    target = res_image[1].numpy()
    loss = np.mean(np.sum(np.power(pred - target, 2), 1))
    assert loss == pytest.approx(gold["Mis-spec. BP"][0], rel=1e-5)
    # after BP_DNS the root's hd_message is hd + ext (root bu aliases hd in the reference, :501-506): (q, B) float64,
    # max over states of (hd_message - ext) == 0 because the upward message is max-shifted before ext is added
    hd = image_tree.root_node.hd_message
    assert hd.shape == (10, B) and hd.dtype == np.float64
    np.testing.assert_allclose((hd - ext).max(0), 0.0, atol=2e-5)


def test_cdm_ood_sweep_reproduces_the_reference_recipe(kat_regen):
    """ghm_b200.sweeps.cdm_ood_sweep (device-resident, one D2H copy) in parity mode against the real reference's
    outputs of figures/eval-cdm-ood.py:98-127 at p = 2, 4, 30 %."""
    from ghm_b200 import sweeps
    gold = kat_regen["cdm-ood.recipe"]
    res = sweeps.cdm_ood_sweep(gold["p_flip"], n_eval=gold["n_eval"], batch_size=gold["batch_size"])
    for j in range(len(gold["p_flip"])):
        assert res["Bayes"][j] == pytest.approx(gold["Bayes"][j], rel=1e-5)
        assert res["Mis-spec. BP"][j] == pytest.approx(gold["Mis-spec. BP"][j], rel=1e-5)


def test_bp_dns_root_hd_message_matches_reference_fixture(G):
    """GHMTree.BP_DNS leaves root_node.hd_message = hd + ext (reference :501-506).  The reference's own root guide
    tensor (guided_info index L, identical halves) carries that row: compare through the facade objects."""
    from conftest import load_tree_case
    for name in ("tree_L3s3q10", "tree_L2s4q16", "tree_L3s2q5_nonTI"):
        c = load_tree_case(name)
        L, s, q, B = c["L"], c["s"], c["q"], c["B"]
        tree = G.GHMTree(L, s, q, c["p_y"], c["p_flip"], c["T"], B, build_tree=True, root=c["root"])
        tree.T_value[-1] = [c[f"val{L}"][i].tolist() for i in range(s ** L)]
        tree.build_tree()
        for tag in ("dns", "dnsx"):
            ext = c["ext"] if tag == "dnsx" else None
            tree.BP_DNS(c["z"], c["sigma"], external_hd_message=ext)
            ref = c[f"{tag}_guide{L}"][:, 0, :q].T                       # (q, B)
            got = tree.root_node.hd_message
            assert got.shape == (q, B) and got.dtype == np.float64
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-5)
            np.testing.assert_allclose(tree.posterior_mean_DNS, c[f"{tag}_mean"], rtol=1e-5, atol=2e-6)


def test_async_get_batch_and_prefetcher(G):
    """Training feed (reference train_CDNS.py:128-141, train_NWP.py:128-141): get_batch(device='cuda', async_=True)
    returns the same values as the blocking call with the posterior as a float64 DEVICE tensor, and BatchPrefetcher
    yields exactly the batches of consecutive direct calls (Philox state = tree_offset)."""
    from ghm_b200.feed import BatchPrefetcher
    mk = {
        "cdm": lambda: G.ConditionalDenoiseSampler([3, 4], [3, 3], [u10, u10], [.1, .1], sigma=0.5, rng="philox", seed=3),
        "nwp": lambda: G.NextWordPredictSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=4),
        "clip": lambda: G.ClipSampler([3, 3], [3, 3], [u10, u10], [.2, .2], K=4, rng="philox", seed=5),
    }

    def flat(b):
        out = []
        for x in b:
            if isinstance(x, (list, tuple)):
                out.extend(flat(x))
            elif x is not None:
                out.append(x)
        return out

    for name, make in mk.items():
        direct = make()
        ref_batches = [direct.get_batch(batch_size=128, device="cuda", guide=True) for _ in range(4)]
        a = make()
        first = a.get_batch(batch_size=128, device="cuda", guide=True, async_=True)
        for x, y in zip(flat(ref_batches[0]), flat(first)):
            if isinstance(x, np.ndarray):                                # blocking call: float64 NumPy; async: device f64
                assert isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == torch.float64 and tuple(y.shape) == x.shape
                assert np.array_equal(x, y.cpu().numpy())
            else:
                assert y.is_cuda and torch.equal(x, y)
        pf = BatchPrefetcher(make(), batch_size=128, guide=True, depth=2)
        for k in range(4):
            got = next(pf)
            for x, y in zip(flat(ref_batches[k]), flat(got)):
                y = y.cpu().numpy() if isinstance(x, np.ndarray) else y
                assert np.array_equal(x, y) if isinstance(x, np.ndarray) else torch.equal(x, y), (name, k)
        assert pf.issued == 4 + 2
    with pytest.raises(ValueError):
        BatchPrefetcher(G.ClipSampler([2, 2], [2, 2], [u10, u10], [.2, .2]), batch_size=8)     # parity mode draws on the host


def test_bayes_checkpoint_key(G, kat, tmp_path):
    """the `bayes` entry of the reference's checkpoints (train_CLIP.py:76,193-200) written from this
    path and read back the way figures/eval-clip-risk.py:28-29 reads it."""
    from ghm_b200 import sweeps
    s = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2])
    path = tmp_path / "checkpoint.pth"
    bayes, std = sweeps.write_bayes_checkpoint(path, s, n_eval=10000, state={"iter": 7, "loss_history": torch.ones(200)})
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ckpt) >= {"iter", "loss_history", "ploss_history", "bayes"} and ckpt["iter"] == 7
    assert float(ckpt["bayes"]) == pytest.approx(kat["clip-risk.json"]["Bayes"][9], rel=1e-5)
    assert float(ckpt["loss_history"][-100:].mean()) == 1.0 and float(bayes) == float(ckpt["bayes"]) and std > 0


def test_lazy_sweep_over_double_buffered_tables_equals_synchronous(G):
    """The e2e pattern of bench.py: per grid point new tables (reparameterize -> ghm_model_update, double buffered) and a
    lazy get_Bayes on one of two alternating streams, results read two calls late.  Every result must equal the
    synchronous evaluation of the same grid point (same Philox trees, same tables: 1e-12), also when the same
    tables are swapped in twice in a row and when evaluations are issued without a swap in between."""
    import torch
    n = 6000
    grid = [0.04, 0.1, 0.16, 0.22, 0.28, 0.34, 0.1, 0.1, 0.4]
    ref = []
    s1 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=11)
    for p in grid:
        s1.reparameterize([p, p])
        s1.tree_offset = 0
        ref.append(s1.get_Bayes(n_eval=n, keep_batch=True))
    s2 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.2, .2], rng="philox", seed=11)
    pend, got = [], []
    for k, p in enumerate(grid):
        s2.reparameterize([p, p])
        if k == 4:
            s2.reparameterize([p, p])                       # two swaps in a row: the fence must still protect evaluation k-1
        s2.tree_offset = 0
        pend.append(s2.get_Bayes(n_eval=n, keep_batch=True, lazy=True))
        if k == 6:                                          # a second evaluation of the same tables, no swap in between
            s2.tree_offset = 0
            extra = s2.get_Bayes(n_eval=n, lazy=True)
        if len(pend) > 2:
            got.append(pend.pop(0).result())
    got += [h.result() for h in pend]
    torch.cuda.synchronize()
    # (the float64 atomics of the risk reduction complete in a different order from run to run: 1e-12, not bit equality)
    for g, r in zip(got, ref):
        assert tuple(map(float, g)) == pytest.approx(tuple(map(float, r)), rel=1e-12)
    assert len(got) == len(ref)
    assert tuple(map(float, extra.result())) == pytest.approx(tuple(map(float, ref[6])), rel=1e-12)
    # the device-side one-call evaluation equals the Python-level sequence (sample both modalities + risk_clip)
    from ghm_b200 import ops
    s3 = G.ClipSampler([4, 4], [3, 3], [u10, u10], [.3, .3], rng="philox", seed=12)
    r = s3._sample_layout(n, want_leaves=True, want_post=True)
    a = ops.risk_clip(r["t"]["post"], r["i"]["post"], n, 4, 10)
    s3.tree_offset = 0
    h = s3.get_Bayes(n_eval=n, keep_batch=True, lazy=True)
    h.result()
    assert torch.equal(s3.last_batch["t"]["leaves"], r["t"]["leaves"]) and torch.equal(s3.last_batch["i"]["leaves"], r["i"]["leaves"])
    assert h.sums() == pytest.approx([float(x) for x in a.tolist()], rel=1e-12)

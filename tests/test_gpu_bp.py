"""GPU parity tests for K3 (BP_DNS), K4 (next-token BP) and K5 (guide tensors) through the C ABI.

Checked against the committed reference fixtures (tests/golden/tree_*.npz, generated from the
real reference) and, on larger seeded batches, against the float64 oracle.  Bars: posterior
means / next-token posteriors within 1e-5 relative (float32), guide tensors (shifted log
messages) within 2e-5 absolute + 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from conftest import TREE_CASES, load_tree_case

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from ghm_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def _model(ops, c):
    return ops.GhmModel(c["T"], c["L"], c["s"], c["q"], p_y=c["p_y"], device="cuda:0")


def _bq(x):     # (q,B) numpy -> device f32 [B,q]
    return torch.from_numpy(np.ascontiguousarray(x.T)).float().cuda()


def _cmp_guides(got, c, tag, n, atol=2e-5):
    assert len(got) == n
    for i, g in enumerate(got):
        ref = c[f"{tag}_guide{i}"]
        g = g.cpu().numpy()
        assert g.shape == ref.shape and g.dtype == np.float32, (i, g.shape, ref.shape)
        # no mask: the kernels work in the log domain with a max shift, so entries far below e^-80 (leaf
        # hd = -(z-k)^2 / 2 sigma^2 reaches -4400 at sigma = 0.1) are compared like every other one
        assert np.isfinite(ref).all() and np.isfinite(g).all(), f"{tag} guide {i}: non-finite entries"
        np.testing.assert_allclose(g, ref, rtol=RTOL, atol=atol, err_msg=f"{tag} guide {i}")


@pytest.mark.parametrize("name", TREE_CASES)
@pytest.mark.parametrize("tag", ["dns", "dnsx"])
def test_bp_dns_vs_reference_fixture(ops, name, tag):
    c = load_tree_case(name)
    m = _model(ops, c)
    z = _bq(c["z"])
    ext = _bq(c["ext"]) if tag == "dnsx" else None
    mean, root_bu = m.bp_dns(z, c["sigma"], ext, want_root_bu=True)
    np.testing.assert_allclose(mean.cpu().numpy(), c[f"{tag}_mean"].T, rtol=RTOL, atol=2e-6)
    # root_node.hd_message after BP_DNS is hd + ext (root bu aliases hd, reference :501-506): the reference's root
    # guide tensor (index L) broadcasts exactly that row over the leaves, identical halves
    ref_root = c[f"{tag}_guide{c['L']}"][:, 0, :c["q"]]
    assert np.array_equal(ref_root, c[f"{tag}_guide{c['L']}"][:, -1, c["q"]:])
    np.testing.assert_allclose(root_bu.cpu().numpy(), ref_root, rtol=RTOL, atol=2e-5)
    guides, mean2 = m.guides_dns(z, c["sigma"], ext)       # independent log-domain implementation
    np.testing.assert_allclose(mean2.cpu().numpy(), c[f"{tag}_mean"].T, rtol=RTOL, atol=2e-6)
    _cmp_guides(guides, c, tag, 2 * c["L"] + 1)


@pytest.mark.parametrize("name", TREE_CASES)
def test_guides_cls_vs_reference_fixture(ops, name):
    c = load_tree_case(name)
    m = _model(ops, c)
    leaves = torch.from_numpy(np.ascontiguousarray(c[f"val{c['L']}"].T)).cuda()
    guides, post, hd = m.guides_cls(leaves)
    np.testing.assert_allclose(post.cpu().numpy(), c["cls_post"].T, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(hd.cpu().numpy(), c["cls_root_hd"].T, rtol=RTOL, atol=2e-5)
    _cmp_guides(guides, c, "cls", c["L"])
    # cross-check the two BP_CLS implementations (linear DFS vs log level-synchronous)
    p2, h2 = m.bp_cls(leaves)
    np.testing.assert_allclose(p2.cpu().numpy(), post.cpu().numpy(), rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("name", [n for n in TREE_CASES])
@pytest.mark.parametrize("tag", ["nwp", "nwpx"])
def test_bp_nwp_vs_reference_fixture(ops, name, tag):
    c = load_tree_case(name)
    m = _model(ops, c)
    leaves = torch.from_numpy(np.ascontiguousarray(c[f"val{c['L']}"].T)).cuda()
    ext = _bq(c["ext"]) if tag == "nwpx" else None
    pp = m.bp_nwp(leaves, ext)
    assert pp.shape == c[f"{tag}_pp"].shape
    np.testing.assert_allclose(pp.cpu().numpy(), c[f"{tag}_pp"], rtol=RTOL, atol=1e-7)
    guides, pp2 = m.guides_nwp(leaves.to(torch.uint8), ext)
    np.testing.assert_allclose(pp2.cpu().numpy(), c[f"{tag}_pp"], rtol=RTOL, atol=1e-7)
    _cmp_guides(guides, c, tag, 2 * c["L"] + 1)


@pytest.mark.parametrize("L,s,q,ti,B,sigma", [(4, 3, 10, True, 2051, 1.0), (3, 4, 10, True, 515, 0.1),
                                               (6, 2, 7, False, 300, 0.5), (2, 8, 16, True, 200, 2.0),
                                               (1, 5, 12, True, 65, 0.7), (5, 3, 4, False, 130, 1.0)])
def test_bp_dns_and_nwp_vs_oracle(ops, L, s, q, ti, B, sigma):
    from oracle import ghm_oracle as O
    rng = np.random.RandomState(L * 100 + s * 10 + q)
    np.random.seed(L * 1000 + s * 10 + q + 1)
    T = O.gen_transition(L, s, q, 0.2, 1.0, ti)
    root = rng.randint(0, q, size=B)
    vals = O.sample_tree(T, L, s, q, B, root=root, U=rng.rand(O.n_edges(L, s), B))
    leaves = vals[-1]
    z = leaves + sigma * rng.randn(*leaves.shape)
    ext = np.log(rng.dirichlet(np.ones(q), size=B).T)
    ext -= ext.max(0)
    m = ops.GhmModel(T, L, s, q, device="cuda:0")
    mean, *_ = O.bp_dns(T, z, sigma, L, s, q, ext=ext)
    got = m.bp_dns(_bq(z), sigma, _bq(ext))
    np.testing.assert_allclose(got.cpu().numpy(), mean.T, rtol=RTOL, atol=5e-6)
    mean0, *_ = O.bp_dns(T, z, sigma, L, s, q)
    got0 = m.bp_dns(_bq(z), sigma, None)
    np.testing.assert_allclose(got0.cpu().numpy(), mean0.T, rtol=RTOL, atol=5e-6)
    if B <= 600:
        pp, _ = O.bp_nwp(T, leaves, L, s, q, ext=ext)
        gotp = m.bp_nwp(torch.from_numpy(np.ascontiguousarray(leaves.T)).cuda(), _bq(ext))
        np.testing.assert_allclose(gotp.cpu().numpy(), pp, rtol=RTOL, atol=1e-7)
    assert m.status() == 0


@pytest.mark.parametrize("L,s,q,sigma", [(3, 3, 10, 0.1), (2, 4, 16, 0.3), (3, 2, 32, 0.25)])
def test_dns_guides_with_outlier_observations(ops, L, s, q, sigma):
    """Defined behaviour where float32 exp underflows: observations up to 3.5 away from every state at small sigma make
    ALL leaf log-likelihoods fall below -87 (f32 exp underflow) while the reference's float64 exp(-0.5 d^2/sigma^2)
    is still > 0 (d < 3.86 sigma-units of 0.1).  The log-domain kernels shift by the max before exponentiating, so every
    guide entry is finite and matches the float64 oracle; the posterior mean matches too."""
    from oracle import ghm_oracle as O
    rng = np.random.RandomState(q + L)
    np.random.seed(5 * q + L)
    T = O.gen_transition(L, s, q, 0.15, 1.0, True)
    B = 96
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    z = vals[-1] + sigma * rng.randn(*vals[-1].shape)
    out = rng.rand(*z.shape) < 0.2                                   # 20 % outliers beyond both ends of the state range
    lim = min(3.5, 36.0 * sigma)                                     # keep the float64 oracle itself finite (hd > -700)
    z[out] = np.where(rng.rand(out.sum()) < 0.5, -rng.uniform(1.5, lim, out.sum()), q - 1 + rng.uniform(1.5, lim, out.sum()))
    ext = np.log(rng.dirichlet(np.ones(q), size=B).T)
    ext -= ext.max(0)
    # identical inputs on both sides: at d = 3.5, sigma = 0.1 the log-likelihood RATIO of neighbouring states moves by
    # 100 * dz, so the float32 rounding of z (7e-7) alone would show up as 7e-5 in the messages
    z = z.astype(np.float32).astype(np.float64)
    ext = ext.astype(np.float32).astype(np.float64)
    mean, hd, qd, bu = O.bp_dns(T, z, sigma, L, s, q, ext=ext)
    ref = O.guides_dns(hd, qd, bu, L, s)
    assert all(np.isfinite(r).all() for r in ref) and min(r.min() for r in ref) < -87.0
    m = ops.GhmModel(T, L, s, q, device="cuda:0")
    # q <= 16: tree-tiled register kernels; q = 32: the warp-per-row log-domain kernels (ghm_wide_lvl.cuh)
    guides, mean2 = m.guides_dns(_bq(z), sigma, _bq(ext))
    for i, (g, r) in enumerate(zip(guides, ref)):
        g = g.cpu().numpy()
        assert np.isfinite(g).all()
        # the deepest tensor carries the leaves' bu = hd + log(T^T exp(bu_parent - qd)) - max, a difference of two
        # numbers of magnitude |leaf hd| (<= 612 here): bounded by a few float32 ulps of that magnitude
        atol = 2.5e-4 if i == 2 * L else 5e-5
        np.testing.assert_allclose(g, r, rtol=1e-5, atol=atol, err_msg=f"guide {i}")
    np.testing.assert_allclose(mean2.cpu().numpy(), mean.T, rtol=2e-5, atol=1e-5)
    got = m.bp_dns(_bq(z), sigma, _bq(ext))
    np.testing.assert_allclose(got.cpu().numpy(), mean.T, rtol=2e-5, atol=1e-5)


def test_risk_cdm_and_ce_vs_numpy(ops):
    rng = np.random.RandomState(1)
    B, nL, q = 1000, 27, 10
    mean = rng.rand(B, nL).astype(np.float32) * 9
    x = rng.randint(0, q, size=(B, nL))
    loss = np.sum((mean.astype(np.float64) - x) ** 2, 1)
    sums = ops.risk_cdm(torch.from_numpy(mean).cuda(), torch.from_numpy(x).cuda())
    m, se = ops.mean_se(sums)
    assert m == pytest.approx(loss.mean(), rel=1e-12) and se == pytest.approx(loss.std() / np.sqrt(B), rel=1e-9)
    pp = rng.dirichlet(np.ones(q), size=(B, nL - 1)).astype(np.float32)
    ce = -np.log(pp[np.arange(B)[:, None], np.arange(nL - 1)[None, :], x[:, 1:]])
    sums = ops.risk_ce(torch.from_numpy(pp).cuda(), torch.from_numpy(x).cuda(), target_stride=nL, target_offset=1,
                       row_group=nL - 1)
    assert ops.mean_se(sums)[0] == pytest.approx(ce.astype(np.float64).mean(), rel=1e-6)


def test_gauss_noise_matches_philox_oracle_and_is_normal(ops):
    from oracle import ghm_oracle as O, philox
    np.random.seed(1)
    L, s, q, B = 3, 3, 10, 4096
    T = O.gen_transition(L, s, q, 0.2, 1.0, True)
    m = ops.GhmModel(T, L, s, q, device="cuda:0")
    out = m.sample(B, seed=3, root_mode=ops.ROOT_UNIFORM)
    z = m.gauss_noise(out["leaves"], 0.5, seed=77, tree_offset=10)
    g = (z - out["leaves"].float()).cpu().numpy() / 0.5
    ref = philox.gauss_noise(77, 10, B, s ** L).T
    np.testing.assert_allclose(g, ref, rtol=0, atol=2e-5)
    assert abs(g.mean()) < 5 / np.sqrt(g.size) and abs(g.std() - 1) < 0.01


def test_model_update_equals_fresh_model(ops):
    """ghm_model_update (p_flip sweeps): updated tables give exactly what a freshly created model gives."""
    from oracle import ghm_oracle as O
    L, s, q, B = 4, 3, 10, 513
    np.random.seed(42)
    T1 = O.gen_transition(L, s, q, 0.2, 1.0, True)
    np.random.seed(42)
    T2 = O.gen_transition(L, s, q, 0.06, 1.0, True)
    py = np.random.dirichlet(np.ones(q))
    m = ops.GhmModel(T1, L, s, q, p_y=py, device="cuda:0")
    a1 = m.sample(B, seed=3, root_mode=ops.ROOT_PRIOR, want_post=True, want_root_hd=True)
    m.update(T2, py)
    fresh = ops.GhmModel(T2, L, s, q, p_y=py, device="cuda:0")
    a2, f2 = (x.sample(B, seed=3, root_mode=ops.ROOT_PRIOR, want_post=True, want_root_hd=True) for x in (m, fresh))
    for k in ("root", "leaves", "post", "root_hd"):
        assert torch.equal(a2[k], f2[k]), k
    assert not torch.equal(a1["leaves"], a2["leaves"])
    z = m.gauss_noise(a2["leaves"], 1.0, seed=9)
    assert torch.equal(m.bp_dns(z, 1.0, a2["root_hd"]), fresh.bp_dns(z, 1.0, f2["root_hd"]))
    assert torch.equal(m.bp_nwp(a2["leaves"], a2["root_hd"]), fresh.bp_nwp(f2["leaves"], f2["root_hd"]))
    assert m.table_bytes > 0 and m.status() == 0

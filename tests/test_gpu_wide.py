"""Wide path (16 < q <= 256): sampling for any q, level-synchronous row-GEMM BP_CLS / BP_DNS, through the C ABI.

FP32 CUDA-core GEMMs must match the float64 oracle to 1e-5 relative like the register-resident kernels;
the tcgen05 TF32 / BF16 variants are held to their stated looser bounds.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from ghm_b200 import ops as _ops
    return _ops


def _model(ops, L, s, q, ti, seed=3, p=0.25):
    from oracle import ghm_oracle as O
    np.random.seed(seed)
    T = O.gen_transition(L, s, q, p, 1.0, ti)
    py = np.random.dirichlet(np.ones(q) * 2)
    return T, py, ops.GhmModel(T, L, s, q, p_y=py, device="cuda:0")


@pytest.mark.parametrize("L,s,q,ti,B", [(3, 3, 20, True, 300), (2, 2, 32, False, 129), (3, 2, 64, True, 257),
                                        (2, 3, 100, True, 150), (2, 2, 256, True, 131)])
def test_wide_sampling_and_bp_cls_vs_oracle(ops, L, s, q, ti, B):
    from oracle import ghm_oracle as O, philox
    T, py, m = _model(ops, L, s, q, ti)
    rng = np.random.RandomState(1)
    root = rng.randint(0, q, size=B)
    U = rng.rand(O.n_edges(L, s), B)
    vals = O.sample_tree(T, L, s, q, B, root=root, U=U)
    out = m.sample(B, root=root, U=torch.from_numpy(U).cuda())
    assert np.array_equal(out["leaves"].cpu().numpy(), vals[-1].T)              # parity mode: bit-exact
    pv = philox.sample_tree_philox(T, L, s, q, B, 77, tree_offset=5, p_y=py)
    po = m.sample(B, seed=77, tree_offset=5, root_mode=ops.ROOT_PRIOR)
    assert np.array_equal(po["root"].cpu().numpy(), pv[0][0])
    assert np.array_equal(po["leaves"].cpu().numpy(), pv[-1].T)                 # Philox mode: bit-exact vs the restatement
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    p, h = m.bp_cls(out["leaves"])
    np.testing.assert_allclose(p.cpu().numpy(), post.T, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(h.cpu().numpy(), hd[0][0].T, rtol=RTOL, atol=2e-5)
    p8, _ = m.bp_cls(out["leaves"].to(torch.uint8))
    assert torch.equal(p8, p)
    assert m.status() == 0


@pytest.mark.parametrize("L,s,q,ti,B,sigma", [(3, 3, 20, True, 200, 1.0), (2, 2, 32, False, 129, 0.5), (3, 2, 64, True, 140, 2.0),
                                              (2, 2, 256, True, 70, 1.0)])
def test_wide_bp_dns_vs_oracle(ops, L, s, q, ti, B, sigma):
    from oracle import ghm_oracle as O
    T, py, m = _model(ops, L, s, q, ti, seed=5)
    rng = np.random.RandomState(2)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    z = vals[-1] + sigma * rng.randn(s ** L, B)
    ext = np.log(rng.dirichlet(np.ones(q), size=B).T)                           # (q, B) log-message
    for e in (None, ext):
        ref = O.bp_dns(T, z, sigma, L, s, q, ext=e)
        mean_ref = ref[0] if isinstance(ref, tuple) else ref
        zt = torch.from_numpy(z.T.astype(np.float32)).cuda().contiguous()
        et = None if e is None else torch.from_numpy(e.T.astype(np.float32)).cuda().contiguous()
        got = m.bp_dns(zt, sigma, et).cpu().numpy()
        np.testing.assert_allclose(got, np.asarray(mean_ref).T, rtol=2e-5, atol=2e-5 * q)


@pytest.mark.parametrize("L,s,q,B", [(3, 2, 64, 300), (2, 3, 128, 200), (2, 2, 180, 129), (2, 2, 256, 260)])
@pytest.mark.parametrize("mode,tol", [("tf32", 3e-3), ("bf16", 3e-2)])
def test_wide_tcgen05_gemm_variants(ops, L, s, q, B, mode, tol):
    """tcgen05 TF32 / BF16 row-GEMMs (FP32 accumulation in TMEM) against the float64 oracle at their stated
    looser bound, and against the FP32 CUDA-core path of the same library."""
    from oracle import ghm_oracle as O
    T, py, m = _model(ops, L, s, q, True, seed=9, p=0.3)
    rng = np.random.RandomState(4)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    leaves = torch.from_numpy(vals[-1].T.copy()).cuda()
    z = torch.from_numpy((vals[-1] + rng.randn(s ** L, B)).T.astype(np.float32)).cuda().contiguous()
    p32, h32 = m.bp_cls(leaves)
    mean32 = m.bp_dns(z, 1.0, h32)
    m.set_gemm_mode(m.GEMM_TF32 if mode == "tf32" else m.GEMM_BF16)
    ptc, htc = m.bp_cls(leaves)
    meantc = m.bp_dns(z, 1.0, h32)
    torch.cuda.synchronize()
    assert not torch.equal(ptc, p32), "tensor-core mode produced bit-identical results: it did not run"
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    np.testing.assert_allclose(ptc.cpu().numpy(), post.T, rtol=tol, atol=tol * 1e-3)
    np.testing.assert_allclose(ptc.cpu().numpy(), p32.cpu().numpy(), rtol=tol, atol=tol * 1e-3)
    np.testing.assert_allclose(meantc.cpu().numpy(), mean32.cpu().numpy(), rtol=tol, atol=tol * q * 0.05)
    m.set_gemm_mode(m.GEMM_F32)
    p_again, _ = m.bp_cls(leaves)
    assert torch.equal(p_again, p32)


# ---- 16 < q <= 256: next-token posteriors and the three guide sets (log-domain warp-per-row kernels) ----------
def _cmp_list(got, ref, tag, atol=3e-5):
    assert len(got) == len(ref), (tag, len(got), len(ref))
    for i, (g, r) in enumerate(zip(got, ref)):
        g = g.cpu().numpy()
        assert g.shape == r.shape and g.dtype == np.float32, (tag, i, g.shape, r.shape)
        fin = np.isfinite(r)
        assert np.isfinite(g[fin]).all(), f"{tag} guide {i}: non-finite where the oracle is finite"
        np.testing.assert_allclose(g[fin], r[fin], rtol=RTOL, atol=atol, err_msg=f"{tag} guide {i}")
        assert (g[~fin] < -80).all(), f"{tag} guide {i}: entries the float64 oracle underflows must stay far below e^-80"


@pytest.mark.parametrize("L,s,q,ti,B", [(3, 2, 20, True, 37), (2, 3, 64, True, 50), (3, 2, 100, False, 21), (2, 2, 256, True, 19)])
def test_wide_bp_nwp_and_guides_vs_oracle(ops, L, s, q, ti, B):
    """BP_NWP_autoregressive for q > 16 (reference :336-463): posteriors at 1e-5, every guide tensor (shifted
    log-messages, :357-459) at 1e-5 relative + 3e-5 absolute against the float64 oracle, with and without ext."""
    from oracle import ghm_oracle as O
    T, py, m = _model(ops, L, s, q, ti, seed=11)
    rng = np.random.RandomState(6)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    leaves = torch.from_numpy(vals[-1].T.copy()).cuda()
    ext = np.log(rng.dirichlet(np.ones(q), size=B).T)
    for e in (None, ext):
        pp_ref, g_ref = O.bp_nwp(T, vals[-1], L, s, q, ext=e, guide=True)
        et = None if e is None else torch.from_numpy(e.T.astype(np.float32)).cuda().contiguous()
        pp = m.bp_nwp(leaves, et)
        np.testing.assert_allclose(pp.cpu().numpy(), pp_ref, rtol=2e-5, atol=1e-7)
        guides, pp2 = m.guides_nwp(leaves, et)
        assert torch.equal(pp2, pp)
        _cmp_list(guides, g_ref, "nwp")
    pp8 = m.bp_nwp(leaves.to(torch.uint8), et)
    assert torch.equal(pp8, pp)
    assert m.status() == 0


@pytest.mark.parametrize("L,s,q,ti,B,sigma", [(3, 2, 20, True, 33, 1.0), (2, 3, 64, True, 40, 0.5), (3, 2, 100, False, 17, 2.0),
                                              (2, 2, 256, True, 12, 1.0)])
def test_wide_guides_cls_dns_vs_oracle(ops, L, s, q, ti, B, sigma):
    """guided_info for q > 16 (reference :526-592): cls set (hd per depth) and dns set ((hd|qd), root (hd|bu),
    (hd|qd|bu)) against the float64 oracle; the posteriors they return equal the GEMM path's to 1e-5."""
    from oracle import ghm_oracle as O
    T, py, m = _model(ops, L, s, q, ti, seed=13)
    rng = np.random.RandomState(8)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    leaves = torch.from_numpy(vals[-1].T.copy()).cuda()
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    guides, p, h = m.guides_cls(leaves)
    np.testing.assert_allclose(p.cpu().numpy(), post.T, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(h.cpu().numpy(), hd[0][0].T, rtol=RTOL, atol=2e-5)
    _cmp_list(guides, O.guides_cls(hd, L, s), "cls")
    z = vals[-1] + sigma * rng.randn(s ** L, B)
    ext = hd[0][0]
    zt = torch.from_numpy(z.T.astype(np.float32)).cuda().contiguous()
    for e in (None, ext):
        mean_ref, dhd, dqd, dbu = O.bp_dns(T, z, sigma, L, s, q, ext=e)
        et = None if e is None else torch.from_numpy(e.T.astype(np.float32)).cuda().contiguous()
        g, mean = m.guides_dns(zt, sigma, et)
        np.testing.assert_allclose(mean.cpu().numpy(), mean_ref.T, rtol=2e-5, atol=2e-5 * q)
        # leaf hd = -(z-k)^2 / 2 sigma^2 reaches -3e4 at q = 256: float32 resolution there is 2e-3, hence the relative bound
        _cmp_list(g, O.guides_dns(dhd, dqd, dbu, L, s), "dns", atol=1e-4)
    assert m.status() == 0


def test_wide_tf32_fused_levels_per_edge_tables(ops):
    """The fused TF32 kernels (likelihoods / cavities / beliefs / leaf-row products computed inside the GEMM) pick the weight
    per NODE: a per-edge (non translation-invariant) model, ragged batch, s = 3, against the float64 oracle."""
    from oracle import ghm_oracle as O
    L, s, q, B = 3, 3, 64, 203
    T, py, m = _model(ops, L, s, q, False, seed=21, p=0.3)
    rng = np.random.RandomState(12)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    leaves = torch.from_numpy(vals[-1].T.copy()).cuda()
    z = vals[-1] + 0.7 * rng.randn(s ** L, B)
    zt = torch.from_numpy(z.T.astype(np.float32)).cuda().contiguous()
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    mean_ref = O.bp_dns(T, z, 0.7, L, s, q, ext=hd[0][0])[0]
    m.set_gemm_mode(m.GEMM_TF32)
    p, h = m.bp_cls(leaves)
    np.testing.assert_allclose(p.cpu().numpy(), post.T, rtol=3e-3, atol=3e-6)
    ext = torch.from_numpy(hd[0][0].T.astype(np.float32)).cuda().contiguous()
    mean = m.bp_dns(zt, 0.7, ext)
    np.testing.assert_allclose(mean.cpu().numpy(), mean_ref.T, rtol=3e-3, atol=3e-3 * q * 0.05)
    assert m.status() == 0


@pytest.mark.parametrize("L,s,q,B", [(3, 3, 20, 300), (2, 3, 32, 257), (3, 2, 96, 200), (2, 2, 150, 129), (2, 3, 224, 140)])
def test_wide_tf32_every_multiple_of_32(ops, L, s, q, B):
    """The TF32 tensor path takes every padded width QW = 32, 64, ..., 256 (UMMA M = 128 needs N % 16 == 0): q = 20 and 32
    run as N = 32 -- the step right after the register-resident kernels (q <= 16)."""
    from oracle import ghm_oracle as O
    T, py, m = _model(ops, L, s, q, True, seed=31, p=0.3)
    rng = np.random.RandomState(14)
    vals = O.sample_tree(T, L, s, q, B, root=rng.randint(0, q, size=B), U=rng.rand(O.n_edges(L, s), B))
    leaves = torch.from_numpy(vals[-1].T.copy()).cuda()
    z = vals[-1] + rng.randn(s ** L, B)
    zt = torch.from_numpy(z.T.astype(np.float32)).cuda().contiguous()
    p32, h32 = m.bp_cls(leaves)
    m.set_gemm_mode(m.GEMM_TF32)
    ptc, htc = m.bp_cls(leaves)
    meantc = m.bp_dns(zt, 1.0, h32)
    torch.cuda.synchronize()
    assert not torch.equal(ptc, p32), "TF32 mode produced bit-identical results: the tensor path did not run"
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    mean_ref = O.bp_dns(T, z, 1.0, L, s, q, ext=hd[0][0])[0]
    np.testing.assert_allclose(ptc.cpu().numpy(), post.T, rtol=3e-3, atol=3e-6)
    np.testing.assert_allclose(meantc.cpu().numpy(), mean_ref.T, rtol=3e-3, atol=3e-3 * q * 0.05)

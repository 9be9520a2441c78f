"""GPU parity tests for K1 (sampler), K2 (root-posterior BP) and K6 (CLIP risk) through the C ABI.

Oracle = oracle/ghm_oracle.py (pinned to the reference by tests/test_oracle_golden.py) plus
the committed reference fixtures in tests/golden/.  Bars: integers bit-exact; BP marginals
within 1e-5 relative (float32), as BASELINE.json's north_star states.
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
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _ops


def _model(ops, c):
    return ops.GhmModel(c["T"], c["L"], c["s"], c["q"], p_y=c["p_y"], device="cuda:0")


@pytest.mark.parametrize("name", TREE_CASES)
def test_parity_sampling_bit_exact_vs_reference_fixture(ops, name):
    """Same root + same uniforms as the reference -> identical leaves (reference :145-165)."""
    c = load_tree_case(name)
    m = _model(ops, c)
    for dt in (torch.int64, torch.uint8):
        out = m.sample(c["B"], root=c["root"], U=torch.from_numpy(c["U"]).cuda(), leaf_dtype=dt)
        assert np.array_equal(out["leaves"].cpu().numpy().astype(np.int64), c[f"val{c['L']}"].T)
        assert np.array_equal(out["root"].cpu().numpy(), c["root"])
    assert m.status() == 0


@pytest.mark.parametrize("name", TREE_CASES)
def test_bp_cls_vs_reference_fixture(ops, name):
    c = load_tree_case(name)
    m = _model(ops, c)
    leaves = torch.from_numpy(np.ascontiguousarray(c[f"val{c['L']}"].T)).cuda()
    for lv in (leaves, leaves.to(torch.uint8)):
        post, hd = m.bp_cls(lv)
        np.testing.assert_allclose(post.cpu().numpy(), c["cls_post"].T, rtol=RTOL, atol=1e-7)
        ref_hd = c["cls_root_hd"].T
        np.testing.assert_allclose(hd.cpu().numpy(), ref_hd, rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("L,s,q,ti,B", [(4, 3, 10, True, 4099), (3, 4, 10, True, 1000), (6, 2, 7, False, 513),
                                         (2, 8, 16, True, 300), (9, 2, 3, True, 200), (4, 3, 10, False, 777),
                                         (1, 5, 12, True, 65), (3, 3, 2, True, 129),
                                         # leaf-memo shapes of every padded q (ghm_common.cuh: ghm_memo_ok), q < padded q included
                                         (3, 3, 16, True, 333), (4, 2, 8, True, 515), (3, 4, 7, True, 260), (5, 2, 13, True, 131)])
def test_sampling_and_bp_vs_oracle(ops, L, s, q, ti, B):
    """Seeded larger batches (ragged tails, n_L > 256 chunking, per-edge tables) against the oracle."""
    from oracle import ghm_oracle as O
    rng = np.random.RandomState(L * 100 + s * 10 + q)
    np.random.seed(L * 1000 + s * 10 + q)
    T = O.gen_transition(L, s, q, 0.2, 1.0, ti)
    py = rng.dirichlet(np.ones(q) * 3)
    root = rng.randint(0, q, size=B)
    U = rng.rand(O.n_edges(L, s), B)
    vals = O.sample_tree(T, L, s, q, B, root=root, U=U)
    post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
    m = ops.GhmModel(T, L, s, q, p_y=py, device="cuda:0")
    assert m.ti == ti
    out = m.sample(B, root=root, U=torch.from_numpy(U).cuda())
    assert np.array_equal(out["leaves"].cpu().numpy(), vals[-1].T)
    p, h = m.bp_cls(out["leaves"])
    np.testing.assert_allclose(p.cpu().numpy(), post.T, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(h.cpu().numpy(), hd[0][0].T, rtol=RTOL, atol=2e-5)
    assert m.status() == 0


@pytest.mark.parametrize("L,s,q,ti,B", [(4, 3, 10, True, 1031), (3, 2, 5, False, 200), (2, 4, 16, True, 97),
                                        (1, 3, 4, True, 70), (2, 2, 8, True, 300), (3, 5, 6, True, 130),
                                        (3, 8, 10, True, 150), (5, 2, 4, False, 100), (3, 4, 10, True, 257),
                                        (3, 3, 16, True, 190), (4, 2, 8, True, 321), (3, 4, 6, True, 129), (4, 3, 9, True, 200)])
def test_philox_sampling_bit_exact_vs_philox_oracle(ops, L, s, q, ti, B):
    """Philox mode: leaves, roots and fused BP must match the NumPy Philox restatement exactly."""
    from oracle import ghm_oracle as O, philox
    np.random.seed(7)
    T = O.gen_transition(L, s, q, 0.25, 1.0, ti)
    py = np.random.dirichlet(np.ones(q) * 2)
    m = ops.GhmModel(T, L, s, q, p_y=py, device="cuda:0")
    seed, off = 0x1234ABCD5678, 1000003
    for mode, kw in ((ops.ROOT_PRIOR, dict(p_y=py)), (ops.ROOT_UNIFORM, dict(root_uniform=True))):
        vals = philox.sample_tree_philox(T, L, s, q, B, seed, tree_offset=off, **kw)
        out = m.sample(B, seed=seed, tree_offset=off, root_mode=mode, want_post=True, want_root_hd=True)
        assert np.array_equal(out["root"].cpu().numpy(), vals[0][0])
        assert np.array_equal(out["leaves"].cpu().numpy(), vals[-1].T)
        post, hd = O.bp_cls(T, vals[-1], L, s, q, py)
        np.testing.assert_allclose(out["post"].cpu().numpy(), post.T, rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(out["root_hd"].cpu().numpy(), hd[0][0].T, rtol=RTOL, atol=2e-5)
    # sharding invariance: the same global trees drawn as two shards
    a = m.sample(B // 2, seed=seed, tree_offset=off, root_mode=ops.ROOT_UNIFORM)
    b = m.sample(B - B // 2, seed=seed, tree_offset=off + B // 2, root_mode=ops.ROOT_UNIFORM)
    assert np.array_equal(torch.cat([a["leaves"], b["leaves"]]).cpu().numpy(), vals[-1].T)


def test_philox_marginals_statistical(ops):
    """Leaf marginals of 2M Philox trees against the analytic p_y * prod T (5 sigma)."""
    from oracle import ghm_oracle as O
    np.random.seed(5)
    L, s, q, B = 3, 3, 10, 1 << 21
    T = O.gen_transition(L, s, q, 0.2, 1.0, True)
    py = np.random.dirichlet(np.ones(q) * 2)
    m = ops.GhmModel(T, L, s, q, p_y=py, device="cuda:0")
    out = m.sample(B, seed=11, leaf_dtype=torch.uint8)
    lv = out["leaves"]
    for leaf in (0, 13, 26):
        p = py.copy()
        idx = leaf
        path = []
        for l in range(L, 0, -1):
            path.append((l - 1, idx))
            idx //= s
        for l, e in reversed(path):
            p = p @ T[l][e]
        emp = torch.bincount(lv[:, leaf].long(), minlength=q).cpu().numpy() / B
        assert np.all(np.abs(emp - p) < 5 * np.sqrt(p * (1 - p) / B)), (leaf, emp, p)


def test_clip_risk_vs_oracle(ops):
    from oracle import ghm_oracle as O
    rng = np.random.RandomState(0)
    n, K, q = 1000, 4, 10
    t = rng.dirichlet(np.ones(q) * 0.5, size=n * (K + 1)).astype(np.float32)
    i = rng.dirichlet(np.ones(q) * 0.5, size=n * (K + 1)).astype(np.float32)
    ref_mean, ref_se = O.clip_loss(t.T.astype(np.float64), i.T.astype(np.float64), n, K, q)
    sums = ops.risk_clip(torch.from_numpy(t).cuda(), torch.from_numpy(i).cuda(), n, K, q)
    mean, se = ops.mean_se(sums)
    assert mean == pytest.approx(ref_mean, rel=1e-12) and se == pytest.approx(ref_se, rel=1e-9)
    # sharded over pair ranges == whole
    s2 = ops.new_sums("cuda:0")
    for lo, hi in ((0, 333), (333, 1000)):
        ops.risk_clip(torch.from_numpy(t).cuda(), torch.from_numpy(i).cuda(), n, K, q, sums=s2, pair_lo=lo, pair_hi=hi)
    assert ops.mean_se(s2)[0] == pytest.approx(ref_mean, rel=1e-12)


def test_host_clip_bayes_matches_device_path(ops):
    """The host-buffer entry point == the same pipeline composed from device entry points."""
    from oracle import ghm_oracle as O
    u = np.ones(10) / 10
    mo = O.PairedModel([4, 4], [3, 3], [u, u], [.2, .2])
    tm = ops.GhmModel(mo.t_T, 4, 3, 10, device="cuda:0")
    im = ops.GhmModel(mo.i_T, 4, 3, 10, device="cuda:0")
    n, K, seed = 3000, 4, 77
    B = n * (K + 1)
    tl = torch.empty((B, 81), dtype=torch.int64).pin_memory()
    il = torch.empty((B, 81), dtype=torch.int64).pin_memory()
    tp = torch.empty((B, 10), dtype=torch.float32).pin_memory()
    ip = torch.empty((B, 10), dtype=torch.float32).pin_memory()
    sums = ops.host_clip_bayes(tm, im, n, K, seed=seed, leaves_out=(tl, il), pp_out=(tp, ip))
    # device composition
    t = tm.sample(B, seed=seed, root_mode=ops.ROOT_UNIFORM, want_post=True)
    i1 = im.sample(2 * n, root=t["root"][:2 * n], seed=seed ^ ops.IMAGE_SEED_XOR, want_post=True)
    i2 = im.sample((K - 1) * n, seed=seed ^ ops.IMAGE_SEED_XOR, tree_offset=2 * n, root_mode=ops.ROOT_UNIFORM,
                   want_post=True)
    assert torch.equal(t["leaves"].cpu(), tl)
    assert torch.equal(torch.cat([i1["leaves"], i2["leaves"]]).cpu(), il)
    ipp = torch.cat([i1["post"], i2["post"]])
    assert torch.equal(t["post"].cpu(), tp) and torch.equal(ipp.cpu(), ip)
    s2 = ops.risk_clip(t["post"], ipp, n, K, 10)
    assert sums[2] == n and sums[0] == pytest.approx(float(s2[0]), rel=1e-12)
    # and against the float64 oracle on the same leaves (BP parity at the risk level)
    tpp, _ = O.bp_cls(mo.t_T, tl.numpy().T, 4, 3, 10, u)
    ipp64, _ = O.bp_cls(mo.i_T, il.numpy().T, 4, 3, 10, u)
    ref, _ = O.clip_loss(tpp, ipp64, n, K, 10)
    assert sums[0] / n == pytest.approx(ref, rel=1e-5)


@pytest.mark.parametrize("L,s,q", [(4, 3, 10), (3, 8, 10), (2, 2, 3)])
def test_leaf_dtype_alignment_and_tpt_variants_agree(ops, L, s, q):
    """uint8 / int64 leaves, 16-byte-unaligned leaf views (row-chunk staging path) and odd batch tails must give
    the same trees and the same posteriors as the aligned int64 call; bp_cls on each form reproduces the fused BP."""
    from oracle import ghm_oracle as O
    np.random.seed(11)
    T = O.gen_transition(L, s, q, 0.2, 1.0, True)
    m = ops.GhmModel(T, L, s, q, device="cuda:0")
    B, nL = 333, s ** L
    ref = m.sample(B, seed=5, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)
    u8 = m.sample(B, seed=5, root_mode=ops.ROOT_UNIFORM, leaf_dtype=torch.uint8, want_post=True)
    assert torch.equal(u8["leaves"].long(), ref["leaves"]) and torch.equal(u8["post"], ref["post"])
    nol = m.sample(B, seed=5, root_mode=ops.ROOT_UNIFORM, want_leaves=False, want_post=True)
    assert torch.equal(nol["post"], ref["post"]) and torch.equal(nol["root"], ref["root"])
    # unaligned int64 / uint8 destinations: a view that starts 8 (resp. 1) bytes into an allocation.  These launches take the
    # generic one-tree-per-thread kernel (per-node row products + matvec), `ref` the memoised fast variant: torch.equal on
    # the posteriors is the bit-identity check of the leaf memo (ghm_common.cuh: GhmDev::leaf_memo)
    for dt in (torch.int64, torch.uint8):
        buf = torch.zeros(B * nL + 1, dtype=dt, device="cuda:0")
        view = buf[1:].view(B, nL)
        post = torch.empty((B, q), dtype=torch.float32, device="cuda:0")
        ops.sample_into(m, B, ops.ROOT_UNIFORM, None, 5, 0, None, view, post, None)
        assert torch.equal(view.long(), ref["leaves"]) and torch.equal(post, ref["post"])
        assert int(buf[0]) == 0
        p2, h2 = m.bp_cls(view)
        np.testing.assert_allclose(p2.cpu().numpy(), ref["post"].cpu().numpy(), rtol=RTOL, atol=1e-7)
    p3, h3 = m.bp_cls(u8["leaves"])
    p4, h4 = m.bp_cls(ref["leaves"])
    assert torch.equal(p3, p4) and torch.equal(h3, h4)
    np.testing.assert_allclose(p4.cpu().numpy(), ref["post"].cpu().numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(h4.cpu().numpy(), ref["root_hd"].cpu().numpy(), rtol=RTOL, atol=2e-5)
    assert m.status() == 0


def test_paired_sampling_redraws_partner_roots(ops):
    """ghm_sample_paired: the image launch re-draws the text roots from the text Philox key -- same trees as feeding the
    text launch's roots in (ghm_sample_mixed), without depending on the text launch."""
    from oracle import ghm_oracle as O
    u = np.ones(10) / 10
    mo = O.PairedModel([4, 3], [3, 3], [u, u], [.2, .3])
    tm = ops.GhmModel(mo.t_T, 4, 3, 10, device="cuda:0")
    im = ops.GhmModel(mo.i_T, 3, 3, 10, device="cuda:0")
    B, n_shared, seed, iseed, off = 1000, 400, 77, 77 ^ ops.IMAGE_SEED_XOR, 12345
    t = tm.sample(B, seed=seed, tree_offset=off, root_mode=ops.ROOT_UNIFORM)
    ref_l = torch.empty((B, 27), dtype=torch.int64, device="cuda:0")
    ref_r = torch.empty(B, dtype=torch.int64, device="cuda:0")
    ref_p = torch.empty((B, 10), dtype=torch.float32, device="cuda:0")
    ops.sample_mixed_into(im, B, n_shared, t["root"], iseed, off, ref_r, ref_l, ref_p, None)
    got_l, got_r, got_p = torch.empty_like(ref_l), torch.empty_like(ref_r), torch.empty_like(ref_p)
    ops.sample_paired_into(im, B, n_shared, seed, iseed, off, got_r, got_l, got_p, None)
    assert torch.equal(got_r, ref_r) and torch.equal(got_l, ref_l) and torch.equal(got_p, ref_p)
    assert torch.equal(got_r[:n_shared], t["root"][:n_shared])
    assert not torch.equal(got_r[n_shared:], t["root"][n_shared:])


def test_full_size_properties_c2(ops):
    """BASELINE configs[1] at its full size (65 536 pairs -> 327 680 trees per modality), through size-independent
    properties: leaves in range, posteriors normalised, BP on the produced leaves reproduces the fused posterior
    (sample -> BP idempotence), uint8 and int64 leaves agree, two shards of the global batch equal the whole, and a
    checksum of the leaves is invariant to how the batch is split across launches."""
    from oracle import ghm_oracle as O
    u = np.ones(10) / 10
    mo = O.PairedModel([4, 4], [3, 3], [u, u], [.2, .2])
    tm = ops.GhmModel(mo.t_T, 4, 3, 10, device="cuda:0")
    B, seed = 5 * 65536, 2024
    a = tm.sample(B, seed=seed, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)
    lv, post = a["leaves"], a["post"]
    assert int(lv.min()) >= 0 and int(lv.max()) <= 9 and int(a["root"].min()) >= 0 and int(a["root"].max()) <= 9
    assert torch.allclose(post.sum(1), torch.ones(B, device="cuda:0"), atol=2e-6) and bool((post >= 0).all())
    assert float(a["root_hd"].max(1).values.abs().max()) < 1e-5                      # max-shifted log-likelihood
    p2, h2 = tm.bp_cls(lv)
    assert torch.allclose(p2, post, rtol=2e-5, atol=1e-7)
    u8 = tm.sample(B, seed=seed, root_mode=ops.ROOT_UNIFORM, leaf_dtype=torch.uint8, want_post=True)
    assert torch.equal(u8["leaves"].long(), lv) and torch.equal(u8["post"], post)
    cut = 123457                                                                      # odd split, not a tile multiple
    s1 = tm.sample(cut, seed=seed, tree_offset=0, root_mode=ops.ROOT_UNIFORM, want_post=True)
    s2 = tm.sample(B - cut, seed=seed, tree_offset=cut, root_mode=ops.ROOT_UNIFORM, want_post=True)
    assert torch.equal(torch.cat([s1["leaves"], s2["leaves"]]), lv)
    assert torch.equal(torch.cat([s1["post"], s2["post"]]), post)
    w = torch.arange(1, 82, device="cuda:0", dtype=torch.int64)
    assert int((lv * w).sum()) == int((s1["leaves"] * w).sum()) + int((s2["leaves"] * w).sum())
    # the root posterior should put most mass on the true root at p_flip = 0.2 (sanity of the whole pipeline)
    acc = float((post.argmax(1) == a["root"]).float().mean())
    assert 0.5 < acc <= 1.0
    assert tm.status() == 0


def test_sample_blocked_equals_block_wise_launches(ops):
    """ghm_sample_blocked: local tree b of a shard has the global Philox index offset + (b // blk_len) * blk_stride + b % blk_len,
    so one blocked launch equals one contiguous launch per block -- roots (uniform and given), leaves and fused posteriors."""
    from oracle import ghm_oracle as O
    L, s, q = 4, 3, 10
    np.random.seed(42)
    T = O.gen_transition(L, s, q, 0.2, 1.0, True)
    m = ops.GhmModel(T, L, s, q, p_y=np.ones(q) / q, device="cuda:0")
    n, nl, lo, nb = 1000, 137, 411, 5
    B = nl * nb
    dev = m.device
    def bufs():
        return (torch.empty(B, dtype=torch.int64, device=dev), torch.empty((B, m.n_leaves), dtype=torch.int64, device=dev),
                torch.empty((B, q), dtype=torch.float32, device=dev))
    r1, l1, p1 = bufs()
    ops.sample_blocked_into(m, B, nl, n, ops.ROOT_UNIFORM, 0, None, 0, 77, 5000 + lo, r1, l1, p1, None)
    r2, l2, p2 = bufs()
    for j in range(nb):
        sl = slice(j * nl, (j + 1) * nl)
        ops.sample_into(m, nl, ops.ROOT_UNIFORM, None, 77, 5000 + lo + j * n, r2[sl], l2[sl], p2[sl], None)
    assert torch.equal(r1, r2) and torch.equal(l1, l2) and torch.equal(p1, p2)
    given = torch.randint(0, q, (B,), device=dev)
    r3, l3, p3 = bufs()
    ops.sample_blocked_into(m, B, nl, n, ops.ROOT_GIVEN, B, given, 0, 78, lo, r3, l3, p3, None)
    r4, l4, p4 = bufs()
    for j in range(nb):
        sl = slice(j * nl, (j + 1) * nl)
        ops.sample_into(m, nl, ops.ROOT_GIVEN, given[sl].contiguous(), 78, lo + j * n, r4[sl], l4[sl], p4[sl], None)
    assert torch.equal(r3, given) and torch.equal(l3, l4) and torch.equal(p3, p4)

"""Pin the NumPy Philox4x32-10 restatement to the Random123 known-answer vectors."""
import numpy as np

from oracle import philox


def test_philox4x32_10_kat():
    # Random123 kat_vectors: philox4x32 10 rounds
    out = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    out = philox.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_sampler_marginals():
    """Philox-mode sampler: leaf marginals match the analytic p_y * prod T within 5 sigma."""
    from oracle import ghm_oracle as O
    np.random.seed(3)
    L, s, q, B = 2, 2, 4, 40000
    T = O.gen_transition(L, s, q, 0.3, 1.0, True)
    py = np.array([.4, .3, .2, .1])
    vals = philox.sample_tree_philox(T, L, s, q, B, seed=99, p_y=py)
    emp = np.bincount(vals[0][0], minlength=q) / B
    assert np.all(np.abs(emp - py) < 5 * np.sqrt(py * (1 - py) / B))
    for leaf in range(s ** L):
        p = py @ T[0][leaf // s] @ T[1][leaf]
        emp = np.bincount(vals[2][leaf], minlength=q) / B
        assert np.all(np.abs(emp - p) < 5 * np.sqrt(p * (1 - p) / B))


def _alias_implied_probs(entries, q):
    """Exact distribution of philox.draw_alias over all 2^32 words r for one alias row.

    Bucket k owns the words r with (r*q) >> 32 == k; inside it frac = r*q mod 2^32 steps by q, so the number of
    accepted words (frac < e_k) is counted exactly with integer arithmetic."""
    p = np.zeros(q)
    for k in range(q):
        r_lo = -((-k << 32) // q)                      # ceil(k * 2^32 / q)
        r_hi = -((-(k + 1) << 32) // q)                # first word of the next bucket
        e = int(entries[k])
        # accepted: r*q - k*2^32 < e  <=>  r < (k*2^32 + e) / q
        r_acc = min(max(-((-((k << 32) + e)) // q), r_lo), r_hi)
        p[k] += r_acc - r_lo
        p[e & 255] += r_hi - r_acc
    return p / 4294967296.0


def test_alias_tables_reproduce_the_transition_rows():
    """Philox mode draws from Walker alias tables with 24-bit thresholds: the implied distribution (counted exactly
    over all 2^32 Philox words) must equal the transition row up to the threshold quantisation, per entry
    |P_alias - T| <= q * 2^-24 -- for the reference's parameter range (p_flip 2 % .. 40 %) and skewed rows."""
    from oracle import ghm_oracle as O
    worst = 0.0
    for q, p_flip, seed in [(10, 0.02, 42), (10, 0.2, 42), (10, 0.4, 42), (4, 0.3, 1), (16, 0.1, 2), (7, 0.25, 3),
                            (64, 0.2, 4), (256, 0.2, 5)]:
        np.random.seed(seed)
        T = O.gen_transition(2, 2, q, p_flip, 1.0, True)
        for mat in (T[0][0], T[1][1]):
            for a in range(0, q, max(1, q // 8)):
                ent = philox.alias_table(mat[a])
                assert all(int(e) & 255 < q for e in ent)
                err = np.abs(_alias_implied_probs(ent, q) - mat[a]).max()
                worst = max(worst, err / (q * 2.0 ** -24))
                assert err <= q * 2.0 ** -24, (q, p_flip, a, err)
    rows = [np.array([1.0, 0.0, 0.0]), np.array([0.5, 0.5]), np.full(5, 0.2), np.array([1e-9, 1 - 1e-9])]
    for row in rows:
        err = np.abs(_alias_implied_probs(philox.alias_table(row), len(row)) - row).max()
        assert err <= len(row) * 2.0 ** -24
    assert worst > 0.0                                  # the quantisation is real, the bound is not vacuous

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

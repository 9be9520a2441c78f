"""NumPy restatement of the library's Philox mode.  TEST INFRASTRUCTURE ONLY.

The reference has no counter-based RNG (it uses NumPy's global MT19937 stream);
Philox mode is this repo's device-side RNG, so there is no reference code to cite.
This file restates ``philox4x32_10`` / ``ghm_rng_block`` of
``multimodal-ghm_b200/csrc/ghm_common.cuh`` and the integer-threshold inverse CDF
of ``csrc/ghm_vec.cuh`` so that Philox-mode samples can be checked bit-for-bit.
Philox4x32-10 itself is pinned to the Random123 known-answer vectors in
``tests/test_oracle_philox.py``.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32 arrays. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def draw_words(seed, trees, level, n_nodes, stream=0):
    """uint32 [n_nodes, B]: word (node & 3) of block (node >> 2) of `level` for every tree."""
    trees = np.asarray(trees, dtype=np.uint64)
    nodes = np.arange(n_nodes, dtype=np.uint64)
    blk = (nodes >> np.uint64(2))[:, None]
    w = philox4x32_10(trees[None, :] & MASK, trees[None, :] >> np.uint64(32), blk,
                      np.uint64(level | (stream << 8)), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    stack = np.stack(np.broadcast_arrays(*w), axis=0)            # [4, n_nodes, B]
    sel = (nodes & np.uint64(3)).astype(np.int64)
    return stack[sel, np.arange(n_nodes), :]


def thresholds(row_probs):
    """u32 thresholds floor(cumsum * 2^32) (saturating) along the last axis, sequential f64 cumsum."""
    cdf = np.cumsum(np.asarray(row_probs, dtype=np.float64), axis=-1)
    thr = np.floor(cdf * 4294967296.0)
    return np.where(thr >= 4294967295.0, 4294967295.0, thr).astype(np.uint64)


def search(thr_rows, r, q):
    """child = #{k < q-1 : r >= thr[k]}; thr_rows [B, q], r [B]."""
    return (np.asarray(r, dtype=np.uint64)[:, None] >= thr_rows[:, :q - 1]).sum(axis=1).astype(np.int64)


def sample_tree_philox(transition, n_layer, n_child, q, batch, seed, tree_offset=0, root=None, p_y=None,
                       root_uniform=False):
    """Philox-mode twin of oracle.ghm_oracle.sample_tree: same traversal, Philox words instead of U."""
    trees = np.arange(batch, dtype=np.uint64) + np.uint64(tree_offset)
    if root is None:
        p = np.full(q, 1.0 / q) if (root_uniform or p_y is None) else np.asarray(p_y, dtype=np.float64)
        r = draw_words(seed, trees, 0, 1)[0]
        root = search(np.broadcast_to(thresholds(p), (batch, q)), r, q)
    root = np.asarray(root, dtype=np.int64)
    values = [root.reshape(1, batch)]
    s = n_child
    for layer in range(1, n_layer + 1):
        prev = values[-1]
        n = s ** layer
        words = draw_words(seed, trees, layer, n)
        cur = np.empty((n, batch), dtype=np.int64)
        for idx in range(n):
            thr = thresholds(transition[layer - 1][idx])        # [q, q]
            cur[idx] = search(thr[prev[idx // s]], words[idx], q)
        values.append(cur)
    return values


def gauss_noise(seed, tree_offset, batch, n_leaves):
    """Box-Muller normals of ghm_gauss_noise (stream 1): returns float32 [n_leaves, B]."""
    trees = np.arange(batch, dtype=np.uint64) + np.uint64(tree_offset)
    words = draw_words(seed, trees, 0, 2 * n_leaves, stream=1)   # two words per leaf
    w1, w2 = words[0::2], words[1::2]
    u1 = ((w1 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    u2 = ((w2 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    return (np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2 * np.pi) * u2)).astype(np.float32)

"""NumPy restatement of the library's Philox mode.  TEST INFRASTRUCTURE ONLY.

The reference has no counter-based RNG (it uses NumPy's global MT19937 stream);
Philox mode is this repo's device-side RNG, so there is no reference code to cite.
This file restates ``philox4x32_10`` / ``ghm_rng_block`` of
``multimodal-ghm_b200/csrc/ghm_common.cuh``, the Walker alias tables built by
``ghm_model_create`` (``csrc/ghm_model.cu``) and the alias draw of ``csrc/ghm_vec2.cuh``
(root draws keep the integer-threshold inverse CDF) so that Philox-mode samples can be
checked bit-for-bit.
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


def alias_table(row_probs):
    """Walker/Vose alias table of one probability row -> uint64 [q] entries (thr24 << 8) | alias8.

    Same float64 operations in the same order as ``ghm_model_create`` (csrc/ghm_model.cu): small /
    large work-lists in ascending index order, popped from the back.
    """
    p = np.asarray(row_probs, dtype=np.float64)
    q = p.shape[0]
    scaled = [float(p[b]) * float(q) for b in range(q)]
    small = [b for b in range(q) if scaled[b] < 1.0]
    large = [b for b in range(q) if not scaled[b] < 1.0]
    entry = [0xFFFFFF00 | b for b in range(q)]
    while small and large:
        sm = small.pop()
        lg = large.pop()
        thr = float(np.floor(scaled[sm] * 16777216.0))
        thr = min(max(thr, 0.0), 16777215.0)
        entry[sm] = (int(thr) << 8) | lg
        scaled[lg] = (scaled[lg] + scaled[sm]) - 1.0
        (small if scaled[lg] < 1.0 else large).append(lg)
    return np.array(entry, dtype=np.uint64)


def draw_alias(entries, r, q):
    """child = frac < e ? k : e & 255 with m = r*q, k = m >> 32, frac = m mod 2^32; entries [B, q], r [B]."""
    m = np.asarray(r, dtype=np.uint64) * np.uint64(q)
    k = (m >> np.uint64(32)).astype(np.int64)
    frac = m & MASK
    e = entries[np.arange(entries.shape[0]), k]
    return np.where(frac < e, k, (e & np.uint64(255)).astype(np.int64)).astype(np.int64)


def draw_words_at(seed, trees, level, blocks, words, stream=0):
    """uint32 [n, B]: word ``words[i]`` of Philox block ``blocks[i]`` of ``level`` for every tree."""
    trees = np.asarray(trees, dtype=np.uint64)
    blocks = np.asarray(blocks, dtype=np.uint64)
    w = philox4x32_10(trees[None, :] & MASK, trees[None, :] >> np.uint64(32), blocks[:, None],
                      np.uint64(level | (stream << 8)), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    stack = np.stack(np.broadcast_arrays(*w), axis=0)            # [4, n, B]
    return stack[np.asarray(words, dtype=np.int64), np.arange(len(blocks)), :]


def sample_tree_philox(transition, n_layer, n_child, q, batch, seed, tree_offset=0, root=None, p_y=None,
                       root_uniform=False):
    """Philox-mode twin of oracle.ghm_oracle.sample_tree: same traversal, Philox words instead of U.

    Counter layout (csrc/ghm_tree.cu): root = word 0 of block (level 0, 0); the s leaves under
    depth-(L-1) node j take words c % 4 of blocks (level L, j*ceil(s/4) + c // 4); when s % 4 != 0 and
    L >= 2 node j itself is drawn from the spare word s % 4 of its last leaf block; every other node
    (l, idx) uses word idx & 3 of block (level l, idx >> 2).
    """
    trees = np.arange(batch, dtype=np.uint64) + np.uint64(tree_offset)
    if root is None:
        p = np.full(q, 1.0 / q) if (root_uniform or p_y is None) else np.asarray(p_y, dtype=np.float64)
        r = draw_words(seed, trees, 0, 1)[0]
        root = search(np.broadcast_to(thresholds(p), (batch, q)), r, q)
    root = np.asarray(root, dtype=np.int64)
    values = [root.reshape(1, batch)]
    s, L = n_child, n_layer
    nb = (s + 3) // 4
    spare = (s % 4 != 0) and L >= 2
    for layer in range(1, L + 1):
        prev = values[-1]
        n = s ** layer
        idx = np.arange(n)
        if layer == L:
            j, c = idx // s, idx % s
            words = draw_words_at(seed, trees, L, j * nb + c // 4, c % 4)
        elif layer == L - 1 and spare:
            words = draw_words_at(seed, trees, L, idx * nb + nb - 1, np.full(n, s % 4))
        else:
            words = draw_words_at(seed, trees, layer, idx >> 2, idx & 3)
        cur = np.empty((n, batch), dtype=np.int64)
        cache = {}
        for i in range(n):
            mat = transition[layer - 1][i]
            key = id(mat)
            if key not in cache:
                cache[key] = np.stack([alias_table(mat[a]) for a in range(q)])      # [q parent, q]
            cur[i] = draw_alias(cache[key][prev[i // s]], words[i], q)
        values.append(cur)
    return values


def gauss_noise(seed, tree_offset, batch, n_leaves):
    """Box-Muller normals of ghm_gauss_noise (stream 1): returns float32 [n_leaves, B]."""
    trees = np.arange(batch, dtype=np.uint64) + np.uint64(tree_offset)
    words = draw_words(seed, trees, 0, 2 * n_leaves, stream=1)   # two words per leaf
    w1, w2 = words[0::2], words[1::2]
    u1 = (w1.astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -32)      # all 32 bits of the radial word
    u2 = ((w2 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    return (np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2 * np.pi) * u2)).astype(np.float32)

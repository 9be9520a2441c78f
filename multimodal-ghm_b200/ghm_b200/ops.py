"""Torch-tensor wrappers over libghm_b200's C ABI (device pointers + current CUDA stream).

Every function allocates its outputs with ``torch.empty`` on the model's device, passes raw
``data_ptr()``s and ``torch.cuda.current_stream().cuda_stream`` and returns torch tensors.
Nothing here computes on the CPU: without the CUDA library or a GPU these raise.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from ._lib import check, get_lib

LEAF_I64, LEAF_U8 = 0, 1
ROOT_GIVEN, ROOT_PRIOR, ROOT_UNIFORM, ROOT_SHARED = 0, 1, 2, 3


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    """Raw handle of torch's current stream on the current device (the private fast path costs ~1 us; the public
    ``torch.cuda.current_stream().cuda_stream`` builds a Stream object per call and showed up in the e2e profile)."""
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_NULL_CTX = contextlib.nullcontext()


def _on(device):
    """Device guard that is a no-op when ``device`` is already current (the usual case: one process per GPU)."""
    return _NULL_CTX if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def _leaf_code(t):
    if t.dtype == torch.int64:
        return LEAF_I64
    if t.dtype == torch.uint8:
        return LEAF_U8
    raise TypeError("leaves must be int64 or uint8, got %s" % t.dtype)


def is_translation_invariant(transition, n_child):
    """True when every level tiles the same n_child matrices (reference GenTransition TI mode, :71-76)."""
    for level in transition:
        for i, m in enumerate(level):
            ref = level[i % n_child]
            if m is not ref and not np.array_equal(m, ref):
                return False
    return True


class GhmModel:
    """Device-side tables of one tree family (transition matrices + prior).  Immutable after creation.

    ``transition`` is the reference's structure: list[L] of list[s**(l+1)] of (q,q) float64
    (``GenTransition`` output / ``sampler.transition``).
    """

    def __init__(self, transition, n_layer, n_child, variable_type, p_y=None, device=None):
        lib = get_lib()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        self.L, self.s, self.q = int(n_layer), int(n_child), int(variable_type)
        assert len(transition) == self.L
        self.ti = is_translation_invariant(transition, self.s)
        T = self._pack(transition)
        py = None if p_y is None else np.ascontiguousarray(np.asarray(p_y, dtype=np.float64))
        self.n_leaves = self.s ** self.L
        self.n_edges = sum(self.s ** l for l in range(1, self.L + 1))
        h = C.c_void_p()
        check(lib.ghm_model_create(C.byref(h), self.L, self.s, self.q, int(self.ti), T.ctypes.data_as(C.c_void_p),
                                   py.ctypes.data_as(C.c_void_p) if py is not None else C.c_void_p(0),
                                   self.device.index))
        self._h = h
        self._lib = lib

    def _pack(self, transition):
        """reference list-of-lists -> float64 [n_mat, q, q] (L*s matrices when translation invariant, else E)."""
        n_mat = self.L * self.s if self.ti else sum(self.s ** l for l in range(1, self.L + 1))
        T = np.empty((n_mat, self.q, self.q), dtype=np.float64)      # (filled row by row: np.stack of the list costs 3x more
        k = 0                                                        #  on the per-grid-point path of the p_flip sweeps)
        for l, level in enumerate(transition):
            assert len(level) == self.s ** (l + 1), "level %d has %d matrices" % (l, len(level))
            for m in (level[:self.s] if self.ti else level):
                T[k] = m                                             # shape mismatch raises
                k += 1
        assert k == n_mat
        assert T.shape[1:] == (self.q, self.q)
        return T

    def update(self, transition, p_y=None):
        """Swap in the tables of another sampler of the same shape (one async H2D copy on the current stream)."""
        assert len(transition) == self.L and is_translation_invariant(transition, self.s) == self.ti
        T = self._pack(transition)
        py = None if p_y is None else np.ascontiguousarray(np.asarray(p_y, dtype=np.float64))
        with _on(self.device):
            check(self._lib.ghm_model_update(self._h, T.ctypes.data_as(C.c_void_p),
                                             py.ctypes.data_as(C.c_void_p) if py is not None else C.c_void_p(0),
                                             _stream()))
        self._transition_ref = transition

    @property
    def table_bytes(self):
        return int(self._lib.ghm_model_table_bytes(self._h))

    def dns_workspace_bytes(self, batch):
        return int(self._lib.ghm_bp_dns_workspace_bytes(self._h, int(batch)))

    def nwp_workspace_bytes(self, batch):
        return int(self._lib.ghm_bp_nwp_workspace_bytes(self._h, int(batch)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.ghm_model_destroy(h)
            except Exception:
                pass

    GEMM_F32, GEMM_TF32, GEMM_BF16 = 0, 1, 2

    def set_gemm_mode(self, mode):
        """Arithmetic of the wide-q (q > 16) row-GEMMs: GEMM_F32 (CUDA cores, default), GEMM_TF32 / GEMM_BF16 (tcgen05)."""
        check(self._lib.ghm_model_set_gemm_mode(self._h, int(mode)))

    def status(self):
        out = C.c_int(0)
        with _on(self.device):
            check(self._lib.ghm_model_status(self._h, _stream(), C.byref(out)))
        return out.value

    # ---- K1 (+ optional fused K2) ------------------------------------------------------
    def sample(self, batch, root=None, U=None, seed=0, tree_offset=0, root_mode=None, leaf_dtype=torch.int64,
               want_leaves=True, want_root=True, want_post=False, want_root_hd=False):
        """Sample ``batch`` trees.  Returns dict(root, leaves, post, root_hd) (absent -> None).

        ``U`` (float64 [E,B], device) selects parity mode (reference uniforms, bit-exact leaves);
        otherwise Philox keyed by (seed, tree_offset + b).
        """
        B = int(batch)
        dev = self.device
        with _on(dev):
            if root is not None:
                root = torch.as_tensor(root).to(device=dev, dtype=torch.int64).contiguous()
                assert root.numel() == B
                mode = ROOT_GIVEN
            else:
                mode = ROOT_PRIOR if root_mode is None else root_mode
            if U is not None:
                U = torch.as_tensor(U).to(device=dev, dtype=torch.float64).contiguous()
                assert tuple(U.shape) == (self.n_edges, B), "U must be [E=%d, B=%d]" % (self.n_edges, B)
            root_out = torch.empty(B, dtype=torch.int64, device=dev) if want_root else None
            leaves = torch.empty((B, self.n_leaves), dtype=leaf_dtype, device=dev) if want_leaves else None
            post = torch.empty((B, self.q), dtype=torch.float32, device=dev) if want_post else None
            hd = torch.empty((B, self.q), dtype=torch.float32, device=dev) if want_root_hd else None
            check(self._lib.ghm_sample(self._h, B, mode, _ptr(root), _ptr(U), seed, tree_offset, _ptr(root_out),
                                       _ptr(leaves), _leaf_code(leaves) if leaves is not None else LEAF_I64,
                                       _ptr(post), _ptr(hd), _stream()))
        return {"root": root_out, "leaves": leaves, "post": post, "root_hd": hd}

    # ---- K2 ---------------------------------------------------------------------------
    def bp_cls(self, leaves):
        """leaves [B, n_L] (int64/uint8, device) -> (post [B,q] f32, root_hd [B,q] f32)."""
        leaves = leaves.contiguous()
        B = leaves.shape[0]
        assert leaves.shape[1] == self.n_leaves and leaves.device == self.device
        with _on(self.device):
            post = torch.empty((B, self.q), dtype=torch.float32, device=self.device)
            hd = torch.empty((B, self.q), dtype=torch.float32, device=self.device)
            ws = self._workspace(self._lib.ghm_bp_cls_workspace_bytes(self._h, B))
            check(self._lib.ghm_bp_cls(self._h, B, _ptr(leaves), _leaf_code(leaves), _ptr(post), _ptr(hd), _ptr(ws),
                                       _stream()))
        return post, hd

    # ---- K3 ---------------------------------------------------------------------------
    def _workspace(self, nbytes):
        """Scratch for one call, allocated on the CURRENT stream through torch's caching allocator (stream-ordered
        reuse: safe when the model is driven from several streams; nothing is held between calls)."""
        nbytes = int(nbytes)
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device) if nbytes > 0 else None

    def bp_dns(self, z, sigma, ext=None, want_root_bu=False):
        """z f32 [B,n_L], ext f32 [B,q] or None -> posterior mean f32 [B,n_L]  (reference BP_DNS, :467-523).

        ``want_root_bu=True`` returns ``(mean, root_bu [B,q])``: the root's hd_message after the pass, i.e. the
        max-shifted upward message plus ``ext`` without a re-shift (root bu aliases hd in the reference, :501-506)."""
        z = z.contiguous()
        B = z.shape[0]
        assert z.dtype == torch.float32 and z.shape[1] == self.n_leaves and z.device == self.device
        if ext is not None:
            ext = ext.contiguous()
            assert ext.dtype == torch.float32 and tuple(ext.shape) == (B, self.q)
        with _on(self.device):
            mean = torch.empty((B, self.n_leaves), dtype=torch.float32, device=self.device)
            ws = self._workspace(self._lib.ghm_bp_dns_workspace_bytes(self._h, B))
            rbu = torch.empty((B, self.q), dtype=torch.float32, device=self.device) if want_root_bu else None
            check(self._lib.ghm_bp_dns(self._h, B, _ptr(z), float(sigma), _ptr(ext), _ptr(mean), _ptr(rbu), _ptr(ws),
                                       _stream()))
        return (mean, rbu) if want_root_bu else mean

    # ---- K4 ---------------------------------------------------------------------------
    def bp_nwp(self, leaves, ext=None):
        """leaves [B,n_L], ext f32 [B,q] or None -> next-token posteriors f32 [B,n_L-1,q]  (reference :336-463)."""
        leaves = leaves.contiguous()
        B = leaves.shape[0]
        assert leaves.shape[1] == self.n_leaves and leaves.device == self.device
        if ext is not None:
            ext = ext.contiguous()
            assert ext.dtype == torch.float32 and tuple(ext.shape) == (B, self.q)
        with _on(self.device):
            pp = torch.empty((B, self.n_leaves - 1, self.q), dtype=torch.float32, device=self.device)
            ws = self._workspace(self._lib.ghm_bp_nwp_workspace_bytes(self._h, B))
            check(self._lib.ghm_bp_nwp(self._h, B, _ptr(leaves), _leaf_code(leaves), _ptr(ext), _ptr(pp), _ptr(ws),
                                       _stream()))
        return pp

    # ---- K5 ---------------------------------------------------------------------------
    @staticmethod
    def _ptr_array(tensors):
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return arr

    def guides_cls(self, leaves):
        """-> (L guide tensors f32 [B,n_L,q] (depth L-1..0), post [B,q], root_hd [B,q])  (reference :533-549)."""
        leaves = leaves.contiguous()
        B = leaves.shape[0]
        with _on(self.device):
            guides = [torch.empty((B, self.n_leaves, self.q), dtype=torch.float32, device=self.device)
                      for _ in range(self.L)]
            post = torch.empty((B, self.q), dtype=torch.float32, device=self.device)
            hd = torch.empty((B, self.q), dtype=torch.float32, device=self.device)
            check(self._lib.ghm_guides_cls(self._h, B, _ptr(leaves), _leaf_code(leaves), self._ptr_array(guides),
                                           _ptr(post), _ptr(hd), _stream()))
        return guides, post, hd

    def guides_dns(self, z, sigma, ext=None):
        """-> (2L+1 guide tensors, mean [B,n_L])  (reference :551-590)."""
        z = z.contiguous()
        B = z.shape[0]
        q, nL = self.q, self.n_leaves
        if ext is not None:
            ext = ext.contiguous()
        with _on(self.device):
            widths = [2 * q] * self.L + [2 * q] + [3 * q] * self.L
            guides = [torch.empty((B, nL, w), dtype=torch.float32, device=self.device) for w in widths]
            mean = torch.empty((B, nL), dtype=torch.float32, device=self.device)
            ws = self._workspace(self._lib.ghm_guides_dns_workspace_bytes(self._h, B))
            check(self._lib.ghm_guides_dns(self._h, B, _ptr(z), float(sigma), _ptr(ext), self._ptr_array(guides),
                                           _ptr(mean), _ptr(ws), _stream()))
        return guides, mean

    def guides_nwp(self, leaves, ext=None):
        """-> (2L+1 guide tensors [B,n_L-1,{q,2q,...,2q,q,...}], pp [B,n_L-1,q])  (reference :357-459)."""
        leaves = leaves.contiguous()
        B = leaves.shape[0]
        q, nL = self.q, self.n_leaves
        if ext is not None:
            ext = ext.contiguous()
        with _on(self.device):
            widths = [q] + [2 * q] * self.L + [q] * self.L
            guides = [torch.empty((B, nL - 1, w), dtype=torch.float32, device=self.device) for w in widths]
            pp = torch.empty((B, nL - 1, q), dtype=torch.float32, device=self.device)
            ws = self._workspace(self._lib.ghm_guides_nwp_workspace_bytes(self._h, B))
            check(self._lib.ghm_guides_nwp(self._h, B, _ptr(leaves), _leaf_code(leaves), _ptr(ext),
                                           self._ptr_array(guides), _ptr(pp), _ptr(ws), _stream()))
        return guides, pp

    # ---- Gaussian observations ---------------------------------------------------------
    def gauss_noise(self, leaves, sigma, seed=0, tree_offset=0):
        """z = leaves + sigma*N(0,1) (Philox stream 1) -> f32 [B, n_L]."""
        leaves = leaves.contiguous()
        B = leaves.shape[0]
        with _on(self.device):
            z = torch.empty((B, self.n_leaves), dtype=torch.float32, device=self.device)
            check(self._lib.ghm_gauss_noise(self._h, B, _ptr(leaves), _leaf_code(leaves), float(sigma), seed,
                                            tree_offset, _ptr(z), _stream()))
        return z


IMAGE_SEED_XOR = 0x1234567887654321


def sample_into(model, batch, root_mode, root_in, seed, tree_offset, root_out, leaves_out, post_out, root_hd_out):
    """Philox-mode ghm_sample into caller-owned device tensors (views allowed when contiguous): no allocation."""
    for t in (root_in, root_out, leaves_out, post_out, root_hd_out):
        assert t is None or t.is_contiguous()
    with _on(model.device):
        check(model._lib.ghm_sample(model._h, int(batch), root_mode, _ptr(root_in), C.c_void_p(0), seed, tree_offset,
                                    _ptr(root_out), _ptr(leaves_out),
                                    _leaf_code(leaves_out) if leaves_out is not None else LEAF_I64,
                                    _ptr(post_out), _ptr(root_hd_out), _stream()))


def sample_mixed_into(model, batch, n_given, root_in, seed, tree_offset, root_out, leaves_out, post_out, root_hd_out):
    """ghm_sample_mixed: trees [0, n_given) take ``root_in``, the rest draw uniform roots (ClipSampler image layout)."""
    for t in (root_in, root_out, leaves_out, post_out, root_hd_out):
        assert t is None or t.is_contiguous()
    with _on(model.device):
        check(model._lib.ghm_sample_mixed(model._h, int(batch), int(n_given), _ptr(root_in), seed, tree_offset,
                                          _ptr(root_out), _ptr(leaves_out),
                                          _leaf_code(leaves_out) if leaves_out is not None else LEAF_I64,
                                          _ptr(post_out), _ptr(root_hd_out), _stream()))


def sample_paired_into(model, batch, n_shared, root_seed, seed, tree_offset, root_out, leaves_out, post_out, root_hd_out):
    """ghm_sample_paired: trees [0, n_shared) re-draw the partner modality's uniform root (Philox key ``root_seed``),
    the rest draw their own; no dependency on the partner's launch, so the two may run on different streams."""
    for t in (root_out, leaves_out, post_out, root_hd_out):
        assert t is None or t.is_contiguous()
    with _on(model.device):
        check(model._lib.ghm_sample_paired(model._h, int(batch), int(n_shared), root_seed, seed, tree_offset,
                                           _ptr(root_out), _ptr(leaves_out),
                                           _leaf_code(leaves_out) if leaves_out is not None else LEAF_I64,
                                           _ptr(post_out), _ptr(root_hd_out), _stream()))


def sample_blocked_into(model, batch, blk_len, blk_stride, root_mode, n_given, root_in, root_seed, seed, tree_offset, root_out,
                        leaves_out, post_out, root_hd_out):
    """ghm_sample_blocked: one launch for a shard of a block-structured batch (local tree b -> global Philox index
    tree_offset + (b // blk_len) * blk_stride + b % blk_len)."""
    for t in (root_in, root_out, leaves_out, post_out, root_hd_out):
        assert t is None or t.is_contiguous()
    with _on(model.device):
        check(model._lib.ghm_sample_blocked(model._h, int(batch), int(blk_len), int(blk_stride), int(root_mode), int(n_given),
                                            _ptr(root_in), root_seed, seed, tree_offset, _ptr(root_out), _ptr(leaves_out),
                                            _leaf_code(leaves_out) if leaves_out is not None else LEAF_I64,
                                            _ptr(post_out), _ptr(root_hd_out), _stream()))


def clip_bayes_into(text, image, n, K, pair_lo, pair_hi, seed, tree_offset, t_root, t_leaves, i_leaves, t_pp, i_pp, sums,
                    side_stream=None):
    """ghm_clip_bayes: sample both modalities (+ fused BP) and accumulate the contrastive risk of pairs [pair_lo, pair_hi)
    into ``sums`` with ONE library call; the image launch runs on ``side_stream`` (a torch.cuda.Stream) when given."""
    code = _leaf_code(t_leaves) if t_leaves is not None else LEAF_I64
    with _on(text.device):
        check(text._lib.ghm_clip_bayes(text._h, image._h, int(n), int(K), int(pair_lo), int(pair_hi), seed, tree_offset,
                                       _ptr(t_root), _ptr(t_leaves), _ptr(i_leaves), code, _ptr(t_pp), _ptr(i_pp), _ptr(sums),
                                       _stream(), C.c_void_p(side_stream.cuda_stream if side_stream is not None else 0)))
    return sums


def new_sums(device):
    """Zeroed {sum, sum of squares, count} accumulator (float64[3]) for the risk kernels."""
    return torch.zeros(3, dtype=torch.float64, device=device)


def risk_clip(t_pp, i_pp, n, K, q, sums=None, pair_lo=0, pair_hi=None):
    """Accumulate the symmetric K-way Bayes CLIP loss of pairs [pair_lo, pair_hi) into ``sums``."""
    lib = get_lib()
    t_pp, i_pp = t_pp.contiguous(), i_pp.contiguous()
    assert t_pp.dtype == torch.float32 and i_pp.dtype == torch.float32
    assert tuple(t_pp.shape) == (n * (K + 1), q) and tuple(i_pp.shape) == (n * (K + 1), q)
    pair_hi = n if pair_hi is None else pair_hi
    with _on(t_pp.device):
        if sums is None:
            sums = new_sums(t_pp.device)
        check(lib.ghm_risk_clip(_ptr(t_pp), _ptr(i_pp), n, K, q, pair_lo, pair_hi, _ptr(sums), _stream()))
    return sums


def risk_cdm(mean, leaves, sums=None):
    """Accumulate per-tree sum_leaf (mean - x)^2 into ``sums``."""
    lib = get_lib()
    mean, leaves = mean.contiguous(), leaves.contiguous()
    assert mean.dtype == torch.float32 and mean.shape == leaves.shape
    with _on(mean.device):
        if sums is None:
            sums = new_sums(mean.device)
        check(lib.ghm_risk_cdm(_ptr(mean), _ptr(leaves), _leaf_code(leaves), mean.shape[0], mean.shape[1],
                               _ptr(sums), _stream()))
    return sums


def risk_ce(pp, target, sums=None, target_stride=1, target_offset=0, row_group=1):
    """Accumulate -log pp[r, target(r)] over rows; target(r) = target[(r//g)*stride + offset + r%g]."""
    lib = get_lib()
    pp, target = pp.contiguous(), target.contiguous()
    q = pp.shape[-1]
    rows = pp.numel() // q
    with _on(pp.device):
        if sums is None:
            sums = new_sums(pp.device)
        check(lib.ghm_risk_ce(_ptr(pp), _ptr(target), _leaf_code(target), rows, q, target_stride, target_offset,
                              row_group, _ptr(sums), _stream()))
    return sums


def risk_zsc(text_model, i_pp, t_leaves, sums=None):
    """Accumulate the zero-shot CE: image root posterior pushed down the text tree's leftmost path vs the first text leaf."""
    i_pp, t_leaves = i_pp.contiguous(), t_leaves.contiguous()
    assert i_pp.dtype == torch.float32 and i_pp.shape[0] == t_leaves.shape[0]
    with _on(i_pp.device):
        if sums is None:
            sums = new_sums(i_pp.device)
        check(text_model._lib.ghm_risk_zsc(text_model._h, i_pp.shape[0], _ptr(i_pp), _ptr(t_leaves), _leaf_code(t_leaves),
                                           _ptr(sums), _stream()))
    return sums


def mean_se(sums, se_count=None):
    """(mean, std/sqrt(se_count)) from {sum, sumsq, count}; population std like np.std (reference :41,817,894)."""
    s1, s2, c = (float(x) for x in sums.tolist())
    mean = s1 / c
    var = max(s2 / c - mean * mean, 0.0)
    return mean, (var ** 0.5) / ((se_count if se_count else c) ** 0.5)


def host_clip_bayes(text, image, n, K=4, seed=0, tree_offset=0, leaves_out=None, pp_out=None):
    """ClipSampler.get_Bayes through the HOST-buffer C entry point (its own stream, H2D/D2H inside).

    ``leaves_out`` = (t_leaves, i_leaves) and ``pp_out`` = (t_pp, i_pp): optional pre-allocated
    (ideally pinned) CPU tensors to receive what ClipSampler.get_batch returns.
    """
    lib = get_lib()
    sums = np.zeros(3, dtype=np.float64)
    tl = il = tp = ip = None
    code = LEAF_I64
    if leaves_out is not None:
        tl, il = leaves_out
        code = _leaf_code(tl)
    if pp_out is not None:
        tp, ip = pp_out
    check(lib.ghm_host_clip_bayes(text._h, image._h, n, K, seed, tree_offset, sums.ctypes.data_as(C.c_void_p),
                                  _ptr(tl), _ptr(il), code, _ptr(tp), _ptr(ip)))
    return sums

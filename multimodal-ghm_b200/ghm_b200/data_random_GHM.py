"""Call-compatible mirror of the reference module ``ghmclip.data.data_random_GHM``.

Same public names, constructor / method signatures, return-tuple structure, shapes and dtypes
as the reference (src/ghmclip/data/data_random_GHM.py; shape dump in SURVEY.md Appendix B), but
sampling and every belief-propagation pass run in libghm_b200's sm_100a kernels.  There is no
CPU fallback: constructing a tree or sampler without the CUDA library / a GPU raises.

Two RNG modes
  rng="numpy"  (default, PARITY): the host draws exactly what the reference draws from NumPy's
      global legacy stream, in the reference's order (seed in the ctor, GenTransition, root
      ``choice``, ``rand`` per edge, ``randn`` noise), uploads the uniforms and the f64-CDF sampler
      kernel reproduces the reference's leaves bit-for-bit.  Downstream draws stay in lock-step.
  rng="philox": roots, trees and noise are drawn on the device by Philox4x32-10 keyed with
      (seed, global tree index); nothing random touches the host.  This is the throughput mode.

Extra keyword arguments (all optional, defaults preserve reference behaviour):
  device=  CUDA device used for the computation (default: current device)
  rng=, seed=  as above.
Deliberate deviations, each mirrored from SURVEY.md section 8 / Appendix A:
  * ``GHMTree.T_value`` holds the root and the leaf level only (interior levels are ``None``): no
    caller in the reference reads them and they never leave the SM in the fused kernel.
  * ``Node`` objects exist for ``root_node`` / ``leaves_nodes`` only (no per-node message graph).
  * ``BP_dummy_NWP`` / ``BP_NWP`` (dead code upstream, :223-334) raise NotImplementedError.
"""
import collections
import weakref

import numpy as np
import torch
import torch.nn as nn  # noqa: F401  (the reference re-exports these through `import *`)
from tqdm import tqdm  # noqa: F401

from . import ops
from .sharding import all_reduce_sums, dist_info, mean_se_from_sums, shard_range

__all__ = ["PPCLIPLoss", "GenTransition", "Node", "GHMTree", "SingleSampler", "DoubleSampler",
           "ClassificationSampler", "DenoiseSampler", "ClipSampler", "clip_loss_compute",
           "ConditionalDenoiseSampler", "NextWordPredictSampler", "np", "torch", "nn", "tqdm"]
# additions (not in the reference module): zeroshot_bayes


def _cuda_device(device=None):
    if device is None or (isinstance(device, str) and device == "cpu") or \
            (isinstance(device, torch.device) and device.type == "cpu"):
        if not torch.cuda.is_available():
            raise RuntimeError("ghm_b200 needs a CUDA device: this path has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


# ------------------------------------------------------------------------------------------
# module-level functions
# ------------------------------------------------------------------------------------------
def _clip_loss_device(t_pp, i_pp, n_eval, K, variable_type, device=None):
    """(q, n(K+1)) posteriors (numpy or torch) -> (mean, se) via the K6 kernel."""
    dev = _cuda_device(device if device is not None else (t_pp.device if isinstance(t_pp, torch.Tensor) else None))
    t = torch.as_tensor(np.asarray(t_pp) if not isinstance(t_pp, torch.Tensor) else t_pp)
    i = torch.as_tensor(np.asarray(i_pp) if not isinstance(i_pp, torch.Tensor) else i_pp)
    t = t.to(dev).T.to(torch.float32).contiguous()
    i = i.to(dev).T.to(torch.float32).contiguous()
    sums = ops.risk_clip(t, i, int(n_eval), int(K), int(variable_type))
    mean, se = mean_se_from_sums(sums)
    return np.float64(mean), np.float64(se)


def PPCLIPLoss(t_pp, i_pp, n_eval, K=4, variable_type=10):
    """Bayes CLIP objective from (q, n(K+1)) posteriors (reference :13-41)."""
    return _clip_loss_device(t_pp, i_pp, n_eval, K, variable_type)


def clip_loss_compute(ttree_pp, itree_pp, n_eval, K, variable_type):
    """Same math as PPCLIPLoss (reference :819-844)."""
    return _clip_loss_device(ttree_pp, itree_pp, n_eval, K, variable_type)


def _softmax_row(x=np.array([[]])):
    """Row softmax (reference :91-96); host-side, used once per sampler to build the tables."""
    e_x = np.exp(x - np.max(x, axis=1, keepdims=True))
    return e_x / e_x.sum(axis=1, keepdims=True)


def GenTransition(n_layer, n_child, variable_type, p_flip=0.3, flip_scale=1.0, translation_invariance=True,
                  verbose=False):
    """Per-edge transition matrices from NumPy's global stream in the reference's draw order (:43-89).

    Host-side one-off (microseconds next to a batch); the result is what ``ops.GhmModel`` uploads.
    Returns list[n_layer] of list[n_child**(l+1)] of (q,q) float64; TI levels share the same objects.
    """
    q = variable_type
    transition, skeleton = [], []
    for layer in range(n_layer):
        level = []
        if translation_invariance:
            skel = np.identity(q)[np.random.permutation(q), :]
            shared = [(1 - p_flip) * skel + p_flip * _softmax_row(np.random.normal(0, flip_scale, [q, q]))
                      for _ in range(n_child)]
            for _ in range(n_child ** layer):
                level.extend(shared)
            skeleton.append(skel)
        else:
            for _ in range(n_child ** layer):
                for _ in range(n_child):
                    skel = np.identity(q)[np.random.permutation(q), :]
                    level.append((1 - p_flip) * skel
                                 + p_flip * _softmax_row(np.random.normal(0, flip_scale, [q, q])))
        transition.append(level)
    return (transition, skeleton) if verbose else transition


# ------------------------------------------------------------------------------------------
# device-model cache: one GhmModel per (transition list, p_y, device)
# ------------------------------------------------------------------------------------------
_MODEL_CACHE = {}
_MODEL_CACHE_MAX = 64


def _model_for(transition, n_layer, n_child, variable_type, p_y, device):
    py = None if p_y is None else np.asarray(p_y, dtype=np.float64)
    key = (id(transition), device.index, None if py is None else py.tobytes())
    hit = _MODEL_CACHE.get(key)
    if hit is not None and hit[0] is transition:
        return hit[1]
    model = ops.GhmModel(transition, n_layer, n_child, variable_type, p_y=py, device=device)
    if len(_MODEL_CACHE) >= _MODEL_CACHE_MAX:
        _MODEL_CACHE.pop(next(iter(_MODEL_CACHE)))
    _MODEL_CACHE[key] = (transition, model)      # the strong ref keeps id(transition) valid
    return model


# ------------------------------------------------------------------------------------------
# tree
# ------------------------------------------------------------------------------------------
class Node:
    """Lightweight node (reference :100-110).  Only root / leaf nodes are materialised here."""

    def __init__(self, value=None, parent=None, children=None):
        self.value = value
        self.parent = parent
        self.children = children
        self.hd_message = 0


class _LeafColumns:
    """``T_value[-1]``: n_L per-leaf value lists, materialised lazily from the device tensor.

    Behaves like the reference's list of lists: ``len``, iteration, ``[k]`` -> Python list of B
    ints, item assignment (marks the tree dirty so ``build_tree`` re-uploads), ``np.array(x)``.
    """

    def __init__(self, dev_leaves):
        self._dev = dev_leaves          # [B, n_L] device tensor (int64)
        self._host = None               # (n_L, B) int64 ndarray, lazily
        self.dirty = False

    def _np(self):
        if self._host is None:
            self._host = np.ascontiguousarray(self._dev.cpu().numpy().T)
        return self._host

    def __len__(self):
        return self._dev.shape[1]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [row.tolist() for row in self._np()[k]]
        return self._np()[k].tolist()

    def __setitem__(self, k, v):
        arr = self._np()
        if not arr.flags.writeable:
            self._host = arr = arr.copy()
        arr[k] = np.asarray(v, dtype=np.int64)
        self.dirty = True

    def __iter__(self):
        return (row.tolist() for row in self._np())

    def __array__(self, dtype=None, copy=None):
        a = self._np()
        return a.astype(dtype) if dtype is not None else a


class _RootLevel:
    """``T_value[0]``: a one-element list holding the (B,) root array, copied from the device on first access so
    that sampling a tree never synchronises the host (the training feed)."""

    def __init__(self, tree):
        # weak: the tree owns this object (tree.T_value[0]); a strong back-reference would make every GHMTree a
        # reference cycle, and its device tensors (170 MB of leaves per 262 144 trees) would then wait for the cyclic
        # garbage collector instead of being recycled by the caching allocator as soon as the tree is dropped
        self._tree_ref = weakref.ref(tree)

    @property
    def _tree(self):
        t = self._tree_ref()
        if t is None:
            raise ReferenceError("the GHMTree this root level belongs to has been released")
        return t

    def __len__(self):
        return 1

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._tree._root_np()][k]
        if k not in (0, -1):
            raise IndexError("list index out of range")
        return self._tree._root_np()

    def __setitem__(self, k, v):
        if k not in (0, -1):
            raise IndexError("list assignment index out of range")
        self._tree._root_override = np.asarray(v)

    def __iter__(self):
        return iter([self._tree._root_np()])


class GHMTree:
    """Sampled GHM tree + exact BP (reference :112-613), backed by device tensors.

    Device state: ``_leaves`` int64 [B, n_L], ``_root`` int64 [B], after BP: ``_post`` / ``_root_hd``
    f32 [B, q], ``_mean`` f32 [B, n_L].  NumPy views are materialised on attribute access.
    """

    def __init__(self, n_layer=4, n_child=3, variable_type=10, p_y=np.ones(10) / 10, p_flip=0.3, transition=None,
                 batch_size=128, build_tree=False, root=None, device=None, rng="numpy", seed=0, tree_offset=0,
                 _model=None):
        self.variable_type = variable_type
        self.posterior_probability_CLS = None
        self.posterior_mean_DNS = None
        self.n_layer = n_layer
        self.n_child = n_child
        self.p_y = p_y
        self.p_flip = p_flip
        self.transition = transition
        self.batch_size = batch_size
        self.root = root
        self.build_tree_flag = build_tree
        self.dns_flag = False
        self.cls_flag = False
        self.device = _cuda_device(device)
        self.rng, self.seed, self.tree_offset = rng, seed, tree_offset
        self._model_hint = _model
        self._root_hd = None          # device f32 [B, q]; root_node.hd_message (BP_CLS: shifted hd; BP_DNS: hd + ext, :504-506)
        self._root_hd_host = None     # its (q, B) float64 host copy, made on first access
        self._post = self._mean = None
        self._dns_state = None        # (z, sigma, ext) of the last BP_DNS, for guided_info
        self._cls_guides = None
        self.gen_values()
        if self.build_tree_flag:
            self.build_tree()

    # -- model -------------------------------------------------------------------------
    @property
    def model(self):
        m = self._model_hint
        if m is not None and getattr(m, "_transition_ref", None) is self.transition:
            return m
        m = _model_for(self.transition, self.n_layer, self.n_child, self.variable_type, self.p_y, self.device)
        m._transition_ref = self.transition
        self._model_hint = m
        return m

    # -- sampling (reference gen_values, :145-165) ----------------------------------------
    def gen_values(self):
        m = self.model
        B = self.batch_size
        if self.rng == "numpy":
            root = self.root
            if root is None:
                root = np.random.choice(self.variable_type, size=B, p=self.p_y)
            root = np.asarray(root)
            # one rand(E, B) == the reference's E sequential rand(B, 1) calls (same stream)
            U = np.random.rand(m.n_edges, B)
            out = m.sample(B, root=torch.from_numpy(root.astype(np.int64)), U=torch.from_numpy(U))
            root_host = root
        elif self.rng == "philox":
            if self.root is not None:
                out = m.sample(B, root=torch.as_tensor(self.root), seed=self.seed, tree_offset=self.tree_offset)
            else:
                out = m.sample(B, seed=self.seed, tree_offset=self.tree_offset, root_mode=ops.ROOT_PRIOR)
            root_host = None
        else:
            raise ValueError("rng must be 'numpy' or 'philox'")
        self._leaves = out["leaves"]
        self._root = out["root"]
        self._root_host = root_host
        self._root_override = None
        self.T_value = [_RootLevel(self)] + [None] * (self.n_layer - 1) + [_LeafColumns(self._leaves)]

    def _root_np(self):
        if self._root_override is not None:
            return self._root_override
        if self._root_host is None:
            self._root_host = self._root.cpu().numpy()
        return self._root_host

    # -- (re)build: pick up caller edits of T_value (reference build_tree, :167-183) -----------
    def build_tree(self):
        self.posterior_probability_CLS = None
        self.posterior_mean_DNS = None
        self._post = self._mean = self._root_hd = self._root_hd_host = None
        self._cls_guides = None
        lv = self.T_value[-1]
        if not isinstance(lv, _LeafColumns) or lv.dirty or lv._dev is not self._leaves:
            arr = np.asarray(lv, dtype=np.int64)                      # (n_L, B)
            if arr.ndim != 2 or arr.shape[0] != self.n_child ** self.n_layer:
                raise ValueError("T_value[-1] must hold %d leaf columns" % (self.n_child ** self.n_layer))
            self._leaves = torch.from_numpy(np.ascontiguousarray(arr.T)).to(self.device)
            self.batch_size = arr.shape[1]
            cols = _LeafColumns(self._leaves)
            cols._host = arr
            self.T_value[-1] = cols
        lvl0 = self.T_value[0]
        r0 = None
        if not isinstance(lvl0, _RootLevel):                 # the caller replaced the whole root level
            r0 = np.asarray(lvl0[0])
        elif self._root_override is not None:                # ... or assigned T_value[0][0]
            r0 = self._root_override
        if r0 is not None:
            self._root_host, self._root_override = np.asarray(r0), None
            self._root = torch.from_numpy(self._root_host.astype(np.int64)).to(self.device)
            self.T_value[0] = _RootLevel(self)
        self.Tree = None

    # -- BP: root posterior (reference BP_CLS, :185-221) --------------------------------------
    def BP_CLS(self):
        self._post, self._root_hd = self.model.bp_cls(self._leaves)
        self._root_hd_host = None
        self.posterior_probability_CLS = self._post.T.double().cpu().numpy()
        self.cls_flag = True
        return self.posterior_probability_CLS

    # -- BP: Gaussian denoiser (reference BP_DNS, :467-523) -------------------------------------
    def _to_dev_bq(self, ext):
        """external message (q,B) numpy/torch -> device f32 [B,q]."""
        if ext is None:
            return None
        e = ext if isinstance(ext, torch.Tensor) else torch.from_numpy(np.asarray(ext))
        return e.to(self.device).T.to(torch.float32).contiguous()

    def BP_DNS(self, z, sigma=1.0, external_hd_message=None):
        zz = z if isinstance(z, torch.Tensor) else torch.from_numpy(np.asarray(z))
        zd = zz.to(self.device).T.to(torch.float32).contiguous()              # [B, n_L]
        ext = self._to_dev_bq(external_hd_message)
        # root_node.hd_message after BP_DNS is hd + ext (the reference's root bu aliases hd, :501-506)
        self._mean, self._root_hd = self.model.bp_dns(zd, float(sigma), ext, want_root_bu=True)
        self._root_hd_host = None
        self._dns_state = (zd, float(sigma), ext)
        self.posterior_mean_DNS = self._mean.T.double().cpu().numpy()
        self.dns_flag = True
        return self.posterior_mean_DNS

    # -- BP: next-token posterior (reference BP_NWP_autoregressive, :336-463) -------------------
    def BP_NWP_autoregressive(self, guide_info=False, device="cpu", external_hd_message=None, verbose=False, pos=3):
        ext = self._to_dev_bq(external_hd_message)
        if guide_info:
            guides, pp = self.model.guides_nwp(self._leaves, ext)
            guides = [g.to(device) for g in guides]
        else:
            pp, guides = self.model.bp_nwp(self._leaves, ext), []
        return pp.to(device), guides

    def BP_dummy_NWP(self, position, external_hd_message=None):
        raise NotImplementedError("BP_dummy_NWP is dead code upstream (reference :223-272); "
                                  "use BP_NWP_autoregressive (no CPU fallback is provided)")

    def BP_NWP(self, position, external_hd_message=None):
        raise NotImplementedError("BP_NWP is dead code upstream (reference :274-334); "
                                  "use BP_NWP_autoregressive (no CPU fallback is provided)")

    # -- guide tensors (reference guided_info, :526-592) ---------------------------------------
    def guided_info(self, device="cpu"):
        if self.cls_flag:
            guides, post, hd = self.model.guides_cls(self._leaves)
        elif self.dns_flag:
            z, sigma, ext = self._dns_state
            guides, _ = self.model.guides_dns(z, sigma, ext)
        else:
            return []
        return [g.to(device) for g in guides]

    # -- properties (reference :595-613) ----------------------------------------------------
    @property
    def root_node(self):
        node = Node(self.root_value)
        if self._root_hd is not None:       # (q, B) float64: the cross-modal "external" message (:871,919)
            if self._root_hd_host is None:
                self._root_hd_host = self._root_hd.T.double().cpu().numpy()
            node.hd_message = self._root_hd_host
        return node

    @property
    def leaves_nodes(self):
        return [Node(v) for v in self.T_value[-1]]

    @property
    def leaves_values(self):
        return self.T_value[-1]

    @property
    def root_value(self):
        return self.T_value[0][0]


# ------------------------------------------------------------------------------------------
# samplers (reference :617-942)
# ------------------------------------------------------------------------------------------
def _posterior_out(t, async_):
    """4th element of the get_batch tuples: float64 NumPy like the reference (a blocking device -> host copy), or --
    ``async_=True``, the training feed -- the same values as a float64 DEVICE tensor with no host synchronisation."""
    t = t.double()
    return t if async_ else t.cpu().numpy()


class LazyRisk:
    """Handle of a risk evaluation that is still running on the device (``get_Bayes(lazy=True)``).

    The {sum, sum of squares, count} accumulator is copied into pinned host memory on the issuing stream and an
    event is recorded; ``result()`` waits for that event only.  A caller that sweeps a grid (one evaluation per
    p_flip, figures/eval-clip-ood.py:73-79) enqueues evaluation k+1 before reading evaluation k, so the GPU never
    idles behind the host.  ``finish`` maps the three doubles to the method's return tuple."""

    # Pinned landing slots are recycled: cudaHostAlloc per evaluation costs more than the evaluation's launches (and
    # synchronises the device).  A slot is handed out again only after POOL further evaluations were issued.
    POOL = 256
    _slots = None
    _next = 0

    @classmethod
    def _slot(cls):
        if cls._slots is None:
            cls._slots = torch.empty((cls.POOL, 3), dtype=torch.float64).pin_memory()
        i = cls._next
        cls._next = (i + 1) % cls.POOL
        return cls._slots[i]

    def __init__(self, sums, finish):
        self._host = self._slot()
        self._host.copy_(sums, non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record(torch.cuda.current_stream(sums.device))
        self._finish = finish
        self._value = None
        self._sums = None

    def done(self):
        return self._ev.query()

    def sums(self):
        """The raw {sum, sum of squares, count} as host floats (waits for the evaluation)."""
        if self._sums is None:
            self._ev.synchronize()
            self._sums = [float(x) for x in self._host.tolist()]     # copied out: the slot is recycled later
        return self._sums

    def result(self):
        if self._value is None:
            self._value = self._finish(torch.tensor(self.sums(), dtype=torch.float64))
        return self._value


class _SamplerBase:
    #: pairs / trees per internal launch of get_Bayes (bounds the device workspace; any n_eval is accepted)
    bayes_chunk = 262144

    def _chunks(self, n, per_unit_bytes=0):
        """[(start, size)] covering n units with at most ``bayes_chunk`` units (and about 4 GiB of scratch) per piece."""
        c = int(self.bayes_chunk)
        if per_unit_bytes > 0:
            c = max(1024, min(c, (4 << 30) // int(per_unit_bytes)))
        return [(s0, min(c, n - s0)) for s0 in range(0, n, c)]

    def _init_backend(self, device, rng, seed):
        self.device = _cuda_device(device)
        if rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' or 'philox'")
        self.rng, self.seed = rng, int(seed)
        self.tree_offset = 0          # global index of the next Philox tree (explicit, resumable RNG state)

    def _eval_stream(self):
        """Lazy risk evaluations alternate between two internal streams, so that evaluation k+1 (after a table swap: the
        model tables are double buffered, ``ghm_model_update``) starts while the tail of evaluation k drains -- the same
        overlap ``bench.py``'s device-resident loop gets from its two pipelines."""
        sts = getattr(self, "_eval_streams", None)
        if sts is None:
            sts = self._eval_streams = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
            self._eval_turn = 0
            self._lazy_events = collections.deque(maxlen=4)       # (event, table version the evaluation read)
        self._eval_turn ^= 1
        return sts[self._eval_turn]

    def _fence_tables(self):
        """Called before the model tables are swapped to version v+1, which overwrites the slab of version v-1 (double
        buffering): the current stream waits for every lazy evaluation that read version v-1 or older (normally long
        complete), then the version counter advances."""
        v = getattr(self, "_table_version", 0)
        cur = None                                               # (looked up once: torch.cuda.current_stream costs ~9 us a call)
        for ev, ver in getattr(self, "_lazy_events", ()):
            if ver <= v - 1:
                if cur is None:
                    cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
        self._table_version = v + 1

    def _side_stream(self):
        """Stream of the image-side launch; lazy evaluations get one per evaluation stream so that they do not queue behind
        each other."""
        sides = getattr(self, "_sides", None)
        if sides is None:
            sides = self._sides = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
        return sides[getattr(self, "_eval_turn", 0)]

    def _advance(self, n):
        off = self.tree_offset
        self.tree_offset += int(n)
        return off


class SingleSampler(_SamplerBase):
    """Single-tree sampler (reference :617-639)."""

    def __init__(self, n_layer, n_child, p_y, p_flip, flip_scale=1.0, variable_type=10, translation_invariance=True,
                 seedtree=42, device=None, rng="numpy", seed=1234):
        self.n_layer, self.n_child, self.p_y, self.p_flip = n_layer, n_child, p_y, p_flip
        self.variable_type, self.translation_invariance = variable_type, translation_invariance
        self.seedtree, self.flip_scale = seedtree, flip_scale
        self._init_backend(device, rng, seed)
        np.random.seed(seedtree)
        self.transition = GenTransition(n_layer, n_child, variable_type, p_flip, flip_scale,
                                        translation_invariance=translation_invariance)
        self.model = ops.GhmModel(self.transition, n_layer, n_child, variable_type, p_y=p_y, device=self.device)
        self.model._transition_ref = self.transition

    def _tree(self, batch_size, root=None):
        return GHMTree(self.n_layer, self.n_child, self.variable_type, self.p_y, self.p_flip, self.transition,
                       batch_size, build_tree=True, root=root, device=self.device, rng=self.rng, seed=self.seed,
                       tree_offset=self._advance(batch_size) if self.rng == "philox" else 0, _model=self.model)

    def get_batch(self, batch_size=128):
        T = self._tree(batch_size)
        return T.T_value[0][0], T.T_value[-1][0]


class DoubleSampler(_SamplerBase):
    """Paired text/image sampler (reference :641-682)."""

    def __init__(self, n_layers, n_childs, p_ys, p_flips, flip_scale=1, variable_type=10, translation_invariance=True,
                 seedtree=42, device=None, rng="numpy", seed=1234):
        self.n_layers, self.n_childs, self.p_ys, self.p_flips = n_layers, n_childs, p_ys, p_flips
        self.flip_scale, self.variable_type, self.seedtree = flip_scale, variable_type, seedtree
        self._init_backend(device, rng, seed)
        np.random.seed(seedtree)
        self.t_transition = GenTransition(n_layers[0], n_childs[0], variable_type, p_flips[0], flip_scale,
                                          translation_invariance=translation_invariance)
        self.i_transition = GenTransition(n_layers[1], n_childs[1], variable_type, p_flips[1], flip_scale,
                                          translation_invariance=translation_invariance)
        self.t_model = ops.GhmModel(self.t_transition, n_layers[0], n_childs[0], variable_type, p_y=p_ys[0],
                                    device=self.device)
        self.i_model = ops.GhmModel(self.i_transition, n_layers[1], n_childs[1], variable_type, p_y=p_ys[1],
                                    device=self.device)
        self.t_model._transition_ref = self.t_transition
        self.i_model._transition_ref = self.i_transition

    def reparameterize(self, p_flips, seedtree=None, flip_scale=None, transitions=None):
        """Swap in the transition tables of another (p_flips, seedtree) IN PLACE: equivalent to constructing a new
        sampler of the same shape (same NumPy draws: seed, GenTransition text, GenTransition image -- reference
        :654-658) but re-using the device models (one pinned H2D table copy per modality).  This is what the
        p_flip sweeps of figures/eval-*-ood.py do once per grid point.  ``transitions=(t, i)`` supplies
        pre-generated tables instead of drawing them."""
        self.p_flips = p_flips
        if seedtree is not None:
            self.seedtree = seedtree
        if flip_scale is not None:
            self.flip_scale = flip_scale
        if transitions is None:
            ti = self.t_model.ti
            np.random.seed(self.seedtree)
            t_tr = GenTransition(self.n_layers[0], self.n_childs[0], self.variable_type, p_flips[0], self.flip_scale,
                                 translation_invariance=ti)
            i_tr = GenTransition(self.n_layers[1], self.n_childs[1], self.variable_type, p_flips[1], self.flip_scale,
                                 translation_invariance=ti)
        else:
            t_tr, i_tr = transitions
        self.t_transition, self.i_transition = t_tr, i_tr
        self._fence_tables()
        self.t_model.update(t_tr, self.p_ys[0])
        self.i_model.update(i_tr, self.p_ys[1])
        return self

    # modality 0 = text, 1 = image; the image modality draws from an independent Philox key
    def _tree(self, which, batch_size, root=None, tree_offset=0):
        tr, mo = (self.t_transition, self.t_model) if which == 0 else (self.i_transition, self.i_model)
        seed = self.seed if which == 0 else self.seed ^ ops.IMAGE_SEED_XOR
        return GHMTree(self.n_layers[which], self.n_childs[which], self.variable_type, self.p_ys[which],
                       self.p_flips[which], tr, batch_size, build_tree=True, root=root, device=self.device,
                       rng=self.rng, seed=seed, tree_offset=tree_offset, _model=mo)

    def _shared_root(self, batch_size, off):
        """np.random.choice(q, size=B): uniform, ignores p_ys (reference :674,758,858,906)."""
        if self.rng == "numpy":
            return np.random.choice(self.variable_type, size=batch_size)
        return None

    def _paired_trees(self, batch_size, text_bp=False, image_bp=False):
        """Shared-root text / image trees (:858-861).  Philox mode, q <= 16: ``text_bp`` / ``image_bp`` fuse BP_CLS into
        the sampling launch of that modality (the leaves are absorbed from registers instead of being re-read) and
        leave ``_post`` / ``_root_hd`` on the tree."""
        off = self._advance(batch_size) if self.rng == "philox" else 0
        if self.rng == "numpy":
            root = np.random.choice(self.variable_type, size=batch_size)
            text_tree = self._tree(0, batch_size, root=root)
            image_tree = self._tree(1, batch_size, root=root)
        else:
            fuse = self.variable_type <= 16
            tb, ib = text_bp and fuse, image_bp and fuse
            out = self.t_model.sample(batch_size, seed=self.seed, tree_offset=off, root_mode=ops.ROOT_UNIFORM,
                                      want_post=tb, want_root_hd=tb)
            text_tree = _tree_from_device(self, 0, out, batch_size)
            iout = self.i_model.sample(batch_size, root=out["root"], seed=self.seed ^ ops.IMAGE_SEED_XOR,
                                       tree_offset=off, want_post=ib, want_root_hd=ib)
            image_tree = _tree_from_device(self, 1, iout, batch_size)
            root = None
        return root, text_tree, image_tree

    def get_batch(self, batch_size=128):
        off = self._advance(batch_size) if self.rng == "philox" else 0
        text_tree = self._tree(0, batch_size, tree_offset=off)
        image_tree = self._tree(1, batch_size, tree_offset=off)
        return text_tree.T_value[0][0], image_tree.T_value[0][0], text_tree.T_value[-1][0], image_tree.T_value[-1][0]

    def get_zeroshot_batch(self, batch_size=128, return_tree=False):
        root_tree, text_tree, image_tree = self._paired_trees(batch_size)
        text_tree.BP_CLS()
        image_tree.BP_CLS()
        if return_tree:
            return text_tree, image_tree
        return (np.array(text_tree.leaves_values).T, np.array(image_tree.leaves_values).T,
                np.array(text_tree.posterior_probability_CLS).T, np.array(image_tree.posterior_probability_CLS).T,
                np.array(text_tree.root_value))


def zeroshot_bayes(sampler, batch_size=7500):
    """Zero-shot classification Bayes risk of a DoubleSampler: the recipe of figures/eval-zsc-risk.py:66-83
    (get_zeroshot_batch, image root posterior pushed through t_transition[l][0], float32 CE against the first text
    leaf) with the projection and the reduction on the device.  Returns (mean, std / sqrt(n))."""
    _, text_tree, image_tree = sampler._paired_trees(batch_size)
    text_tree.BP_CLS()
    image_tree.BP_CLS()
    sums = ops.risk_zsc(sampler.t_model, image_tree._post, text_tree._leaves)
    mean, se = mean_se_from_sums(sums)
    return np.float64(mean), np.float64(se)


def _tree_from_device(sampler, which, out, batch_size):
    """Wrap already-sampled device tensors (Philox path) in a GHMTree without re-sampling."""
    t = GHMTree.__new__(GHMTree)
    tr, mo = (sampler.t_transition, sampler.t_model) if which == 0 else (sampler.i_transition, sampler.i_model)
    t.variable_type = sampler.variable_type
    t.posterior_probability_CLS = t.posterior_mean_DNS = None
    t.n_layer, t.n_child = sampler.n_layers[which], sampler.n_childs[which]
    t.p_y, t.p_flip, t.transition = sampler.p_ys[which], sampler.p_flips[which], tr
    t.batch_size, t.root, t.build_tree_flag = batch_size, None, True
    t.dns_flag = t.cls_flag = False
    t.device, t.rng, t.seed, t.tree_offset = sampler.device, sampler.rng, sampler.seed, 0
    t._model_hint = mo
    t._root_hd_host = t._mean = t._dns_state = t._cls_guides = None
    t._post, t._root_hd = out.get("post"), out.get("root_hd")
    t._leaves, t._root, t._root_host, t._root_override = out["leaves"], out["root"], None, None
    t.T_value = [_RootLevel(t)] + [None] * (t.n_layer - 1) + [_LeafColumns(t._leaves)]
    t.Tree = None
    return t


class ClassificationSampler(SingleSampler):
    """Root classification from all leaves (reference :685-720)."""

    def __init__(self, n_layer, n_child, p_y, p_flip=0.3, flip_scale=1, variable_type=10, translation_invariance=True,
                 seedtree=42, device=None, rng="numpy", seed=1234):
        super().__init__(n_layer, n_child, p_y, p_flip, flip_scale, variable_type, translation_invariance, seedtree,
                         device=device, rng=rng, seed=seed)

    def get_batch(self, batch_size=128, guide=False, device="cpu", async_=False):
        tree = self._tree(batch_size)
        if not guide:
            # the reference evaluates `None.T` here (:705) -> AttributeError after sampling; keep the error behaviour
            raise AttributeError("'NoneType' object has no attribute 'T' (ClassificationSampler.get_batch needs guide=True)")
        guides, post, _ = self.model.guides_cls(tree._leaves)         # BP_CLS + guided_info in one fused kernel
        return (tree._leaves.to(device), tree._root.to(device), [g.to(device) for g in guides],
                _posterior_out(post, async_))

    def get_Bayes(self, n_eval=10000):
        """Bayes CE of the root (reference :707-720): float32 loss, torch.std (unbiased) / sqrt(n)."""
        tree = self._tree(n_eval)
        tree.BP_CLS()
        sums = ops.risk_ce(tree._post, tree._root)
        s1, s2, c = (float(x) for x in sums.tolist())
        mean = s1 / c
        var = max((s2 - c * mean * mean) / max(c - 1, 1), 0.0)
        return mean, (var ** 0.5) / np.sqrt(n_eval)


class DenoiseSampler(SingleSampler):
    """Denoising noisy leaves of one tree (reference :722-742)."""

    def __init__(self, n_layer, n_child, p_y, p_flip=0.3, sigma=1, flip_scale=1, variable_type=10,
                 translation_invariance=True, seedtree=42, device=None, rng="numpy", seed=1234):
        super().__init__(n_layer, n_child, p_y, p_flip, flip_scale, variable_type, translation_invariance, seedtree,
                         device=device, rng=rng, seed=seed)
        self.sigma = sigma

    def get_batch(self, batch_size=128, guide=False, device="cpu", async_=False):
        tree = self._tree(batch_size)
        if self.rng == "numpy":
            zs = np.random.randn(self.n_child ** self.n_layer, batch_size) * self.sigma + np.asarray(tree.leaves_values)
            zs_dev = torch.from_numpy(zs).to(self.device).T.to(torch.float32).contiguous()
        else:
            zs_dev = self.model.gauss_noise(tree._leaves, self.sigma, seed=self.seed, tree_offset=tree.tree_offset)
        if not guide:                                        # reference: `None.T` after sampling and noise (:742)
            raise AttributeError("'NoneType' object has no attribute 'T' (DenoiseSampler.get_batch needs guide=True)")
        xs = tree._leaves.to(torch.float32).to(device)
        guides, mean = self.model.guides_dns(zs_dev, float(self.sigma), None)
        # the reference does not forward `device` to guided_info here (:737): its guides stay on the CPU.  The
        # training feed (async_=True) keeps them on the device they were computed on.
        guided_info = guides if async_ else [g.to("cpu") for g in guides]
        return zs_dev.to(device), xs, guided_info, _posterior_out(mean, async_)


class ClipSampler(DoubleSampler):
    """Matched / mismatched pairs for the K-way CLIP objective (reference :746-817)."""

    def __init__(self, n_layers, n_childs, p_ys, p_flips, K=4, flip_scale=1, variable_type=10,
                 translation_invariance=True, seedtree=42, device=None, rng="numpy", seed=1234):
        super().__init__(n_layers, n_childs, p_ys, p_flips, flip_scale, variable_type, translation_invariance,
                         seedtree, device=device, rng=rng, seed=seed)
        self.K = K

    def _sample_layout(self, batch_size, want_leaves=True, want_post=False, pair_lo=0, pair_hi=None):
        """Device-side sampling of the block layout [match1 | match2 | K-1 negatives] (:758-764).

        Returns dict(t=..., i=...) of ops.sample outputs (``post`` fused in Philox mode when asked).
        Philox mode draws block ``j`` of pairs [pair_lo, pair_hi) only (multi-GPU sharding on the pair
        index; global tree index = offset + j*n + pair).
        """
        n, K, q = batch_size, self.K, self.variable_type
        B = n * (K + 1)
        if self.rng == "numpy":
            text_root = np.random.choice(q, size=n * (K + 1))
            image_root = np.random.choice(q, size=n * (K - 1))
            image_root = np.append(text_root[:2 * n], image_root)
            Ut = np.random.rand(self.t_model.n_edges, B)
            Ui = np.random.rand(self.i_model.n_edges, B)
            t = self.t_model.sample(B, root=torch.from_numpy(text_root.astype(np.int64)), U=torch.from_numpy(Ut))
            i = self.i_model.sample(B, root=torch.from_numpy(image_root.astype(np.int64)), U=torch.from_numpy(Ui))
            if want_post:
                t["post"], t["root_hd"] = self.t_model.bp_cls(t["leaves"])
                i["post"], i["root_hd"] = self.i_model.bp_cls(i["leaves"])
            return {"t": t, "i": i, "n_local": n}
        pair_hi = n if pair_hi is None else pair_hi
        nl = pair_hi - pair_lo
        off = self._advance(B)
        dev = self.device
        if nl == n:      # whole layout on this device: one launch per modality over contiguous global tree ranges
            t = {"leaves": torch.empty((B, self.t_model.n_leaves), dtype=torch.int64, device=dev) if want_leaves else None,
                 "root": torch.empty(B, dtype=torch.int64, device=dev),
                 "post": torch.empty((B, q), dtype=torch.float32, device=dev) if want_post else None}
            i = {"leaves": torch.empty((B, self.i_model.n_leaves), dtype=torch.int64, device=dev) if want_leaves else None,
                 "root": torch.empty(B, dtype=torch.int64, device=dev),
                 "post": torch.empty((B, q), dtype=torch.float32, device=dev) if want_post else None}
            iseed = self.seed ^ ops.IMAGE_SEED_XOR
            # The image side re-draws the shared roots from the text key (ghm_sample_paired), so the two launches are
            # independent: they go to two streams and their CTAs fill each other's tail waves.
            cur = torch.cuda.current_stream(dev)
            side = self._side_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ops.sample_paired_into(self.i_model, B, 2 * n, self.seed, iseed, off, i["root"], i["leaves"], i["post"], None)
            ops.sample_into(self.t_model, B, ops.ROOT_UNIFORM, None, self.seed, off, t["root"], t["leaves"], t["post"], None)
            cur.wait_stream(side)
            return {"t": t, "i": i, "n_local": n}
        Bl = nl * (K + 1)
        t = {"leaves": torch.empty((Bl, self.t_model.n_leaves), dtype=torch.int64, device=dev) if want_leaves else None,
             "root": torch.empty(Bl, dtype=torch.int64, device=dev),
             "post": torch.empty((Bl, q), dtype=torch.float32, device=dev) if want_post else None}
        i = {"leaves": torch.empty((Bl, self.i_model.n_leaves), dtype=torch.int64, device=dev) if want_leaves else None,
             "root": torch.empty(Bl, dtype=torch.int64, device=dev),
             "post": torch.empty((Bl, q), dtype=torch.float32, device=dev) if want_post else None}
        iseed = self.seed ^ ops.IMAGE_SEED_XOR
        # One launch per modality for the whole shard (ghm_sample_blocked: block j of the local batch is pairs [lo, hi) of
        # block j of the global layout), image roots of the two matched blocks re-drawn from the text key: the two launches
        # are independent and overlap on two streams exactly like the unsharded path.
        cur = torch.cuda.current_stream(dev)
        side = self._side_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ops.sample_blocked_into(self.i_model, Bl, nl, n, ops.ROOT_SHARED, 2 * nl, None, self.seed, iseed, off + pair_lo,
                                    i["root"], i["leaves"], i["post"], None)
        ops.sample_blocked_into(self.t_model, Bl, nl, n, ops.ROOT_UNIFORM, 0, None, 0, self.seed, off + pair_lo,
                                t["root"], t["leaves"], t["post"], None)
        cur.wait_stream(side)
        return {"t": t, "i": i, "n_local": nl}

    def get_batch(self, device="cpu", batch_size=128, guide=False, async_=False):
        r = self._sample_layout(batch_size, want_leaves=True, want_post=False)
        t, i = r["t"], r["i"]
        if guide:
            tg, t_post, _ = self.t_model.guides_cls(t["leaves"])
            ig, i_post, _ = self.i_model.guides_cls(i["leaves"])
            text_guided_info = [g.to(device) for g in tg]
            image_guided_info = [g.to(device) for g in ig]
            t_pp = _posterior_out(t_post, async_)
            i_pp = _posterior_out(i_post, async_)
        else:
            text_guided_info = image_guided_info = t_pp = i_pp = None
        return [t["leaves"].to(device), t["root"].to(device), text_guided_info, t_pp], \
               [i["leaves"].to(device), i["root"].to(device), image_guided_info, i_pp]

    def get_Bayes(self, n_eval=10000, distributed=False, group=None, lazy=False, keep_batch=False):
        """Bayes CLIP loss (reference :786-817).  ``distributed=True`` (Philox only) shards the pair index
        over the initialised process group and all-reduces the 24-byte risk sums.  ``keep_batch=True`` also
        materialises the int64 leaves of the evaluated batch on the device (the reference's get_Bayes builds the
        whole batch through get_batch, :790) and leaves the layout in ``self.last_batch``.  ``lazy=True`` returns a
        ``LazyRisk`` handle instead of blocking on the 24-byte device -> host read."""
        K, q = self.K, self.variable_type
        if distributed:
            if self.rng != "philox":
                raise ValueError("distributed get_Bayes needs rng='philox' (the NumPy stream is serial)")
            rank, world = dist_info(group)
            lo, hi = shard_range(n_eval, rank, world)
        else:
            lo, hi = 0, n_eval
        def finish(s):
            mean, se = mean_se_from_sums(s)
            return np.float64(mean), np.float64(se)

        def evaluate():
            if self.rng == "philox" and q <= 16:
                # ONE library call (ghm_clip_bayes): both sampling launches with the BP fused, fork / join of the side
                # stream and the contrastive reduction -- the Python-level sequence below costs 0.1 ms more per evaluation
                dev, nl = self.device, hi - lo
                Bl = nl * (K + 1)
                off = self._advance(n_eval * (K + 1))
                sums = ops.new_sums(dev)
                t = {"root": torch.empty(Bl, dtype=torch.int64, device=dev), "leaves": None,
                     "post": torch.empty((Bl, q), dtype=torch.float32, device=dev)}
                i = {"root": None, "leaves": None, "post": torch.empty((Bl, q), dtype=torch.float32, device=dev)}
                if keep_batch:
                    t["leaves"] = torch.empty((Bl, self.t_model.n_leaves), dtype=torch.int64, device=dev)
                    i["leaves"] = torch.empty((Bl, self.i_model.n_leaves), dtype=torch.int64, device=dev)
                side = self._side_stream()
                ops.clip_bayes_into(self.t_model, self.i_model, n_eval, K, lo, hi, self.seed, off, t["root"], t["leaves"],
                                    i["leaves"], t["post"], i["post"], sums, side_stream=side)
                # (the call joins `side` back into the current stream, so the allocator's stream-ordered reuse is safe)
                self.last_batch = {"t": t, "i": i, "n_local": nl} if keep_batch else None
            else:
                r = self._sample_layout(n_eval, want_leaves=keep_batch, want_post=True, pair_lo=lo, pair_hi=hi)
                self.last_batch = r if keep_batch else None
                nl = r["n_local"]
                sums = ops.new_sums(self.device)
                if nl > 0:
                    ops.risk_clip(r["t"]["post"], r["i"]["post"], nl, K, q, sums=sums)
            if distributed:
                all_reduce_sums(sums, group)
            return sums
        if not (lazy and self.rng == "philox"):
            sums = evaluate()
            return LazyRisk(sums, finish) if lazy else finish(sums)
        # lazy: the whole evaluation runs on one of two internal streams (ordered after everything already enqueued on the
        # caller's stream, e.g. a table upload); `last_batch` and the result are valid once the handle's event has completed
        cur = torch.cuda.current_stream(self.device)
        ev_stream = self._eval_stream()
        ev_stream.wait_stream(cur)
        with torch.cuda.stream(ev_stream):
            sums = evaluate()
            handle = LazyRisk(sums, finish)
        if len(self._lazy_events) == self._lazy_events.maxlen:
            self._lazy_events[0][0].synchronize()            # at most four evaluations in flight: the fence above stays exact
        self._lazy_events.append((handle._ev, getattr(self, "_table_version", 0)))
        return handle


class ConditionalDenoiseSampler(DoubleSampler):
    """Denoise image leaves conditioned on text (reference :846-894)."""

    def __init__(self, n_layers, n_childs, p_ys, p_flips, sigma=1, flip_scale=1, variable_type=10,
                 translation_invariance=True, seedtree=42, device=None, rng="numpy", seed=1234):
        super().__init__(n_layers, n_childs, p_ys, p_flips, flip_scale, variable_type, translation_invariance,
                         seedtree, device=device, rng=rng, seed=seed)
        self.sigma = sigma

    def _noise(self, image_tree, batch_size):
        """z = x + sigma * N(0, 1) for the image leaves (reference :867): host randn in parity mode, Philox on the device."""
        nLi = self.n_childs[1] ** self.n_layers[1]
        if self.rng == "numpy":
            noise = np.random.randn(nLi, batch_size) * self.sigma + np.asarray(image_tree.leaves_values)
            return torch.from_numpy(noise).to(self.device).T.to(torch.float32).contiguous()
        return self.i_model.gauss_noise(image_tree._leaves, self.sigma, seed=self.seed ^ ops.IMAGE_SEED_XOR,
                                        tree_offset=self.tree_offset - batch_size)

    def _run(self, batch_size):
        """sample pair -> noise -> text BP_CLS -> ext -> image BP_DNS; everything stays on the device."""
        _, text_tree, image_tree = self._paired_trees(batch_size, text_bp=True)
        z = self._noise(image_tree, batch_size)
        if text_tree._root_hd is not None:                   # Philox, q <= 16: BP_CLS ran inside the sampling launch
            t_post, t_hd = text_tree._post, text_tree._root_hd
        else:
            t_post, t_hd = self.t_model.bp_cls(text_tree._leaves)
        mean = self.i_model.bp_dns(z, float(self.sigma), t_hd)
        return text_tree, image_tree, z, t_post, t_hd, mean

    def get_batch(self, batch_size=128, device="cpu", guide=False, async_=False):
        if guide:
            # the guide kernels run the same BP and also return the posteriors: one pass per modality
            _, text_tree, image_tree = self._paired_trees(batch_size)
            z = self._noise(image_tree, batch_size)
            tg, t_post, t_hd = self.t_model.guides_cls(text_tree._leaves)
            ig, mean = self.i_model.guides_dns(z, float(self.sigma), t_hd)
            text_guided_info = [g.to(device) for g in tg]
            image_guided_info = [g.to(device) for g in ig]
        else:
            text_tree, image_tree, z, t_post, t_hd, mean = self._run(batch_size)
            text_guided_info = image_guided_info = None
        return (text_tree._leaves.to(device), text_tree._root.to(device), text_guided_info,
                _posterior_out(t_post.T, async_)), \
               (z.to(device), image_tree._leaves.to(device), image_guided_info, _posterior_out(mean, async_))

    def get_Bayes(self, n_eval=30000, distributed=False, group=None, lazy=False):
        """Bayes MSE (reference :886-894): mean_b sum_leaf (m - x)^2 and np.std / sqrt(n).  Philox mode evaluates
        any ``n_eval`` in pieces of ``bayes_chunk`` pairs (bounded workspace); ``distributed`` / ``lazy`` as in
        ``ClipSampler.get_Bayes``."""
        if distributed:
            if self.rng != "philox":
                raise ValueError("distributed get_Bayes needs rng='philox'")
            rank, world = dist_info(group)
            lo, hi = shard_range(n_eval, rank, world)
            base = self._advance(n_eval)
            self.tree_offset = base + lo
            n_loc = hi - lo
        else:
            n_loc = n_eval
        sums = ops.new_sums(self.device)
        if self.rng == "numpy":                              # parity mode: one draw like the reference
            _, image_tree, _, _, _, mean = self._run(n_loc)
            ops.risk_cdm(mean, image_tree._leaves, sums=sums)
        else:                                                # Philox: bounded pieces, consecutive global tree indices
            per_pair = self.i_model.dns_workspace_bytes(1) + 16 * (self.t_model.n_leaves + self.i_model.n_leaves)
            for _, size in self._chunks(n_loc, per_pair):
                _, image_tree, _, _, _, mean = self._run(size)
                ops.risk_cdm(mean, image_tree._leaves, sums=sums)
        if distributed:
            self.tree_offset = base + n_eval
            all_reduce_sums(sums, group)

        def finish(s):
            mean_v, se = mean_se_from_sums(s)
            return np.float64(mean_v), np.float64(se)
        return LazyRisk(sums, finish) if lazy else finish(sums)


class NextWordPredictSampler(DoubleSampler):
    """Image-conditioned next-word prediction (reference :896-942)."""

    def __init__(self, n_layers, n_childs, p_ys, p_flips, flip_scale=1, variable_type=10, translation_invariance=True,
                 seedtree=42, device=None, rng="numpy", seed=1234):
        super().__init__(n_layers, n_childs, p_ys, p_flips, flip_scale, variable_type, translation_invariance,
                         seedtree, device=device, rng=rng, seed=seed)

    def get_batch(self, batch_size=128, device="cpu", guide=False, async_=False):
        _, text_tree, image_tree = self._paired_trees(batch_size)
        text_leaves = text_tree._leaves.to(device)
        if guide:
            ig, i_post, i_hd = self.i_model.guides_cls(image_tree._leaves)
            tg, pp = self.t_model.guides_nwp(text_tree._leaves, i_hd)
            image_guided_info = [g.to(device) for g in ig]
            text_guided_info = [g.to(device) for g in tg]
        else:
            i_post, i_hd = self.i_model.bp_cls(image_tree._leaves)
            pp = self.t_model.bp_nwp(text_tree._leaves, i_hd)
            image_guided_info = text_guided_info = None
        return (text_leaves[:, :-1], text_leaves[:, 1:], text_guided_info, pp.to(device)), \
               (image_tree._leaves.to(device), image_tree._root.to(device), image_guided_info,
                _posterior_out(i_post, async_))

    def get_Bayes(self, n_eval=30000, distributed=False, group=None, lazy=False):
        """Bayes token CE (reference :931-942): float32 mean; "SE" = torch.std / sqrt(n_eval) (sic).  Chunked /
        ``distributed`` / ``lazy`` as in ``ConditionalDenoiseSampler.get_Bayes``."""
        if distributed:
            if self.rng != "philox":
                raise ValueError("distributed get_Bayes needs rng='philox'")
            rank, world = dist_info(group)
            lo, hi = shard_range(n_eval, rank, world)
            base = self._advance(n_eval)
            self.tree_offset = base + lo
            n_loc = hi - lo
        else:
            n_loc = n_eval
        sums = ops.new_sums(self.device)
        nL = self.t_model.n_leaves
        per_pair = 4 * (nL - 1) * self.variable_type + self.t_model.nwp_workspace_bytes(1) + 16 * (nL + self.i_model.n_leaves)
        pieces = [(0, n_loc)] if self.rng == "numpy" else self._chunks(n_loc, per_pair)
        for _, size in pieces:
            _, text_tree, image_tree = self._paired_trees(size, image_bp=True)
            i_hd = image_tree._root_hd
            if i_hd is None:
                _, i_hd = self.i_model.bp_cls(image_tree._leaves)
            pp = self.t_model.bp_nwp(text_tree._leaves, i_hd)
            ops.risk_ce(pp, text_tree._leaves, sums=sums, target_stride=nL, target_offset=1, row_group=nL - 1)
        if distributed:
            self.tree_offset = base + n_eval
            all_reduce_sums(sums, group)

        def finish(s):
            s1, s2, c = (float(x) for x in s.tolist())
            mean = s1 / c
            var = max((s2 - c * mean * mean) / max(c - 1, 1), 0.0)      # torch.std is the unbiased estimator
            return (torch.tensor(mean, dtype=torch.float32),
                    torch.tensor((var ** 0.5) / np.sqrt(n_eval), dtype=torch.float32))
        return LazyRisk(sums, finish) if lazy else finish(sums)

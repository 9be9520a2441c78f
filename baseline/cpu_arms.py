"""CPU arms of bench.py: the reference's own NumPy implementation of the hot path, timed on the host cores.

Two implementations of every task, selected by ``impl``:

* ``"reference"`` -- the UNMODIFIED reference package installed under the git-ignored ``baseline/_ref`` by
  ``baseline/install_reference.py`` (``ghmclip.data.data_random_GHM``, called through its own public classes);
* ``"port"``      -- ``oracle/ghm_oracle.py``, the array-based NumPy restatement (about 4-5x faster per core than the
  reference because it carries no Python node graph; pinned to the reference bit for bit by tests/test_oracle_golden.py).

Only ``bench.py`` (its ``cpu_baseline`` leg and ``--impl reference``) imports this module; nothing under
``multimodal-ghm_b200/`` does.  Every task returns ``(trees, seconds)`` for ONE process with one BLAS thread;
``fan_out`` runs one task per host core (the reference's own ``&`` fan-out, scripts/experiments/exp_clip_guidedTF.sh:41)
and reports aggregate trees/s = total trees / wall time of the slowest worker.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")

SIGMAS_C3 = (0.1, 0.25, 0.5, 1.0, 2.0, 4.0)


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "ghmclip", "data", "data_random_GHM.py"))


def _ref():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ghmclip.data.data_random_GHM as R          # noqa: E402  (the unmodified reference)
    return R


def _port():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ghm_oracle as O
    return O


def _u(q):
    return np.ones(q) / q


# ---------------------------------------------------------------------------------------------------------------
# tasks: (impl, size, seed) -> (trees, seconds).  Sampler / table construction is outside the timed span on both
# sides; the global NumPy RNG is re-seeded after construction so workers draw different batches.
# ---------------------------------------------------------------------------------------------------------------
def c2_clip(impl, n_eval, seed):
    """ClipSampler([4,4],[3,3],p=.2,K=4).get_Bayes(n_eval): 2*(K+1)*n_eval trees sampled + BP_CLS + contrastive loss."""
    q, K = 10, 4
    if impl == "reference":
        s = _ref().ClipSampler([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], K=K)
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_Bayes(n_eval=n_eval)
    else:
        O = _port()
        m = O.PairedModel([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        O.clip_bayes(m, n_eval, K)
    return 2 * (K + 1) * n_eval, time.perf_counter() - t0


def c1_cdm(impl, B, seed):
    """tests/test_data_randomghm.py:41-42: ConditionalDenoiseSampler([3,4],[3,3],p=.1,sigma=.1).get_batch(B, guide=True)."""
    q = 10
    if impl == "reference":
        s = _ref().ConditionalDenoiseSampler([3, 4], [3, 3], [_u(q), _u(q)], [0.1, 0.1], sigma=0.1)
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_batch(batch_size=B, guide=True)
    else:
        O = _port()
        m = O.PairedModel([3, 4], [3, 3], [_u(q), _u(q)], [0.1, 0.1], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        r = O.cdm_get_batch(m, B, sigma=0.1)
        O.guides_cls(r["t_hd"], 3, 3)
        O.guides_dns(r["hd"], r["qd"], r["bu"], 4, 3)
    return 2 * B, time.perf_counter() - t0


def c1_dns(impl, B, seed):
    """tests/test_data_randomghm.py:50-51: DenoiseSampler(3,3,p=.1,sigma=.1).get_batch(B, guide=True)."""
    q = 10
    if impl == "reference":
        s = _ref().DenoiseSampler(3, 3, _u(q), p_flip=0.1, sigma=0.1)
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_batch(batch_size=B, guide=True)
    else:
        O = _port()
        m = O.SingleModel(3, 3, _u(q), 0.1)
        np.random.seed(seed)
        t0 = time.perf_counter()
        r = O.dns_get_batch(m, B, 0.1)
        O.guides_dns(r["hd"], r["qd"], r["bu"], 3, 3)
    return B, time.perf_counter() - t0


def c3_sigma(impl, B, seed):
    """CDM sigma sweep: one paired sample, one text BP_CLS -> ext; per sigma: z = x + sigma*N(0,1), image BP_DNS(z, sigma,
    ext), risk sum_leaf (m - x)^2.  trees = B * len(SIGMAS_C3) denoiser passes."""
    q, L, s_ = 10, 4, 3
    nL = s_ ** L
    if impl == "reference":
        R = _ref()
        sm = R.ConditionalDenoiseSampler([L, L], [s_, s_], [_u(q), _u(q)], [0.2, 0.2], sigma=1)
        np.random.seed(seed)
        t0 = time.perf_counter()
        root = np.random.choice(q, size=B)
        tt = R.GHMTree(L, s_, q, _u(q), 0.2, sm.t_transition, B, build_tree=True, root=root)
        it = R.GHMTree(L, s_, q, _u(q), 0.2, sm.i_transition, B, build_tree=True, root=root)
        tt.BP_CLS()
        ext = tt.root_node.hd_message
        x = np.asarray(it.leaves_values)
        for sg in SIGMAS_C3:
            z = np.random.randn(nL, B) * sg + x
            it.BP_DNS(z, sg, external_hd_message=ext.copy())
            np.sum(np.power(it.posterior_mean_DNS - x, 2), 0).mean()
    else:
        O = _port()
        m = O.PairedModel([L, L], [s_, s_], [_u(q), _u(q)], [0.2, 0.2], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        root = np.random.choice(q, size=B)
        tv = O.sample_tree(m.t_T, L, s_, q, B, root=root)
        iv = O.sample_tree(m.i_T, L, s_, q, B, root=root)
        _, t_hd = O.bp_cls(m.t_T, tv[-1], L, s_, q, _u(q))
        ext = t_hd[0][0]
        for sg in SIGMAS_C3:
            z = np.random.randn(nL, B) * sg + iv[-1]
            mean = O.bp_dns(m.i_T, z, sg, L, s_, q, ext=ext)[0]
            np.sum(np.power(mean - iv[-1], 2), 0).mean()
    return B * len(SIGMAS_C3), time.perf_counter() - t0


def _c4(impl, B, seed, guide):
    q = 10
    if impl == "reference":
        s = _ref().NextWordPredictSampler([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2])
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_batch(batch_size=B, guide=guide)
    else:
        O = _port()
        m = O.PairedModel([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        O.nwp_get_batch(m, B, guide=guide)
    return 2 * B, time.perf_counter() - t0


def c4_nwp(impl, B, seed):
    """NextWordPredictSampler([4,4],[3,3],p=.2).get_batch(B, guide=False): image BP_CLS -> ext -> text next-token BP."""
    return _c4(impl, B, seed, False)


def c4_nwp_guides(impl, B, seed):
    """Same with guide=True (the 2L+1 per-position guide tensors)."""
    return _c4(impl, B, seed, True)


def _c5(impl, B, seed, L, s_, q):
    """C5 op: sample one tree, BP_CLS (-> root_hd), z = x + N(0,1), BP_DNS(z, 1, ext = root_hd)."""
    nL = s_ ** L
    if impl == "reference":
        R = _ref()
        np.random.seed(42)
        T = R.GenTransition(L, s_, q, 0.2, 1.0)
        np.random.seed(seed)
        t0 = time.perf_counter()
        root = np.random.choice(q, size=B)
        tr = R.GHMTree(L, s_, q, _u(q), 0.2, T, B, build_tree=True, root=root)
        tr.BP_CLS()
        ext = tr.root_node.hd_message.copy()
        z = np.random.randn(nL, B) + np.asarray(tr.leaves_values)
        tr.BP_DNS(z, 1.0, external_hd_message=ext)
    else:
        O = _port()
        np.random.seed(42)
        T = O.gen_transition(L, s_, q, 0.2, 1.0, True)
        np.random.seed(seed)
        t0 = time.perf_counter()
        root = np.random.choice(q, size=B)
        v = O.sample_tree(T, L, s_, q, B, root=root)
        _, hd = O.bp_cls(T, v[-1], L, s_, q, _u(q))
        z = np.random.randn(nL, B) + v[-1]
        O.bp_dns(T, z, 1.0, L, s_, q, ext=hd[0][0])
    return B, time.perf_counter() - t0


def c5_q10(impl, B, seed):
    return _c5(impl, B, seed, 4, 3, 10)


def c5_q256(impl, B, seed):
    return _c5(impl, B, seed, 4, 3, 256)


def c5_pair(impl, B, seed, q=10):
    """Strong-scaling op: ConditionalDenoiseSampler([4,4],[3,3],p=.2,sigma=1).get_Bayes(B) = paired sample + text BP_CLS +
    image BP_DNS(ext) + risk; trees = 2 * B."""
    if impl == "reference":
        s = _ref().ConditionalDenoiseSampler([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], sigma=1, variable_type=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_Bayes(n_eval=B)
    else:
        O = _port()
        m = O.PairedModel([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        O.cdm_bayes(m, B, 1.0)
    return 2 * B, time.perf_counter() - t0


def feed_cdm(impl, B, seed):
    """Training-loop feed (training/train_CDNS.py:128): ConditionalDenoiseSampler([4,4],[3,3],p=.2).get_batch(B, guide=True)."""
    q = 10
    if impl == "reference":
        s = _ref().ConditionalDenoiseSampler([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], sigma=1)
        np.random.seed(seed)
        t0 = time.perf_counter()
        s.get_batch(batch_size=B, guide=True)
    else:
        O = _port()
        m = O.PairedModel([4, 4], [3, 3], [_u(q), _u(q)], [0.2, 0.2], q=q)
        np.random.seed(seed)
        t0 = time.perf_counter()
        r = O.cdm_get_batch(m, B, sigma=1.0)
        O.guides_cls(r["t_hd"], 4, 3)
        O.guides_dns(r["hd"], r["qd"], r["bu"], 4, 3)
    return 2 * B, time.perf_counter() - t0


def feed_vlm(impl, B, seed):
    """Training-loop feed (training/train_NWP.py:128): NextWordPredictSampler([4,4],[3,3],p=.2).get_batch(B, guide=True)."""
    return _c4(impl, B, seed, True)


TASKS = {f.__name__: f for f in (feed_cdm, feed_vlm, c2_clip, c1_cdm, c1_dns, c3_sigma, c4_nwp, c4_nwp_guides, c5_q10, c5_q256, c5_pair)}


def _worker(args):
    name, impl, size, seed, reps = args
    trees, secs = 0, 0.0
    for r in range(reps):
        t, s_ = TASKS[name](impl, size, seed + 1009 * r)
        trees += t
        secs += s_
    return trees, secs


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def fan_out(pool, cores, name, impl, size, seed0=0, reps=1):
    """``reps`` calls of the task per core, all cores at once; returns dict(trees_per_s, cores, seconds, trees).
    ``pool`` is a multiprocessing Pool with at least ``cores`` workers (``cores=1`` times a single process: the
    ``ref_1core`` figure of SURVEY 8(d)).  ``seconds`` is the wall time of the whole fan-out (slowest worker), sampler
    construction included in it but not in the per-task spans."""
    t0 = time.perf_counter()
    res = pool.map(_worker, [(name, impl, size, seed0 + 17 * i, reps) for i in range(cores)], chunksize=1)
    wall = time.perf_counter() - t0
    trees = sum(r[0] for r in res)
    busy = max(r[1] for r in res)                       # timed span of the slowest worker (construction excluded)
    return {"trees_per_s": trees / busy, "cores": cores, "seconds": round(busy, 3), "wall_seconds": round(wall, 3),
            "trees": trees, "per_core_size": size, "calls_per_core": reps, "impl": impl}

"""p_flip sweeps of the out-of-distribution evaluations: Bayes risk and mis-specified-BP risk per grid point.

Reference: the BP-only portion of ``figures/eval-clip-ood.py:58-94``, ``eval-vlm-ood.py:98-132`` and
``eval-cdm-ood.py:98-127``.  There every grid point builds a new sampler (re-seeding NumPy and re-drawing the
tables), evaluates ``get_Bayes`` on it, draws one more batch from it and runs BP with the tables of the
*training* distribution (p_model) on that batch -- through a host round trip of the leaves
(``T_value[-1] = [...tolist()...]; build_tree()``).

Here one sampler is re-parameterised in place per grid point (``DoubleSampler.reparameterize``: same NumPy draws
as the constructor, one pinned table upload per modality), the leaves never leave the device, the BP of the
p_model tables runs on a second resident model, and every risk accumulates into one ``[n_p, 2, 3]`` float64
tensor that is copied to the host ONCE at the end -- no synchronisation inside the sweep.

``rng="numpy"`` reproduces the reference's columns of ``figures/data/ghm-data/{ood-clip,vlm-ood}.json`` (the
per-point NumPy call order is the reference's: constructor, ``get_Bayes``, ``get_batch``); ``rng="philox"``
draws everything on the device.

``write_reference_json`` stores the result in the on-disk layout the reference's plotting notebooks read
(``eval-clip-ood.py:107-109``: ``{"p_flip": [...], "Bayes": [...], "Mis-spec. BP": [...]}``, ``indent=4``).
"""
import json

import numpy as np
import torch

from . import ops
from . import data_random_GHM as G
from .sharding import mean_se_from_sums

__all__ = ["clip_ood_sweep", "vlm_ood_sweep", "cdm_ood_sweep", "cdm_sigma_sweep", "write_reference_json",
           "write_bayes_checkpoint", "DEFAULT_P_GRID"]

DEFAULT_P_GRID = tuple(int(p) for p in np.arange(2, 42, 2))          # percent, as the reference stores it


def _grid(p_list):
    p_list = DEFAULT_P_GRID if p_list is None else p_list
    return [int(p) if float(p).is_integer() else float(p) for p in p_list]


def _finish(p_list, sums, vlm=False, n_eval=None):
    host = sums.cpu()                                                   # the sweep's only device -> host copy
    res = {"p_flip": list(p_list), "Bayes": [], "Bayes SE": [], "Mis-spec. BP": []}
    for k in range(len(p_list)):
        if vlm:
            # float32 mean like the reference's torch.mean over float32 token losses (:931-942, eval-vlm-ood.py:129)
            s1, s2, c = (float(x) for x in host[k, 0].tolist())
            mean = s1 / c
            var = max((s2 - c * mean * mean) / max(c - 1, 1), 0.0)
            res["Bayes"].append(float(np.float32(mean)))
            res["Bayes SE"].append(float(np.float32((var ** 0.5) / np.sqrt(n_eval))))
            s1, _, c = (float(x) for x in host[k, 1].tolist())
            res["Mis-spec. BP"].append(float(np.float32(s1 / c)))
        else:
            m, se = mean_se_from_sums(host[k, 0])
            res["Bayes"].append(float(m))
            res["Bayes SE"].append(float(se))
            res["Mis-spec. BP"].append(float(mean_se_from_sums(host[k, 1])[0]))
    return res


def _uniform(q):
    return np.ones(q) / q


def clip_ood_sweep(p_list=None, p_model=0.2, n_eval=10000, batch_size=5000, n_layers=(4, 4), n_childs=(3, 3), K=4,
                   variable_type=10, rng="numpy", seed=1234, device=None):
    """CLIP: Bayes contrastive risk at each test p, and the risk of BP run with the p_model tables on test-p data.

    ``p_list`` in percent (default 2, 4, .., 40 as ``eval-clip-ood.py:69``).  Returns the reference's result dict
    (plus ``"Bayes SE"``)."""
    p_list = _grid(p_list)
    q = variable_type
    py = [_uniform(q), _uniform(q)]
    nl, nc = list(n_layers), list(n_childs)
    ref = G.DoubleSampler(nl, nc, py, [p_model, p_model], variable_type=q, device=device, rng=rng, seed=seed)
    sampler = None
    sums = torch.zeros((len(p_list), 2, 3), dtype=torch.float64, device=ref.device)
    for k, p in enumerate(p_list):
        pf = [p / 100, p / 100]
        if sampler is None:
            sampler = G.ClipSampler(nl, nc, py, pf, K=K, variable_type=q, device=device, rng=rng, seed=seed)
        else:
            sampler.reparameterize(pf)
        r = sampler._sample_layout(n_eval, want_leaves=False, want_post=True)             # get_Bayes (:78)
        ops.risk_clip(r["t"]["post"], r["i"]["post"], n_eval, K, q, sums=sums[k, 0])
        r = sampler._sample_layout(batch_size, want_leaves=True, want_post=False)         # get_batch (:82)
        t_post, _ = ref.t_model.bp_cls(r["t"]["leaves"])                                 # BP with the p_model tables (:83-90)
        i_post, _ = ref.i_model.bp_cls(r["i"]["leaves"])
        ops.risk_clip(t_post, i_post, batch_size, K, q, sums=sums[k, 1])
    return _finish(p_list, sums)


def vlm_ood_sweep(p_list=None, p_model=0.2, n_eval=10000, batch_size=1000, n_layers=(4, 4), n_childs=(3, 3),
                  variable_type=10, rng="numpy", seed=1234, device=None):
    """Next-token prediction: Bayes token cross-entropy at each test p and the mis-specified-BP cross-entropy
    (``eval-vlm-ood.py:104-132``): image BP_CLS -> external message -> text next-token BP, both with the p_model
    tables."""
    p_list = _grid(p_list)
    q = variable_type
    py = [_uniform(q), _uniform(q)]
    nl, nc = list(n_layers), list(n_childs)
    ref = G.DoubleSampler(nl, nc, py, [p_model, p_model], variable_type=q, device=device, rng=rng, seed=seed)
    sampler = None
    sums = torch.zeros((len(p_list), 2, 3), dtype=torch.float64, device=ref.device)
    nL = ref.t_model.n_leaves

    def token_ce(t_model, i_model, text_leaves, image_leaves, out):
        _, i_hd = i_model.bp_cls(image_leaves)
        pp = t_model.bp_nwp(text_leaves, i_hd)
        ops.risk_ce(pp, text_leaves, sums=out, target_stride=nL, target_offset=1, row_group=nL - 1)

    for k, p in enumerate(p_list):
        pf = [p / 100, p / 100]
        if sampler is None:
            sampler = G.NextWordPredictSampler(nl, nc, py, pf, variable_type=q, device=device, rng=rng, seed=seed)
        else:
            sampler.reparameterize(pf)
        _, tt, it = sampler._paired_trees(n_eval)                                         # get_Bayes
        token_ce(sampler.t_model, sampler.i_model, tt._leaves, it._leaves, sums[k, 0])
        _, tt, it = sampler._paired_trees(batch_size)                                     # get_batch
        token_ce(ref.t_model, ref.i_model, tt._leaves, it._leaves, sums[k, 1])
    return _finish(p_list, sums, vlm=True, n_eval=n_eval)


def cdm_ood_sweep(p_list=None, p_model=0.2, sigma=1.0, n_eval=10000, batch_size=1000, n_layers=(4, 4),
                  n_childs=(3, 3), variable_type=10, rng="numpy", seed=1234, device=None):
    """Conditional denoising: Bayes MSE at each test p and the MSE of the denoiser that runs BP with the p_model
    tables (``eval-cdm-ood.py:104-127``): text BP_CLS -> external message -> image BP_DNS."""
    p_list = _grid(p_list)
    q = variable_type
    py = [_uniform(q), _uniform(q)]
    nl, nc = list(n_layers), list(n_childs)
    ref = G.DoubleSampler(nl, nc, py, [p_model, p_model], variable_type=q, device=device, rng=rng, seed=seed)
    sampler = None
    sums = torch.zeros((len(p_list), 2, 3), dtype=torch.float64, device=ref.device)
    for k, p in enumerate(p_list):
        pf = [p / 100, p / 100]
        if sampler is None:
            sampler = G.ConditionalDenoiseSampler(nl, nc, py, pf, sigma=sigma, variable_type=q, device=device, rng=rng,
                                                  seed=seed)
        else:
            sampler.reparameterize(pf)
        _, image_tree, _, _, _, mean = sampler._run(n_eval)                               # get_Bayes
        ops.risk_cdm(mean, image_tree._leaves, sums=sums[k, 0])
        text_tree, image_tree, z, _, _, _ = sampler._run(batch_size)                      # get_batch
        _, t_hd = ref.t_model.bp_cls(text_tree._leaves)
        mean = ref.i_model.bp_dns(z, float(sigma), t_hd)
        ops.risk_cdm(mean, image_tree._leaves, sums=sums[k, 1])
    return _finish(p_list, sums)


def cdm_sigma_sweep(sigmas=(0.1, 0.25, 0.5, 1.0, 2.0, 4.0), p_flip=0.2, n_eval=65536, n_layers=(4, 4), n_childs=(3, 3),
                    variable_type=10, seed=1234, device=None, sampler=None):
    """Bayes denoising risk across diffusion noise levels (BASELINE config 3): ONE paired sample and ONE text BP_CLS
    give the external root message; per sigma: z = x + sigma * N(0, 1) (Philox normal), image BP_DNS conditioned on the
    text, risk sum_leaf (m - x)^2 (reference ConditionalDenoiseSampler.get_Bayes :886-894 at that sigma).  Everything
    stays on the device; the [n_sigma, 3] accumulator is copied back once.  Philox mode only (the reference has no
    sigma loop whose NumPy stream could be mirrored).  ``sampler`` re-uses a Philox-mode ConditionalDenoiseSampler (its
    device tables) instead of constructing one."""
    q = variable_type
    py = [_uniform(q), _uniform(q)]
    if sampler is None:
        sampler = G.ConditionalDenoiseSampler(list(n_layers), list(n_childs), py, [p_flip, p_flip], sigma=1.0,
                                              variable_type=q, device=device, rng="philox", seed=seed)
    elif sampler.rng != "philox":
        raise ValueError("cdm_sigma_sweep needs a Philox-mode sampler")
    seed = sampler.seed
    _, text_tree, image_tree = sampler._paired_trees(n_eval, text_bp=True)
    off = sampler.tree_offset - n_eval
    t_hd = text_tree._root_hd
    if t_hd is None:
        _, t_hd = sampler.t_model.bp_cls(text_tree._leaves)
    sums = torch.zeros((len(sigmas), 3), dtype=torch.float64, device=sampler.device)
    for k, sg in enumerate(sigmas):
        z = sampler.i_model.gauss_noise(image_tree._leaves, float(sg), seed=(seed ^ ops.IMAGE_SEED_XOR) + 7919 * (k + 1),
                                        tree_offset=off)
        mean = sampler.i_model.bp_dns(z, float(sg), t_hd)
        ops.risk_cdm(mean, image_tree._leaves, sums=sums[k])
    host = sums.cpu()
    res = {"sigma": [float(x) for x in sigmas], "Bayes": [], "Bayes SE": []}
    for k in range(len(sigmas)):
        m, se = mean_se_from_sums(host[k])
        res["Bayes"].append(float(m))
        res["Bayes SE"].append(float(se))
    return res


def write_reference_json(res, path, extra=None):
    """Write a sweep result in the reference's ``figures/data/ghm-data/*.json`` layout (``json.dump(res, f, indent=4)``
    with the columns ``p_flip``, ``Bayes``, ``Mis-spec. BP``); ``extra`` adds model columns computed elsewhere."""
    out = {"p_flip": list(res["p_flip"]), "Bayes": list(res["Bayes"]), "Mis-spec. BP": list(res["Mis-spec. BP"])}
    if extra:
        for name, col in extra.items():
            if len(col) != len(out["p_flip"]):
                raise ValueError("column %r has %d entries for %d grid points" % (name, len(col), len(out["p_flip"])))
            out[name] = list(col)
    with open(path, "w") as f:
        json.dump(out, f, indent=4)
    return out


def write_bayes_checkpoint(path, sampler, n_eval=10000, state=None):
    """Write the ``bayes`` key of the reference's training checkpoints from this path.

    The reference's training scripts evaluate ``Bayes_loss, _ = sampler.get_Bayes(n_eval=10000)`` once and store it
    in every ``checkpoint.pth`` (``training/train_CLIP.py:76,193-200``, ``train_CDNS.py:75,165-173``,
    ``train_NWP.py:74``); the risk figures read only ``ckpt["bayes"]`` and ``ckpt["loss_history"]``
    (``figures/eval-clip-risk.py:28-29``, ``eval-cdm-risk.py:28``, ``eval-vlm-risk.py:28``).  ``state`` carries the
    caller's model / optimiser entries (``*_state_dict``, ``iter``, ``loss_history``, ``ploss_history`` ...); the
    file is a plain ``torch.save`` dict, loadable with ``torch.load(path, weights_only=False)`` like the reference's.
    Returns ``(bayes, bayes_std)``."""
    bayes, bayes_std = sampler.get_Bayes(n_eval=n_eval)
    ckpt = {"iter": 0, "loss_history": torch.zeros(0), "ploss_history": torch.zeros(0)}
    if state:
        ckpt.update(state)
    ckpt["bayes"] = bayes
    torch.save(ckpt, path)
    return bayes, bayes_std

"""Golden values for recipes whose SHIPPED fixture is stale: run the real reference here (build container only).

    python tests/golden/make_golden_recipes.py            # needs /root/reference

``figures/data/ghm-data/cdm-ood.json`` ("Mis-spec. BP" 10.0038 at p = 2 %) is not reproducible from the shipped
``figures/eval-cdm-ood.py``, so the BP-only portion of that script (:98-127) is executed with
the unmodified reference classes and its outputs are stored in ``tests/golden/kat_regenerated.json``.
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("GHM_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))

from ghmclip.data.data_random_GHM import ConditionalDenoiseSampler, DoubleSampler  # noqa: E402


def cdm_ood(p_list, batch_size=5000, n_eval=10000):
    """figures/eval-cdm-ood.py:98-127 without the model columns (they need checkpoints and consume no NumPy RNG)."""
    n_layers, n_childs = [4, 4], [3, 3]
    p_ys = [np.ones(10) / 10, np.ones(10) / 10]
    tree_sampler = DoubleSampler(n_layers, n_childs, p_ys, [0.2, 0.2])
    text_tree, image_tree = tree_sampler.get_zeroshot_batch(batch_size=batch_size, return_tree=True)
    res = {"p_flip": [int(p) for p in p_list], "Bayes": [], "Mis-spec. BP": [], "batch_size": batch_size, "n_eval": n_eval}
    for p in p_list:
        sampler = ConditionalDenoiseSampler(n_layers, n_childs, p_ys, [p / 100, p / 100])
        bayes, _ = sampler.get_Bayes(n_eval=n_eval)
        res["Bayes"].append(float(bayes))
        res_text, res_image = sampler.get_batch(device="cpu", batch_size=batch_size, guide=False)
        text_tree.T_value[-1] = [res_text[0][:, idx].tolist() for idx in range(81)]
        image_tree.T_value[-1] = [res_image[1][:, idx].tolist() for idx in range(81)]
        text_tree.build_tree()
        image_tree.build_tree()
        text_tree.BP_CLS()
        ext = text_tree.root_node.hd_message
        image_tree.BP_DNS(res_image[0].T.numpy(), 1, external_hd_message=ext)
        pred = image_tree.posterior_mean_DNS.T
        target = res_image[1].numpy()
        res["Mis-spec. BP"].append(float(np.mean(np.sum(np.power(pred - target, 2), 1))))
        print(p, res["Bayes"][-1], res["Mis-spec. BP"][-1], flush=True)
    return res


if __name__ == "__main__":
    out = {"_source": "real reference executed by tests/golden/make_golden_recipes.py (BP-only portion of "
                      "figures/eval-cdm-ood.py:98-127; the shipped cdm-ood.json column is stale)",
           "cdm-ood.recipe": cdm_ood([2, 4, 30])}
    with open(os.path.join(HERE, "kat_regenerated.json"), "w") as f:
        json.dump(out, f, indent=1)

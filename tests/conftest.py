"""pytest configuration: marker registration, import paths, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multimodal-ghm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_tree_case(name):
    """npz fixture -> dict with ``T`` rebuilt as list[L] of list[edges] of (q,q)."""
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    L = int(d["L"])
    d["T"] = [[m for m in d[f"T{l}"]] for l in range(L)]
    for k in ("L", "s", "q", "ti", "B", "given_root"):
        d[k] = int(d[k])
    for k in ("p_flip", "sigma"):
        d[k] = float(d[k])
    return d


TREE_CASES = ["tree_L1s2q3", "tree_L2s2q3_nonTI_py", "tree_L3s3q10", "tree_L4s3q10",
              "tree_L3s2q5_nonTI", "tree_L2s4q16", "tree_L5s2q4"]


@pytest.fixture(params=TREE_CASES)
def tree_case(request):
    return load_tree_case(request.param)

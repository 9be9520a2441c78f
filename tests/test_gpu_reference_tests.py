"""The reference's OWN unit tests (reference tests/test_data_randomghm.py), run UNCHANGED against the facade.

`baseline/install_reference.py` (called by `__graft_entry__.build()` in the build container) places the unmodified
test file under the git-ignored `baseline/_ref/_reference_tests/`; it travels to the GPU box with the snapshot.  The
file does `from ghmclip.data.data_random_GHM import ConditionalDenoiseSampler, DenoiseSampler`: that module name is
aliased to `ghm_b200.data_random_GHM` for the duration of the test, so the reference's test code drives the CUDA
path through the reference's API.
"""
import importlib.util
import os
import sys
import types
import unittest

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF_TEST = os.path.join(ROOT, "baseline", "_ref", "_reference_tests", "test_data_randomghm.py")


def test_reference_unit_tests_unchanged():
    if not os.path.exists(REF_TEST):
        pytest.skip("baseline/_ref/_reference_tests is absent (run python baseline/install_reference.py where /root/reference exists)")
    assert torch.cuda.is_available()
    from ghm_b200 import data_random_GHM as facade
    saved = {k: sys.modules.get(k) for k in ("ghmclip", "ghmclip.data", "ghmclip.data.data_random_GHM")}
    pkg, sub = types.ModuleType("ghmclip"), types.ModuleType("ghmclip.data")
    pkg.__path__, sub.__path__ = [], []
    pkg.data, sub.data_random_GHM = sub, facade
    sys.modules.update({"ghmclip": pkg, "ghmclip.data": sub, "ghmclip.data.data_random_GHM": facade})
    try:
        spec = importlib.util.spec_from_file_location("reference_test_data_randomghm", REF_TEST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.ConditionalDenoiseSampler is facade.ConditionalDenoiseSampler     # the alias took effect
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        assert suite.countTestCases() == 2
        result = unittest.TextTestRunner(verbosity=0).run(suite)
        assert result.wasSuccessful(), (result.failures, result.errors)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

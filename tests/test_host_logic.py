"""CPU tests (no GPU): C-ABI surface, host-side facade logic, sharding over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_tree_case

PKG = os.path.join(ROOT, "multimodal-ghm_b200")


@pytest.fixture(scope="module")
def lib():
    """Build (nvcc cross-compiles sm_100a without a GPU) and load the shared library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ghm_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = mod.build()
    return ctypes.CDLL(path), path


def test_library_exports_every_declared_symbol(lib):
    """Every prototype in include/ghm_b200.h is exported, and the ctypes table covers exactly that set."""
    cdll, path = lib
    header = open(os.path.join(ROOT, "include", "ghm_b200.h")).read()
    declared = set(re.findall(r"GHM_API\s+[\w\s\*]+?\b(ghm_\w+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(cdll, name), "missing export " + name
    from ghm_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    # built for sm_100a only
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device model creation fails loudly (GHM_ECUDA) instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ghm_b200 import ops
    c = load_tree_case("tree_L1s2q3")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.GhmModel(c["T"], c["L"], c["s"], c["q"], p_y=c["p_y"], device="cuda:0")
    from ghm_b200 import data_random_GHM as G
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        G.ClipSampler([2, 2], [2, 2], [np.ones(10) / 10] * 2, [.2, .2])


def test_argument_validation_needs_no_gpu(lib):
    cdll, _ = lib
    cdll.ghm_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    T = np.eye(3)[None].repeat(2, 0).copy()
    rc = cdll.ghm_model_create(ctypes.byref(h), 0, 2, 3, 1, T.ctypes.data_as(ctypes.c_void_p), None, 0)
    assert rc == 1 and b"n_layer" in cdll.ghm_last_error()
    rc = cdll.ghm_model_create(ctypes.byref(h), 2, 2, 300, 1, T.ctypes.data_as(ctypes.c_void_p), None, 0)
    assert rc == 1 and b"variable_type" in cdll.ghm_last_error()
    assert cdll.ghm_risk_clip(None, None, 1, 4, 10, 0, 1, None, None) == 1


def test_facade_gen_transition_matches_reference_tables():
    """Host-side GenTransition consumes NumPy's stream like the reference (:43-89): identical tables."""
    from ghm_b200.data_random_GHM import GenTransition
    from ghm_b200.ops import is_translation_invariant
    for name, seed in (("tree_L3s3q10", 3), ("tree_L3s2q5_nonTI", 5), ("tree_L5s2q4", 7)):
        c = load_tree_case(name)
        np.random.seed(seed)
        T = GenTransition(c["L"], c["s"], c["q"], c["p_flip"], 1.0, translation_invariance=bool(c["ti"]))
        for l in range(c["L"]):
            assert np.array_equal(np.stack(T[l]), c[f"T{l}"])
        assert is_translation_invariant(T, c["s"]) == bool(c["ti"])
        if c["ti"]:
            assert T[1][0] is T[1][c["s"]]          # TI levels share the same objects (:75-76)


def test_leaf_columns_list_semantics():
    import torch
    from ghm_b200.data_random_GHM import _LeafColumns
    dev = torch.arange(12, dtype=torch.int64).reshape(4, 3)      # [B=4, n_L=3]
    cols = _LeafColumns(dev)
    assert len(cols) == 3 and cols[1] == [1, 4, 7, 10] and isinstance(cols[1], list)
    assert np.array(cols).shape == (3, 4) and [r for r in cols][2] == [2, 5, 8, 11]
    assert not cols.dirty
    cols[0] = [9, 9, 9, 9]
    assert cols.dirty and cols[0] == [9, 9, 9, 9] and np.array(cols).T[:, 0].tolist() == [9, 9, 9, 9]


def test_shard_range_partitions():
    from ghm_b200.sharding import mean_se_from_sums, shard_range
    for n in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    x = np.array([1.0, 2.0, 4.0, 7.0])
    m, se = mean_se_from_sums([x.sum(), (x ** 2).sum(), 4])
    assert m == pytest.approx(x.mean()) and se == pytest.approx(x.std() / 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "multimodal-ghm_b200")); sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from ghm_b200.sharding import shard_range, dist_info, all_reduce_sums, mean_se_from_sums
from oracle import ghm_oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
assert dist_info() == (rank, world)
# every rank owns a slice of the pair index and reduces only {sum, sumsq, count}
rng = np.random.RandomState(0)
n, K, q = 501, 4, 10
t = rng.dirichlet(np.ones(q), size=n * (K + 1)).T
i = rng.dirichlet(np.ones(q), size=n * (K + 1)).T
S = O.clip_loss_terms(t, i, n, K, q)               # per-pair losses (the CPU oracle stands in for the kernel)
lo, hi = shard_range(n, rank, world)
mine = S[lo:hi]
sums = torch.tensor([mine.sum(), (mine ** 2).sum(), float(hi - lo)], dtype=torch.float64)
all_reduce_sums(sums)
mean, se = mean_se_from_sums(sums)
ref = O.clip_loss(t, i, n, K, q)
assert abs(mean - ref[0]) < 1e-12 and abs(se - ref[1]) < 1e-12, (mean, ref)
# global Philox tree indices of the shards tile the batch with no gap or overlap
spans = [None] * world
dist.all_gather_object(spans, (lo, hi))
assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_risk_reduction_gloo_world2(tmp_path):
    """N>1 path on CPU: 2 gloo ranks shard the pair index, all-reduce 3 doubles, agree with the whole."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` runs the CPU arm (the unmodified reference under baseline/_ref when installed,
    else the oracle port) and prints the JSON line the driver parses."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-n-eval", "300"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "trees/s" and line["value"] > 0
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "ghmclip", "data", "data_random_GHM.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_cpu_arms_reference_and_port_agree():
    """baseline/cpu_arms.py: every task runs in both implementations and counts the same trees."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import cpu_arms as A
    impls = ["port"] + (["reference"] if A.reference_available() else [])
    for name, size in (("c2_clip", 40), ("c1_cdm", 16), ("c1_dns", 16), ("c3_sigma", 16), ("c4_nwp", 8),
                       ("c4_nwp_guides", 4), ("c5_q10", 16), ("c5_pair", 16)):
        trees = {impl: A.TASKS[name](impl, size, 3)[0] for impl in impls}
        assert len(set(trees.values())) == 1 and trees["port"] > 0, (name, trees)


def test_reference_json_writer_layout(tmp_path):
    """write_reference_json keeps the reference's on-disk contract (figures/eval-clip-ood.py:107-109): the
    columns p_flip / Bayes / Mis-spec. BP, extra model columns appended, indent=4; rejects ragged columns."""
    import json
    from ghm_b200 import sweeps
    res = {"p_flip": [2, 4], "Bayes": [0.4, 0.41], "Bayes SE": [0.01, 0.01], "Mis-spec. BP": [0.45, 0.43]}
    p = tmp_path / "ood.json"
    sweeps.write_reference_json(res, p, extra={"Guided TF": [0.5, 0.6]})
    text = p.read_text()
    assert text.startswith('{\n    "p_flip": [\n        2,')
    assert list(json.loads(text)) == ["p_flip", "Bayes", "Mis-spec. BP", "Guided TF"]
    import pytest
    with pytest.raises(ValueError):
        sweeps.write_reference_json(res, p, extra={"bad": [1.0]})
    assert sweeps.DEFAULT_P_GRID == tuple(range(2, 42, 2))

"""Build libghm_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

    python multimodal-ghm_b200/build.py [--force] [--verbose]

Each csrc/*.cu is compiled to an object in parallel, then linked into
multimodal-ghm_b200/ghm_b200/libghm_b200.so (git-ignored; it travels to the GPU box).
"""
import concurrent.futures as cf
import glob
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "ghm_b200")
BUILD = os.environ.get("GHM_BUILD_DIR") or os.path.join(HERE, "build")          # (development aids: an A/B library
LIB = os.environ.get("GHM_LIB_OUT") or os.path.join(OUT_DIR, "libghm_b200.so")     #  built with other flags elsewhere)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("GHM_EXTRA_FLAGS", "").split()


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _closure(path, seen):
    """Transitive closure of the quoted includes of ``path`` (only headers of this repo are quoted)."""
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    with open(path, "r", errors="replace") as fh:
        text = fh.read()
    for inc in _INC.findall(text):
        _closure(os.path.join(os.path.dirname(path), inc), seen)


def _deps_hash(src):
    """Hash of the source, the headers it (transitively) includes and the flags: a header edit rebuilds only its users."""
    seen = set()
    _closure(src, seen)
    h = hashlib.sha1()
    for f in sorted(seen):
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, force, verbose):
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha1"
    hh = _deps_hash(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == hh:
        return obj, False, ""
    only = os.environ.get("GHM_BUILD_ONLY")              # development aid: recompile only the sources whose name contains
    if only and only not in os.path.basename(src) and os.path.exists(obj):   # this string, link the other (stale) objects
        return obj, False, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(hh)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not srcs:
        raise RuntimeError("no CUDA sources under " + CSRC)

    def cost(path):                                      # longest translation units first: they are the long pole of a
        name = os.path.basename(path)                    # parallel build (q = 16 instantiations take 2-3 minutes each)
        return -(4 if "q16" in name else 3 if ("guides" in name or "q10" in name) else 2 if ("_inst_" in name or "_fast_" in name) else 1)
    srcs.sort(key=lambda p: (cost(p), p))
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in res]
    if verbose:
        for r in res:
            if r[2]:
                sys.stderr.write(r[2])
    if any(r[1] for r in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Install the UNMODIFIED reference under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).

    python baseline/install_reference.py            # build container only: needs /root/reference

``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`` -- the
reference is a pure-Python package (``ghmclip``), so the "build" is a wheel of its sources; ``--no-deps`` because its
dependency list names packages that are absent here and irrelevant to the hot path (s3fs, matplotlib, ipykernel ...);
NumPy / torch / tqdm, the only imports of ``ghmclip.data.data_random_GHM``, are in the image.  /root/reference is
read-only, so the wheel is built from a copy under /tmp.  The reference's own unit-test file (not part of the
package) is placed next to it so that ``tests/test_gpu_reference_tests.py`` can run it unchanged against the facade.

Nothing here is product code: baseline/_ref is used by ``bench.py --impl reference`` / ``cpu_baseline`` (the CPU arm)
and by tests as the checker.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REF = os.environ.get("GHM_REFERENCE", "/root/reference")


def installed():
    return os.path.exists(os.path.join(TARGET, "ghmclip", "data", "data_random_GHM.py"))


def install(force=False):
    """Returns a one-line outcome string; raises only when the reference is present and pip fails."""
    if installed() and not force:
        return "already installed"
    if not os.path.isdir(REF):
        return "reference tree %s not present (GPU box): nothing to install" % REF
    tmp = tempfile.mkdtemp(prefix="ghmref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("__pycache__", "*.pdf", "*.ipynb", "figures"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        tdir = os.path.join(TARGET, "_reference_tests")
        os.makedirs(tdir, exist_ok=True)
        shutil.copy(os.path.join(REF, "tests", "test_data_randomghm.py"), tdir)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return "installed ghmclip into baseline/_ref (pip --no-deps --target)"


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

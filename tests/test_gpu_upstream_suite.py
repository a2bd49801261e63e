"""The reference's OWN hot-path tests, unmodified, against this package.

Wherever the upstream tree is present next to a CUDA device, ``tests/test_bmc.py`` and
``tests/test_inference_utils.py`` of the reference are run as they are, in a subprocess whose ``pybmc`` is this
package (``pybmc_b200.install_as_pybmc()`` in a ``sitecustomize`` shim) -- nothing of the reference's package is
importable there, only its test files are read.  Skipped when the tree is absent (the GPU box of the driver has
no /root/reference; the authoring container has no GPU).  The same checks, re-typed, are in test_gpu_bmc.py.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UPSTREAM = os.environ.get("PYBMC_REFERENCE_TREE", "/root/reference")
FILES = ["test_bmc.py", "test_inference_utils.py"]


def upstream_command(tmp_path, files):
    """(argv, env) that runs the upstream test files with ``import pybmc`` resolving to pybmc_b200."""
    shim = tmp_path / "shim"
    shim.mkdir(exist_ok=True)
    (shim / "sitecustomize.py").write_text(
        "import sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import pybmc_b200\n"
        "pybmc_b200.install_as_pybmc()\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(shim), ROOT])        # NOT the upstream tree: its package stays unimportable
    argv = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path)]
    argv += [os.path.join(UPSTREAM, "tests", f) for f in files]
    return argv, env


@pytest.mark.skipif(not os.path.isdir(os.path.join(UPSTREAM, "tests")), reason="upstream tree not present")
def test_upstream_hot_path_tests_pass_unmodified(tmp_path):
    argv, env = upstream_command(tmp_path, FILES)
    r = subprocess.run(argv, env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout

"""
The reference's own test script, UNMODIFIED, run against this repository's packages
(build container only: /root/reference does not exist on the GPU box).  It is executed
where it lies through a scratch package made of symlinks - nothing is copied.
"""
import os
import subprocess
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted")
def test_reference_test_nashconv_runs_unchanged(tmp_path):
    pkg = tmp_path / "refpkg"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    for name in ("environment", "nn", "learn", "util", "_b200.py", "lib"):
        os.symlink(os.path.join(REPO, "r-nad_b200", name), pkg / name)
    os.symlink(os.path.join(REFERENCE, "tests"), pkg / "tests")
    # The script draws UNSEEDED random trees and asserts exact float equalities (sum of reach probabilities == 2), which
    # holds or not depending on how a mixed equilibrium's probabilities round: pin the generators from outside the
    # script (sitecustomize runs at interpreter start) so that the run is reproducible.
    (pkg / "sitecustomize.py").write_text("import random, numpy\nrandom.seed(7)\nnumpy.random.seed(7)\n")
    env = dict(os.environ, PYTHONPATH=str(pkg))
    proc = subprocess.run([sys.executable, "-m", "refpkg.tests.test_nashconv"], cwd=tmp_path, env=env,
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert proc.stdout.split() == ["None"] * 5      # its __main__ prints the (None) result of five parametrisations

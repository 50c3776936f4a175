"""
The reference's own test scripts, UNMODIFIED (tests/test_optimize.py, tests/test_newton.py of cselab/odil; taken from
/root/reference or the staged copy oracle/_ref), run as they are run upstream -- `python test_*.py` -- with `import
odil` resolving to this repository.  They exercise Field / MultigridField at every location, Array and NeuralNet
unknowns, ctx.field(loc=...) location changes, L-BFGS-B and Adam, Problem.linearize and the SciPy-matrix surface of
the Jacobian, and judge themselves (exit code = number of failed checks; bars 1e-2 and 1e-6).
"""
import os
import subprocess
import sys
import tempfile

import pytest

from tests import refsrc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refsrc.available(), reason="reference scripts absent")]
ROOT = refsrc.ROOT


def run_script(relpath, *argv):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["ODIL_DTYPE"] = "float64" if "newton" in relpath else env.get("ODIL_DTYPE", "float32")
    with tempfile.TemporaryDirectory() as tmp:
        # a stand-in for matplotlib (absent from the image); the scripts import it but never plot in these runs
        os.makedirs(os.path.join(tmp, "matplotlib"))
        for name in ["__init__.py", "pyplot.py"]:
            with open(os.path.join(tmp, "matplotlib", name), "w") as f:
                f.write("def __getattr__(name):\n    return lambda *a, **k: None\n")
        env["PYTHONPATH"] = tmp + os.pathsep + env["PYTHONPATH"]
        r = subprocess.run([sys.executable, os.path.join(refsrc.root(), relpath), *argv], cwd=tmp, env=env,
                           capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    return r


def test_reference_test_optimize_runs_unmodified():
    r = run_script("tests/test_optimize.py")
    assert r.returncode == 0, "failed checks: %d" % r.returncode
    assert r.stdout.count("PASS") == 2 and "FAIL" not in r.stdout


def test_reference_test_newton_runs_unmodified():
    r = run_script("tests/test_newton.py")
    assert r.returncode == 0, "failed checks: %d" % r.returncode
    assert r.stdout.count("PASS") == 4 and "FAIL" not in r.stdout

"""The real boundary under test: integration/ngsb200_ngla.cpp compiled against the reference build (oracle/_ref/ngs, which
travels to the GPU box) and an UNCHANGED NGSolve script run on it (integration/run_ngsolve_dropin.py): netgen mesh,
BilinearForm.Assemble, CreateSmoother, `CGSolver(mat, pre)` / `GMRESSolver(mat, pre)` behind CreateDeviceMatrix() /
CreateDeviceVector(), plus the reference's own python/krylovspace.py solvers on the device objects.

Parity bars (BASELINE.json north_star): same step count as NGSolve's CPU solve within +-2, solutions to 1e-8 relative."""
import glob
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

PFX = os.path.join(ROOT, "oracle", "_ref", "ngs")
SITE = os.path.join(PFX, "lib", "python3.12", "site-packages")


def _env():
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    libs = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs") if spec and spec.origin else ""
    env = dict(os.environ)
    env["PYTHONPATH"] = SITE + os.pathsep + env.get("PYTHONPATH", "")
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(PFX, "lib"), os.path.join(SITE, "netgen"), libs, env.get("LD_LIBRARY_PATH", "")])
    return env


@pytest.fixture(scope="module")
def dropin():
    if not os.path.isdir(os.path.join(SITE, "ngsolve")):
        pytest.skip("the reference build oracle/_ref/ngs is not on this box (oracle/build_reference.sh)")
    if not glob.glob(os.path.join(ROOT, "integration", "_build", "_ngsb200*.so")):
        r = subprocess.run(["bash", os.path.join(ROOT, "integration", "build_adapter.sh")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "run_ngsolve_dropin.py"), "--maxh", "0.08", "--order", "3", "--reorder", "1"],
                       env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_creators_return_device_objects(dropin):
    assert dropin["is_host_object"] == [False, False]          # not the registry's silent host fallback
    assert dropin["reorder_info"][0] is True                   # netgen numbering -> products run on P A P^T


def test_unchanged_cgsolver_runs_the_fused_loop(dropin):
    d = dropin
    assert d["dev_fused"] is True and d["dev_solver_type"] == "KrylovSpaceSolver"
    assert d["host_factory_fused"] is False                    # host operands: the reference's own solver, untouched
    assert abs(d["dev_steps"] - d["cpu_steps"]) <= 2
    assert d["rel_diff"] <= 1e-8
    assert abs(d["fused_steps"] - d["cpu_steps"]) <= 2 and d["fused_rel_diff"] <= 1e-8


def test_block_jacobi_is_not_silently_dropped(dropin):
    d = dropin
    assert d["bj_fused"] is False
    assert abs(d["bj_dev_steps"] - d["bj_cpu_steps"]) <= 2 and d["bj_rel_diff"] <= 1e-7


def test_reference_python_krylov_solvers_on_device_objects(dropin):
    d = dropin
    assert abs(d["pycg_iterations"] - d["pycg_host_iterations"]) <= 2 and d["pycg_rel_diff"] <= 1e-8
    assert abs(d["pygmres_iterations"] - d["pygmres_host_iterations"]) <= 2 and d["pygmres_rel_diff"] <= 1e-6


def test_other_entry_kinds(dropin):
    d = dropin
    assert "b3_error" not in d and "z_error" not in d, (d.get("b3_error"), d.get("z_error"))
    assert d["b3_fused"] is True and abs(d["b3_dev_steps"] - d["b3_cpu_steps"]) <= 2 and d["b3_rel_diff"] <= 1e-8
    assert d["z_fused"] is True and abs(d["z_dev_gmres_steps"] - d["z_cpu_gmres_steps"]) <= 2 and d["z_rel_diff"] <= 1e-8


def test_range_views_are_coherent_both_ways(dropin):
    assert dropin["range_coherence"] == [True, True, True, True]


def test_device_scalars(dropin):
    d = dropin
    assert "scalar_error" not in d, d.get("scalar_error")
    assert d["scalar_axpy_rel_diff"] <= 1e-14

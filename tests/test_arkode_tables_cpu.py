"""CPU checks of ReactorArkode's data and control logic (no GPU):
 * pelec_b200/csrc/prx_ark_tables.inc, the Butcher tables the CUDA library is compiled with, satisfies the row-sum and order-1/2
   conditions for every table and equals what tools/extract_erk_tables.py regenerates from SUNDIALS' definitions (build
   container only);
 * the numpy prototype of the ARKODE ERKStep control logic restated in prx_arkode.cuh / react_arkode (tools/arkode_prototype.py,
   test infrastructure: tables read from that same .inc, right-hand side from the pinned C oracle) reproduces the reference's
   right-hand-side evaluation counts and results for tables of every dense-output degree (cubic, quartic, quintic) on the
   committed fixture of the unmodified reference (tests/golden/make_golden6.py)."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "pelec_b200", "csrc", "prx_ark_tables.inc")


def load_prototype():
    spec = importlib.util.spec_from_file_location("arkode_prototype", os.path.join(ROOT, "tools", "arkode_prototype.py"))
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["arkode_prototype"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


@pytest.fixture(scope="module")
def proto():
    return load_prototype()


@pytest.mark.parametrize("method", [20, 30, 31, 40, 41, 42, 43, 50, 51, 52, 53, 54, 60, 80])
def test_table_consistency(proto, method):
    s, A, b, d, c, q, p = proto.table(method)
    assert np.allclose(A, np.tril(A, -1))                      # explicit
    assert np.abs(A.sum(axis=1) - c).max() < 1e-14             # c_i = sum_j a_ij
    assert abs(b.sum() - 1.0) < 1e-14 and abs(d.sum() - 1.0) < 1e-14
    if q >= 2:
        assert abs(b @ c - 0.5) < 1e-14
    if p >= 2:
        assert abs(d @ c - 0.5) < 1e-14
    assert 1 <= p < q <= 8 and s <= 13


def test_tables_file_is_what_the_extractor_generates(tmp_path):
    ref = os.path.join(os.environ.get("PELE_REFERENCE", "/root/reference"),
                       "Submodules/PelePhysics/Submodules/sundials/src/arkode/arkode_butcher_erk.def")
    if not os.path.exists(ref):
        pytest.skip("reference not present")
    out = str(tmp_path / "tables.inc")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "extract_erk_tables.py"), out], stdout=subprocess.DEVNULL)
    assert open(out).read() == open(INC).read()


@pytest.mark.parametrize("method", [42, 52, 60])
def test_prototype_reproduces_reference_counts(proto, method):
    z = np.load(os.path.join(ROOT, "tests", "golden", "arkode_tables_LiDryer.npz"))
    so = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)  # the oracle library prints its banner
    try:
        orc = proto.Oracle("LiDryer")
    finally:
        os.dup2(so, 1)
    ctl = proto.Ctl(0)
    for i in range(3):
        pi, po = "in_%d_" % i, "m%d_%d_" % (method, i)
        st = proto.BoxState(z[pi + "rY0"].copy(), z[pi + "T0"].copy(), z[pi + "rE0"].copy(), z[pi + "rYsrc"].copy(), z[pi + "rEsrc"].copy())
        y, nfe, nst, netf = proto.arkode_box(orc, st, float(z["dt"]), ctl, 1e-6, 1e-10, method)
        assert nfe == int(z["m%d_nfe" % method][i])
        lim = 1e-10 + 1e-6 * np.abs(z[po + "rY"])
        assert (np.abs(y[:, :orc.N].T - z[po + "rY"]) / lim).max() < 1e-5

"""The drop-in IN THE APPLICATION: the unmodified PeleC (PMF regression case, two AMR levels, hydro +
diffusion + reactions, SDC) is advanced with the reference's ReactorCvode / ReactorRK64 and, in a second process, with
pelec.chem_integrator = ReactorCvodeB200 / ReactorRK64B200 - the reference's own factory creating this repo's reactor, and
PeleC::react_state (PC/Source/React.cpp:222-230) driving the CUDA kernels through the reference's virtual react().
The two runs must produce the same hierarchy and the same state.

Needs oracle/_ref/libpelec_pmf_LiDryer.so (built in the build container by oracle/ref_pelec/Makefile; travels to the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpelec_pmf_LiDryer.so")
NSTEPS = 4


def run(integ, tmp_path, extra=(), cuda=False):
    out = str(tmp_path / (integ + ("_cuda" if cuda else "") + ".npz"))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pelec_pmf_run.py"), integ, str(NSTEPS), out] + list(extra)
                       + (["--cuda"] if cuda else []), capture_output=True, text=True, timeout=900)
    assert "PELEC_PMF_DONE" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
    assert "Creating ReactorBase instance: " + integ in p.stdout  # the reference's factory (Factory.H) announces what it creates
    z = np.load(out)
    return {k: z[k] for k in z.files}


LIB_CUDA = os.path.join(ROOT, "oracle", "_ref", "libpelec_pmf_cuda_LiDryer.so")
CASES = [
    # (stock integrator, stock build is CUDA, our integrator, our build is CUDA, bar on the state)
    ("ReactorCvode", False, "ReactorCvodeB200", False, 1e-7),
    ("ReactorRK64", False, "ReactorRK64B200", False, 1e-9),
    # the reference's third reactor (the CI regression case pmf-lidryer-arkode): box-coupled ERKStep, controller history carried
    # from box to box in MFIter order
    ("ReactorArkode", False, "ReactorArkodeB200", False, 1e-9),
    # the CUDA build of the application (AMReX + PeleC + PelePhysics compiled by nvcc for sm_100a, state in device
    # memory): the shim's AMREX_USE_GPU branch hands PeleC's device FABs and AMReX's stream to prx_react_box.  ReactorCvode is
    # compared with the CPU build's per-cell ReactorCvode (the GPU ReactorCvode of the reference integrates each box as one
    # coupled system: other semantics); RK64 with the reference's own GPU ReactorRK64 in the same CUDA build.
    ("ReactorCvode", False, "ReactorCvodeB200", True, 1e-7),
    ("ReactorRK64", True, "ReactorRK64B200", True, 1e-9),
]


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libpelec_pmf_LiDryer.so not built")
@pytest.mark.parametrize("stock,stock_cuda,ours,ours_cuda,tol", CASES)
def test_pmf_with_b200_reactor_matches_stock_pelec(stock, stock_cuda, ours, ours_cuda, tol, tmp_path):
    if (stock_cuda or ours_cuda) and not os.path.exists(LIB_CUDA):
        pytest.skip("oracle/_ref/libpelec_pmf_cuda_LiDryer.so not built")
    ode = ["ode.rtol=1e-8", "ode.atol=1e-10", "cvode.max_order=3"]
    a = run(stock, tmp_path, ode, stock_cuda)
    b = run(ours, tmp_path, ode, ours_cuda)
    assert a["finest"] == b["finest"] >= 1 and a["time"] == b["time"]  # same dt sequence (CFL + reactions)
    worst, hr_worst = {}, 0.0
    for lev in range(int(a["finest"]) + 1):
        assert np.array_equal(a["boxes%d" % lev], b["boxes%d" % lev])  # same regridding
        sa, sb = a["S%d" % lev], b["S%d" % lev]
        for c in range(sa.shape[0]):
            scale = np.abs(sa[c]).max()
            if c >= 7:
                scale = max(scale, 1e-6 * np.abs(sa[0]).max())  # species: against max(rhoY_k, 1e-6 rho)
            if c in (1, 2):
                scale = max(scale, np.abs(sa[3]).max())  # transverse momenta: against the axial one
            worst[(lev, c)] = float(np.abs(sa[c] - sb[c]).max() / scale)
        # heat release (I_R's last component): the quantity PeleC plots as the flame marker
        hr = np.abs(a["IR%d" % lev][-1]).max()
        hrel = float(np.abs(a["IR%d" % lev][-1] - b["IR%d" % lev][-1]).max() / hr)
        # I_R = (new - old) / dt - src over a 6e-9 s step amplifies the reactor's tolerance-level differences by |rhoY| / |change|
        assert hrel <= 1e-4, (lev, hrel)
        hr_worst = max(hr_worst, hrel)
    print("PMF %s%s vs %s%s after %d coarse steps: worst relative difference of the state %.2e (%s), of the heat release %.2e; wall %.1f s vs %.1f s" % (
        stock, " (CUDA build)" if stock_cuda else "", ours, " (CUDA build)" if ours_cuda else "", NSTEPS, max(worst.values()), max(worst, key=worst.get), hr_worst, float(a["wall_s"]), float(b["wall_s"])))
    assert max(worst.values()) <= tol, worst


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libpelec_pmf_LiDryer.so not built")
def test_shipped_pmf_input_with_only_the_integrator_key_changed(tmp_path):
    """The shipped pmf-lidryer-cvode.inp (cvode.solve_type = GMRES, default tolerances) with pelec.chem_integrator as the ONLY
    change, plus the two PeleC options that feed the reactor: typical values (PeleC::set_typical_values_chem,
    PC/Source/React.cpp:8-27 -> ReactorBase::set_typ_vals_ode) and the reaction work estimate (React.cpp:316-322).  The stock run
    solves every cell's Newton systems with matrix-free GMRES, the B200 reactor with a dense LU: same BDF, answers within the
    integrator's tolerance (rtol 1e-6)."""
    extra = ["cvode.solve_type=GMRES", "pelec.use_typ_vals_chem=1", "amr.loadbalance_with_workestimates=1"]
    a = run("ReactorCvode", tmp_path, extra)
    b = run("ReactorCvodeB200", tmp_path, extra)
    assert a["finest"] == b["finest"] >= 1 and a["time"] == b["time"]
    worst = 0.0
    for lev in range(int(a["finest"]) + 1):
        assert np.array_equal(a["boxes%d" % lev], b["boxes%d" % lev])
        sa, sb = a["S%d" % lev], b["S%d" % lev]
        for c in range(sa.shape[0]):
            scale = max(np.abs(sa[c]).max(), 1e-6 * np.abs(sa[0]).max() if c >= 7 else 0.0, np.abs(sa[3]).max() if c in (1, 2) else 0.0)
            worst = max(worst, float(np.abs(sa[c] - sb[c]).max() / scale))
    print("PMF, shipped input (GMRES, typical values, load balance): worst relative difference of the state %.2e" % worst)
    assert worst <= 1e-5

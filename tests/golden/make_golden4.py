#!/usr/bin/env python
"""Generate the fixtures of the pointwise kernels either side of the reactor from the UNMODIFIED reference
(run in the build container only):

  tests/golden/ctoprim_pmf_LiDryer.npz   pc_ctoprim (PC/Source/Utilities.H:242-321) of the PeleC application library
                                         (oracle/_ref/libpelec_pmf_LiDryer.so, pelecref_ctoprim) on the 628 PMF flame cells of
                                         the react_state fixture, with randomly perturbed temperature guesses
  tests/golden/transport_<mech>.npz      SimpleTransport::transport (PP/Source/Transport/Simple.H:230-383) of
                                         oracle/_ref/libpeleref_<mech>.so (peleref_transport) on 384 flame-table states
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CASE = "/root/reference/Exec/RegTests/PMF"


def ctoprim_worker(out):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpelec_pmf_LiDryer.so"))
    args = ["pelec", CASE + "/pmf-lidryer-cvode.inp", "amr.plot_files_output=0", "amr.checkpoint_files_output=0",
            "pelec.xNormPlane.int=100000", "prob.pmf_datafile=" + CASE + "/LiDryer_H2_p1_phi0_4000tu0300.dat",
            "amr.max_level=0", "pelec.v=0", "amr.v=0", "pelec.sum_interval=-1"]
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    lib.pelecref_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.pelecref_init(len(args), argv)  # EOS / problem parameters live in the application
    g = np.load(os.path.join(HERE, "react_state_pmf_LiDryer_rk64.npz"))
    u = np.ascontiguousarray(g["sold"])
    rng = np.random.default_rng(11)
    u[6] += rng.uniform(-30.0, 30.0, u.shape[1])  # the T guess of the Newton iteration
    u[7 + 3, ::7] = 1e-22 * u[0, ::7]             # |Y_H| below 1e-4 eps: flushed to zero by pc_ctoprim
    n = u.shape[1]
    qvar, nqaux = lib.pelecref_qvar(), lib.pelecref_nqaux()
    q, qa = np.zeros((qvar, n)), np.zeros((nqaux, n))
    lib.pelecref_ctoprim.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pelecref_ctoprim(n, u.ctypes.data, q.ctypes.data, qa.ctypes.data)
    np.savez_compressed(out, u=u, q=q, qa=qa, note="pc_ctoprim of the unmodified PeleC (tests/golden/make_golden4.py)")
    print("ctoprim LiDryer: %d cells, T %.0f..%.0f K, p %.4g..%.4g, c %.4g..%.4g" % (
        n, q[7].min(), q[7].max(), q[5].min(), q[5].max(), qa[1].min(), qa[1].max()))


def transport_fixture(mech):
    from oracle.pyoracle import Reference
    ref = Reference(mech)
    z = np.load(os.path.join(HERE, "pmf_states_%s.npz" % mech))
    rY, T = z["rY"], z["T"]
    rho = rY.sum(axis=0)
    idx = np.unique(np.linspace(0, len(T) - 1, 384).astype(int))
    Y = np.ascontiguousarray((rY / rho).T[idx])
    Y[::5, 0] = 0.0  # a species missing altogether: the trace-species regularisation of Simple.H:262-264
    D, mu, xi, lam = ref.transport(T[idx], rho[idx], Y)
    np.savez_compressed(os.path.join(HERE, "transport_%s.npz" % mech), T=T[idx], rho=rho[idx], Y=Y, D=D, mu=mu, xi=xi, lam=lam,
                        note="SimpleTransport::transport of the unmodified reference (tests/golden/make_golden4.py)")
    return "transport %s: %d states, T %.0f..%.0f K, mu %.3g..%.3g, lam %.3g..%.3g" % (
        mech, len(idx), T[idx].min(), T[idx].max(), mu.min(), mu.max(), lam.min(), lam.max())


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "ctoprim":
        ctoprim_worker(sys.argv[2])
        return
    tmp = "/tmp/pelec_golden_ctoprim"
    os.makedirs(tmp, exist_ok=True)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "ctoprim", os.path.join(HERE, "ctoprim_pmf_LiDryer.npz")],
                       cwd=tmp, capture_output=True, text=True)
    print(p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-2000:])
    so = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    msgs = [transport_fixture(m) for m in ("LiDryer", "drm19", "dodecane_lu")]
    sys.stdout.flush()
    os.dup2(so, 1)
    print("\n".join(msgs))


if __name__ == "__main__":
    main()

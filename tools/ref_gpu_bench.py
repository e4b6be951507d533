#!/usr/bin/env python
"""MEASUREMENT INFRASTRUCTURE: time the UNMODIFIED reference's own GPU reactors on this box's GPU.

Loads oracle/_ref/libpeleref_gpu_<mech>.so (PelePhysics + AMReX + SUNDIALS compiled for CUDA sm_100a by
oracle/ref_gpu/Makefile) and calls ReactorBase::react(box, Array4..., stream) - the AMREX_USE_GPU region of
PP/Source/Reactions/ReactorCvode.cpp:1315-1376 (whole box = ONE coupled CVODE system on N_Vector_Cuda) or
ReactorRK64.cpp:166-320 - on the ReactEval states of bench.py's workload, CUDA-event timed around the react call.
Prints ONE JSON line.  The second stated baseline of VERDICT r1 item 9; bench.py runs it in a subprocess
(`ref_gpu_baseline`), so AMReX's device arena and signal handling never share a process with the product library.

  python tools/ref_gpu_bench.py --mech LiDryer --n 64 --reactor ReactorCvode --solve-type GMRES
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mech", default="LiDryer")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--reactor", default="ReactorCvode", choices=["ReactorCvode", "ReactorRK64", "ReactorArkode", "ReactorCvodeB200", "ReactorRK64B200", "ReactorArkodeB200"],
                    help="...B200: this repo's reactor through the shim's AMREX_USE_GPU branch (device pointers, AMReX's stream)")
    ap.add_argument("--solve-type", default="GMRES")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--atol", type=float, default=1e-10)
    ap.add_argument("--max-order", type=int, default=4)
    ap.add_argument("--dt", type=float, default=1e-6)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare a j-line with the pinned CPU oracle")
    a = ap.parse_args()

    path = os.path.join(ROOT, "oracle", "_ref", "libpeleref_gpu_%s.so" % a.mech)
    if not os.path.exists(path):
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libpeleref_gpu_%s.so not built" % a.mech}))
        return
    # states from the CPU oracle (same generator bench.py's CPU leg uses)
    so = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # AMReX banners
    try:
        from oracle.pyoracle import Oracle, reacteval_states
        cpu = Oracle(a.mech)  # the pinned C restatement: no second copy of AMReX in this process
        n = a.n
        st = reacteval_states(cpu, (n, n, n))
        N = cpu.N
        lib = C.CDLL(path)
        lib.peleref_gpu_reactor_create.restype = C.c_void_p
        lib.peleref_gpu_reactor_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.peleref_gpu_set_parm.argtypes = [C.c_char_p] * 3
        lib.peleref_gpu_react_box_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7 + \
            [C.c_double, C.c_double, C.c_int, C.c_void_p]
        lib.peleref_gpu_init()

        def parm(p, k, v):
            lib.peleref_gpu_set_parm(p.encode(), k.encode(), str(v).encode())
        parm("ode", "rtol", repr(a.rtol))
        parm("ode", "atol", repr(a.atol))
        parm("ode", "verbose", 0)
        parm("ode", "analytical_jacobian", 1)
        parm("ode", "clean_init_massfrac", 0)
        parm("cvode", "solve_type", a.solve_type)
        parm("cvode", "max_order", a.max_order)
        if a.reactor.endswith("B200"):
            parm("cvode", "solve_type", "denseAJ_direct")
        h = lib.peleref_gpu_reactor_create(a.reactor.encode(), 1, n * n * n)
        idx = np.arange(0, n * n, n)  # one cell per j (i = 0, k = 0)
        s1 = st.take(idx)  # taken BEFORE the call: react works in place
        rY = np.ascontiguousarray(st.rY, dtype=np.float64)
        T = np.ascontiguousarray(st.T, dtype=np.float64)
        rE = np.ascontiguousarray(st.rE, dtype=np.float64)
        rYsrc = np.zeros_like(rY)
        rEsrc = np.zeros_like(T)
        FC = np.zeros_like(T)
        mask = np.ones(T.shape, dtype=np.int32)
        ms = np.zeros(a.reps, dtype=np.float32)
        rc = lib.peleref_gpu_react_box_host(
            h, n, n, n, rY.ctypes.data, rYsrc.ctypes.data, T.ctypes.data, rE.ctypes.data, rEsrc.ctypes.data,
            FC.ctypes.data, mask.ctypes.data, a.dt, 0.0, a.reps, ms.ctypes.data)
        check = None
        if a.check:
            if a.reactor.startswith("ReactorCvode"):
                cpu.react_cvode(s1, a.dt, rtol=a.rtol, atol=a.atol, max_order=a.max_order, solve_type="denseAJ_direct")
            else:
                cpu.react_rk64(s1, a.dt, atol=a.atol)
            rho = s1.rY.sum(axis=0)
            check = {
                "max_drY_over_bar": float(np.max(np.abs(rY[:, idx] - s1.rY) / (10.0 * (a.atol + a.rtol * np.abs(s1.rY))))),
                "max_rel_dT_vs_cpu_oracle": float(np.max(np.abs(T[idx] - s1.T) / s1.T)),
                "max_drY_over_rho_vs_cpu_oracle": float(np.max(np.abs(rY[:, idx] - s1.rY) / rho[None, :])),
            }
    finally:
        sys.stdout.flush()
        os.dup2(so, 1)
    best = float(ms.min())
    line = {
        "impl": "reference-gpu" if not a.reactor.endswith("B200") else "b200-in-reference-gpu-build", "kind": "reference", "mech": a.mech, "reactor": a.reactor,
        "solve_type": (a.solve_type if a.reactor == "ReactorCvode" else "denseAJ_direct" if a.reactor == "ReactorCvodeB200" else None),
        "cells": n ** 3, "box": "%d^3" % n, "ms_per_call": [float(x) for x in ms], "ms_best": best,
        "value": n ** 3 / (best * 1e-3), "unit": "cells/s", "react_return": int(rc),
        "mean_FC": float(FC.mean()), "rtol": a.rtol, "atol": a.atol, "max_order": a.max_order, "dt": a.dt,
        "check": check,
        "note": "unmodified PelePhysics GPU path built for sm_100a (oracle/ref_gpu); CUDA events around ReactorBase::react "
                "with device-resident FABs; the GPU ReactorCvode integrates the box as one coupled system",
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

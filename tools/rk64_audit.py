#!/usr/bin/env python3
"""RK64 parity audit on the GPU ("prove the FMA argument on the device").

Runs ReactorRK64 of the library selected by PRX_LIB_PATH (default build: FMA contraction on; `make audit` build:
-fmad=false) on (a) the ReactEval golden states, compared with the UNMODIFIED reference's output stored in
tests/golden/reactor_<mech>.npz, and (b) 2048 flame-table states (tests/golden/pmf_states_<mech>.npz), compared with the C
oracle (bit-exact restatement of the reference's RK64, tests/test_oracle.py).  Prints one JSON object: per mechanism and state
set the PURE relative error |d rhoY_k| / |rhoY_k| over all entries with Y_k >= 1e-15 (max, and a histogram of its decimal
exponent per species), the relative temperature error, and the error relative to max(Y_k, 1e-2) (the floor bar of tests/util.py).
usage: [PRX_LIB_PATH=...] python tools/rk64_audit.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def errors(got_rY, got_T, ref_rY, ref_T, species):
    rho = ref_rY.sum(axis=0, keepdims=True)
    Y = ref_rY / rho
    live = Y >= 1e-15
    rel = np.where(live, np.abs(got_rY - ref_rY) / np.maximum(np.abs(ref_rY), 1e-300), 0.0)
    resolved = np.abs(ref_rY) >= 1e-10  # entries above the integrator's own absolute tolerance (ode.atol, g/cm^3)
    rel_res = np.where(resolved, rel, 0.0)
    floor = np.abs(got_rY - ref_rY) / np.maximum(np.abs(ref_rY), 1e-2 * rho)
    hist = {}
    for k, sp in enumerate(species):
        e = rel[k][live[k] & (rel[k] > 0)]
        if e.size:
            ex, cnt = np.unique(np.floor(np.log10(e)).astype(int), return_counts=True)
            hist[sp] = {str(int(a)): int(b) for a, b in zip(ex, cnt)}
    worst = int(np.argmax(rel.max(axis=1)))
    return {"cells": int(ref_rY.shape[1]), "max_rel_species": float(rel.max()), "worst_species": species[worst],
            "max_rel_species_above_atol": float(rel_res.max()), "entries_above_atol": int(resolved.sum()),
            "worst_species_Y": float(Y[worst, int(np.argmax(rel[worst]))]),
            "max_rel_T": float((np.abs(got_T - ref_T) / ref_T).max()), "max_rel_species_floor_1e-2": float(floor.max()),
            "frac_entries_above_1e-10": float((rel[live] > 1e-10).mean()), "hist_log10_rel_by_species": hist}


def main():
    from oracle.pyoracle import BoxState, Oracle
    from pelec_b200.reactor import LIB_PATH, ReactorB200
    out = {"library": os.path.basename(LIB_PATH)}
    for mech in ("LiDryer", "drm19"):
        orc = Oracle(mech)
        r = ReactorB200.create("ReactorRK64", mech, ode__atol=1e-10)
        r.init(1, 1)
        g = np.load(os.path.join(ROOT, "tests", "golden", "reactor_%s.npz" % mech))
        s = BoxState(g["rY0"], g["T0"], g["rE0"], g["rYsrc"], g["rEsrc"])
        r.react_host(s.ncells, 1, 1, s.rY, s.rYsrc, s.T, s.rE, s.rEsrc, s.FC, s.mask, float(g["dt"]), 0.0)
        out[mech + ":reacteval_vs_reference"] = errors(s.rY, s.T, g["rk64_rY"], g["rk64_T"], orc.species)
        out[mech + ":reacteval_vs_reference"]["steps_equal"] = bool(np.array_equal(s.FC, g["rk64_FC"]))
        z = np.load(os.path.join(ROOT, "tests", "golden", "pmf_states_%s.npz" % mech))
        a = BoxState(z["rY"].copy(), z["T"].copy(), z["rE"].copy())
        b = a.copy()
        r.react_host(a.ncells, 1, 1, a.rY, a.rYsrc, a.T, a.rE, a.rEsrc, a.FC, a.mask, 1e-6, 0.0)
        orc.react_rk64(b, 1e-6, atol=1e-10)
        ok = np.isfinite(b.rY).all(axis=0) & np.isfinite(b.T)
        out[mech + ":flame_vs_oracle"] = errors(a.rY[:, ok], a.T[ok], b.rY[:, ok], b.T[ok], orc.species)
        out[mech + ":flame_vs_oracle"]["steps_equal_frac"] = float((a.FC[ok] == b.FC[ok]).mean())
        r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

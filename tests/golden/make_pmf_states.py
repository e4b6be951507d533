"""Generates tests/golden/pmf_states_<mech>.npz: "Set B" of SURVEY.md 8(d) - flame-like thermochemical states sampled
from the reference's premixed-flame tables (PC/Exec/RegTests/PMF/*.dat) at seeded pseudo-random positions, with the
reference's own interpolation rule (linear between table points, PC/Exec/RegTests/PMF/prob.H pmf()), mole fractions
-> mass fractions -> (rho, e) at 1 atm with the oracle's X2Y / PYT2RE (bit-exact restatements of the reference EOS).
Unlike ReactEval's set A (128 distinct states, identical along i), neighbouring cells here sit at unrelated flame
positions - cold reactants, the reaction zone, hot products - so the lanes of a warp take different numbers of steps.

Run in the build container (needs /root/reference):  python tests/golden/make_pmf_states.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle  # noqa: E402

TABLES = {"LiDryer": "LiDryer_H2_p1_phi0_4000tu0300.dat", "drm19": "PMF_CH4_1bar_300K_DRM_MixAvg.dat",
          "dodecane_lu": "PMF_NC12H26_1bar_300K_DodecaneLu.dat"}
NSTATES_OF = {"dodecane_lu": 512}
NSTATES = 2048


def lcg_uniform(n, seed=12345):
    """64-bit LCG (Knuth MMIX constants), upper 53 bits -> [0, 1)."""
    out, x = np.zeros(n), seed
    for i in range(n):
        x = (6364136223846793005 * x + 1442695040888963407) % (1 << 64)
        out[i] = (x >> 11) / float(1 << 53)
    return out


def main():
    only = sys.argv[1:]
    for mech, fname in TABLES.items():
        if only and mech not in only:
            continue
        NSTATES = NSTATES_OF.get(mech, 2048)
        path = os.path.join("/root/reference/Exec/RegTests/PMF", fname)
        lines = open(path).read().splitlines()
        names = [t.strip('"') for t in lines[0].split("=")[1].split()]
        data = np.array([[float(v) for v in ln.split()] for ln in lines[2:] if ln.strip()])
        orc = Oracle(mech)
        x = data[:, 0]
        u = lcg_uniform(NSTATES)
        xs = x[0] + u * (x[-1] - x[0])
        T = np.interp(xs, x, data[:, names.index("temp")])
        X = np.zeros((NSTATES, orc.N))
        for k, sp in enumerate(orc.species):
            col = sp if sp in names else sp.replace("*", "S")  # the table spells CH2* as CH2S
            X[:, k] = np.interp(xs, x, data[:, names.index(col)])
        X = np.maximum(X, 0.0)
        X /= X.sum(axis=1, keepdims=True)
        Y = orc.x2y(X)
        rho, e = orc.pyt2re(np.full(NSTATES, 1013250.0), Y, T)
        out = os.path.join(ROOT, "tests", "golden", "pmf_states_%s.npz" % mech)
        np.savez_compressed(out, rY=(Y * rho[:, None]).T.copy(), T=T, rE=rho * e, x=xs)
        print(out, os.path.getsize(out), "bytes; T range", T.min(), T.max())


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""TEST INFRASTRUCTURE (build container only): does the mechanism pipeline - tools/extract_mechanism.py (mechanism.H ->
JSON) -> pelec_b200/mechgen/gen.py (JSON -> code) - reproduce the reference for a mechanism it was not written for?

For every requested mechanism of PP/Mechanisms: extract, generate the C flavour, build oracle/_build/liboracle_<mech>.so and the
reference library oracle/_ref/libpeleref_<mech>.so (oracle/ref/Makefile), and compare on random reacting states
  * the ODE right-hand side fKernelSpec (production rates, thermo, T-Newton)   - expected: bit for bit,
  * the analytic Jacobian (re-derived here, not transliterated)                 - expected: ~1e-13 of the column scale,
  * SimpleTransport::transport                                                  - expected: bit for bit.
Prints one line per mechanism and writes profiles/mechanism_coverage.json.

  python tools/validate_mechanisms.py [mech ...]        (default: every mechanism the extractor accepts)
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import extract_mechanism as X  # noqa: E402
from pelec_b200.mechgen import gen  # noqa: E402

FUELS = ("H2", "CH4", "C2H4", "C3H8", "NC7H16", "NC12H26", "IC8H18", "CH3OCH3", "NC10H22", "POSF10325", "CO")


def one(mech, max_species):
    try:
        d = X.extract(mech)
    except Exception as e:
        return {"mech": mech, "status": "not extracted", "why": "%s: %s" % (type(e).__name__, str(e)[:80])}
    n, nr = len(d["species"]), len(d["reactions"])
    res = {"mech": mech, "species": n, "reactions": nr}
    if n > max_species:
        res["status"] = "skipped (more than %d species)" % max_species
        return res
    ident = mech.replace("-", "_")  # file name / C identifier
    path = os.path.join(gen.MECH_JSON_DIR, ident + ".json")
    fresh = not os.path.exists(path)
    with open(path, "w") as f:
        json.dump(d, f, indent=0, separators=(",", ":"))
    try:
        gen.generate_all(ROOT, (ident,))
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "MECHS=" + ident], stdout=subprocess.DEVNULL)
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref"), "-j8", "MECHS=" + mech],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        code = r'''
import sys, json, os
import numpy as np
sys.path.insert(0, %r)
os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
from oracle.pyoracle import Oracle, Reference
mech = %r
o, r = Oracle(mech.replace("-", "_")), Reference(mech)
rng = np.random.default_rng(5)
N = o.N
fuel = next((f for f in %r if f in o.species), o.species[0])
n = 48
X = np.full((n, N), 1e-4) + rng.uniform(0, 2e-3, (n, N))
X[:, o.species.index(fuel)] += 0.1
if "O2" in o.species: X[:, o.species.index("O2")] += 0.2
if "N2" in o.species: X[:, o.species.index("N2")] += 0.7
X /= X.sum(axis=1, keepdims=True)
Y = o.x2y(X)
T = rng.uniform(900.0, 2600.0, n)
rho, e = o.pyt2re(np.full(n, 1013250.0), Y, T)
y = np.concatenate([Y * rho[:, None], T[:, None]], axis=1)
z = np.zeros(n)
fa = o.fkernelspec(0.0, y, rho * e, z, np.zeros((n, N)))
fb = r.fkernelspec(0.0, y, rho * e, z, np.zeros((n, N)))
scale = np.abs(fb).max(axis=0) + 1e-300
Ja, Jb = o.rty2jac(rho, T, Y), r.rty2jac(rho, T, Y)
Ja, Jb = Ja.reshape(n, N + 1, N + 1), Jb.reshape(n, N + 1, N + 1)
jsc = np.abs(Jb).max(axis=2, keepdims=True) + 1e-300
out = {"rhs_bit_exact": bool(np.array_equal(fa, fb)), "rhs_max_rel": float((np.abs(fa - fb) / scale).max()),
       "jac_max_rel_of_column": float((np.abs(Ja - Jb) / jsc).max()), "fuel": fuel}
try:
    ta, tb = o.transport(T, rho, Y), r.transport(T, rho, Y)
    out["transport_bit_exact"] = bool(all(np.array_equal(a, b) for a, b in zip(ta, tb)))
except Exception as ex:
    out["transport_bit_exact"] = "n/a: " + type(ex).__name__
sys.stderr.write("RESULT " + json.dumps(out) + "\n")
''' % (ROOT, mech, FUELS)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
        line = [l for l in p.stderr.splitlines() if l.startswith("RESULT ")]
        if not line:
            res.update(status="comparison failed", why=p.stderr[-300:])
        else:
            res.update(json.loads(line[0][7:]))
            good = res["rhs_max_rel"] <= 1e-12 and res["jac_max_rel_of_column"] <= 1e-9
            res["status"] = "reproduces the reference" if good else "MISMATCH"
    except Exception as e:
        res.update(status="pipeline failed", why="%s: %s" % (type(e).__name__, str(e)[:120]))
    finally:
        if fresh and "--keep" not in sys.argv:  # leave nothing behind for a mechanism the repo does not ship
            os.remove(path)
            for f in (os.path.join(ROOT, "oracle", "gen", "mech_%s.h" % ident), os.path.join(ROOT, "oracle", "_build", "liboracle_%s.so" % ident),
                      os.path.join(ROOT, "pelec_b200", "csrc", "gen", "mech_%s.cuh" % ident),
                      os.path.join(ROOT, "pelec_b200", "csrc", "gen", "mech_%s.mk" % ident),
                      os.path.join(ROOT, "oracle", "_ref", "libpeleref_%s.so" % mech)):
                if os.path.exists(f):
                    os.remove(f)
            subprocess.call(["rm", "-rf", os.path.join(ROOT, "oracle", "_ref", "obj", "pp_" + mech)])
    return res


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    max_species = 64
    mechs = args or sorted(m for m in os.listdir(X.MECH_DIR) if os.path.exists(os.path.join(X.MECH_DIR, m, "mechanism.H")))
    out = []
    for m in mechs:
        r = one(m, max_species)
        out.append(r)
        print(json.dumps(r))
        sys.stdout.flush()
    if not args:
        with open(os.path.join(ROOT, "profiles", "mechanism_coverage.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Cantera YAML -> mechanism table (``pelec_b200/mechanisms/<name>.json``): the from-YAML front end of the mechanism
pipeline (the reference's counterpart is CEPTR, ``PP/Support/ceptr``, which needs Cantera and pint -
neither is used here: the YAML is read with PyYAML and converted with the arithmetic below).

    python tools/yaml_to_mechanism.py path/to/mechanism.yaml NAME [--transport-from other.json]

What it produces is the same table ``tools/extract_mechanism.py`` extracts from a CEPTR-generated ``mechanism.H``:
species / molecular weights, the NASA-7 polynomial forms the kernels evaluate (cv_R, cp_R, gibbs, internal energy,
enthalpy, d cv/dT), and per reaction the Arrhenius / third-body / Troe / Lindemann parameters in the units the generated
code computes in (mol, m^3, K), the stoichiometry, and the reference's reaction ORDER (Troe falloff, SRI, Lindemann,
third-body - including Cantera 3's implicit "explicit collider" reactions - then the rest; ``reaction_info.py:36-128``).
``pelec_b200/mechgen/gen.py`` then emits the CUDA / C code from it.

Two routes, two purposes.  The mechanism.H route carries the reference's own rounded literals and is what the parity tests
pin (RK64 to 1e-10).  This route makes a mechanism usable WITHOUT a CEPTR run; its numbers agree with the
header's to the last printed digits or ~1e-15 relative (``tests/test_yaml_frontend.py`` compares the two tables for the five
shipped mechanisms and the generated right-hand side against the reference library).  Not covered (rejected with an error):
SRI falloff, PLOG / Chebyshev rates, non-integer or explicit reaction orders, QSS species, NASA-9 / Shomate thermo.
The Simple-transport fit coefficients (COFETA / COFLAM / COFD) are NOT derived here (CEPTR fits them from collision
integrals): ``--transport-from`` copies the section from an existing table; without it the reactor path (react, react_state,
computeTemp, ctoprim) works and the transport kernel is unavailable for the mechanism.
"""
import argparse
import json
import os
import re
import sys

import yaml

# standard atomic weights as Cantera 3 carries them (the reference's mechanism.H prints their sums with 6 decimals)
ATOMIC_WEIGHTS = {
    "H": 1.008, "D": 2.0141017781, "He": 4.002602, "C": 12.011, "N": 14.007, "O": 15.999, "F": 18.998403163,
    "Ne": 20.1797, "S": 32.06, "Cl": 35.45, "Ar": 39.95, "Si": 28.085, "E": 0.000548579909,
}
GAS_CONSTANT = 8314.46261815324  # J / (kmol K)
# The reference converts activation energies to cal/mol and divides by its own constant Rc (cal/mol/K, ceptr/constants.py:11;
# 5.7e-6 relative above CODATA's R / 4.184): reproduced, since the generated code is what the parity tests pin.
RC_CAL = 1.98721558317399615845
RU_CGS = "8.31446261815324e+07"
CAL = 4.184

LENGTH_TO_M = {"m": 1.0, "cm": 1e-2, "mm": 1e-3}
QUANTITY_TO_MOL = {"mol": 1.0, "kmol": 1e3, "molec": 1.0 / 6.02214076e23}
ENERGY_TO_J_PER_KMOL = {  # activation-energy unit -> J/kmol
    "cal/mol": CAL * 1e3, "kcal/mol": CAL * 1e6, "J/mol": 1e3, "kJ/mol": 1e6, "J/kmol": 1.0, "K": GAS_CONSTANT,
}


def g15(x):
    return "%.15g" % x


def atom_weight(el):
    key = el if el in ATOMIC_WEIGHTS else el.capitalize()
    if key not in ATOMIC_WEIGHTS:
        raise NotImplementedError("atomic weight of element %r" % el)
    return ATOMIC_WEIGHTS[key]


# ---- thermo ----------------------------------------------------------------------------------------------------------
def nasa7_forms(lo, hi, tmid):
    """The six polynomial forms of mechanism.H for one species from its NASA-7 coefficient rows (a0..a6)."""
    def forms(a):
        a0, a1, a2, a3, a4, a5, a6 = a
        e8, e15 = (lambda v: "%.8e" % v), (lambda v: "%.15e" % v)
        return {
            "cv_R": [[e8(a0 - 1.0), "1"], [e8(a1), "T"], [e8(a2), "T2"], [e8(a3), "T3"], [e8(a4), "T4"]],
            "cp_R": [[e8(a0), "1"], [e8(a1), "T"], [e8(a2), "T2"], [e8(a3), "T3"], [e8(a4), "T4"]],
            "gibbs": [[e15(a5), "invT"], [e15(a0 - a6), "1"], [e15(-a0), "logT"], [e15(-a1 / 2.0), "T"],
                      [e15(-a2 / 6.0), "T2"], [e15(-a3 / 12.0), "T3"], [e15(-a4 / 20.0), "T4"]],
            "speciesInternalEnergy": [[e8(a0 - 1.0), "1"], [e8(a1 / 2.0), "T"], [e8(a2 / 3.0), "T2"], [e8(a3 / 4.0), "T3"],
                                      [e8(a4 / 5.0), "T4"], [e8(a5), "invT"]],
            "speciesEnthalpy": [[e8(a0), "1"], [e8(a1 / 2.0), "T"], [e8(a2 / 3.0), "T2"], [e8(a3 / 4.0), "T3"],
                                [e8(a4 / 5.0), "T4"], [e8(a5), "invT"]],
            "dcvpRdT": [[e8(a1), "1"], [e8(2.0 * a2), "T"], [e8(3.0 * a3), "T2"], [e8(4.0 * a4), "T3"]],
        }
    # the reference treats a species whose two coefficient rows agree to rtol 1e-5 (numpy allclose) as one range and
    # evaluates the HIGH-temperature row everywhere ("species with no change across T", ceptr/thermo.py:243-247)
    same = all(abs(x - y) <= 1e-22 + 1e-5 * abs(y) for x, y in zip(hi, lo))
    if same:
        lo, tmid = hi, None
    f_lo, f_hi = forms(lo), forms(hi)
    out = {}
    for k in f_lo:
        # exact-zero terms are not printed
        t_lo = [t for t in f_lo[k] if float(t[0]) != 0.0] or [["0.0", "1"]]
        t_hi = [t for t in f_hi[k] if float(t[0]) != 0.0] or [["0.0", "1"]]
        out[k] = {"tmid": tmid, "lo": t_lo, "hi": t_hi}
    return out


def species_thermo(sp):
    th = sp["thermo"]
    if th.get("model") != "NASA7":
        raise NotImplementedError("thermo model %r of species %s" % (th.get("model"), sp["name"]))
    rows = [[float(v) for v in row] for row in th["data"]]
    tr = [float(v) for v in th["temperature-ranges"]]
    if len(rows) == 1:
        return nasa7_forms(rows[0], rows[0], None)
    if len(rows) != 2 or len(tr) != 3:
        raise NotImplementedError("NASA7 with %d ranges (species %s)" % (len(rows), sp["name"]))
    return nasa7_forms(rows[0], rows[1], float(tr[1]))


# ---- reactions -------------------------------------------------------------------------------------------------------
def parse_side(text, names):
    """'2 OH + M' -> ({species: coefficient}, has_M, has_falloff_M)."""
    text = text.strip()
    falloff = False
    m = re.search(r"\(\s*\+\s*([^)\s]+)\s*\)\s*$", text)
    collider = None
    if m:
        falloff = True
        collider = m.group(1)
        text = text[:m.start()].strip()
    out = {}
    third = False
    for tok in re.split(r"\s\+\s", text):
        tok = tok.strip()
        if tok == "M":
            third = True
            continue
        mm = re.match(r"^(\d+(?:\.\d*)?)\s+(\S+)$", tok)
        coef, name = (float(mm.group(1)), mm.group(2)) if mm else (1.0, tok)
        if name not in names:
            raise ValueError("unknown species %r in %r" % (name, text))
        if coef != int(coef):
            raise NotImplementedError("non-integer stoichiometric coefficient in %r" % text)
        out[name] = out.get(name, 0.0) + coef
    return out, third, falloff, collider


def arrhenius(rate, order, conc_to_si, e_to_k):
    """{A, b, Ea} in the file's units -> strings in (m^3/mol)^(order-1)/s, -, K (None for an exact zero, as the header omits
    the term).  e_to_k: activation-energy unit -> K."""
    A = float(rate["A"]) * conc_to_si ** (order - 1)
    b = float(rate.get("b", 0.0))
    ea = float(rate.get("Ea", 0.0)) * e_to_k
    return g15(A), (g15(b) if b != 0.0 else None), (g15(ea) if ea != 0.0 else None)


def convert(path, name, transport_from=None):
    # BaseLoader: every scalar stays a string (YAML 1.1 would read the species NO as a boolean); numbers are converted where used
    doc = yaml.load(open(path), Loader=yaml.BaseLoader)
    units = doc.get("units", {})
    length = LENGTH_TO_M[units.get("length", "m")]
    quantity = QUANTITY_TO_MOL[units.get("quantity", "kmol")]
    conc_to_si = length ** 3 / quantity  # (file volume per file quantity) in m^3/mol
    e_unit = units.get("activation-energy", "J/kmol")
    if e_unit not in ENERGY_TO_J_PER_KMOL:
        raise NotImplementedError("activation-energy unit %r" % e_unit)
    phase = doc["phases"][0]
    by_name = {s["name"]: s for s in doc["species"]}
    names = phase["species"]
    if names == "all" or names == ["all"]:
        names = [s["name"] for s in doc["species"]]
    names = [str(n) for n in names]
    index = {n: i for i, n in enumerate(names)}

    mw, imw = [], []
    thermo = {k: [] for k in ("cv_R", "cp_R", "gibbs", "speciesInternalEnergy", "speciesEnthalpy", "dcvpRdT")}
    for n in names:
        sp = by_name[n]
        w = sum(float(c) * atom_weight(el) for el, c in sp["composition"].items())
        mw.append("%f" % w)
        imw.append("%.16f" % (1.0 / w))
        forms = species_thermo(sp)
        for k in thermo:
            thermo[k].append(forms[k])

    records = []
    for idx, r in enumerate(doc["reactions"]):
        eq = r["equation"]
        for key in ("orders", "negative-orders", "nonreactant-orders", "rate-constants", "data", "SRI", "coverage-dependencies"):
            if key in r:
                raise NotImplementedError("reaction %d (%s): %r is not supported" % (idx, eq, key))
        m = re.match(r"^(.*?)\s(<=>|=>|=)\s(.*)$", eq)
        if not m:
            raise ValueError("cannot parse equation %r" % eq)
        lhs, thirdL, fallL, colL = parse_side(m.group(1), index)
        rhs, thirdR, fallR, colR = parse_side(m.group(3), index)
        reversible = m.group(2) != "=>"
        rtype = r.get("type", "elementary")
        if rtype not in ("elementary", "three-body", "falloff"):
            raise NotImplementedError("reaction %d (%s): type %r" % (idx, eq, rtype))
        if (fallL or fallR) and (colL != "M" or colR != "M"):
            raise NotImplementedError("reaction %d (%s): falloff with a specific collider" % (idx, eq))
        full_lhs, full_rhs = dict(lhs), dict(rhs)  # what multiplies the rate constant (concentration products)
        order = sum(lhs.values())
        # Cantera 3: an untyped reaction with ONE species on both sides and three molecules on a side is a third-body
        # reaction with that species as explicit collider: the collider leaves the stoichiometry (Gibbs sum, wdot) but
        # stays in the concentration products, and the reaction sorts with the third-body group
        explicit_collider = None
        if rtype == "elementary" and not (thirdL or thirdR):
            both = [s for s in lhs if s in rhs]
            nl, nr = sum(lhs.values()), sum(rhs.values())
            if len(both) == 1 and (nl == 3 or nr == 3):
                explicit_collider = both[0]
                for side in (lhs, rhs):
                    side[explicit_collider] -= 1.0
                    if side[explicit_collider] == 0.0:
                        del side[explicit_collider]
        if rtype == "three-body":
            if not (thirdL and (thirdR or not reversible or thirdR == thirdL)):
                raise ValueError("reaction %d (%s): three-body without M" % (idx, eq))
            order += 1
        rec = {"idx": idx, "eq": eq}
        e_to_k_rc = ENERGY_TO_J_PER_KMOL[e_unit] / (CAL * 1e3) / RC_CAL
        if rtype == "falloff":
            hi = r["high-P-rate-constant"]
            lo = r["low-P-rate-constant"]
            rec["A"], rec["beta"], rec["Ea"] = arrhenius(hi, order, conc_to_si, e_to_k_rc)
        else:
            rec["A"], rec["beta"], rec["Ea"] = arrhenius(r["rate-constant"], order, conc_to_si, e_to_k_rc)
        if rtype in ("three-body", "falloff"):
            if float(r.get("default-efficiency", 1.0)) != 1.0:
                raise NotImplementedError("reaction %d (%s): default-efficiency != 1" % (idx, eq))
            eff = sorted((index[s], float(e)) for s, e in (r.get("efficiencies") or {}).items() if s in index)
            rec["eff"] = [[k, g15(e - 1.0)] for k, e in eff if e != 1.0]
        if rtype == "falloff":
            A0, b0, E0 = arrhenius(lo, order + 1, conc_to_si, e_to_k_rc)
            rec["low"] = {"A": A0, "beta": b0, "Ea": E0}
            if "Troe" in r:
                t = r["Troe"]
                a = float(t["A"])
                # a term with a zero coefficient or T = 0 (exp(-T/0) = 0) is not printed
                terms = [[g15(c), g15(1.0 / float(t[key]))] for c, key in ((1.0 - a, "T3"), (a, "T1"))
                         if c != 0.0 and float(t[key]) != 0.0]
                t2 = t.get("T2")
                rec["fcent"] = {"exp_terms": terms, "invT_coef": (g15(-float(t2)) if t2 is not None else None),
                                "const": ("0.0" if t2 is None else None)}
                rec["falloff"] = "troe"
            else:
                rec["falloff"] = "lindemann"

        def expand(side):
            out = []
            for s in sorted(side, key=lambda n_: index[n_]):
                out += [index[s]] * int(side[s])
            return out

        def coef(c):
            return "1" if c == 1.0 else "%f" % c

        rec["reactants"] = expand(full_lhs)
        rec["reversible"] = reversible
        if reversible:
            # Gibbs sum and the reference-concentration power of the equilibrium constant (per species: reactants +, products -)
            dg = []
            for k in range(len(names)):
                s = names[k]
                if s in lhs:
                    dg.append([coef(lhs[s]), k])
                if s in rhs:
                    dg.append(["-" + coef(rhs[s]), k])
            rec["dG"] = dg
            rec["refC"] = int(round(sum(lhs.values()) - sum(rhs.values())))
            if abs(rec["refC"]) > 2:
                raise NotImplementedError("reaction %d (%s): |delta n| > 2" % (idx, eq))
            rec["products"] = expand(full_rhs)
        else:
            rec["dG"] = []
            rec["refC"] = 0
            rec["products"] = []
        rec["type"] = {"elementary": "elementary", "three-body": "three_body", "falloff": "falloff"}[rtype]
        wdot = []
        for k in range(len(names)):
            s = names[k]
            if s in lhs:
                wdot.append([k, "-" + coef(lhs[s])])
            if s in rhs:
                wdot.append([k, coef(rhs[s])])
        rec["wdot"] = wdot
        rec["_group"] = (0 if rec.get("falloff") == "troe" else 2 if rec.get("falloff") == "lindemann" else
                         3 if (rtype == "three-body" or explicit_collider is not None) else 4)
        records.append(rec)
    records.sort(key=lambda q: (q["_group"], q["idx"]))  # stable: original order inside a group
    for q in records:
        del q["_group"]

    out = {
        "name": name,
        "source": "Cantera YAML %s (tools/yaml_to_mechanism.py)" % os.path.basename(os.path.abspath(os.path.dirname(path))),
        "species": names,
        "mw": mw,
        "imw": imw,
        "Ru": RU_CGS,
        "refC_num": "101325",
        "refC_den": "8.31446",
        "T_newton": {"tmin": "90", "tmax": "4000", "tol": "1.e-6", "maxiter": 200, "clamp": "100."},
        "thermo": thermo,
        "reactions": records,
    }
    if transport_from:
        other = json.load(open(transport_from))
        if other["species"] != names:
            raise ValueError("--transport-from: species lists differ")
        out["transport"] = other["transport"]
    return out


# ---- comparison of two tables (test infrastructure for tests/test_yaml_frontend.py) -----------------------------------
def compare_tables(a, b, rtol=1e-12):
    """Largest relative difference of the numbers of two mechanism tables with identical structure; raises AssertionError
    with a message on the first structural difference."""
    worst = [0.0, ""]

    def num(x, y, where):
        x, y = float(x), float(y)
        d = abs(x - y) / max(abs(x), abs(y), 1e-300)
        if d > worst[0]:
            worst[0], worst[1] = d, where
        assert d <= rtol, "%s: %r vs %r (rel %.2e)" % (where, x, y, d)

    def opt(x, y, where):
        if x is None or y is None:
            assert (x is None or float(x) == 0.0) and (y is None or float(y) == 0.0), "%s: %r vs %r" % (where, x, y)
        else:
            num(x, y, where)

    assert a["species"] == b["species"], "species"
    for k, (x, y) in enumerate(zip(a["mw"], b["mw"])):
        num(x, y, "mw[%d]" % k)
    for k, (x, y) in enumerate(zip(a["imw"], b["imw"])):
        num(x, y, "imw[%d]" % k)
    for f in a["thermo"]:
        for k, (x, y) in enumerate(zip(a["thermo"][f], b["thermo"][f])):
            assert (x["tmid"] is None) == (y["tmid"] is None) and (x["tmid"] is None or float(x["tmid"]) == float(y["tmid"])), (f, k)
            for part in ("lo", "hi"):
                tx = {bs: c for c, bs in x[part]}
                ty = {bs: c for c, bs in y[part]}
                for bs in set(tx) | set(ty):
                    # 9 printed digits (cv_R ...) / 16 (gibbs): compare at the printed precision
                    xv, yv = float(tx.get(bs, 0.0)), float(ty.get(bs, 0.0))
                    tol = 2e-9 if f != "gibbs" else 1e-14
                    d = abs(xv - yv) / max(abs(xv), abs(yv), 1e-300) if (xv != 0.0 or yv != 0.0) else 0.0
                    assert d <= tol, "%s[%d].%s %s: %r vs %r" % (f, k, part, bs, xv, yv)
    assert len(a["reactions"]) == len(b["reactions"]), "number of reactions"
    for i, (x, y) in enumerate(zip(a["reactions"], b["reactions"])):
        w = "reaction %d (%s)" % (i, y["eq"])
        assert x["idx"] == y["idx"], w + ": order (%d vs %d)" % (x["idx"], y["idx"])
        for key in ("type", "reversible", "refC", "reactants", "products"):
            assert x[key] == y[key], "%s: %s %r vs %r" % (w, key, x[key], y[key])
        assert x.get("falloff") == y.get("falloff"), w + ": falloff"
        num(x["A"], y["A"], w + " A")
        opt(x["beta"], y["beta"], w + " beta")
        opt(x["Ea"], y["Ea"], w + " Ea")
        for key in ("dG", "wdot"):
            sx = [(int(k), float(c)) if key == "wdot" else (int(c2), float(k2)) for k, c in x[key] for k2, c2 in [(k, c)]]
            sy = [(int(k), float(c)) if key == "wdot" else (int(c2), float(k2)) for k, c in y[key] for k2, c2 in [(k, c)]]
            assert sx == sy, "%s: %s %r vs %r" % (w, key, x[key], y[key])
        ex = {int(k): float(c) for k, c in x.get("eff", [])}
        ey = {int(k): float(c) for k, c in y.get("eff", [])}
        assert set(ex) == set(ey), w + ": third-body species"
        for k in ex:
            num(ex[k] + 1.0, ey[k] + 1.0, w + " efficiency")
        if "low" in y:
            num(x["low"]["A"], y["low"]["A"], w + " low A")
            opt(x["low"]["beta"], y["low"]["beta"], w + " low beta")
            opt(x["low"]["Ea"], y["low"]["Ea"], w + " low Ea")
        if "fcent" in y:
            fx, fy = x["fcent"], y["fcent"]
            assert len(fx["exp_terms"]) == len(fy["exp_terms"]), w + ": Troe terms"
            for (c1, t1), (c2, t2) in zip(fx["exp_terms"], fy["exp_terms"]):
                num(c1, c2, w + " Troe coefficient")
                num(t1, t2, w + " Troe 1/T")
            opt(fx["invT_coef"], fy["invT_coef"], w + " Troe T2")
    return worst


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("yaml")
    ap.add_argument("name")
    ap.add_argument("--transport-from", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    d = convert(args.yaml, args.name, args.transport_from)
    out = args.out or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pelec_b200", "mechanisms", args.name + ".json")
    with open(out, "w") as f:
        json.dump(d, f, indent=0, separators=(",", ":"))
    kinds = {}
    for r in d["reactions"]:
        kinds[r["type"]] = kinds.get(r["type"], 0) + 1
    print(args.name, len(d["species"]), "species", len(d["reactions"]), "reactions", kinds, "->", os.path.relpath(out))


if __name__ == "__main__":
    sys.exit(main())

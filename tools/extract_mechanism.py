#!/usr/bin/env python3
"""Extract the mechanism DATA (species, NASA-7 literals, Arrhenius / third-body / Troe
parameters, stoichiometry) that the reference's CEPTR-generated
``PP/Mechanisms/<mech>/mechanism.H`` encodes, into ``pelec_b200/mechanisms/<mech>.json``.

Why from mechanism.H and not from mechanism.yaml: the generated header carries the
*pre-divided, rounded* literals the reference actually computes with (Ea/R such as
``8352.89343569254``; 9-digit ``cv_R`` vs 16-digit ``gibbs`` coefficients, see
SURVEY.md H1).  Regenerating them from the YAML would differ at ~1e-9 and break the
1e-10 RK64 parity target.  Only numbers and indices are extracted - no code is copied;
``pelec_b200/mechgen`` emits its own CUDA / C from this table.

Run in the build container (needs /root/reference):
    python tools/extract_mechanism.py LiDryer drm19 dodecane_lu
The JSON files are committed (they are the mechanism specification, like golden
fixtures) so that nothing at build or run time on the GPU box needs /root/reference.

Reference anchors (LiDryer line numbers): species/mw ``mechanism.H:120-186``; thermo
``cv_R :199``, ``cp_R :265``, ``gibbs :331``, ``speciesInternalEnergy :532``,
``speciesEnthalpy :615``, ``dcvpRdT :781``; kinetics ``productionRate :2285-2603``.
"""
import json
import os
import re
import sys

REF = os.environ.get("PELE_REFERENCE", "/root/reference")
MECH_DIR = os.path.join(REF, "Submodules/PelePhysics/Mechanisms")
NUM = r"\d+\.?\d*(?:[eE][+-]?\d+)?"


def func_body(src, name):
    """Return the text of the top-level function `name(`...`}` in the generated header."""
    m = re.search(r"^" + re.escape(name) + r"\(", src, re.M)
    if not m:
        raise KeyError(name)
    start = m.start()
    end = src.index("\n}\n", start)
    return src[start:end]


def parse_terms(expr):
    """'+2.29e+00 + 8.2e-04 * T - 1.0e+03 * invT' -> [[signed literal, basis], ...]"""
    expr = expr.replace("\n", " ").strip()
    out = []
    pos = 0
    pat = re.compile(r"\s*([+-])?\s*(" + NUM + r")(?:\s*\*\s*(invT2|invT|logT|T2|T3|T4|T))?")
    while pos < len(expr):
        m = pat.match(expr, pos)
        if not m or m.end() == pos:
            raise ValueError("cannot parse thermo terms: %r at %d" % (expr, pos))
        sign = "-" if m.group(1) == "-" else ""
        out.append([sign + m.group(2), m.group(3) or "1"])
        pos = m.end()
        while pos < len(expr) and expr[pos].isspace():
            pos += 1
    return out


def parse_thermo(src, fname, nsp):
    """Per species: {'tmid': float|None, 'lo': terms, 'hi': terms}."""
    body = func_body(src, fname)
    res = [None] * nsp
    # split into the 'no change' group and the 'midpoint' groups
    chunks = re.split(r"// species with (no change across T|midpoint at T=(\d+) kelvin)", body)
    # chunks: [pre, label, tmid, text, label, tmid, text ...]
    i = 1
    while i < len(chunks):
        label, tmid, text = chunks[i], chunks[i + 1], chunks[i + 2]
        i += 3
        if label.startswith("no change"):
            for m in re.finditer(r"species\[(\d+)\]\s*=\s*([^;]*);", text):
                k = int(m.group(1))
                t = parse_terms(m.group(2))
                res[k] = {"tmid": None, "lo": t, "hi": t}
        else:
            m = re.search(r"if \(T < (\d+)\) \{(.*?)\} else \{(.*)\}", text, re.S)
            assert m and m.group(1) == tmid, (fname, tmid)
            for which, blk in (("lo", m.group(2)), ("hi", m.group(3))):
                for mm in re.finditer(r"species\[(\d+)\]\s*=\s*([^;]*);", blk):
                    k = int(mm.group(1))
                    if res[k] is None:
                        res[k] = {"tmid": float(tmid)}
                    res[k][which] = parse_terms(mm.group(2))
    assert all(r is not None and "lo" in r and "hi" in r for r in res), fname
    return res


def parse_kf(stmt):
    """k_f = A * exp((b) * logT - (E) * invT)  ->  A, b, E  (exponent = b*logT - E*invT)."""
    s = stmt.replace("\n", " ")
    m = re.match(r"\s*(" + NUM + r")\s*(?:\*\s*exp\((.*)\))?\s*$", s)
    assert m, stmt
    A, arg = m.group(1), m.group(2)
    beta = Ea = None
    if arg is not None:
        mb = re.search(r"\((-?" + NUM + r")\)\s*\*\s*logT", arg)
        if mb:
            beta = mb.group(1)
        me = re.search(r"-\s*\((-?" + NUM + r")\)\s*\*\s*invT", arg)
        if me:
            Ea = me.group(1)
        assert mb or me, stmt
        # make sure nothing else hides in the exponent
        rest = arg
        if mb:
            rest = rest.replace(mb.group(0), "")
        if me:
            rest = rest.replace(me.group(0), "")
        assert rest.strip() == "", (stmt, rest)
    return A, beta, Ea


def parse_corr(stmt):
    """Corr = mixture + (c) * sc[i] + sc[j] ...  -> [[i, 'c'], [j, '1'], ...]"""
    s = stmt.replace("\n", " ").strip()
    assert s.startswith("mixture"), stmt
    s = s[len("mixture"):]
    eff = []
    pos = 0
    pat = re.compile(r"\s*\+\s*(?:\((-?" + NUM + r")\)\s*\*\s*)?sc\[(\d+)\]")
    while pos < len(s):
        m = pat.match(s, pos)
        assert m, (stmt, s[pos:])
        eff.append([int(m.group(2)), m.group(1) or "1"])
        pos = m.end()
        while pos < len(s) and s[pos].isspace():
            pos += 1
    return eff


def parse_redp(stmt):
    """redP = Corr / k_f * A0 * exp(c1 * logT - c2 * invT) -> exponent = b0*logT - E0*invT."""
    s = stmt.replace("\n", " ")
    m = re.match(r"\s*Corr / k_f \* (" + NUM + r") \* exp\((.*)\)\s*$", s)
    assert m, stmt
    A0, arg = m.group(1), m.group(2).strip()
    b0 = E0 = None
    mm = re.match(r"^(-?" + NUM + r") \* logT - (-?" + NUM + r") \* invT$", arg)
    if mm:
        b0, E0 = mm.group(1), mm.group(2)
    else:
        mm = re.match(r"^-(-?" + NUM + r") \* invT$", arg)
        if mm:
            E0 = mm.group(1)
        else:
            mm = re.match(r"^(-?" + NUM + r") \* logT$", arg)
            assert mm, stmt
            b0 = mm.group(1)
    return {"A": A0, "beta": b0, "Ea": E0}


def parse_fcent(stmt):
    """logFcent = log10(a*exp(-T*x) + b*exp(-T*y) + exp(c*invT) | const); c is stored signed."""
    s = stmt.replace("\n", " ")
    m = re.match(r"\s*log10\((.*)\)\s*$", s)
    assert m, stmt
    arg = m.group(1).strip()
    terms = []
    pos = 0
    # CEPTR prints every Troe term; absent ones appear as a literal 0.0 (dropping "+ 0.0" is exact), negative coefficients as
    # "+ -0.569 * exp(...)", the leading one sometimes in parentheses
    pat_e = re.compile(r"\s*([+-])?\s*\(?(-?" + NUM + r")\)? \* exp\(-T \* (-?" + NUM + r")\)")
    pat_t2 = re.compile(r"\s*\+\s*exp\((-?" + NUM + r") \* invT\)")
    pat_c = re.compile(r"\s*\+?\s*(" + NUM + r")\s*(?=\+|$)")
    t2 = None
    const = None
    while pos < len(arg):
        m = pat_e.match(arg, pos)
        if m:
            coef = m.group(2)
            if m.group(1) == "-":
                coef = coef[1:] if coef.startswith("-") else "-" + coef
            terms.append([coef, m.group(3)])
            pos = m.end()
            continue
        m = pat_t2.match(arg, pos)
        if m:
            t2 = m.group(1)
            pos = m.end()
            continue
        m = pat_c.match(arg, pos)
        if m:
            if m.end() == len(arg) or float(m.group(1)) != 0.0:  # the trailing constant is kept as printed, even 0.0
                if const is not None and float(const) != 0.0:
                    raise ValueError("Fcent: two constants in %r" % arg)
                const = m.group(1)
            pos = m.end()
            continue
        raise ValueError("Fcent: %r" % arg[pos:])
    return {"exp_terms": terms, "invT_coef": t2, "const": const}


def parse_sc_product(expr):
    return [int(i) for i in re.findall(r"sc\[(\d+)\]", expr)]


def parse_dg(expr):
    """-(g_RT[1] + g_RT[3] - 2.000000 * g_RT[6]) -> [[coef, idx], ...] (coef signed string)."""
    s = expr.replace("\n", " ").strip()
    out = []
    pos = 0
    pat = re.compile(r"\s*([+-])?\s*(?:(" + NUM + r") \* )?g_RT\[(\d+)\]")
    while pos < len(s):
        m = pat.match(s, pos)
        assert m, (expr, s[pos:])
        sign = "-" if m.group(1) == "-" else ""
        out.append([sign + (m.group(2) or "1"), int(m.group(3))])
        pos = m.end()
        while pos < len(s) and s[pos].isspace():
            pos += 1
    return out


def parse_reactions(src, nreac):
    body = func_body(src, "productionRate")
    blocks = re.split(r"\{\s*// reaction (\d+):\s*", body)[1:]
    reactions = []
    for bi in range(0, len(blocks), 2):
        idx = int(blocks[bi])
        text = blocks[bi + 1]
        eq, rest = text.split("\n", 1)
        stm = {}
        order = []
        for s in rest.split(";"):
            s = s.strip()
            if not s or s == "}" or s.startswith("}"):
                s = s.lstrip("} \n")
                if not s:
                    continue
            s = re.sub(r"^(const\s+)?amrex::Real\s+", "", s)
            m = re.match(r"(\w+(?:\[\d+\])?)\s*([-+]?=)\s*(.*)$", s, re.S)
            assert m, s
            order.append((m.group(1), m.group(2), m.group(3)))
        r = {"idx": idx, "eq": eq.strip()}
        kind = "elementary"
        wdot = []
        corr_seen = 0
        for lhs, op, rhs in order:
            rhs1 = " ".join(rhs.split())
            if lhs == "k_f":
                r["A"], r["beta"], r["Ea"] = parse_kf(rhs1)
            elif lhs == "Corr":
                if corr_seen == 0:
                    r["eff"] = parse_corr(rhs1)
                    kind = "three_body"
                else:
                    if rhs1 == "F * F_troe":
                        r["falloff"] = "troe"
                    elif rhs1 == "redP / (1.0 + redP)":
                        r["falloff"] = "lindemann"
                    else:
                        raise ValueError(rhs1)
                corr_seen += 1
            elif lhs == "redP":
                r["low"] = parse_redp(rhs1)
                kind = "falloff"
            elif lhs == "logFcent":
                r["fcent"] = parse_fcent(rhs1)
            elif lhs in ("F", "logPred", "troe_c", "troe_n", "troe", "F_troe", "qdot"):
                # fixed-form Troe algebra / qdot = qf - qr: check the canonical constants
                canon = {
                    "F": "redP / (1.0 + redP)",
                    "logPred": "log10(redP)",
                    "troe_c": "-0.4 - 0.67 * logFcent",
                    "troe_n": "0.75 - 1.27 * logFcent",
                    "troe": "(troe_c + logPred) / (troe_n - 0.14 * (troe_c + logPred))",
                    "F_troe": "exp(M_LN10 * logFcent / (1.0 + troe * troe))",
                    "qdot": "qf - qr",
                }[lhs]
                assert rhs1 == canon, (lhs, rhs1)
            elif lhs == "qf":
                mm = re.match(r"^(Corr \* )?k_f \* \((.*)\)$", rhs1)
                assert mm, rhs1
                r["reactants"] = parse_sc_product(mm.group(2))
                assert bool(mm.group(1)) == (kind != "elementary"), rhs1
            elif lhs == "qr":
                if rhs1 in ("0.0", "0"):
                    r["reversible"] = False
                    r["products"] = []
                    r["dG"] = []
                    r["refC"] = 0
                else:
                    r["reversible"] = True
                    mm = re.match(
                        r"^(Corr \* )?k_f \* exp\(-\((.*?)\)\) \* (?:\(+(refC|refCinv|refC \* refC|refCinv \* refCinv)\)+ \* )?\((.*)\)$",
                        rhs1)
                    assert mm, rhs1
                    assert "refC" not in mm.group(4), rhs1  # |delta n| = 2 is printed with doubled parentheses
                    r["dG"] = parse_dg(mm.group(2))
                    r["refC"] = {None: 0, "refC": 1, "refCinv": -1, "refC * refC": 2,
                                 "refCinv * refCinv": -2}[mm.group(3)]
                    r["products"] = parse_sc_product(mm.group(4))
                    if mm.group(4).strip() in ("0.0", "0"):
                        # irreversible third-body / falloff reactions are printed as "... * (0.0)": no reverse rate
                        r["reversible"] = False
                        r["products"] = []
                        r["dG"] = []
                        r["refC"] = 0
            elif lhs.startswith("wdot["):
                k = int(lhs[5:-1])
                mm = re.match(r"^(?:(" + NUM + r") \* )?qdot$", rhs1)
                assert mm, rhs1
                c = mm.group(1) or "1"
                wdot.append([k, ("-" if op == "-=" else "") + c])
            else:
                raise ValueError("unknown statement %s %s %s" % (lhs, op, rhs1))
        r["type"] = kind
        r["wdot"] = wdot
        if kind == "falloff":
            assert "falloff" in r
        reactions.append(r)
    assert len(reactions) == nreac, (len(reactions), nreac)
    return reactions


def parse_transport(src, nsp):
    """Simple-transport fit data: the literal tables of the egtranset* functions (PP/Source/Transport/TransportParams.H:128-176
    copies them into TransParm: wt, eps, sig, dip, pol, zrot, nlin, fitmu = COFETA, fitlam = COFLAM, fitdbin = COFD)."""
    def table(fn, count, integer=False):
        body = func_body(src, "egtranset" + fn)
        vals = [None] * count
        for m in re.finditer(fn + r"\[(\d+)\] = (-?" + NUM + r");", body):
            vals[int(m.group(1))] = int(m.group(2)) if integer else m.group(2)
        assert None not in vals, fn
        return vals
    # RU / mw_k as CEPTR printed it (CKCVMS / CKCPMS "multiply by R/molecularweight"): not reproducible from the rounded mw
    ru_mw = re.findall(r"cvms\[\d+\] \*= (" + NUM + r");", func_body(src, "CKCVMS"))
    assert len(ru_mw) == nsp
    no = int(re.search(r"\*NO = (\d+);", func_body(src, "egtransetNO")).group(1))
    return {
        "NO": no, "ru_mw": ru_mw, "patm": re.search(r"\*PATM = (" + NUM + r");", func_body(src, "egtransetPATM")).group(1),
        "wt": table("WT", nsp), "eps": table("EPS", nsp), "sig": table("SIG", nsp), "dip": table("DIP", nsp),
        "pol": table("POL", nsp), "zrot": table("ZROT", nsp), "nlin": table("NLIN", nsp, True),
        "cofeta": table("COFETA", no * nsp), "coflam": table("COFLAM", no * nsp), "cofd": table("COFD", no * nsp * nsp),
    }


def extract(mech):
    src = open(os.path.join(MECH_DIR, mech, "mechanism.H")).read()
    nsp = int(re.search(r"#define NUM_GAS_SPECIES (\d+)", src).group(1))
    nreac = int(re.search(r"#define NUM_GAS_REACTIONS (\d+)", src).group(1))
    names = [None] * nsp
    for m in re.finditer(r"#define (\w+)_ID (\d+)", src):
        names[int(m.group(2))] = m.group(1)
    mw = [None] * nsp
    imw = [None] * nsp
    for m in re.finditer(r"mw_new\[(\d+)\] = (" + NUM + r");\s*// (\S+)", func_body(src, "get_mw")):
        mw[int(m.group(1))] = m.group(2)
        names[int(m.group(1))] = m.group(3)
    for m in re.finditer(r"imw_new\[(\d+)\] = (" + NUM + r");", func_body(src, "get_imw")):
        imw[int(m.group(1))] = m.group(2)
    assert None not in mw and None not in imw and None not in names

    thermo = {}
    for f in ("cv_R", "cp_R", "gibbs", "speciesInternalEnergy", "speciesEnthalpy", "dcvpRdT"):
        thermo[f] = parse_thermo(src, f, nsp)

    # CKRP constants and the reference-concentration literal used by productionRate
    ckrp = func_body(src, "CKRP")
    ru = re.search(r"ru = (" + NUM + r");", ckrp).group(1)
    prod = func_body(src, "productionRate")
    if "pow(" in prod:
        raise NotImplementedError("global reactions with non-integer orders (pow in productionRate) are not represented")
    refc = re.search(r"refC = (" + NUM + r") / (" + NUM + r") \* invT;", prod)
    tmin = re.search(r"tmin = (" + NUM + r");", func_body(src, "GET_T_GIVEN_EY")).group(1)
    tmax = re.search(r"tmax = (" + NUM + r");", func_body(src, "GET_T_GIVEN_EY")).group(1)

    return {
        "name": mech,
        "source": "PelePhysics 57614815 Mechanisms/%s/mechanism.H (CEPTR-generated literals)" % mech,
        "species": names,
        "mw": mw,
        "imw": imw,
        "Ru": ru,
        "refC_num": refc.group(1),
        "refC_den": refc.group(2),
        "T_newton": {"tmin": tmin, "tmax": tmax, "tol": "1.e-6", "maxiter": 200, "clamp": "100."},
        "thermo": thermo,
        "reactions": parse_reactions(src, nreac),
        "transport": parse_transport(src, nsp),
    }


def main():
    outdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pelec_b200", "mechanisms")
    os.makedirs(outdir, exist_ok=True)
    for mech in sys.argv[1:] or ["LiDryer", "drm19", "dodecane_lu"]:
        d = extract(mech)
        path = os.path.join(outdir, mech.replace("-", "_") + ".json")  # file / C identifier; "name" keeps the directory name
        with open(path, "w") as f:
            json.dump(d, f, indent=0, separators=(",", ":"))
        kinds = {}
        for r in d["reactions"]:
            kinds[r["type"]] = kinds.get(r["type"], 0) + 1
        print(mech, len(d["species"]), "species", len(d["reactions"]), "reactions", kinds,
              "->", os.path.relpath(path))


if __name__ == "__main__":
    main()

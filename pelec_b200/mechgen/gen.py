"""Mechanism code generator: ``pelec_b200/mechanisms/<mech>.json`` -> source code.

Two flavours from the same table:

* ``emit(mech, "cuda")``: a ``__device__`` header for the sm_100a kernels
  (``pelec_b200/csrc/gen/mech_<mech>.cuh``) - the product.
* ``emit(mech, "c")``: a plain-C header for the CPU restatement
  (``oracle/gen/mech_<mech>.h``) - test infrastructure.

What is generated (reference function each one restates; LiDryer line numbers of
``PP/Mechanisms/LiDryer/mechanism.H``):

  cv_R / cp_R / gibbs / e_RT / h_RT / dcvpRdT   NASA-7 polynomials          (:199/:265/:331/:532/:615/:781)
  ckcvbs / ckubms / ckcpbs / ckhbms             mixture cv, e, cp, h [CGS]   (:981/:1238/:862/:1100)
  production_rate(wdot, sc, T)                   molar rates, SI             (:2285-2603)
  ajacobian(J, sc, T, consP)                     analytic d(wdot,Tdot)/d(C,T)  (:3905-5138)

``production_rate`` keeps the reference's reaction order and arithmetic association so
that the CPU flavour reproduces the reference bit-for-bit up to libm; the Jacobian is
derived here from the rate law (same mathematics as the reference's ``aJacobian``, not
the same text): it only steers Newton convergence, never the converged answer.

Jacobian storage is behind a macro ``PRX_J(row, col)`` = d f_row / d x_col so the kernel
can choose its own shared-memory layout; default is column-major ``J[col*(N+1)+row]`` as
in the reference (``J[k*neq+m] = d wdot_m / d C_k``).
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
MECH_JSON_DIR = os.path.join(HERE, "..", "mechanisms")


def load(mech):
    with open(os.path.join(MECH_JSON_DIR, mech + ".json")) as f:
        return json.load(f)


def lit(s):
    """A double literal that C parses to the same value as the reference's token."""
    s = str(s)
    if not any(c in s for c in ".eE"):
        s += ".0"
    return s


def signed_sum(terms, var_of):
    """[[coef, key], ...] -> '+c0*v0 - c1*v1 ...' keeping order and association."""
    out = []
    for i, (coef, key) in enumerate(terms):
        neg = coef.startswith("-")
        mag = coef[1:] if neg else coef
        v = var_of(key)
        body = lit(mag) if v is None else "%s * %s" % (lit(mag), v)
        if i == 0:
            out.append(("-" if neg else "") + body)
        else:
            out.append((" - " if neg else " + ") + body)
    return "".join(out)


def poly_expr(terms):
    return signed_sum(terms, lambda b: None if b == "1" else b)


class Emitter:
    def __init__(self, d, flavour):
        self.d = d
        self.fl = flavour
        self.N = len(d["species"])
        self.L = []
        self.exp2 = os.environ.get("PRX_GEN_EXP2", "1") != "0"
        self.q = "PRX_DEVFN " if flavour == "cuda" else "static inline "

    def w(self, s=""):
        self.L.append(s)

    # ------------------------------------------------------------------ thermo
    def thermo_groups(self, fname):
        """species grouped by Tmid, in ascending Tmid, 'no change' species first."""
        th = self.d["thermo"][fname]
        groups = {}
        for k, t in enumerate(th):
            groups.setdefault(t["tmid"], []).append(k)
        keys = sorted(groups, key=lambda x: (-1 if x is None else x))
        return th, [(k, groups[k]) for k in keys]

    def emit_species_fn(self, name, fname, need):
        """void name(double* sp, double T): sp[k] = poly_k(T)."""
        w = self.w
        th, groups = self.thermo_groups(fname)
        w("%svoid %s(double* PRX_RESTRICT sp, const double T)" % (self.q, name))
        w("{")
        self.emit_T_powers(need)
        for tmid, ks in groups:
            if tmid is None:
                for k in ks:
                    w("  sp[%d] = %s;" % (k, poly_expr(th[k]["lo"])))
            else:
                w("  if (T < %s) {" % lit(int(tmid)))
                for k in ks:
                    w("    sp[%d] = %s;" % (k, poly_expr(th[k]["lo"])))
                w("  } else {")
                for k in ks:
                    w("    sp[%d] = %s;" % (k, poly_expr(th[k]["hi"])))
                w("  }")
        w("}")
        w()

    def emit_T_powers(self, need):
        w = self.w
        if "T2" in need:
            w("  const double T2 = T * T;")
        if "T3" in need:
            w("  const double T3 = T * T * T;")
        if "T4" in need:
            w("  const double T4 = T * T * T * T;")
        if "invT" in need:
            w("  const double invT = 1.0 / T;")
        if "logT" in need:
            w("  const double logT = log(T);")

    def emit_mix_fn(self, name, fname, need, scale_RT):
        """double name(T, y): (sum_k y_k * poly_k(T) * imw_k) * (Ru*T | Ru)."""
        w = self.w
        th, groups = self.thermo_groups(fname)
        imw = self.d["imw"]
        w("%sdouble %s(const double T, const double* PRX_RESTRICT y)" % (self.q, name))
        w("{")
        w("  double result = 0.0;")
        self.emit_T_powers(need)
        for tmid, ks in groups:
            if tmid is None:
                for k in ks:
                    w("  result += y[%d] * (%s) * %s;" % (k, poly_expr(th[k]["lo"]), imw[k]))
            else:
                w("  if (T < %s) {" % lit(int(tmid)))
                for k in ks:
                    w("    result += y[%d] * (%s) * %s;" % (k, poly_expr(th[k]["lo"]), imw[k]))
                w("  } else {")
                for k in ks:
                    w("    result += y[%d] * (%s) * %s;" % (k, poly_expr(th[k]["hi"]), imw[k]))
                w("  }")
        if scale_RT:
            w("  const double RT = %s * T;" % self.d["Ru"])
            w("  return result * RT;")
        else:
            w("  return result * %s;" % self.d["Ru"])
        w("}")
        w()

    # ------------------------------------------------------------------ kinetics helpers
    @staticmethod
    def arrh(A, beta, Ea, style="ref"):
        """A * exp((beta) * logT - (Ea) * invT) with the reference's variants."""
        if beta is None and Ea is None:
            return lit(A)
        if beta is not None and Ea is not None:
            return "%s * exp((%s) * logT - (%s) * invT)" % (lit(A), beta, Ea)
        if beta is not None:
            return "%s * exp((%s) * logT)" % (lit(A), beta)
        return "%s * exp(-(%s) * invT)" % (lit(A), Ea)

    @staticmethod
    def arrh_parts(A, beta, Ea):
        """(A literal, argument of the exponential or None): the pieces of arrh()."""
        if beta is None and Ea is None:
            return lit(A), None
        if beta is not None and Ea is not None:
            return lit(A), "(%s) * logT - (%s) * invT" % (beta, Ea)
        if beta is not None:
            return lit(A), "(%s) * logT" % beta
        return lit(A), "-(%s) * invT" % Ea

    @staticmethod
    def conc_product(idx, var="sc"):
        return " * ".join("%s[%d]" % (var, i) for i in idx)

    @staticmethod
    def corr_expr(eff):
        s = "mixture"
        for i, c in eff:
            if c == "1":
                s += " + sc[%d]" % i
            else:
                s += " + (%s) * sc[%d]" % (lit(c), i)
        return s

    @staticmethod
    def low_expr(low):
        A0, b0, E0 = low["A"], low["beta"], low["Ea"]
        if b0 is not None and E0 is not None:
            return "%s * exp(%s * logT - %s * invT)" % (lit(A0), b0, E0)
        if E0 is not None:
            return "%s * exp(-%s * invT)" % (lit(A0), E0)
        return "%s * exp(%s * logT)" % (lit(A0), b0)

    @staticmethod
    def fcent_expr(fc):
        parts = []
        for a, x in fc["exp_terms"]:
            neg = a.startswith("-")
            mag = a[1:] if neg else a
            t = "%s * exp(-T * %s)" % (lit(mag), lit(x))
            parts.append((neg, t))
        if fc["invT_coef"] is not None:
            parts.append((False, "exp(%s * invT)" % fc["invT_coef"]))
        if fc["const"] is not None:
            parts.append((False, lit(fc["const"])))
        s = ""
        for i, (neg, t) in enumerate(parts):
            if i == 0:
                s = ("-" if neg else "") + t
            else:
                s += (" - " if neg else " + ") + t
        return s

    @staticmethod
    def dg_expr(dG, var="g_RT"):
        return signed_sum([[c, k] for c, k in dG],
                          lambda k: None)  # placeholder, replaced below

    def dg_sum(self, dG, var):
        out = ""
        for i, (c, k) in enumerate(dG):
            neg = c.startswith("-")
            mag = c[1:] if neg else c
            t = "%s[%d]" % (var, k) if mag == "1" else "%s * %s[%d]" % (lit(mag), var, k)
            if i == 0:
                out = ("-" if neg else "") + t
            else:
                out += (" - " if neg else " + ") + t
        return out

    @staticmethod
    def refc_factor(p):
        return {0: None, 1: "refC", -1: "refCinv", 2: "(refC * refC)", -2: "(refCinv * refCinv)"}[p]

    # ------------------------------------------------------------------ production rate
    def emit_production_rate(self):
        w = self.w
        d = self.d
        N = self.N
        w("// molar production rates [mol/m^3/s] from concentrations sc [mol/m^3] and T [K]")
        w("%svoid production_rate(double* PRX_RESTRICT wdot, const double* PRX_RESTRICT sc, const double T)" % self.q)
        w("{")
        w("  const double invT = 1.0 / T;")
        w("  const double logT = log(T);")
        w("  const double refC = %s / %s * invT;" % (lit(d["refC_num"]), d["refC_den"]))
        w("  const double refCinv = 1 / refC;")
        w("  for (int i = 0; i < %d; ++i) wdot[i] = 0.0;" % N)
        w("  double mixture = 0.0;")
        w("  for (int i = 0; i < %d; ++i) mixture += sc[i];" % N)
        w("  double g_RT[%d];" % N)
        w("  gibbs(g_RT, T);")
        w("  (void)refC; (void)refCinv; (void)mixture;")
        for r in d["reactions"]:
            w("  { // R%d: %s" % (r["idx"], r["eq"]))
            # device flavour: the two independent exponentials of a reversible reaction (rate constant, equilibrium constant)
            # go through ONE call that evaluates both chains interleaved (PRX_EXP2, prx_chem.cuh): same arguments, same
            # per-value arithmetic, half the dependent-latency exposure (ncu r02c: a single out-of-line exp is a chain of
            # 13 dependent DFMAs and the warp waits for its latency, not for issue slots)
            A_l, k_arg = self.arrh_parts(r["A"], r["beta"], r["Ea"])
            paired = (self.fl == "cuda" and self.exp2 and k_arg is not None and r["reversible"] and r["type"] != "falloff")
            if paired:
                w("    const prx_d2 ee = PRX_EXP2(%s, -(%s));" % (k_arg, self.dg_sum(r["dG"], "g_RT")))
                w("    const double k_f = %s * ee.x;" % A_l)
            else:
                w("    const double k_f = %s;" % self.arrh(r["A"], r["beta"], r["Ea"]))
            pre = ""
            if r["type"] in ("three_body", "falloff"):
                w("    double Corr = %s;" % self.corr_expr(r["eff"]))
                pre = "Corr * "
            if r["type"] == "falloff":
                w("    const double redP = Corr / k_f * %s;" % self.low_expr(r["low"]))
                if r["falloff"] == "troe":
                    w("    const double F = redP / (1.0 + redP);")
                    w("    const double logPred = log10(redP);")
                    w("    const double logFcent = log10(%s);" % self.fcent_expr(r["fcent"]))
                    w("    const double troe_c = -0.4 - 0.67 * logFcent;")
                    w("    const double troe_n = 0.75 - 1.27 * logFcent;")
                    w("    const double troe = (troe_c + logPred) / (troe_n - 0.14 * (troe_c + logPred));")
                    w("    const double F_troe = exp(PRX_LN10 * logFcent / (1.0 + troe * troe));")
                    w("    Corr = F * F_troe;")
                else:
                    w("    Corr = redP / (1.0 + redP);")
            w("    const double qf = %sk_f * (%s);" % (pre, self.conc_product(r["reactants"])))
            if r["reversible"]:
                rc = self.refc_factor(r["refC"])
                rcs = "" if rc is None else ("%s * " % (rc if rc.startswith("(") else "(" + rc + ")"))
                kc = "ee.y" if paired else "exp(-(%s))" % self.dg_sum(r["dG"], "g_RT")
                w("    const double qr = %sk_f * %s * %s(%s);" % (pre, kc, rcs, self.conc_product(r["products"])))
            else:
                w("    const double qr = 0.0;")
            w("    const double qdot = qf - qr;")
            for k, c in r["wdot"]:
                neg = c.startswith("-")
                mag = c[1:] if neg else c
                if mag == "1":
                    w("    wdot[%d] %s qdot;" % (k, "-=" if neg else "+="))
                else:
                    w("    wdot[%d] %s %s * qdot;" % (k, "-=" if neg else "+=", lit(mag)))
            w("  }")
        w("}")
        w()

    # ------------------------------------------------------------------ analytic Jacobian
    def emit_ajacobian(self):
        """d(wdot_m)/d(C_k), d(wdot_m)/dT, d(Tdot)/d(C_k), d(Tdot)/dT; SI concentrations.

        Derivation (per reaction, q = Corr * (kf*Pf - kr*Pr), kr = kf*Kinv):
          dln(kf)/dT   = beta/T + Ea/T^2
          dln(Kinv)/dT = (sum_k c_k h_RT_k - p)/T        (Kinv = exp(-sum c_k g_k) refC^p)
          three-body   : Corr = alpha = sum_k (1+eff_k) C_k
          falloff      : Pr = alpha k0/kf ; Corr = Pr/(1+Pr) * F_troe(T, Pr)
        """
        w = self.w
        d = self.d
        N = self.N
        neq = N + 1
        w("// analytic Jacobian, SI concentrations; PRX_J(row, col) = d f_row / d x_col,")
        w("// rows/cols 0..N-1 species (wdot_m, C_k), N = temperature (Tdot, T).")
        w("%svoid ajacobian(double* PRX_RESTRICT J, const double* PRX_RESTRICT sc, const double T, const int consP)" % self.q)
        w("{")
        w("  for (int i = 0; i < %d; ++i) for (int j = 0; j < %d; ++j) PRX_J(i, j) = 0.0;" % (neq, neq))
        w("  double wdot[%d];" % N)
        w("  for (int i = 0; i < %d; ++i) wdot[i] = 0.0;" % N)
        w("  const double invT = 1.0 / T;")
        w("  const double invT2 = invT * invT;")
        w("  const double logT = log(T);")
        w("  const double refC = %s / %s / T;" % (lit(d["refC_num"]), d["refC_den"]))
        w("  const double refCinv = 1.0 / refC;")
        w("  double mixture = 0.0;")
        w("  for (int k = 0; k < %d; ++k) mixture += sc[k];" % N)
        w("  double g_RT[%d], h_RT[%d];" % (N, N))
        w("  gibbs(g_RT, T);")
        w("  h_RT_fn(h_RT, T);")
        w("  const double ln10 = PRX_LN10;")
        w("  const double log10e = 1.0 / PRX_LN10;")
        w("  (void)refC; (void)refCinv; (void)mixture; (void)ln10; (void)log10e; (void)invT2; (void)logT;")
        w("  double dqdc[%d];" % N)
        w("  for (int k = 0; k < %d; ++k) dqdc[k] = 0.0;" % N)
        for r in d["reactions"]:
            w("  { // R%d: %s" % (r["idx"], r["eq"]))
            # forward
            w("    const double phi_f = %s;" % self.conc_product(r["reactants"]))
            w("    double k_f = %s;" % self.arrh(r["A"], r["beta"], r["Ea"]))
            dl = []
            if r["beta"] is not None:
                dl.append("(%s) * invT" % r["beta"])
            if r["Ea"] is not None:
                dl.append("(%s) * invT2" % r["Ea"])
            w("    const double dlnkfdT = %s;" % (" + ".join(dl) if dl else "0.0"))
            # reverse
            if r["reversible"]:
                rc = self.refc_factor(r["refC"])
                w("    const double phi_r = %s;" % self.conc_product(r["products"]))
                w("    const double Kinv = %sexp(-(%s));" % ((rc + " * ") if rc else "", self.dg_sum(r["dG"], "g_RT")))
                w("    double k_r = k_f * Kinv;")
                w("    const double dlnKinvdT = invT * ((%s) - %s);" % (self.dg_sum(r["dG"], "h_RT"), lit(r["refC"])))
                w("    const double dkrdT = (dlnkfdT + dlnKinvdT) * k_r;")
            else:
                w("    const double phi_r = 0.0; double k_r = 0.0; const double dkrdT = 0.0;")
            w("    const double q_nocor = k_f * phi_f - k_r * phi_r;")
            typ = r["type"]
            if typ == "elementary":
                w("    const double q = q_nocor;")
                w("    const double dqdT = dlnkfdT * k_f * phi_f - dkrdT * phi_r;")
            elif typ == "three_body":
                w("    const double alpha = %s;" % self.corr_expr(r["eff"]))
                w("    const double q = alpha * q_nocor;")
                w("    const double dqdT = alpha * (dlnkfdT * k_f * phi_f - dkrdT * phi_r);")
                w("    const double dcdc_fac = q_nocor;")
                w("    k_f *= alpha; k_r *= alpha;")
            else:
                low = r["low"]
                w("    const double alpha = %s;" % self.corr_expr(r["eff"]))
                w("    const double k_0 = %s;" % self.low_expr(low))
                w("    const double Pr = alpha / k_f * k_0;")
                w("    const double fPr = Pr / (1.0 + Pr);")
                dl0 = []
                if low["beta"] is not None:
                    dl0.append("(%s) * invT" % low["beta"])
                if low["Ea"] is not None:
                    dl0.append("(%s) * invT2" % low["Ea"])
                w("    const double dlnk0dT = %s;" % (" + ".join(dl0) if dl0 else "0.0"))
                w("    const double dlogPrdT = log10e * (dlnk0dT - dlnkfdT);")
                w("    const double dlogfPrdT = dlogPrdT / (1.0 + Pr);")
                if r["falloff"] == "troe":
                    fc = r["fcent"]
                    w("    const double logPr = log10(Pr);")
                    terms = []
                    dterms = []
                    for i, (a, x) in enumerate(fc["exp_terms"]):
                        w("    const double Fc%d = (%s) * exp(-T * %s);" % (i, lit(a), lit(x)))
                        terms.append("Fc%d" % i)
                        dterms.append("-Fc%d * %s" % (i, lit(x)))
                    if fc["invT_coef"] is not None:
                        w("    const double FcT = exp(%s * invT);" % fc["invT_coef"])
                        terms.append("FcT")
                        dterms.append("-FcT * (%s) * invT2" % fc["invT_coef"])
                    if fc["const"] is not None:
                        terms.append(lit(fc["const"]))
                    w("    const double Fcent = %s;" % " + ".join(terms))
                    w("    const double logFcent = log10(Fcent);")
                    w("    const double troe_c = -0.4 - 0.67 * logFcent;")
                    w("    const double troe_n = 0.75 - 1.27 * logFcent;")
                    w("    const double troePr_den = 1.0 / (troe_n - 0.14 * (troe_c + logPr));")
                    w("    const double troePr = (troe_c + logPr) * troePr_den;")
                    w("    const double troe = 1.0 / (1.0 + troePr * troePr);")
                    w("    const double F = exp(ln10 * logFcent * troe);")
                    w("    const double dlogFcentdT = log10e / Fcent * (%s);" % " + ".join("(%s)" % t for t in dterms))
                    w("    const double dlogFdcn_fac = 2.0 * logFcent * troe * troe * troePr * troePr_den;")
                    w("    const double dlogFdc = -troe_n * dlogFdcn_fac * troePr_den;")
                    w("    const double dlogFdn = dlogFdcn_fac * troePr;")
                    w("    const double dlogFdlogPr = dlogFdc;")
                    w("    const double dlogFdT = dlogFcentdT * (troe - 0.67 * dlogFdc - 1.27 * dlogFdn) + dlogFdlogPr * dlogPrdT;")
                else:
                    w("    const double F = 1.0; const double dlogFdlogPr = 0.0; const double dlogFdT = 0.0;")
                w("    const double Corr = fPr * F;")
                w("    const double q = Corr * q_nocor;")
                w("    const double dlnCorrdT = ln10 * (dlogfPrdT + dlogFdT);")
                w("    const double dqdT = Corr * (dlnkfdT * k_f * phi_f - dkrdT * phi_r) + dlnCorrdT * q;")
                w("    const double dcdc_fac = q / alpha * (1.0 / (Pr + 1.0) + dlogFdlogPr);")
                w("    k_f *= Corr; k_r *= Corr;")
            # stoichiometry
            nu = {}
            for k, c in r["wdot"]:
                nu[k] = nu.get(k, 0.0) + float(c)
            for k, c in sorted(nu.items()):
                w("    wdot[%d] += (%s) * q;" % (k, lit(repr(c))))
            # dq/dC_k, direct part
            direct = {}

            def powers(lst):
                p = {}
                for i in lst:
                    p[i] = p.get(i, 0) + 1
                return p

            pf = powers(r["reactants"])
            pr = powers(r["products"]) if r["reversible"] else {}
            for k, n in pf.items():
                others = []
                for i, m in pf.items():
                    others += ["sc[%d]" % i] * (m - 1 if i == k else m)
                e = "%sk_f%s" % ("%d.0 * " % n if n > 1 else "", "".join(" * " + o for o in others))
                direct.setdefault(k, []).append("+ " + e)
            for k, n in pr.items():
                others = []
                for i, m in pr.items():
                    others += ["sc[%d]" % i] * (m - 1 if i == k else m)
                e = "%sk_r%s" % ("%d.0 * " % n if n > 1 else "", "".join(" * " + o for o in others))
                direct.setdefault(k, []).append("- " + e)
            rows = sorted(nu.items())
            if typ == "elementary":
                for k in sorted(direct):
                    w("    { const double dqdc = %s;" % " ".join(direct[k]).lstrip("+ "))
                    for m, c in rows:
                        w("      PRX_J(%d, %d) += (%s) * dqdc;" % (m, k, lit(repr(c))))
                    w("    }")
            else:
                eff = {}
                for i, c in r["eff"]:
                    eff[i] = eff.get(i, 0.0) + float(c)
                w("    if (consP) {")
                ks = sorted(set(direct) | set(eff))
                for k in ks:
                    parts = []
                    if k in eff:
                        parts.append("(%s) * dcdc_fac" % lit(repr(eff[k])))
                    if k in direct:
                        parts.append(" ".join(direct[k]))
                    expr = " ".join(parts).lstrip("+ ")
                    w("      { const double dqdc = %s;" % expr)
                    for m, c in rows:
                        w("        PRX_J(%d, %d) += (%s) * dqdc;" % (m, k, lit(repr(c))))
                    w("      }")
                w("    } else {")
                # NOTE (reference quirk, reproduced on purpose): CEPTR omits the assignment of
                # dqdc[k] when species k has third-body efficiency 0 and no direct term, so the
                # reference's loop below adds the STALE dqdc[k] left by an earlier reaction
                # (e.g. drm19 mechanism.H aJacobian "reaction 80: HCO + M": dqdc[5] missing).
                # dqdc lives at function scope for that reason.
                for k in range(N):
                    tot = 1.0 + eff.get(k, 0.0)
                    parts = []
                    if tot != 0.0:
                        parts.append("dcdc_fac" if tot == 1.0 else "(%s) * dcdc_fac" % lit(repr(tot)))
                    if k in direct:
                        parts.append(" ".join(direct[k]))
                    if not parts:
                        continue
                    w("      dqdc[%d] = %s;" % (k, " ".join(parts).lstrip("+ ")))
                w("      for (int k = 0; k < %d; ++k) {" % N)
                for m, c in rows:
                    w("        PRX_J(%d, k) += (%s) * dqdc[k];" % (m, lit(repr(c))))
                w("      }")
                w("    }")
            for m, c in rows:
                w("    PRX_J(%d, %d) += (%s) * dqdT;" % (m, N, lit(repr(c))))
            w("  }")
        # temperature row
        w("  double c_R[%d], dcRdT[%d], eh_RT[%d];" % (N, N, N))
        w("  if (consP) { cp_R(c_R, T); for (int k = 0; k < %d; ++k) eh_RT[k] = h_RT[k]; }" % N)
        w("  else { cv_R(c_R, T); e_RT_fn(eh_RT, T); }")
        w("  dcvpRdT(dcRdT, T);")
        w("  double cmix = 0.0, ehmix = 0.0, dcmixdT = 0.0, dehmixdT = 0.0;")
        w("  for (int k = 0; k < %d; ++k) {" % N)
        w("    cmix += c_R[k] * sc[k];")
        w("    dcmixdT += dcRdT[k] * sc[k];")
        w("    ehmix += eh_RT[k] * wdot[k];")
        w("    dehmixdT += invT * (c_R[k] - eh_RT[k]) * wdot[k] + eh_RT[k] * PRX_J(k, %d);" % N)
        w("  }")
        w("  const double cmixinv = 1.0 / cmix;")
        w("  const double tmp1 = ehmix * cmixinv;")
        w("  const double tmp3 = cmixinv * T;")
        w("  const double tmp2 = tmp1 * tmp3;")
        w("  for (int k = 0; k < %d; ++k) {" % N)
        w("    double dehmixdc = 0.0;")
        w("    for (int m = 0; m < %d; ++m) dehmixdc += eh_RT[m] * PRX_J(m, k);" % N)
        w("    PRX_J(%d, k) = tmp2 * c_R[k] - tmp3 * dehmixdc;" % N)
        w("  }")
        w("  PRX_J(%d, %d) = -tmp1 + tmp2 * dcmixdT - tmp3 * dehmixdT;" % (N, N))
        w("}")
        w()


    # ------------------------------------------------------------------ tables
    def emit_tables(self):
        """Mechanism as DATA (rolled, table-driven evaluation: prx_tchem.cuh).  Same literals and the same
        evaluation order as the straight-line code above; each array is emitted in two address spaces -
        __constant__ (uniform index: every lane of a warp evaluates the same reaction, thread-per-cell
        kernels) and __device__ global (per-lane index: one warp per cell, lanes over reactions/species) -
        behind the accessor structs TabC / TabG."""
        w = self.w
        d = self.d
        N = self.N
        R = d["reactions"]
        NR = len(R)
        th = d["thermo"]
        arrays = []  # (ctype, name, [values as strings])

        def coefs(fname, bases):
            out = []
            for k in range(N):
                t = th[fname][k]
                for rng in ("lo", "hi"):
                    terms = t[rng] if t[rng] is not None else t["lo"]
                    m = {b: c for c, b in terms}
                    out += [lit(m.get(b, "0.0")) for b in bases]
            return out

        tm = [t["tmid"] for t in th["cv_R"]]
        arrays.append(("double", "tmid", [lit(repr(float(x))) if x is not None else "1.0e300" for x in tm]))
        arrays.append(("double", "cv", coefs("cv_R", ["1", "T", "T2", "T3", "T4"])))
        arrays.append(("double", "cp", coefs("cp_R", ["1", "T", "T2", "T3", "T4"])))
        arrays.append(("double", "gb", coefs("gibbs", ["invT", "1", "logT", "T", "T2", "T3", "T4"])))
        arrays.append(("double", "ei", coefs("speciesInternalEnergy", ["1", "T", "T2", "T3", "T4", "invT"])))
        arrays.append(("double", "hi", coefs("speciesEnthalpy", ["1", "T", "T2", "T3", "T4", "invT"])))
        arrays.append(("double", "dc", coefs("dcvpRdT", ["1", "T", "T2", "T3"])))
        arrays.append(("double", "mw", list(d["mw"])))
        arrays.append(("double", "imw", list(d["imw"])))
        # species accumulation order of the mixture sums (ckubms ...): grouped by Tmid like the reference
        _, groups = self.thermo_groups("cv_R")
        order = [k for _, ks in groups for k in ks]
        arrays.append(("unsigned char", "sp_order", [str(k) for k in order]))

        # reactions
        TYPE = {"elementary": 0, "three_body": 1}
        meta, fo_of, tb_of = [], [], []
        A, beta, Ea = [], [], []
        reac, prod, gidx, gcoef, widx, wcoef = [], [], [], [], [], []
        eff_ptr, eff_idx, eff_val = [0], [], []
        lowA, lowb, lowE, fc_a, fc_x, fc_invT, fc_const, fo_flags = [], [], [], [], [], [], [], []
        for r in R:
            typ = TYPE.get(r["type"], 2 if r.get("falloff") == "troe" else 3)
            noexp = r["beta"] is None and r["Ea"] is None
            assert len(r["reactants"]) <= 3 and len(r["products"]) <= 3 and len(r["wdot"]) <= 5
            dG = r["dG"] if r["reversible"] else []
            assert len(dG) <= 6 and -2 <= r["refC"] <= 2
            m = (typ | (int(bool(r["reversible"])) << 2) | (int(noexp) << 3) | (len(r["reactants"]) << 4)
                 | (len(r["products"]) << 6) | (len(dG) << 8) | (len(r["wdot"]) << 11) | ((r["refC"] + 2) << 14))
            meta.append(m)
            A.append(lit(r["A"]))
            beta.append(lit(r["beta"]) if r["beta"] is not None else "0.0")
            Ea.append(lit(r["Ea"]) if r["Ea"] is not None else "0.0")
            reac += [str(i) for i in r["reactants"]] + ["0"] * (3 - len(r["reactants"]))
            prod += [str(i) for i in r["products"]] + ["0"] * (3 - len(r["products"]))
            gidx += [str(k) for c, k in dG] + ["0"] * (6 - len(dG))
            gcoef += [lit(c) for c, k in dG] + ["0.0"] * (6 - len(dG))
            widx += [str(k) for k, c in r["wdot"]] + ["0"] * (5 - len(r["wdot"]))
            wcoef += [lit(c) for k, c in r["wdot"]] + ["0.0"] * (5 - len(r["wdot"]))
            if typ >= 1:
                tb_of.append(len(eff_ptr) - 1)
                for i, c in r["eff"]:
                    eff_idx.append(str(i))
                    eff_val.append(lit(c))
                eff_ptr.append(len(eff_idx))
            else:
                tb_of.append(-1)
            if typ >= 2:
                fo_of.append(len(lowA))
                low = r["low"]
                lowA.append(lit(low["A"]))
                lowb.append(lit(low["beta"]) if low["beta"] is not None else "0.0")
                lowE.append(lit(low["Ea"]) if low["Ea"] is not None else "0.0")
                fc = r.get("fcent") or {"exp_terms": [], "invT_coef": None, "const": None}
                assert len(fc["exp_terms"]) <= 2
                et = list(fc["exp_terms"]) + [["0.0", "0.0"]] * (2 - len(fc["exp_terms"]))
                fc_a += [lit(a) for a, x in et]
                fc_x += [lit(x) for a, x in et]
                fc_invT.append(lit(fc["invT_coef"]) if fc["invT_coef"] is not None else "0.0")
                fc_const.append(lit(fc["const"]) if fc["const"] is not None else "0.0")
                fo_flags.append(str(len(fc["exp_terms"]) | (int(fc["invT_coef"] is not None) << 2)
                                    | (int(fc["const"] is not None) << 3)))
            else:
                fo_of.append(-1)
        arrays += [("int", "meta", [str(x) for x in meta]), ("short", "tb_of", [str(x) for x in tb_of]),
                   ("short", "fo_of", [str(x) for x in fo_of]),
                   ("double", "A", A), ("double", "beta", beta), ("double", "Ea", Ea),
                   ("unsigned char", "reac", reac), ("unsigned char", "prod", prod), ("unsigned char", "gidx", gidx),
                   ("float", "gcoef", gcoef), ("unsigned char", "widx", widx), ("float", "wcoef", wcoef),
                   ("short", "eff_ptr", [str(x) for x in eff_ptr]), ("unsigned char", "eff_idx", eff_idx or ["0"]),
                   ("double", "eff_val", eff_val or ["0.0"]),
                   ("double", "lowA", lowA or ["0.0"]), ("double", "lowb", lowb or ["0.0"]), ("double", "lowE", lowE or ["0.0"]),
                   ("double", "fc_a", fc_a or ["0.0"]), ("double", "fc_x", fc_x or ["0.0"]),
                   ("double", "fc_invT", fc_invT or ["0.0"]), ("double", "fc_const", fc_const or ["0.0"]),
                   ("unsigned char", "fo_flags", fo_flags or ["0"])]
        # species -> (reaction, coefficient) lists in reaction order (gather form of the wdot accumulation)
        sp_ptr, sp_r, sp_c = [0], [], []
        for k in range(N):
            for ir, r in enumerate(R):
                for kk, c in r["wdot"]:
                    if kk == k:
                        sp_r.append(str(ir))
                        sp_c.append(lit(c))
            sp_ptr.append(len(sp_r))
        arrays += [("short", "sp_ptr", [str(x) for x in sp_ptr]), ("short", "sp_r", sp_r), ("float", "sp_c", sp_c)]
        # ---- Jacobian tables (prx_tchem.cuh ajacobian_tab / prx_coop.cuh jac_coop; same structure as emit_ajacobian)
        # jnu: merged stoichiometric rows sorted by species; jd: "direct" dq/dC terms per distinct species
        # (forward term n*k_f*others, reverse term n*k_r*others, packed n | o1 << 8 | o2 << 16, 255 = none);
        # tb_tot / tb_eff / tb_has: dense third-body efficiency rows (1 + eff, eff, eff present)
        jmeta, jnu_idx, jnu_c, jd_k, jd_f, jd_r = [], [], [], [], [], []
        tb_tot, tb_eff, tb_has = [], [], []

        def powers(lst):
            pw = {}
            for i in lst:
                pw[i] = pw.get(i, 0) + 1
            return pw

        def pack(pw, k):
            others = []
            for i, m in pw.items():
                others += [i] * (m - 1 if i == k else m)
            assert len(others) <= 2
            others += [255] * (2 - len(others))
            return pw[k] | (others[0] << 8) | (others[1] << 16)

        for r in R:
            nu = {}
            for k, c in r["wdot"]:
                nu[k] = nu.get(k, 0.0) + float(c)
            rows = sorted(nu.items())
            pf = powers(r["reactants"])
            pr = powers(r["products"]) if r["reversible"] else {}
            ks = sorted(set(pf) | set(pr))
            assert len(rows) <= 5 and len(ks) <= 6
            jmeta.append(len(rows) | (len(ks) << 4))
            jnu_idx += [str(k) for k, c in rows] + ["0"] * (5 - len(rows))
            jnu_c += [repr(c) for k, c in rows] + ["0.0"] * (5 - len(rows))
            jd_k += [str(k) for k in ks] + ["0"] * (6 - len(ks))
            jd_f += [str(pack(pf, k)) if k in pf else "0" for k in ks] + ["0"] * (6 - len(ks))
            jd_r += [str(pack(pr, k)) if k in pr else "0" for k in ks] + ["0"] * (6 - len(ks))
            if r["type"] != "elementary":
                eff = {}
                for i, c in r["eff"]:
                    eff[i] = eff.get(i, 0.0) + float(c)
                tb_tot += [repr(1.0 + eff.get(k, 0.0)) for k in range(N)]
                tb_eff += [repr(eff.get(k, 0.0)) for k in range(N)]
                tb_has += ["1" if k in eff else "0" for k in range(N)]
        for c in jnu_c:
            import struct
            assert struct.unpack("f", struct.pack("f", float(c)))[0] == float(c), c
        n_common = len(arrays)  # everything after this index is emitted in global space only (64 KB constant bank)
        arrays += [("short", "jmeta", [str(x) for x in jmeta]), ("unsigned char", "jnu_idx", jnu_idx), ("float", "jnu_c", jnu_c),
                   ("unsigned char", "jd_k", jd_k), ("int", "jd_f", jd_f), ("int", "jd_r", jd_r),
                   ("double", "tb_tot", tb_tot or ["0.0"]), ("double", "tb_eff", tb_eff or ["0.0"]),
                   ("unsigned char", "tb_has", tb_has or ["0"])]
        for c in gcoef + wcoef + sp_c:  # stoichiometric coefficients must be exact in float
            import struct
            assert struct.unpack("f", struct.pack("f", float(c)))[0] == float(c), c

        w("// ---- mechanism tables (see Emitter.emit_tables) ----")
        w("#define PRX_NUM_FALLOFF %d" % len(lowA))
        w("#define PRX_REFC_NUM %s" % lit(d["refC_num"]))
        w("#define PRX_REFC_DEN %s" % d["refC_den"])
        if self.fl == "cuda":
            w("#ifdef PRX_HOST_TEST")
            w("#define PRX_TABC static const")
            w("#define PRX_TABG static const")
            w("#define PRX_TABFN static inline")
            w("#else")
            w("#define PRX_TABC static __constant__ const")
            w("#define PRX_TABG static __device__ const")
            w("#define PRX_TABFN static __device__ __forceinline__")
            w("#endif")
        else:
            w("#define PRX_TABC static const")
            w("#define PRX_TABG static const")
        for space, pre in (("PRX_TABC", "tc_"), ("PRX_TABG", "tg_")):
            for ct, nm, vals in (arrays[:n_common] if pre == "tc_" else arrays):
                w("%s %s %s%s[%d] = {%s};" % (space, ct, pre, nm, len(vals), ", ".join(vals)))
        if self.fl == "cuda":
            for sname, pre in (("TabC", "tc_"), ("TabG", "tg_")):
                w("struct %s {" % sname)
                for ct, nm, vals in (arrays[:n_common] if pre == "tc_" else arrays):
                    w("  PRX_TABFN %s %s(const int i) { return %s%s[i]; }" % (ct, nm, pre, nm))
                w("};")
        w()

    # ------------------------------------------------------------------ Simple-transport fit tables
    def emit_transport(self):
        """The data TransParm<Fuego, SimpleTransport> holds (PP/Source/Transport/TransportParams.H:237-256): wt, iwt = 1 / wt,
        eps, zrot, nlin, fitmu = COFETA, fitlam = COFLAM, fitdbin = COFD, as global-memory tables (fitdbin of dodecane_lu is
        90 KB: no constant memory); the kernels index them with warp-uniform indices."""
        t = self.d.get("transport")
        w = self.w
        if t is None:
            return
        N = self.N
        assert t["NO"] == 4 and len(t["cofd"]) == 4 * N * N
        cofd = [float(x) for x in t["cofd"]]
        sym = all(cofd[4 * (i + N * j) + c] == cofd[4 * (j + N * i) + c] for i in range(N) for j in range(N) for c in range(4))
        w("// ---- Simple transport (see Emitter.emit_transport) ----")
        w("#define PRX_HAVE_TRANSPORT 1")
        w("#define PRX_TR_DBIN_SYMMETRIC %d  // fitdbin(i,j) == fitdbin(j,i) bit for bit: one exponential per species pair" % (1 if sym else 0))
        w("#define PRX_TR_PATM %s" % lit(t["patm"]))
        for nm, ct, vals in (("wt", "double", [lit(x) for x in t["wt"]]),
                             ("iwt", "double", [repr(1.0 / float(x)) for x in t["wt"]]),
                             # RU / mw: the per-species literals of CKCVMS / CKCPMS (mechanism.H "multiply by R/molecularweight")
                             ("ru_mw", "double", [lit(x) for x in t["ru_mw"]]),
                             ("eps", "double", [lit(x) for x in t["eps"]]),
                             ("zrot", "double", [lit(x) for x in t["zrot"]]),
                             ("nlin", "int", [str(x) for x in t["nlin"]]),
                             ("fitmu", "double", [lit(x) for x in t["cofeta"]]),
                             ("fitlam", "double", [lit(x) for x in t["coflam"]]),
                             ("fitdbin", "double", [lit(x) for x in t["cofd"]])):
            w("PRX_TABG %s tr_%s[%d] = {%s};" % (ct, nm, len(vals), ", ".join(vals)))
        w()

    # ------------------------------------------------------------------ whole file
    def emit(self):
        d = self.d
        N = self.N
        w = self.w
        guard = "PRX_MECH_%s_GEN" % d["name"].upper().replace("-", "_")
        w("// GENERATED by pelec_b200/mechgen/gen.py from pelec_b200/mechanisms/%s.json - do not edit." % d["name"])
        w("// flavour: %s" % self.fl)
        w("#ifndef %s" % guard)
        w("#define %s" % guard)
        if self.fl == "c":
            w("#include <math.h>")
        else:
            w("#ifndef PRX_DEVFN")
            w("#define PRX_DEVFN __device__ __forceinline__")
            w("#define PRX_DEVCONST __device__ static const")
            w("#endif")
        w("#ifndef PRX_RESTRICT")
        w("#define PRX_RESTRICT %s" % ("__restrict__"))
        w("#endif")
        w("#define PRX_LN10 2.30258509299404568402")
        w("#define PRX_MECH_NAME \"%s\"" % d["name"])
        w("#define PRX_NUM_SPECIES %d" % N)
        w("#define PRX_NUM_REACTIONS %d" % len(d["reactions"]))
        w("#define PRX_RU %s" % d["Ru"])
        w("#define PRX_TMIN %s" % lit(d["T_newton"]["tmin"]))
        w("#define PRX_TMAX %s" % lit(d["T_newton"]["tmax"]))
        w("#ifndef PRX_J")
        w("#define PRX_J(row, col) J[(col) * %d + (row)]" % (N + 1))
        w("#endif")
        cq = "PRX_DEVCONST" if self.fl == "cuda" else "static const"
        w("%s double prx_mw[%d] = {%s};" % (cq, N, ", ".join(d["mw"])))
        w("%s double prx_imw[%d] = {%s};" % (cq, N, ", ".join(d["imw"])))
        if True:  # host-side table in both flavours (prx_species_name / orc_species_name)
            w("static const char* const prx_species_names[%d] = {%s};" % (N, ", ".join('"%s"' % s for s in d["species"])))
        w()
        if self.fl == "cuda":
            # optional: transcendentals of the straight-line code through out-of-line functions (prx_chem.cuh),
            # which cuts ~40 % of the RHS instruction footprint (the inlined libm expansions)
            w("#ifdef PRX_MATH_REDIRECT")
            w("#define exp prx_exp")
            w("#define log prx_log")
            w("#define log10 prx_log10")
            w("#endif")
        self.emit_species_fn("cv_R", "cv_R", {"T2", "T3", "T4"})
        self.emit_species_fn("cp_R", "cp_R", {"T2", "T3", "T4"})
        self.emit_species_fn("gibbs", "gibbs", {"T2", "T3", "T4", "invT", "logT"})
        self.emit_species_fn("e_RT_fn", "speciesInternalEnergy", {"T2", "T3", "T4", "invT"})
        self.emit_species_fn("h_RT_fn", "speciesEnthalpy", {"T2", "T3", "T4", "invT"})
        self.emit_species_fn("dcvpRdT", "dcvpRdT", {"T2", "T3"})
        self.emit_mix_fn("ckcvbs", "cv_R", {"T2", "T3", "T4"}, False)
        self.emit_mix_fn("ckcpbs", "cp_R", {"T2", "T3", "T4"}, False)
        self.emit_mix_fn("ckubms", "speciesInternalEnergy", {"T2", "T3", "T4", "invT"}, True)
        self.emit_mix_fn("ckhbms", "speciesEnthalpy", {"T2", "T3", "T4", "invT"}, True)
        self.emit_production_rate()
        self.emit_ajacobian()
        if self.fl == "cuda":
            w("#ifdef PRX_MATH_REDIRECT")
            w("#undef exp")
            w("#undef log")
            w("#undef log10")
            w("#endif")
        self.emit_tables()
        self.emit_transport()
        w("#endif")
        return "\n".join(self.L) + "\n"


def emit(mech, flavour):
    return Emitter(load(mech), flavour).emit()


# ---------------------------------------------------------------------------------------------
# Literal pool of the straight-line device code
# ---------------------------------------------------------------------------------------------
# ncu on the round-1 kernels (profiles/r02a_prof_rk64_lidryer_64_opcodes.txt): 50 % of the executed instructions of the
# straight-line chemistry were UMOV / MOV / IMAD.MOV pairs materialising 64-bit literals (sm_100a FP64 instructions take
# a register, a uniform register or a 32-bit immediate = the HIGH word of a double, never a full 64-bit immediate), and the
# RK64 kernel was issue-bound at 60 % with the FP64 pipe 40 % busy.  constify() moves every literal whose low word is not
# zero into one __constant__ table, in order of first use: neighbouring literals of an expression become neighbouring table
# entries, which ptxas fetches with ONE LDCU.128 per pair (uniform registers) instead of four UMOVs.
_LIT_RE = None


def _cheap_literal(tok):
    import struct
    return (struct.unpack("<Q", struct.pack("<d", float(tok)))[0] & 0xffffffff) == 0


def constify(text, max_entries=6000):
    """Rewrite the straight-line functions of a cuda-flavour header so that expensive literals come from prx_k[]."""
    import re
    global _LIT_RE
    if _LIT_RE is None:
        _LIT_RE = re.compile(r"(?<![\w.\]\)])(\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
    begin = text.index("PRX_DEVFN void cv_R(")
    end = text.index("#ifdef PRX_MATH_REDIRECT\n#undef exp")
    head, body, tail = text[:begin], text[begin:end], text[end:]
    pool, order = {}, []

    def sub(m):
        tok = m.group(1)
        if _cheap_literal(tok):
            return tok
        if tok not in pool:
            if len(order) >= max_entries:
                return tok
            pool[tok] = len(order)
            order.append(tok)
        return "PRX_K(%d)" % pool[tok]

    body = body.replace("PRX_LN10", "2.30258509299404568402")
    # quotients of two literals were folded by the compiler (IEEE division, same bits as Python's): keep them folded
    num = r"(\d+\.\d*(?:[eE][-+]?\d+)?)"
    body = re.sub(r"= %s / %s (?=[*/])" % (num, num), lambda m: "= %r " % (float(m.group(1)) / float(m.group(2))), body)
    body = _LIT_RE.sub(sub, body)
    if len(order) % 2:
        order.append("0.0")
    tab = ["// literal pool of the straight-line code (see gen.py constify)",
           "#ifdef PRX_HOST_TEST",
           "static const double prx_k[%d] = {%s};" % (len(order), ", ".join(order)),
           "#else",
           "static __constant__ __align__(16) double prx_k[%d] = {%s};" % (len(order), ", ".join(order)),
           "#endif",
           "#define PRX_K(i) prx_k[i]",
           ""]
    return head + "\n".join(tab) + "\n" + body + tail


def write_if_changed(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path) and open(path).read() == text:
        return False
    with open(path, "w") as f:
        f.write(text)
    return True


def shipped_mechanisms():
    """the three BASELINE mechanisms first (ABI ids 0-2), then every other pelec_b200/mechanisms/<mech>.json"""
    core = ["LiDryer", "drm19", "dodecane_lu"]
    extra = sorted(f[:-5] for f in os.listdir(MECH_JSON_DIR) if f.endswith(".json") and f[:-5] not in core)
    return tuple(core + extra)


def generate_all(repo_root, mechs=None):
    mechs = mechs or shipped_mechanisms()
    out = []
    for m in mechs:
        p = os.path.join(repo_root, "oracle", "gen", "mech_%s.h" % m)
        write_if_changed(p, emit(m, "c"))
        out.append(p)
        p = os.path.join(repo_root, "pelec_b200", "csrc", "gen", "mech_%s.cuh" % m)
        text = emit(m, "cuda")
        # Literal pool: OFF by default.  Measured on B200 (profiles/README.md, session r02d): 21 % fewer executed instructions
        # but 12 % SLOWER BDF (75.1 vs 66.7 ms per 128^3: the LDC'd literals live in registers, 460 B more spills per thread)
        # and 2 % slower RK64 - the chemistry is bound by dependent FP64 chains, and the UMOV pairs it removes were filling
        # issue slots that are otherwise idle.  (dodecane_lu's mechanism tables fill the 64 KB constant bank anyway.)
        if os.environ.get("PRX_GEN_CONSTIFY", "0") == "1" and len(load(m)["species"]) <= 32:
            text = constify(text)
        write_if_changed(p, text)
        out.append(p)
        # build settings: a mechanism without its own entries in pelec_b200/csrc/Makefile takes those of the nearest of the three
        n = len(load(m)["species"])
        cls = m if m in ("LiDryer", "drm19", "dodecane_lu") else ("LiDryer" if n <= 14 else "drm19" if n <= 32 else "dodecane_lu")
        write_if_changed(os.path.join(repo_root, "pelec_b200", "csrc", "gen", "mech_%s.mk" % m), "CLASS_%s := %s\n" % (m, cls))
    return out


if __name__ == "__main__":
    import sys
    root = os.path.abspath(os.path.join(HERE, "..", ".."))
    for p in generate_all(root, tuple(sys.argv[1:]) or None):
        print(os.path.relpath(p, root), os.path.getsize(p), "bytes")


# ---------------------------------------------------------------------------------------------
# Algorithmic operation counts of the generated code (the FLOP model of bench.py / DESIGN.md)
# ---------------------------------------------------------------------------------------------
def _function_bodies(text):
    import re
    out = {}
    for m in re.finditer(r"^PRX_DEVFN (?:void|double) (\w+)\([^)]*\)\n\{\n(.*?)\n\}\n", text, re.S | re.M):
        out[m.group(1)] = m.group(2)
    return out


def _count(body):
    import re
    body = re.sub(r"//.*", "", body)
    trans = len(re.findall(r"\b(?:exp|log10|log)\(", body)) + 2 * len(re.findall(r"\bPRX_EXP2\(", body))
    # binary arithmetic operators (unary signs of literals are not operations)
    arith = len(re.findall(r"(?<=[\w\)\]]) [-+*/] (?=[\w\(\-])", body)) + len(re.findall(r"[-+*/]=", body))
    return arith, trans


def op_counts(mech):
    """{'F_rhs', 'T_rhs', 'F_jac', 'T_jac', 'F_lu', 'F_slv', 'F_vec', ...}: arithmetic operations (+,-,*,/ = 1)
    and transcendental calls (exp/log/log10, listed separately and costed at 1 flop in the algorithmic figure,
    SURVEY.md 8d) of ONE evaluation, counted on the text the generator emits (consP = 0 branch of the
    Jacobian).  T-Newton: 2 bound evaluations + 3 polynomial passes (2 updates + the pass that breaks)."""
    d = load(mech)
    N = len(d["species"])
    neq = N + 1
    text = emit(mech, "cuda")
    fb = _function_bodies(text)
    # thermo functions carry both temperature ranges; only one branch executes
    def one_branch(body):
        return re.sub(r"\} else \{.*?\n  \}", "}", body, flags=re.S)
    import re
    c = {k: _count(one_branch(v) if k not in ("production_rate", "ajacobian") else v) for k, v in fb.items()}
    # the Jacobian text holds both consP branches; keep the generic part + the consP = 0 branch
    import re
    jb = fb["ajacobian"]
    jb0 = re.sub(r"if \(consP\) \{.*?\} else \{", "{", jb, flags=re.S)
    cj = _count(jb0)
    newton_passes = 3
    f_rhs = (c["production_rate"][0] + c["gibbs"][0] + (2 + newton_passes) * c["ckubms"][0]
             + (newton_passes + 1) * c["ckcvbs"][0] + c["e_RT_fn"][0] + 12 * N)
    t_rhs = c["production_rate"][1] + c["gibbs"][1]
    f_jac = (cj[0] + c["gibbs"][0] + c["h_RT_fn"][0] + c["cv_R"][0] + c["dcvpRdT"][0] + c["e_RT_fn"][0]
             + 3 * neq * neq + 2 * N * N + 8 * N)
    t_jac = cj[1] + c["gibbs"][1]
    return {
        "N": N, "neq": neq, "reactions": len(d["reactions"]),
        "F_rhs": f_rhs, "T_rhs": t_rhs, "F_jac": f_jac, "T_jac": t_jac,
        "F_lu": round(2.0 / 3.0 * neq ** 3), "F_slv": 2 * neq * neq, "F_vec": 20 * neq,
        "F_rk_update": 6 * 7 * neq,
        "bytes_per_cell": (2 * N + 3) * 8 + (N + 3) * 8 + 4,
    }

// prx_tchem.cuh - TABLE-DRIVEN thermo-chemistry: the mechanism as data (gen/mech_<mech>.cuh,
// arrays tc_* / tg_*) evaluated by small rolled loops instead of straight-line generated code.
//
// Why (ncu, profiles/r01a_*): the straight-line RHS + Jacobian + integrator of even the smallest
// mechanism is 478 KB of SASS (bdf_kernel_LiDryer) - far beyond the SM's instruction caches - and
// 76 % of the warp-stall samples were "no instruction".  The table form keeps the hot code at a
// few KB, reads its constants through the constant cache (thread-per-cell: all lanes of a warp
// evaluate the same reaction, so every table access is a broadcast) or L1 (warp-per-cell: lanes
// evaluate different reactions), and keeps per-cell species vectors in shared memory where the
// data-dependent species indices of a reaction can address them.
//
// The arithmetic is the reference's, operation for operation, in the same order:
//   thermo polynomials / mixture sums   PP/Mechanisms/<mech>/mechanism.H  (cv_R, gibbs, speciesInternalEnergy,
//                                       CKCVBS, CKUBMS, ... e.g. LiDryer :199-1357)
//   GET_T_GIVEN_EY / GET_T_GIVEN_HY     mechanism.H:1419-1520
//   productionRate                      mechanism.H:2285-2603 (k_f, Corr, Troe, K_c, qf - qr, wdot scatter)
//   utils::fKernelSpec                  PP/Source/Reactions/ReactorUtils.H:358-415
// so the host instantiation (tests/host_harness, -ffp-contract=off) reproduces the oracle bit for bit.
#ifndef PRX_TCHEM_CUH
#define PRX_TCHEM_CUH

#include "prx_chem.cuh"

#ifndef PRX_SP_UNROLL
#define PRX_SP_UNROLL _Pragma("unroll")
#endif
#ifdef PRX_HOST_TEST
#define PRX_ROLLED
#else
#define PRX_ROLLED _Pragma("unroll 1")
#endif
#ifndef PRX_RX_UNROLL
#define PRX_RX_UNROLL _Pragma("unroll 2")
#endif

namespace prx {

// bit layout of T*::meta(r) (gen.py emit_tables)
PRX_DEV int rx_type(const int m) { return m & 3; } // 0 elementary, 1 three-body, 2 Troe falloff, 3 Lindemann falloff
PRX_DEV int rx_rev(const int m) { return (m >> 2) & 1; }
PRX_DEV int rx_noexp(const int m) { return (m >> 3) & 1; }
PRX_DEV int rx_nreac(const int m) { return (m >> 4) & 3; }
PRX_DEV int rx_nprod(const int m) { return (m >> 6) & 3; }
PRX_DEV int rx_ng(const int m) { return (m >> 8) & 7; }
PRX_DEV int rx_nw(const int m) { return (m >> 11) & 7; }
PRX_DEV int rx_refc(const int m) { return ((m >> 14) & 7) - 2; }

struct TPow
{
  double T, T2, T3, T4, invT, logT;
};
PRX_DEV TPow tpow_poly(const double T) // what cv_R / e_RT need
{
  TPow p;
  p.T = T;
  p.T2 = T * T;
  p.T3 = T * T * T;
  p.T4 = T * T * T * T;
  p.invT = 1.0 / T;
  p.logT = 0.0;
  return p;
}
PRX_DEV TPow tpow_full(const double T)
{
  TPow p = tpow_poly(T);
  p.logT = log(T);
  return p;
}

template <class TB>
PRX_DEV int sp_range(const int k, const double T)
{
  return (T < TB::tmid(k)) ? 0 : 1;
}
template <class TB>
PRX_DEV double sp_cv_R(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 5;
  return TB::cv(o) + TB::cv(o + 1) * p.T + TB::cv(o + 2) * p.T2 + TB::cv(o + 3) * p.T3 + TB::cv(o + 4) * p.T4;
}
template <class TB>
PRX_DEV double sp_cp_R(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 5;
  return TB::cp(o) + TB::cp(o + 1) * p.T + TB::cp(o + 2) * p.T2 + TB::cp(o + 3) * p.T3 + TB::cp(o + 4) * p.T4;
}
template <class TB>
PRX_DEV double sp_gibbs(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 7;
  return TB::gb(o) * p.invT + TB::gb(o + 1) + TB::gb(o + 2) * p.logT + TB::gb(o + 3) * p.T + TB::gb(o + 4) * p.T2 +
         TB::gb(o + 5) * p.T3 + TB::gb(o + 6) * p.T4;
}
template <class TB>
PRX_DEV double sp_e_RT(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 6;
  return TB::ei(o) + TB::ei(o + 1) * p.T + TB::ei(o + 2) * p.T2 + TB::ei(o + 3) * p.T3 + TB::ei(o + 4) * p.T4 +
         TB::ei(o + 5) * p.invT;
}
template <class TB>
PRX_DEV double sp_h_RT(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 6;
  return TB::hi(o) + TB::hi(o + 1) * p.T + TB::hi(o + 2) * p.T2 + TB::hi(o + 3) * p.T3 + TB::hi(o + 4) * p.T4 +
         TB::hi(o + 5) * p.invT;
}
template <class TB>
PRX_DEV double sp_dcvpRdT(const int k, const TPow& p)
{
  const int o = (k * 2 + sp_range<TB>(k, p.T)) * 4;
  return TB::dc(o) + TB::dc(o + 1) * p.T + TB::dc(o + 2) * p.T2 + TB::dc(o + 3) * p.T3;
}

// CKUBMS / CKHBMS (ETYPE 1 / 2) and CKCVBS / CKCPBS: mass-fraction weighted mixture sums, species
// accumulated in the reference's order (grouped by Tmid).
template <class TB, int ETYPE, class YV>
PRX_DEV_NOINLINE double mix_energy(const double T, YV Y)
{
  const TPow p = tpow_poly(T);
  double result = 0.0;
  PRX_SP_UNROLL
  for (int j = 0; j < N; j++) {
    const int k = TB::sp_order(j);
    const double poly = (ETYPE == 1) ? sp_e_RT<TB>(k, p) : sp_h_RT<TB>(k, p);
    result += Y[k] * poly * TB::imw(k);
  }
  const double RT = PRX_RU * T;
  return result * RT;
}
template <class TB, int ETYPE, class YV>
PRX_DEV_NOINLINE double mix_cvp(const double T, YV Y)
{
  const TPow p = tpow_poly(T);
  double result = 0.0;
  PRX_SP_UNROLL
  for (int j = 0; j < N; j++) {
    const int k = TB::sp_order(j);
    const double poly = (ETYPE == 1) ? sp_cv_R<TB>(k, p) : sp_cp_R<TB>(k, p);
    result += Y[k] * poly * TB::imw(k);
  }
  return result * PRX_RU;
}

// GET_T_GIVEN_EY / GET_T_GIVEN_HY - same control flow as t_from_energy (prx_chem.cuh)
template <class TB, int ETYPE, class YV>
PRX_DEV void t_from_energy_tab_impl(const double e, YV y, double& t)
{
  const double tmin = PRX_TMIN, tmax = PRX_TMAX;
  const double emin = mix_energy<TB, ETYPE>(tmin, y);
  const double emax = mix_energy<TB, ETYPE>(tmax, y);
  if (e < emin) {
    t = tmin - (emin - e) / mix_cvp<TB, ETYPE>(tmin, y);
    return;
  }
  if (e > emax) {
    t = tmax - (emax - e) / mix_cvp<TB, ETYPE>(tmax, y);
    return;
  }
  double t1 = t;
  if (t1 < tmin || t1 > tmax) {
    t1 = tmin + (tmax - tmin) / (emax - emin) * (e - emin);
  }
  PRX_ROLLED
  for (int i = 0; i < 200; ++i) {
    const double e1 = mix_energy<TB, ETYPE>(t1, y);
    const double cv = mix_cvp<TB, ETYPE>(t1, y);
    double dt = (e - e1) / cv;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (fabs(dt) < 1.e-6) {
      break;
    }
    t1 += dt;
  }
  t = t1;
}
template <class TB, class YV>
PRX_DEV void t_from_energy_tab(const int reactor_type, const double e, YV y, double& t)
{
  if (reactor_type == 1) {
    t_from_energy_tab_impl<TB, 1>(e, y, t);
  } else {
    t_from_energy_tab_impl<TB, 2>(e, y, t);
  }
}

// temperature-only factors shared by every reaction of one productionRate call
struct RxCommon
{
  double invT, logT, T, refC, refCinv, mixture;
};
PRX_DEV RxCommon rx_common(const TPow& p, const double mixture)
{
  RxCommon c;
  c.T = p.T;
  c.invT = p.invT;
  c.logT = p.logT;
  c.refC = PRX_REFC_NUM / PRX_REFC_DEN * p.invT;
  c.refCinv = 1 / c.refC;
  c.mixture = mixture;
  return c;
}

// Net rate of progress qdot of reaction r [mol/m^3/s]; sc = concentrations (SI), g = g/RT.
template <class TB, class CV, class GV>
PRX_DEV double reaction_q(const int r, const RxCommon& c, CV sc, GV g)
{
  const int m = TB::meta(r);
  double k_f = TB::A(r);
  if (!rx_noexp(m)) {
    k_f = k_f * exp(TB::beta(r) * c.logT - TB::Ea(r) * c.invT);
  }
  double pre = k_f; // (Corr *) k_f
  const int typ = rx_type(m);
  if (typ != 0) {
    const int tb = TB::tb_of(r);
    double Corr = c.mixture;
    for (int j = TB::eff_ptr(tb); j < TB::eff_ptr(tb + 1); j++) {
      Corr = Corr + TB::eff_val(j) * sc[TB::eff_idx(j)];
    }
    if (typ >= 2) {
      const int f = TB::fo_of(r);
      const double redP = Corr / k_f * TB::lowA(f) * exp(TB::lowb(f) * c.logT - TB::lowE(f) * c.invT);
      const double F = redP / (1.0 + redP);
      if (typ == 2) {
        const int fl = TB::fo_flags(f);
        const int nt = fl & 3;
        const double logPred = log10(redP);
        double fcent = TB::fc_a(2 * f) * exp(-c.T * TB::fc_x(2 * f));
        if (nt > 1) {
          fcent = fcent + TB::fc_a(2 * f + 1) * exp(-c.T * TB::fc_x(2 * f + 1));
        }
        if (fl & 4) {
          fcent = fcent + exp(TB::fc_invT(f) * c.invT);
        }
        fcent = fcent + TB::fc_const(f);
        const double logFcent = log10(fcent);
        const double troe_c = -0.4 - 0.67 * logFcent;
        const double troe_n = 0.75 - 1.27 * logFcent;
        const double troe = (troe_c + logPred) / (troe_n - 0.14 * (troe_c + logPred));
        const double F_troe = exp(PRX_LN10 * logFcent / (1.0 + troe * troe));
        Corr = F * F_troe;
      } else {
        Corr = F;
      }
    }
    pre = Corr * k_f;
  }
  double phi_f = sc[TB::reac(3 * r)];
  const int nr = rx_nreac(m);
  if (nr > 1) {
    phi_f = phi_f * sc[TB::reac(3 * r + 1)];
  }
  if (nr > 2) {
    phi_f = phi_f * sc[TB::reac(3 * r + 2)];
  }
  const double qf = pre * phi_f;
  double qr = 0.0;
  if (rx_rev(m)) {
    const int ng = rx_ng(m);
    double dG = (double)TB::gcoef(6 * r) * g[TB::gidx(6 * r)];
    for (int j = 1; j < ng; j++) {
      dG = dG + (double)TB::gcoef(6 * r + j) * g[TB::gidx(6 * r + j)];
    }
    double phi_r = sc[TB::prod(3 * r)];
    const int np = rx_nprod(m);
    if (np > 1) {
      phi_r = phi_r * sc[TB::prod(3 * r + 1)];
    }
    if (np > 2) {
      phi_r = phi_r * sc[TB::prod(3 * r + 2)];
    }
    double kr = pre * exp(-(dG));
    const int rc = rx_refc(m);
    if (rc == 1) {
      kr = kr * (c.refC);
    } else if (rc == -1) {
      kr = kr * (c.refCinv);
    } else if (rc == 2) {
      kr = kr * (c.refC * c.refC);
    } else if (rc == -2) {
      kr = kr * (c.refCinv * c.refCinv);
    }
    qr = kr * (phi_r);
  }
  return qf - qr;
}

// =====================================================================================
// Table-driven analytic Jacobian (aJacobian, e.g. LiDryer mechanism.H:3905-5138 + DWDOT :5141-5162):
// the same per-reaction derivation the generator writes out as straight-line code (gen.py emit_ajacobian),
// evaluated from the tables.  reaction_jac is shared by the serial form below (host-tested bit-for-bit
// against the generated ajacobian) and by the cooperative form (prx_coop.cuh jac_coop).
// =====================================================================================
struct JacCommon
{
  double T, invT, invT2, logT, refC, refCinv, mixture;
};
PRX_DEV JacCommon jac_common(const double T, const double mixture)
{
  JacCommon c;
  c.T = T;
  c.invT = 1.0 / T;
  c.invT2 = c.invT * c.invT;
  c.logT = log(T);
  c.refC = PRX_REFC_NUM / PRX_REFC_DEN / T;
  c.refCinv = 1.0 / c.refC;
  c.mixture = mixture;
  return c;
}

struct RxJac
{
  double q, dqdT, kf, kr, dcdc; // kf/kr already carry the third-body / falloff correction
};

template <class TB, class CV, class GV, class HV>
PRX_DEV RxJac reaction_jac(const int r, const JacCommon& c, CV sc, GV g, HV h)
{
  const int m = TB::meta(r);
  RxJac o;
  double phi_f = sc[TB::reac(3 * r)];
  const int nr = rx_nreac(m);
  if (nr > 1) {
    phi_f = phi_f * sc[TB::reac(3 * r + 1)];
  }
  if (nr > 2) {
    phi_f = phi_f * sc[TB::reac(3 * r + 2)];
  }
  double k_f = TB::A(r);
  if (!rx_noexp(m)) {
    k_f = k_f * exp(TB::beta(r) * c.logT - TB::Ea(r) * c.invT);
  }
  const double dlnkfdT = TB::beta(r) * c.invT + TB::Ea(r) * c.invT2;
  double phi_r = 0.0, k_r = 0.0, dkrdT = 0.0;
  if (rx_rev(m)) {
    phi_r = sc[TB::prod(3 * r)];
    const int np = rx_nprod(m);
    if (np > 1) {
      phi_r = phi_r * sc[TB::prod(3 * r + 1)];
    }
    if (np > 2) {
      phi_r = phi_r * sc[TB::prod(3 * r + 2)];
    }
    const int ng = rx_ng(m);
    double dG = (double)TB::gcoef(6 * r) * g[TB::gidx(6 * r)];
    double dH = (double)TB::gcoef(6 * r) * h[TB::gidx(6 * r)];
    for (int j = 1; j < ng; j++) {
      dG = dG + (double)TB::gcoef(6 * r + j) * g[TB::gidx(6 * r + j)];
      dH = dH + (double)TB::gcoef(6 * r + j) * h[TB::gidx(6 * r + j)];
    }
    double Kinv = exp(-(dG));
    const int rc = rx_refc(m);
    if (rc == 1) {
      Kinv = c.refC * Kinv;
    } else if (rc == -1) {
      Kinv = c.refCinv * Kinv;
    } else if (rc == 2) {
      Kinv = (c.refC * c.refC) * Kinv;
    } else if (rc == -2) {
      Kinv = (c.refCinv * c.refCinv) * Kinv;
    }
    k_r = k_f * Kinv;
    const double dlnKinvdT = c.invT * ((dH) - (double)rc);
    dkrdT = (dlnkfdT + dlnKinvdT) * k_r;
  }
  const double q_nocor = k_f * phi_f - k_r * phi_r;
  const int typ = rx_type(m);
  if (typ == 0) {
    o.q = q_nocor;
    o.dqdT = dlnkfdT * k_f * phi_f - dkrdT * phi_r;
    o.dcdc = 0.0;
  } else {
    const int tb = TB::tb_of(r);
    double alpha = c.mixture;
    for (int j = TB::eff_ptr(tb); j < TB::eff_ptr(tb + 1); j++) {
      alpha = alpha + TB::eff_val(j) * sc[TB::eff_idx(j)];
    }
    if (typ == 1) {
      o.q = alpha * q_nocor;
      o.dqdT = alpha * (dlnkfdT * k_f * phi_f - dkrdT * phi_r);
      o.dcdc = q_nocor;
      k_f *= alpha;
      k_r *= alpha;
    } else {
      const int f = TB::fo_of(r);
      const double ln10 = PRX_LN10;
      const double log10e = 1.0 / PRX_LN10;
      const double k_0 = TB::lowA(f) * exp(TB::lowb(f) * c.logT - TB::lowE(f) * c.invT);
      const double Pr = alpha / k_f * k_0;
      const double fPr = Pr / (1.0 + Pr);
      const double dlnk0dT = TB::lowb(f) * c.invT + TB::lowE(f) * c.invT2;
      const double dlogPrdT = log10e * (dlnk0dT - dlnkfdT);
      const double dlogfPrdT = dlogPrdT / (1.0 + Pr);
      double F = 1.0, dlogFdlogPr = 0.0, dlogFdT = 0.0;
      if (typ == 2) {
        const int fl = TB::fo_flags(f);
        const int nt = fl & 3;
        const double logPr = log10(Pr);
        const double Fc0 = (TB::fc_a(2 * f)) * exp(-c.T * TB::fc_x(2 * f));
        double Fcent = Fc0;
        double dsum = -Fc0 * TB::fc_x(2 * f);
        if (nt > 1) {
          const double Fc1 = (TB::fc_a(2 * f + 1)) * exp(-c.T * TB::fc_x(2 * f + 1));
          Fcent = Fcent + Fc1;
          dsum = dsum + (-Fc1 * TB::fc_x(2 * f + 1));
        }
        if (fl & 4) {
          const double FcT = exp(TB::fc_invT(f) * c.invT);
          Fcent = Fcent + FcT;
          dsum = dsum + (-FcT * (TB::fc_invT(f)) * c.invT2);
        }
        Fcent = Fcent + TB::fc_const(f);
        const double logFcent = log10(Fcent);
        const double troe_c = -0.4 - 0.67 * logFcent;
        const double troe_n = 0.75 - 1.27 * logFcent;
        const double troePr_den = 1.0 / (troe_n - 0.14 * (troe_c + logPr));
        const double troePr = (troe_c + logPr) * troePr_den;
        const double troe = 1.0 / (1.0 + troePr * troePr);
        F = exp(ln10 * logFcent * troe);
        const double dlogFcentdT = log10e / Fcent * (dsum);
        const double dlogFdcn_fac = 2.0 * logFcent * troe * troe * troePr * troePr_den;
        const double dlogFdc = -troe_n * dlogFdcn_fac * troePr_den;
        const double dlogFdn = dlogFdcn_fac * troePr;
        dlogFdlogPr = dlogFdc;
        dlogFdT = dlogFcentdT * (troe - 0.67 * dlogFdc - 1.27 * dlogFdn) + dlogFdlogPr * dlogPrdT;
      }
      const double Corr = fPr * F;
      o.q = Corr * q_nocor;
      const double dlnCorrdT = ln10 * (dlogfPrdT + dlogFdT);
      o.dqdT = Corr * (dlnkfdT * k_f * phi_f - dkrdT * phi_r) + dlnCorrdT * o.q;
      o.dcdc = o.q / alpha * (1.0 / (Pr + 1.0) + dlogFdlogPr);
      k_f *= Corr;
      k_r *= Corr;
    }
  }
  o.kf = k_f;
  o.kr = k_r;
  return o;
}

// one "direct" term n * k * sc[o1] * sc[o2] (packed n | o1 << 8 | o2 << 16, 255 = no factor)
template <class CV>
PRX_DEV double direct_term(const int packed, const double k, CV sc)
{
  const int n = packed & 255, o1 = (packed >> 8) & 255, o2 = (packed >> 16) & 255;
  double t = (n > 1) ? (double)n * k : k;
  if (o1 != 255) {
    t = t * sc[o1];
  }
  if (o2 != 255) {
    t = t * sc[o2];
  }
  return t;
}
// dq/dC of direct species d of reaction r: [+ forward term] [- reverse term]
template <class TB, class CV>
PRX_DEV double direct_dqdc(const int r, const int d, const RxJac& x, CV sc)
{
  const int pf = TB::jd_f(6 * r + d), pr = TB::jd_r(6 * r + d);
  double v = 0.0;
  if (pf != 0) {
    v = direct_term(pf, x.kf, sc);
    if (pr != 0) {
      v = v - direct_term(pr, x.kr, sc);
    }
  } else {
    v = -direct_term(pr, x.kr, sc);
  }
  return v;
}

// Serial form (one thread): J column-major NEQ x NEQ, PRX_J(row, col) = J[col*NEQ + row]; sc = SI
// concentrations; scratch vectors g, h, wdot, dqdc of N doubles each.
template <class TB>
PRX_DEV void ajacobian_tab(double* J, const double* sc, const double T, const int consP, double* g, double* h,
                           double* wdot, double* dqdc)
{
  for (int i = 0; i < NEQ * NEQ; i++) {
    J[i] = 0.0;
  }
  double mixture = 0.0;
  for (int k = 0; k < N; k++) {
    wdot[k] = 0.0;
    dqdc[k] = 0.0;
    mixture += sc[k];
  }
  const JacCommon c = jac_common(T, mixture);
  TPow p = tpow_poly(T);
  p.logT = c.logT;
  for (int k = 0; k < N; k++) {
    g[k] = sp_gibbs<TB>(k, p);
    h[k] = sp_h_RT<TB>(k, p);
  }
  for (int r = 0; r < PRX_NUM_REACTIONS; r++) {
    const RxJac x = reaction_jac<TB>(r, c, sc, g, h);
    const int jm = TB::jmeta(r);
    const int nnu = jm & 15, nd = (jm >> 4) & 15;
    for (int j = 0; j < nnu; j++) {
      wdot[TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * x.q;
    }
    const int typ = rx_type(TB::meta(r));
    if (typ == 0) {
      for (int d = 0; d < nd; d++) {
        const int k = TB::jd_k(6 * r + d);
        const double v = direct_dqdc<TB>(r, d, x, sc);
        for (int j = 0; j < nnu; j++) {
          J[k * NEQ + TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * v;
        }
      }
    } else {
      const int tb = TB::tb_of(r);
      for (int k = 0; k < N; k++) {
        int d = -1;
        for (int dd = 0; dd < nd; dd++) {
          if (TB::jd_k(6 * r + dd) == k) {
            d = dd;
          }
        }
        if (consP) {
          if (d < 0 && !TB::tb_has(tb * N + k)) {
            continue;
          }
          double v;
          if (TB::tb_has(tb * N + k)) {
            v = (TB::tb_eff(tb * N + k)) * x.dcdc;
            if (d >= 0) {
              const int pf = TB::jd_f(6 * r + d), pr = TB::jd_r(6 * r + d);
              if (pf != 0) {
                v = v + direct_term(pf, x.kf, sc);
              }
              if (pr != 0) {
                v = v - direct_term(pr, x.kr, sc);
              }
            }
          } else {
            v = direct_dqdc<TB>(r, d, x, sc);
          }
          for (int j = 0; j < nnu; j++) {
            J[k * NEQ + TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * v;
          }
        } else {
          // reference quirk (gen.py emit_ajacobian): dqdc[k] keeps its value from an earlier reaction when
          // 1 + eff_k == 0 and k has no direct term
          const double tot = TB::tb_tot(tb * N + k);
          if (tot != 0.0) {
            double v = (tot == 1.0) ? x.dcdc : (tot)*x.dcdc;
            if (d >= 0) {
              const int pf = TB::jd_f(6 * r + d), pr = TB::jd_r(6 * r + d);
              if (pf != 0) {
                v = v + direct_term(pf, x.kf, sc);
              }
              if (pr != 0) {
                v = v - direct_term(pr, x.kr, sc);
              }
            }
            dqdc[k] = v;
          } else if (d >= 0) {
            dqdc[k] = direct_dqdc<TB>(r, d, x, sc);
          }
        }
      }
      if (!consP) {
        for (int k = 0; k < N; k++) {
          for (int j = 0; j < nnu; j++) {
            J[k * NEQ + TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * dqdc[k];
          }
        }
      }
    }
    for (int j = 0; j < nnu; j++) {
      J[N * NEQ + TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * x.dqdT;
    }
  }
  // temperature row (g is reused for c_R, dqdc for dcRdT, h becomes eh_RT)
  double cmix = 0.0, ehmix = 0.0, dcmixdT = 0.0, dehmixdT = 0.0;
  for (int k = 0; k < N; k++) {
    const double c_R = consP ? sp_cp_R<TB>(k, p) : sp_cv_R<TB>(k, p);
    const double eh = consP ? h[k] : sp_e_RT<TB>(k, p);
    const double dcR = sp_dcvpRdT<TB>(k, p);
    g[k] = c_R;
    h[k] = eh;
    cmix += c_R * sc[k];
    dcmixdT += dcR * sc[k];
    ehmix += eh * wdot[k];
    dehmixdT += c.invT * (c_R - eh) * wdot[k] + eh * J[N * NEQ + k];
  }
  const double cmixinv = 1.0 / cmix;
  const double tmp1 = ehmix * cmixinv;
  const double tmp3 = cmixinv * T;
  const double tmp2 = tmp1 * tmp3;
  for (int k = 0; k < N; k++) {
    double dehmixdc = 0.0;
    for (int mm = 0; mm < N; mm++) {
      dehmixdc += h[mm] * J[k * NEQ + mm];
    }
    J[k * NEQ + N] = tmp2 * g[k] - tmp3 * dehmixdc;
  }
  J[N * NEQ + N] = -tmp1 + tmp2 * dcmixdT - tmp3 * dehmixdT;
}

// jac_analytic (prx_chem.cuh) with the table-driven aJacobian: serial form used by the host tests
template <class TB, class Store>
PRX_DEV void jac_analytic_tab(const int reactor_type, const double* ydata, double* Jt, Store&& JM)
{
  double rho = 0.0;
  for (int i = 0; i < N; i++) {
    rho = rho + ydata[i];
  }
  const double temp = ydata[N];
  double c[N], g[N], h[N], wdot[N], dqdc[N];
  for (int i = 0; i < N; i++) {
    const double Yi = ydata[i] / rho;
    c[i] = 1.e6 * (rho * Yi * TB::imw(i));
  }
  ajacobian_tab<TB>(Jt, c, temp, reactor_type == 2, g, h, wdot, dqdc);
  for (int i = 0; i < N; i++) {
    for (int k = 0; k < N; k++) {
      JM(k, i, Jt[i * NEQ + k] * TB::mw(k) * TB::imw(i));
    }
    JM(N, i, (Jt[i * NEQ + N] * 1.e6) * TB::imw(i));
  }
  for (int i = 0; i < N; i++) {
    JM(i, N, (Jt[N * NEQ + i] * 1.e-6) * TB::mw(i));
  }
  JM(N, N, Jt[N * NEQ + N]);
}

// per-cell constants of the ODE that do not fit in registers stay in HBM/L2: the external species
// source is re-read through this accessor at every RHS (N coalesced loads, L2-resident)
struct SrcScalars
{
  double rhoe_init, rhoesrc;
};

// ODE right-hand side, one cell per thread.  y/ydot: N+1 vectors (any accessor), sc/gs: N-vectors of
// scratch with dynamic (warp-uniform) indexing - shared memory columns in the kernels.
template <class TB, class VIn, class VOut, class VS, class VG, class SrcF>
PRX_DEV void rhs_tab(
  const double t, const int reactor_type, VIn y, VOut ydot, VS sc, VG gs, const SrcScalars& c, SrcF rYsrc)
{
  double rho = 0.0;
  PRX_SP_UNROLL
  for (int n = 0; n < N; n++) {
    rho += y[n];
  }
  const double rho_inv = 1.0 / rho;
  PRX_SP_UNROLL
  for (int n = 0; n < N; n++) {
    sc[n] = y[n] * rho_inv; // mass fractions
  }
  const double nrg = (c.rhoe_init + c.rhoesrc * t) * rho_inv;
  double temp = y[N];
  t_from_energy_tab<TB>(reactor_type, nrg, sc, temp);
  const double Cv = (reactor_type == 1) ? mix_cvp<TB, 1>(temp, sc) : mix_cvp<TB, 2>(temp, sc);
  const TPow p = tpow_full(temp);
  double mixture = 0.0;
  PRX_SP_UNROLL
  for (int i = 0; i < N; i++) {
    double C = rho * sc[i] * TB::imw(i);
    C *= 1.0e6;
    sc[i] = C;
    mixture += C;
    gs[i] = sp_gibbs<TB>(i, p);
    ydot[i] = 0.0;
  }
  const RxCommon rc = rx_common(p, mixture);
  PRX_RX_UNROLL
  for (int r = 0; r < PRX_NUM_REACTIONS; r++) {
    const double q = reaction_q<TB>(r, rc, sc, gs);
    const int nw = rx_nw(TB::meta(r));
    for (int j = 0; j < nw; j++) {
      const int k = TB::widx(5 * r + j);
      ydot[k] = ydot[k] + (double)TB::wcoef(5 * r + j) * q;
    }
  }
  const double RT = PRX_RU * temp;
  double rhoesrc = c.rhoesrc;
  PRX_SP_UNROLL
  for (int n = 0; n < N; n++) {
    double w = ydot[n] * 1.0e-6;
    w *= TB::mw(n);
    const double cdot_rYs = w + rYsrc(n);
    ydot[n] = cdot_rYs;
    double ei = (reactor_type == 1) ? sp_e_RT<TB>(n, p) : sp_h_RT<TB>(n, p);
    ei *= RT * TB::imw(n);
    rhoesrc -= cdot_rYs * ei;
  }
  ydot[N] = rhoesrc * (rho_inv / Cv);
}

} // namespace prx
#endif

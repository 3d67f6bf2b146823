// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// The surface-forcing and vertical-mixing physics the shipped BENCHMARK cpp set (ROMS/Include/benchmark.h) adds to the
// main3d chain: BULK_FLUXES (ROMS/Nonlinear/bulk_flux.F), LMD_MIXING with LMD_RIMIX, LMD_CONVEC, LMD_SKPP, LMD_NONLOCAL and
// RI_SPLINES (ROMS/Nonlinear/lmd_vmix.F, lmd_skpp.F, lmd_swfrac.F), and the analytical atmosphere set_data.F asks the
// ana_*.h functionals for (ANA_WINDS, ANA_TAIR, ANA_PAIR, ANA_HUMIDITY, ANA_RAIN, ANA_CLOUD, ANA_SRFLUX + ALBEDO).
// Parity status: no known answers for these routines exist in the reference tree -> PARITY UNPINNED.
#include "roms_oracle.hpp"

namespace orc {

static const double pi = 3.14159265358979323846;       // mod_scalars.F:788
static const double deg2rad = pi / 180.0;              // mod_scalars.F:789
// mod_scalars.F:431-444
static const double Cp = 3985.0, Csolar = 1353.0, StefBo = 5.67e-8, emmiss = 0.97, rhow = 1000.0, vonKar = 0.41;
// mod_scalars.F:1415-1422
static const double blk_Cpa = 1004.67, blk_Cpw = 4000.0, blk_Rgas = 287.1, blk_Zabl = 600.0, blk_beta = 1.2;
// mod_scalars.F:1502-1512 (Jerlov water types), :1552-1629 (LMD constants; lmd_nuwm/lmd_nuws are unused by lmd_vmix.F)
static const double lmd_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};
static const double lmd_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};
static const double lmd_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};
static const double lmd_Ri0 = 0.7, lmd_bvfcon = -2.0e-5, lmd_nu0c = 0.01, lmd_nu0m = 10.0e-4, lmd_nu0s = 10.0e-4;
static const double lmd_Cstar = 10.0, lmd_Cv = 1.25, lmd_Ric = 0.3, lmd_am = 1.257, lmd_as = -28.86, lmd_betaT = -0.2,
                    lmd_cekman = 0.7, lmd_cmonob = 1.0, lmd_cm = 8.36, lmd_cs = 98.96, lmd_epsilon = 0.1, lmd_zetam = -0.2,
                    lmd_zetas = -1.0;

// ---------------------------------------------------------------------------------------------------------------------
// set_data.F:197-394, :628 for the BENCHMARK analytical atmosphere: ana_cloud.h (Cval = 0.6), ana_tair.h (4 degC),
// ana_humid.h (0.8), ana_srflux.h ALBEDO branch, ana_winds.h BENCHMARK branch, ana_rain.h (0), ana_pair.h (1025 mb); each
// functional ends with the periodic exchange of its array.  The day of the year and the hour come from caldate
// (Utility/dateclock.F:161-172, TIME_REF = 0, DSTART = 0): yday = 1 + whole days (of the first year) + day fraction.
void ana_atmosphere(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  for (int j = JstrT; j <= JendT; ++j)
    for (int i = IstrT; i <= IendT; ++i) { m.cloud(i, j) = 0.6; m.Tair(i, j) = 4.0; m.Hair(i, j) = 0.8; }
  exchange_r2d(m, b, m.cloud); exchange_r2d(m, b, m.Tair); exchange_r2d(m, b, m.Hair);
  {
    const double alb_w = 0.06;
    const double whole = std::floor(m.tdays), frac = std::fabs(m.tdays - whole);
    const double yday = (double)(1 + ((long)whole) % 365) + frac, hour = 24.0 * frac;
    double Dangle = 23.44 * std::cos((172.0 - yday) * 2.0 * pi / 365.2425);
    Dangle = Dangle * deg2rad;
    const double Hangle = (12.0 - hour) * pi / 12.0;
    const double Rsolar = Csolar / (c.rho0 * Cp);
    for (int j = JstrT; j <= JendT; ++j)
      for (int i = IstrT; i <= IendT; ++i) {
        const double LatRad = m.latr(i, j) * deg2rad;
        const double cff1 = std::sin(LatRad) * std::sin(Dangle);
        const double cff2 = std::cos(LatRad) * std::cos(Dangle);
        m.srflx(i, j) = 0.0;
        const double zenith = cff1 + cff2 * std::cos(Hangle - m.lonr(i, j) * deg2rad);
        if (zenith > 0.0) {
          const double cff = (0.7859 + 0.03477 * m.Tair(i, j)) / (1.0 + 0.00412 * m.Tair(i, j));
          const double e_sat = std::pow(10.0, cff);
          const double vap_p = e_sat * m.Hair(i, j);
          const double cl = m.cloud(i, j);
          m.srflx(i, j) = Rsolar * zenith * zenith * (1.0 - 0.6 * (cl * cl * cl)) / ((zenith + 2.7) * vap_p * 1.0e-3 + 1.085 * zenith + 0.1);
        }
        m.srflx(i, j) = (1.0 - alb_w) * m.srflx(i, j);
      }
    exchange_r2d(m, b, m.srflx);
  }
  {
    const double Wmag = 15.0;
    for (int j = JstrT; j <= JendT; ++j)
      for (int i = IstrT; i <= IendT; ++i) {
        const double cff = 0.2 * (60.0 + m.latr(i, j));
        m.Uwind(i, j) = Wmag * std::exp(-cff * cff);
        m.Vwind(i, j) = 0.0;
      }
    exchange_r2d(m, b, m.Uwind); exchange_r2d(m, b, m.Vwind);
  }
  for (int j = JstrT; j <= JendT; ++j)
    for (int i = IstrT; i <= IendT; ++i) { m.rain(i, j) = 0.0; m.Pair(i, j) = 1025.0; }
  exchange_r2d(m, b, m.rain); exchange_r2d(m, b, m.Pair);
}

// ---------------------------------------------------------------------------------------------------------------------
// bulk_flux.F:950-1007 (bulk_psiu) and :1009-1066 (bulk_psit)
static double bulk_psiu(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = std::pow(1.0 - 15.0 * ZoL, 0.25);
    const double psik = 2.0 * std::log(0.5 * (1.0 + x)) + std::log(0.5 * (1.0 + x * x)) - 2.0 * std::atan(x) + 0.5 * pi;
    double cff = std::sqrt(3.0);
    const double y = std::pow(1.0 - 10.15 * ZoL, r3);
    const double psic = 1.5 * std::log(r3 * (1.0 + y + y * y)) - cff * std::atan((1.0 + 2.0 * y) / cff) + pi / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = std::min(50.0, 0.35 * ZoL);
  return -((1.0 + ZoL) + 0.6667 * (ZoL - 14.28) / std::exp(cff) + 8.525);
}
static double bulk_psit(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = std::pow(1.0 - 15.0 * ZoL, 0.5);
    const double psik = 2.0 * std::log(0.5 * (1.0 + x));
    double cff = std::sqrt(3.0);
    const double y = std::pow(1.0 - 34.15 * ZoL, r3);
    const double psic = 1.5 * std::log(r3 * (1.0 + y + y * y)) - cff * std::atan((1.0 + 2.0 * y) / cff) + pi / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = std::min(50.0, 0.35 * ZoL);
  return -(std::pow(1.0 + 2.0 * ZoL, 1.5) + 0.6667 * (ZoL - 14.28) / std::exp(cff) + 8.525);
}

// bulk_flux_tile (bulk_flux.F:381-948) with LONGWAVE; not COOL_SKIN, EMINUSP, WIND_MINUS_CURRENT, COARE_*: the COARE 3.0
// iteration at every rho point of Istr-1:IendR x Jstr-1:JendR, then the kinematic fluxes and the stresses at u / v points.
// The per-row 1-D work arrays of the reference carry no dependence between points, so a point is done at a time.
void bulk_flux(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, nrhs = m.nrhs;
  const double g = c.g, rho0 = c.rho0;
  const double eps = 1.0e-20, r3 = 1.0 / 3.0;
  const double blk_ZW = c.blk_ZW, blk_ZT = c.blk_ZT, blk_ZQ = c.blk_ZQ;
  const int IterMax = 3;
  S2 LHeat(IminS, ImaxS, JminS, JmaxS), LRad(IminS, ImaxS, JminS, JmaxS), SHeat(IminS, ImaxS, JminS, JmaxS), Taux(IminS, ImaxS, JminS, JmaxS),
      Tauy(IminS, ImaxS, JminS, JmaxS);
  F3 T = m.t[nrhs][c.itemp - 1];
  for (int j = Jstr - 1; j <= JendR; ++j)
    for (int i = Istr - 1; i <= IendR; ++i) {
      const double Uair = m.Uwind(i, j), Vair = m.Vwind(i, j);                 // :396-401
      const double Wmag = std::sqrt(Uair * Uair + Vair * Vair);                // :413
      const double PairM = m.Pair(i, j);
      const double TairC = m.Tair(i, j), TairK = TairC + 273.16;
      const double TseaC = T(i, j, N), TseaK = TseaC + 273.16;
      const double RH = m.Hair(i, j);
      // LONGWAVE: Berliand (:448-458)
      double cff = (0.7859 + 0.03477 * TairC) / (1.0 + 0.00412 * TairC);
      const double e_sat = std::pow(10.0, cff);
      const double vap_p = e_sat * RH;
      double cff2 = TairK * TairK * TairK;
      double cff1 = cff2 * TairK;
      LRad(i, j) = -emmiss * StefBo * (cff1 * (0.39 - 0.05 * std::sqrt(vap_p)) * (1.0 - 0.6823 * m.cloud(i, j) * m.cloud(i, j)) + cff2 * 4.0 * (TseaK - TairK));
      // specific humidities (:514-544)
      cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * std::exp(17.502 * TairC / (240.97 + TairC));
      const double Qair = 0.62197 * (cff / (PairM - 0.378 * cff));
      double Q;
      if (RH < 2.0) { cff = cff * RH; Q = 0.62197 * (cff / (PairM - 0.378 * cff)); }
      else Q = RH / 1000.0;
      cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * std::exp(17.502 * TseaC / (240.97 + TseaC));
      cff = cff * 0.98;
      const double Qsea = 0.62197 * (cff / (PairM - 0.378 * cff));
      // :554-578
      const double rhoAir = PairM * 100.0 / (blk_Rgas * TairK * (1.0 + 0.61 * Q));
      const double VisAir = 1.326e-5 * (1.0 + TairC * (6.542e-3 + TairC * (8.301e-6 - 4.84e-9 * TairC)));
      const double Hlv = (2.501 - 0.00237 * TseaC) * 1.0e+6;
      double Wgus = 0.5;
      double delW = std::sqrt(Wmag * Wmag + Wgus * Wgus);
      const double delQ = Qsea - Q, delT = TseaC - TairC;
      // neutral coefficients (:582-592)
      double ZoW = 0.0001;
      const double u10 = delW * std::log(10.0 / ZoW) / std::log(blk_ZW / ZoW);
      double Wstar = 0.035 * u10;
      const double Zo10 = 0.011 * Wstar * Wstar / g + 0.11 * VisAir / Wstar;
      double t_ = vonKar / std::log(10.0 / Zo10);
      const double Cd10 = t_ * t_;
      const double Ch10 = 0.00115;
      const double Ct10 = Ch10 / std::sqrt(Cd10);
      const double ZoT10 = 10.0 / std::exp(vonKar / Ct10);
      t_ = vonKar / std::log(blk_ZW / Zo10);
      double Cd = t_ * t_;
      // Richardson number (:596-612)
      const double Ct = vonKar / std::log(blk_ZT / ZoT10);
      const double CC = vonKar * Ct / Cd;
      const double delTc = 0.0, delQc = 0.0;
      const double Ribcu = -blk_ZW / (blk_Zabl * 0.004 * (blk_beta * blk_beta * blk_beta));
      const double Ri = -g * blk_ZW * ((delT - delTc) + 0.61 * TairK * delQ) / (TairK * delW * delW);
      double Zetu;
      if (Ri < 0.0) Zetu = CC * Ri / (1.0 + Ri / Ribcu);
      else Zetu = CC * Ri / (1.0 + 3.0 * Ri / CC);
      const double L10 = blk_ZW / Zetu;
      // first guesses (:616-624)
      Wstar = delW * vonKar / (std::log(blk_ZW / Zo10) - bulk_psiu(blk_ZW / L10));
      double Tstar = -(delT - delTc) * vonKar / (std::log(blk_ZT / ZoT10) - bulk_psit(blk_ZT / L10));
      double Qstar = -(delQ - delQc) * vonKar / (std::log(blk_ZQ / ZoT10) - bulk_psit(blk_ZQ / L10));
      // Charnock (:629-637)
      double charn;
      if (delW > 18.0) charn = 0.018;
      else if (10.0 < delW && delW <= 18.0) charn = 0.011 + 0.125 * (0.018 - 0.011) * (delW - 10.0);
      else charn = 0.011;
      // iteration (:655-715)
      for (int Iter = 1; Iter <= IterMax; ++Iter) {
        ZoW = charn * Wstar * Wstar / g + 0.11 * VisAir / (Wstar + eps);
        const double Rr = ZoW * Wstar / VisAir;
        const double ZoQ = std::min(1.15e-4, 5.5e-5 / std::pow(Rr, 0.6));
        const double ZoT = ZoQ;
        const double ZoL = vonKar * g * blk_ZW * (Tstar * (1.0 + 0.61 * Q) + 0.61 * TairK * Qstar) / (TairK * Wstar * Wstar * (1.0 + 0.61 * Q) + eps);
        const double L = blk_ZW / (ZoL + eps);
        const double Wpsi = bulk_psiu(ZoL);
        const double Tpsi = bulk_psit(blk_ZT / L);
        const double Qpsi = bulk_psit(blk_ZQ / L);
        Wstar = std::max(eps, delW * vonKar / (std::log(blk_ZW / ZoW) - Wpsi));
        Tstar = -(delT - delTc) * vonKar / (std::log(blk_ZT / ZoT) - Tpsi);
        Qstar = -(delQ - delQc) * vonKar / (std::log(blk_ZQ / ZoQ) - Qpsi);
        const double Bf = -g / TairK * Wstar * (Tstar + 0.61 * TairK * Qstar);
        if (Bf > 0.0) Wgus = blk_beta * std::pow(Bf * blk_Zabl, r3);
        else Wgus = 0.2;
        delW = std::sqrt(Wmag * Wmag + Wgus * Wgus);
      }
      // fluxes (:775-855)
      const double Wspeed = std::sqrt(Wmag * Wmag + Wgus * Wgus);
      Cd = Wstar * Wstar / (Wspeed * Wspeed + eps);
      const double Hs = -blk_Cpa * rhoAir * Wstar * Tstar;
      const double diffw = 2.11e-5 * std::pow(TairK / 273.16, 1.94);
      const double diffh = 0.02411 * (1.0 + TairC * (3.309e-3 - 1.44e-6 * TairC)) / (rhoAir * blk_Cpa);
      cff = Qair * Hlv / (blk_Rgas * TairK * TairK);
      const double wet_bulb = 1.0 / (1.0 + 0.622 * (cff * Hlv * diffw) / (blk_Cpa * diffh));
      const double Hsr = m.rain(i, j) * wet_bulb * blk_Cpw * ((TseaC - TairC) + (Qsea - Q) * Hlv / blk_Cpa);
      SHeat(i, j) = (Hs + Hsr);
      const double Hl = -Hlv * rhoAir * Wstar * Qstar;
      const double upvel = -1.61 * Wstar * Qstar - (1.0 + 1.61 * Q) * Wstar * Tstar / TairK;
      const double Hlw = rhoAir * Hlv * upvel * Q;
      LHeat(i, j) = (Hl + Hlw);
      const double Taur = 0.85 * m.rain(i, j) * Wmag;
      cff = rhoAir * Cd * Wspeed;
      Taux(i, j) = (cff * Uair + Taur * std::copysign(1.0, Uair));
      Tauy(i, j) = (cff * Vair + Taur * std::copysign(1.0, Vair));
    }
  // kinematic fluxes (:868-877), stresses (:904-924)
  const double Hscale = 1.0 / (rho0 * Cp);
  for (int j = JstrR; j <= JendR; ++j)
    for (int i = IstrR; i <= IendR; ++i) {
      m.lrflx(i, j) = LRad(i, j) * Hscale;
      m.lhflx(i, j) = -LHeat(i, j) * Hscale;
      m.shflx(i, j) = -SHeat(i, j) * Hscale;
      m.stflux[c.itemp - 1](i, j) = (m.srflx(i, j) + m.lrflx(i, j) + m.lhflx(i, j) + m.shflx(i, j));
    }
  (void)rhow;                                                           // EMINUSP only (:869, :888)
  const double cff = 0.5 / rho0;
  for (int j = JstrR; j <= JendR; ++j)
    for (int i = Istr; i <= IendR; ++i) m.sustr(i, j) = cff * (Taux(i - 1, j) + Taux(i, j));
  for (int j = Jstr; j <= JendR; ++j)
    for (int i = IstrR; i <= IendR; ++i) m.svstr(i, j) = cff * (Tauy(i, j - 1) + Tauy(i, j));
  // :930-960
  exchange_r2d(m, b, m.lrflx); exchange_r2d(m, b, m.lhflx); exchange_r2d(m, b, m.shflx); exchange_r2d(m, b, m.stflux[c.itemp - 1]);
  exchange_u2d(m, b, m.sustr); exchange_v2d(m, b, m.svstr);
}

// ---------------------------------------------------------------------------------------------------------------------
// lmd_swfrac_tile (lmd_swfrac.F:66-80) at one point
static inline double lmd_swfrac(double Zscale, double Z, int Jindex) {
  const double fac1 = Zscale / lmd_mu1[Jindex - 1], fac2 = Zscale / lmd_mu2[Jindex - 1], fac3 = lmd_r1[Jindex - 1];
  return std::exp(Z * fac1) * fac3 + std::exp(Z * fac2) * (1.0 - fac3);
}

// turbulent velocity scales (lmd_skpp.F:454-476, :700-722, :835-857: the same block three times)
static inline void lmd_wscale(double Ustar, double sigma, double Bf, double& wm, double& ws) {
  const double r3 = 1.0 / 3.0, small = 1.0e-20;
  const double Ustar3 = Ustar * Ustar * Ustar;
  const double zetahat = vonKar * sigma * Bf;
  const double zetapar = zetahat / (Ustar3 + small);
  if (zetahat >= 0.0) {
    wm = vonKar * Ustar / (1.0 + 5.0 * zetapar);
    ws = wm;
  } else {
    if (zetapar > lmd_zetam) wm = vonKar * Ustar * std::pow(1.0 - 16.0 * zetapar, 0.25);
    else wm = vonKar * std::pow(lmd_am * Ustar3 - lmd_cm * zetahat, r3);
    if (zetapar > lmd_zetas) ws = vonKar * Ustar * std::pow(1.0 - 16.0 * zetapar, 0.5);
    else ws = vonKar * std::pow(lmd_as * Ustar3 - lmd_cs * zetahat, r3);
  }
}

// lmd_vmix_tile (lmd_vmix.F:100-432): LMD_RIMIX with RI_SPLINES (no RI_HORAVG / RI_VERAVG, no LMD_DDMIX).  The spline of
// the density difference (dR, :205-206) never reaches a result (Rig uses bvf, :233) and is left out.
static void lmd_vmix_interior(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, nstp = m.nstp, Lm = c.Lm, Mm = c.Mm;
  const double eps = 1.0e-14;
  F3 u = m.u[nstp], v = m.v[nstp];
  S3 Rig(IminS, ImaxS, JminS, JmaxS, 0, N);
  SK FC(IminS, ImaxS, 0, N), dU(IminS, ImaxS, 0, N), dV(IminS, ImaxS, 0, N);
  const int i0 = std::max(1, Istr - 1), i1 = std::min(Iend + 1, Lm);
  for (int j = std::max(1, Jstr - 1); j <= std::min(Jend + 1, Mm); ++j) {
    for (int i = i0; i <= i1; ++i) { FC(i, 0) = 0.0; dU(i, 0) = 0.0; dV(i, 0) = 0.0; }
    for (int k = 1; k <= N - 1; ++k)
      for (int i = i0; i <= i1; ++i) {
        const double cff = 1.0 / (2.0 * m.Hz(i, j, k + 1) + m.Hz(i, j, k) * (2.0 - FC(i, k - 1)));
        FC(i, k) = cff * m.Hz(i, j, k + 1);
        dU(i, k) = cff * (3.0 * (u(i, j, k + 1) - u(i, j, k) + u(i + 1, j, k + 1) - u(i + 1, j, k)) - m.Hz(i, j, k) * dU(i, k - 1));
        dV(i, k) = cff * (3.0 * (v(i, j, k + 1) - v(i, j, k) + v(i, j + 1, k + 1) - v(i, j + 1, k)) - m.Hz(i, j, k) * dV(i, k - 1));
      }
    for (int i = i0; i <= i1; ++i) { dU(i, N) = 0.0; dV(i, N) = 0.0; }
    for (int k = N - 1; k >= 1; --k)
      for (int i = i0; i <= i1; ++i) {
        dU(i, k) = dU(i, k) - FC(i, k) * dU(i, k + 1);
        dV(i, k) = dV(i, k) - FC(i, k) * dV(i, k + 1);
      }
    for (int k = 1; k <= N - 1; ++k)
      for (int i = i0; i <= i1; ++i) {
        const double shear2 = dU(i, k) * dU(i, k) + dV(i, k) * dV(i, k);
        Rig(i, j, k) = m.bvf(i, j, k) / (shear2 + eps);
      }
  }
  // :309-347
  for (int k = 1; k <= N - 1; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = std::min(1.0, std::max(0.0, Rig(i, j, k)) / lmd_Ri0);
        double nu_sx = 1.0 - cff * cff;
        nu_sx = nu_sx * nu_sx * nu_sx;
        const double shear2 = m.bvf(i, j, k) / (Rig(i, j, k) + eps);
        cff = shear2 * shear2 / (shear2 * shear2 + 16.0e-10);
        nu_sx = cff * nu_sx;
        cff = 1.0 / std::sqrt(std::max(m.bvf(i, j, k), 1.0e-7));
        const double lmd_iwm = 1.0e-6 * cff, lmd_iws = 1.0e-7 * cff;
        m.Akv(i, j, k) = lmd_iwm + lmd_nu0m * nu_sx;
        m.Akt[c.itemp - 1](i, j, k) = lmd_iws + lmd_nu0s * nu_sx;
        if (c.salinity) m.Akt[c.isalt - 1](i, j, k) = m.Akt[c.itemp - 1](i, j, k);
      }
}

// lmd_skpp_tile (lmd_skpp.F:246-928): SASHA criterion, RI_SPLINES, LMD_NONLOCAL, SALINITY; no LMD_SHAPIRO, LMD_BOUND, MASKING
static void lmd_skpp(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, nstp = m.nstp, itemp = c.itemp - 1, isalt = c.isalt - 1;
  const double g = c.g, gorho0 = c.g / c.rho0;
  const double eps = 1.0e-10;
  const double lmd_Cg = lmd_Cstar * vonKar * std::pow(lmd_cs * vonKar * lmd_epsilon, 1.0 / 3.0);          // mod_scalars.F:4330
  F3 u = m.u[nstp], v = m.v[nstp];
  const double Vtc = lmd_Cv * std::sqrt(-lmd_betaT) / (std::sqrt(lmd_cs * lmd_epsilon) * lmd_Ric * vonKar * vonKar);      // :246
  S2 sl_dpth(IminS, ImaxS, JminS, JmaxS), Ustar(IminS, ImaxS, JminS, JmaxS), Bo(IminS, ImaxS, JminS, JmaxS), Bosol(IminS, ImaxS, JminS, JmaxS),
      Bfsfc(IminS, ImaxS, JminS, JmaxS), Gm1(IminS, ImaxS, JminS, JmaxS), Gt1(IminS, ImaxS, JminS, JmaxS), Gs1(IminS, ImaxS, JminS, JmaxS),
      dGm1dS(IminS, ImaxS, JminS, JmaxS), dGt1dS(IminS, ImaxS, JminS, JmaxS), dGs1dS(IminS, ImaxS, JminS, JmaxS), f1(IminS, ImaxS, JminS, JmaxS),
      wm(IminS, ImaxS, JminS, JmaxS), ws(IminS, ImaxS, JminS, JmaxS);
  S3 Bflux(IminS, ImaxS, JminS, JmaxS, 0, N);
  SK FC(IminS, ImaxS, 0, N), dR(IminS, ImaxS, 0, N), dU(IminS, ImaxS, 0, N), dV(IminS, ImaxS, 0, N);
  std::vector<double> Rref_(ImaxS - IminS + 1), Uref_(ImaxS - IminS + 1), Vref_(ImaxS - IminS + 1);
  auto Rref = [&](int i) -> double& { return Rref_[i - IminS]; };
  auto Uref = [&](int i) -> double& { return Uref_[i - IminS]; };
  auto Vref = [&](int i) -> double& { return Vref_[i - IminS]; };
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      sl_dpth(i, j) = lmd_epsilon * (m.z_w(i, j, N) - m.hsbl(i, j));                                         // :256
      const double a = 0.5 * (m.sustr(i, j) + m.sustr(i + 1, j)), bb = 0.5 * (m.svstr(i, j) + m.svstr(i, j + 1));
      Ustar(i, j) = std::sqrt(std::sqrt(a * a + bb * bb));                                                   // :267-268
      if (c.salinity) Bo(i, j) = g * (m.alpha(i, j) * (m.stflx[itemp](i, j) - m.srflx(i, j)) - m.beta(i, j) * m.stflx[isalt](i, j));   // :286-287
      else Bo(i, j) = g * m.alpha(i, j) * (m.stflx[itemp](i, j) - m.srflx(i, j));
      Bosol(i, j) = g * m.alpha(i, j) * m.srflx(i, j);
    }
  // :302-328
  for (int k = 0; k <= N; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        const double zgrid = m.z_w(i, j, N) - m.z_w(i, j, k);
        const double swdk = lmd_swfrac(-1.0, zgrid, (int)m.Jwtype(i, j));
        Bflux(i, j, k) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk));
        const double cff = 1.0 - (0.5 + std::copysign(0.5, Bflux(i, j, k)));
        m.ghats[itemp](i, j, k) = -cff * (m.stflx[itemp](i, j) - m.srflx(i, j) + m.srflx(i, j) * (1.0 - swdk));
        if (c.salinity) m.ghats[isalt](i, j, k) = cff * m.stflx[isalt](i, j);
      }
  // :336-545
  for (int j = Jstr; j <= Jend; ++j) {
    for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; dR(i, 0) = 0.0; dU(i, 0) = 0.0; dV(i, 0) = 0.0; }
    for (int k = 1; k <= N - 1; ++k)
      for (int i = Istr; i <= Iend; ++i) {
        const double cff = 1.0 / (2.0 * m.Hz(i, j, k + 1) + m.Hz(i, j, k) * (2.0 - FC(i, k - 1)));
        FC(i, k) = cff * m.Hz(i, j, k + 1);
        dR(i, k) = cff * (6.0 * (m.pden(i, j, k + 1) - m.pden(i, j, k)) - m.Hz(i, j, k) * dR(i, k - 1));
        dU(i, k) = cff * (3.0 * (u(i, j, k + 1) - u(i, j, k) + u(i + 1, j, k + 1) - u(i + 1, j, k)) - m.Hz(i, j, k) * dU(i, k - 1));
        dV(i, k) = cff * (3.0 * (v(i, j, k + 1) - v(i, j, k) + v(i, j + 1, k + 1) - v(i, j + 1, k)) - m.Hz(i, j, k) * dV(i, k - 1));
      }
    for (int i = Istr; i <= Iend; ++i) { dR(i, N) = 0.0; dU(i, N) = 0.0; dV(i, N) = 0.0; }
    for (int k = N - 1; k >= 1; --k)
      for (int i = Istr; i <= Iend; ++i) {
        dR(i, k) = dR(i, k) - FC(i, k) * dR(i, k + 1);
        dU(i, k) = dU(i, k) - FC(i, k) * dU(i, k + 1);
        dV(i, k) = dV(i, k) - FC(i, k) * dV(i, k + 1);
      }
    const double cff1 = 1.0 / 3.0, cff2 = 1.0 / 6.0;
    for (int i = Istr; i <= Iend; ++i) {
      Rref(i) = m.pden(i, j, N) + m.Hz(i, j, N) * (cff1 * dR(i, N) + cff2 * dR(i, N - 1));
      Uref(i) = 0.5 * (u(i, j, N) + u(i + 1, j, N)) + m.Hz(i, j, N) * (cff1 * dU(i, N) + cff2 * dU(i, N - 1));
      Vref(i) = 0.5 * (v(i, j, N) + v(i, j + 1, N)) + m.Hz(i, j, N) * (cff1 * dV(i, N) + cff2 * dV(i, N - 1));
    }
    for (int i = Istr; i <= Iend; ++i) {
      FC(i, N) = 0.0;
      for (int k = N; k >= 1; --k) {
        const double depth = m.z_w(i, j, N) - m.z_w(i, j, k - 1);
        double sigma;
        if (Bflux(i, j, k - 1) < 0.0) sigma = std::min(sl_dpth(i, j), depth);
        else sigma = depth;
        lmd_wscale(Ustar(i, j), sigma, Bflux(i, j, k - 1), wm(i, j), ws(i, j));
        const double Rk = m.pden(i, j, k) - m.Hz(i, j, k) * (cff1 * dR(i, k - 1) + cff2 * dR(i, k));
        const double Uk = 0.5 * (u(i, j, k) + u(i + 1, j, k)) - m.Hz(i, j, k) * (cff1 * dU(i, k - 1) + cff2 * dU(i, k));
        const double Vk = 0.5 * (v(i, j, k) + v(i, j + 1, k)) - m.Hz(i, j, k) * (cff1 * dV(i, k - 1) + cff2 * dV(i, k));
        const double Ritop = -gorho0 * (Rref(i) - Rk) * depth;
        const double du = Uref(i) - Uk, dv = Vref(i) - Vk;
        const double Ribot = du * du + dv * dv + Vtc * depth * ws(i, j) * std::sqrt(std::fabs(m.bvf(i, j, k - 1)));
        FC(i, k - 1) = Ritop - lmd_Ric * Ribot;                                                              // SASHA
      }
    }
    for (int i = Istr; i <= Iend; ++i) { m.ksbl(i, j) = 1.0; m.hsbl(i, j) = m.z_w(i, j, 1); }
    for (int k = N; k >= 2; --k)
      for (int i = Istr; i <= Iend; ++i)
        if (m.ksbl(i, j) == 1.0 && FC(i, k - 1) > 0.0) {
          m.hsbl(i, j) = (m.z_w(i, j, k) * FC(i, k - 1) - m.z_w(i, j, k - 1) * FC(i, k)) / (FC(i, k - 1) - FC(i, k));
          m.ksbl(i, j) = (double)k;
        }
  }
  // :551-589
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      const double zgrid = m.z_w(i, j, N) - m.hsbl(i, j);
      const double swdk = lmd_swfrac(-1.0, zgrid, (int)m.Jwtype(i, j));
      Bfsfc(i, j) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk));
    }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      if (Ustar(i, j) > 0.0 && Bfsfc(i, j) > 0.0) {
        const double hekman = lmd_cekman * Ustar(i, j) / std::max(std::fabs(m.f(i, j)), eps);
        const double hmonob = lmd_cmonob * Ustar(i, j) * Ustar(i, j) * Ustar(i, j) / std::max(vonKar * Bfsfc(i, j), eps);
        m.hsbl(i, j) = (m.z_w(i, j, N) - std::min(std::min(hekman, hmonob), m.z_w(i, j, N) - m.hsbl(i, j)));
      }
      m.hsbl(i, j) = std::min(m.hsbl(i, j), m.z_w(i, j, N));
      m.hsbl(i, j) = std::max(m.hsbl(i, j), m.z_w(i, j, 0));
    }
  bc_r2d(m, b, m.hsbl);                                                                                     // :636-647
  // :651-687
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      m.ksbl(i, j) = 1.0;
      for (int k = N; k >= 2; --k)
        if (m.ksbl(i, j) == 1.0 && m.z_w(i, j, k - 1) < m.hsbl(i, j)) m.ksbl(i, j) = (double)k;
    }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      const double zgrid = m.z_w(i, j, N) - m.hsbl(i, j);
      const double swdk = lmd_swfrac(-1.0, zgrid, (int)m.Jwtype(i, j));
      Bfsfc(i, j) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk));
    }
  // :697-737
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      sl_dpth(i, j) = lmd_epsilon * (m.z_w(i, j, N) - m.hsbl(i, j));
      const double cff = (Bfsfc(i, j) > 0.0) ? 1.0 : lmd_epsilon;
      const double sigma = cff * (m.z_w(i, j, N) - m.hsbl(i, j));
      lmd_wscale(Ustar(i, j), sigma, Bfsfc(i, j), wm(i, j), ws(i, j));
    }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i)
      f1(i, j) = 5.0 * std::max(0.0, Bfsfc(i, j)) * vonKar / (Ustar(i, j) * Ustar(i, j) * Ustar(i, j) * Ustar(i, j) + eps);
  // :739-818
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      const double zbl = m.z_w(i, j, N) - m.hsbl(i, j);
      if (m.hsbl(i, j) > m.z_w(i, j, 1)) {
        const int k = (int)m.ksbl(i, j);
        const double cff = 1.0 / (m.z_w(i, j, k) - m.z_w(i, j, k - 1));
        const double cff_dn = cff * (m.hsbl(i, j) - m.z_w(i, j, k - 1));
        const double cff_up = cff * (m.z_w(i, j, k) - m.hsbl(i, j));
        double K_bl = cff_dn * m.Akv(i, j, k) + cff_up * m.Akv(i, j, k - 1);
        double dK_bl = cff * (m.Akv(i, j, k) - m.Akv(i, j, k - 1));
        Gm1(i, j) = K_bl / (zbl * wm(i, j) + eps);
        dGm1dS(i, j) = std::min(0.0, -dK_bl / (wm(i, j) + eps) - K_bl * f1(i, j));
        K_bl = cff_dn * m.Akt[itemp](i, j, k) + cff_up * m.Akt[itemp](i, j, k - 1);
        dK_bl = cff * (m.Akt[itemp](i, j, k) - m.Akt[itemp](i, j, k - 1));
        Gt1(i, j) = K_bl / (zbl * ws(i, j) + eps);
        dGt1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
        if (c.salinity) {
          K_bl = cff_dn * m.Akt[isalt](i, j, k) + cff_up * m.Akt[isalt](i, j, k - 1);
          dK_bl = cff * (m.Akt[isalt](i, j, k) - m.Akt[isalt](i, j, k - 1));
          Gs1(i, j) = K_bl / (zbl * ws(i, j) + eps);
          dGs1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
        }
      } else {
        m.ksbl(i, j) = 0.0;
        const double a = 0.5 * (m.bustr(i, j) + m.bustr(i + 1, j)), bb = 0.5 * (m.bvstr(i, j) + m.bvstr(i, j + 1));
        const double Ustarb = std::sqrt(std::sqrt(a * a + bb * bb));
        const double dK_bl = vonKar * Ustarb;
        const double K_bl = dK_bl * (m.hsbl(i, j) - m.z_w(i, j, 0));
        Gm1(i, j) = K_bl / (zbl * wm(i, j) + eps);
        dGm1dS(i, j) = std::min(0.0, -dK_bl / (wm(i, j) + eps) - K_bl * f1(i, j));
        Gt1(i, j) = K_bl / (zbl * ws(i, j) + eps);
        dGt1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
        if (c.salinity) { Gs1(i, j) = Gt1(i, j); dGs1dS(i, j) = dGt1dS(i, j); }
      }
    }
  // :826-923
  for (int k = 1; k <= N - 1; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        const double zbl = m.z_w(i, j, N) - m.hsbl(i, j);
        if ((double)k > m.ksbl(i, j)) {
          const double depth = m.z_w(i, j, N) - m.z_w(i, j, k);
          double sigma;
          if (Bflux(i, j, k) < 0.0) sigma = std::min(sl_dpth(i, j), depth);
          else sigma = depth;
          lmd_wscale(Ustar(i, j), sigma, Bflux(i, j, k), wm(i, j), ws(i, j));
          sigma = depth / (zbl + eps);
          const double a1 = sigma - 2.0, a2 = 3.0 - 2.0 * sigma, a3 = sigma - 1.0;
          const double Gm = a1 + a2 * Gm1(i, j) + a3 * dGm1dS(i, j);
          const double Gt = a1 + a2 * Gt1(i, j) + a3 * dGt1dS(i, j);
          m.Akv(i, j, k) = depth * wm(i, j) * (1.0 + sigma * Gm);
          m.Akt[itemp](i, j, k) = depth * ws(i, j) * (1.0 + sigma * Gt);
          if (c.salinity) {
            const double Gs = a1 + a2 * Gs1(i, j) + a3 * dGs1dS(i, j);
            m.Akt[isalt](i, j, k) = depth * ws(i, j) * (1.0 + sigma * Gs);
          }
          const double cff = lmd_Cg * (1.0 - (0.5 + std::copysign(0.5, Bflux(i, j, k)))) / (zbl * ws(i, j) + eps);
          m.ghats[itemp](i, j, k) = cff * m.ghats[itemp](i, j, k);
          if (c.salinity) m.ghats[isalt](i, j, k) = cff * m.ghats[isalt](i, j, k);
        } else {
          m.ghats[itemp](i, j, k) = 0.0;
          if (c.salinity) m.ghats[isalt](i, j, k) = 0.0;
        }
      }
}

// lmd_finish_tile (lmd_vmix.F:464-659): LMD_CONVEC, then the boundary copies (the edge switches do not depend on the
// periodicity, get_bounds.F:530-551; the eastern-edge copy writes column Iend-1 as the reference does, :568-575), bc_w3d.
static void lmd_finish(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, NAT = c.salinity ? 2 : 1;
  for (int k = 1; k <= N - 1; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = std::max(m.bvf(i, j, k), lmd_bvfcon);
        cff = std::min(1.0, (lmd_bvfcon - cff) / lmd_bvfcon);
        double nu_sxc = 1.0 - cff * cff;
        nu_sxc = nu_sxc * nu_sxc * nu_sxc;
        m.Akv(i, j, k) = m.Akv(i, j, k) + lmd_nu0c * nu_sxc;
        m.Akt[c.itemp - 1](i, j, k) = m.Akt[c.itemp - 1](i, j, k) + lmd_nu0c * nu_sxc;
        if (c.salinity) m.Akt[c.isalt - 1](i, j, k) = m.Akt[c.isalt - 1](i, j, k) + lmd_nu0c * nu_sxc;
      }
  for (int k = 0; k <= N; ++k) {
    if (b.Western_Edge)
      for (int j = Jstr; j <= Jend; ++j) {
        for (int it = 0; it < NAT; ++it) m.Akt[it](Istr - 1, j, k) = m.Akt[it](Istr, j, k);
        m.Akv(Istr - 1, j, k) = m.Akv(Istr, j, k);
      }
    if (b.Eastern_Edge)
      for (int j = Jstr; j <= Jend; ++j) {
        for (int it = 0; it < NAT; ++it) m.Akt[it](Iend - 1, j, k) = m.Akt[it](Iend, j, k);
        m.Akv(Iend - 1, j, k) = m.Akv(Iend, j, k);
      }
    if (b.Southern_Edge)
      for (int i = Istr; i <= Iend; ++i) {
        for (int it = 0; it < NAT; ++it) m.Akt[it](i, Jstr - 1, k) = m.Akt[it](i, Jstr, k);
        m.Akv(i, Jstr - 1, k) = m.Akv(i, Jstr, k);
      }
    if (b.Northern_Edge)
      for (int i = Istr; i <= Iend; ++i) {
        for (int it = 0; it < NAT; ++it) m.Akt[it](i, Jend + 1, k) = m.Akt[it](i, Jend, k);
        m.Akv(i, Jend + 1, k) = m.Akv(i, Jend, k);
      }
    if (b.SouthWest_Corner) {
      for (int it = 0; it < NAT; ++it) m.Akt[it](Istr - 1, Jstr - 1, k) = 0.5 * (m.Akt[it](Istr, Jstr - 1, k) + m.Akt[it](Istr - 1, Jstr, k));
      m.Akv(Istr - 1, Jstr - 1, k) = 0.5 * (m.Akv(Istr, Jstr - 1, k) + m.Akv(Istr - 1, Jstr, k));
    }
    if (b.SouthEast_Corner) {
      for (int it = 0; it < NAT; ++it) m.Akt[it](Iend + 1, Jstr - 1, k) = 0.5 * (m.Akt[it](Iend, Jstr - 1, k) + m.Akt[it](Iend + 1, Jstr, k));
      m.Akv(Iend + 1, Jstr - 1, k) = 0.5 * (m.Akv(Iend, Jstr - 1, k) + m.Akv(Iend + 1, Jstr, k));
    }
    if (b.NorthWest_Corner) {
      for (int it = 0; it < NAT; ++it) m.Akt[it](Istr - 1, Jend + 1, k) = 0.5 * (m.Akt[it](Istr, Jend + 1, k) + m.Akt[it](Istr - 1, Jend, k));
      m.Akv(Istr - 1, Jend + 1, k) = 0.5 * (m.Akv(Istr, Jend + 1, k) + m.Akv(Istr - 1, Jend, k));
    }
    if (b.NorthEast_Corner) {
      for (int it = 0; it < NAT; ++it) m.Akt[it](Iend + 1, Jend + 1, k) = 0.5 * (m.Akt[it](Iend, Jend + 1, k) + m.Akt[it](Iend + 1, Jend, k));
      m.Akv(Iend + 1, Jend + 1, k) = 0.5 * (m.Akv(Iend, Jend + 1, k) + m.Akv(Iend + 1, Jend, k));
    }
  }
}

// bvf_mix_tile (bvf_mix.F:92-127; constants mod_scalars.F:1793-1796): diffusivity ~ 1/N, convective value where unstable
void bvf_mix(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, itemp = c.itemp - 1, isalt = c.isalt - 1;
  const double bvf_numax = 4.0e-4, bvf_numin = 3.0e-5, bvf_nu0 = 1.0e-7, bvf_nu0c = 1.0;
  for (int k = 1; k <= N - 1; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        m.Akv(i, j, k) = c.Akv_bak;
        if (m.bvf(i, j, k) < 0.0) {
          m.Akv(i, j, k) = bvf_nu0c;
          m.Akt[itemp](i, j, k) = bvf_nu0c;
          if (c.salinity) m.Akt[isalt](i, j, k) = bvf_nu0c;
        } else if (m.bvf(i, j, k) == 0.0) {
          m.Akv(i, j, k) = c.Akv_bak;
          m.Akt[itemp](i, j, k) = c.Akt_bak[itemp];
          if (c.salinity) m.Akt[isalt](i, j, k) = c.Akt_bak[isalt];
        } else {
          const double cff = bvf_nu0 / std::sqrt(m.bvf(i, j, k));
          m.Akt[itemp](i, j, k) = std::min(bvf_numax, std::max(bvf_numin, cff));
          m.Akv(i, j, k) = m.Akt[itemp](i, j, k);
          if (c.salinity) m.Akt[isalt](i, j, k) = m.Akt[itemp](i, j, k);
        }
      }
  exchange_w3d(m, b, m.Akv);
  for (int it = 0; it < (c.salinity ? 2 : 1); ++it) exchange_w3d(m, b, m.Akt[it]);
}

// point functions for the known-answer tests (tests/test_oracle_cpu.py)
void physics_point(int which, const double* in, double* out) {
  if (which == 0) { out[0] = bulk_psiu(in[0]); out[1] = bulk_psit(in[0]); }                                 // stability functions at Z/L
  else if (which == 1) out[0] = lmd_swfrac(-1.0, in[0], (int)in[1]);                                       // shortwave fraction at depth Z
  else if (which == 2) lmd_wscale(in[0], in[1], in[2], out[0], out[1]);                                    // wm, ws (Ustar, sigma, Bflux)
}

// lmd_vmix (lmd_vmix.F:33-96): interior scheme, surface boundary layer, convective adjustment + boundary copies ...
void lmd_vmix(Model& m, const Bnd& b) {
  lmd_vmix_interior(m, b);
  lmd_skpp(m, b);
  lmd_finish(m, b);
}
// ... and the end of lmd_finish_tile (lmd_vmix.F:636-655): bc_w3d = closed-wall copies + periodic / halo exchange.  A separate
// stage, run after every tile has done lmd_vmix: the western-edge copy of one tile and the periodic copy of the eastern tile
// write the same ghost column (0), and step3d_uv reads it, so in the reference's shared-memory mode the result depends on the
// order of the tiles; the serial and the distributed-memory runs have the exchange last, which is what this ordering gives.
void lmd_vmix_bc(Model& m, const Bnd& b) {
  const int NAT = m.c.salinity ? 2 : 1;
  bc_w3d(m, b, m.Akv);
  for (int it = 0; it < NAT; ++it) bc_w3d(m, b, m.Akt[it]);
}

}  // namespace orc

// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// Baroclinic pressure gradient: ROMS/Nonlinear/prsgrd32.h:106-421 (DJ_GRADPS, spline density Jacobian) and
// ROMS/Nonlinear/prsgrd31.h:97-362 (standard density Jacobian, RHO_SURF on: globaldefs.h:130).
// WJ_GRADP variant of prsgrd31 (:236-254, :317-335); ROMS/Nonlinear/prsgrd40.h:176-270 (PJ_GRADP, finite-volume pressure Jacobian).
// Dispatch: ROMS/Nonlinear/prsgrd.F:16-26.  ru,rv(:,:,1:N,nrhs) are overwritten.
#include "roms_oracle.hpp"

namespace orc {

static void prsgrd32(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double g = c.g, rho0 = c.rho0;
  F3 &rho = m.rho, &z_r = m.z_r, &z_w = m.z_w, &Hz = m.Hz; F3 ru = m.ru[m.nrhs], rv = m.rv[m.nrhs];
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  const double GRho = g / rho0, HalfGRho = 0.5 * GRho;
  S3 P(IminS, ImaxS, JminS, JmaxS, 1, N);
  SK dR(IminS, ImaxS, 0, N), dZ(IminS, ImaxS, 0, N);
  S2 FC(IminS, ImaxS, JminS, JmaxS), aux(IminS, ImaxS, JminS, JmaxS), dRx(IminS, ImaxS, JminS, JmaxS), dZx(IminS, ImaxS, JminS, JmaxS);
  // :236-290
  for (int j = JstrV - 1; j <= Jend; ++j) {
    for (int k = 1; k <= N - 1; ++k)
      for (int i = IstrU - 1; i <= Iend; ++i) { dR(i, k) = rho(i, j, k + 1) - rho(i, j, k); dZ(i, k) = z_r(i, j, k + 1) - z_r(i, j, k); }
    for (int i = IstrU - 1; i <= Iend; ++i) { dR(i, N) = dR(i, N - 1); dZ(i, N) = dZ(i, N - 1); dR(i, 0) = dR(i, 1); dZ(i, 0) = dZ(i, 1); }
    for (int k = N; k >= 1; --k)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff = 2.0 * dR(i, k) * dR(i, k - 1);
        if (cff > eps) dR(i, k) = cff / (dR(i, k) + dR(i, k - 1)); else dR(i, k) = 0.0;
        dZ(i, k) = 2.0 * dZ(i, k) * dZ(i, k - 1) / (dZ(i, k) + dZ(i, k - 1));
      }
    for (int i = IstrU - 1; i <= Iend; ++i) {
      double cff1 = 1.0 / (z_r(i, j, N) - z_r(i, j, N - 1));
      double cff2 = 0.5 * (rho(i, j, N) - rho(i, j, N - 1)) * (z_w(i, j, N) - z_r(i, j, N)) * cff1;
      if (c.atm_press)                                                          // ATM_PRESS :229-232, :265-267
        P(i, j, N) = g * z_w(i, j, N) + (100.0 / rho0) * (m.Pair(i, j) - 1013.25) + GRho * (rho(i, j, N) + cff2) * (z_w(i, j, N) - z_r(i, j, N));
      else
      P(i, j, N) = g * z_w(i, j, N) + GRho * (rho(i, j, N) + cff2) * (z_w(i, j, N) - z_r(i, j, N));
    }
    for (int k = N - 1; k >= 1; --k)
      for (int i = IstrU - 1; i <= Iend; ++i)
        P(i, j, k) = P(i, j, k + 1) +
                     HalfGRho * ((rho(i, j, k + 1) + rho(i, j, k)) * (z_r(i, j, k + 1) - z_r(i, j, k)) -
                                 OneFifth * ((dR(i, k + 1) - dR(i, k)) * (z_r(i, j, k + 1) - z_r(i, j, k) - OneTwelfth * (dZ(i, k + 1) + dZ(i, k))) -
                                             (dZ(i, k + 1) - dZ(i, k)) * (rho(i, j, k + 1) - rho(i, j, k) - OneTwelfth * (dR(i, k + 1) + dR(i, k)))));
  }
  // :296-354  XI-component
  for (int k = N; k >= 1; --k) {
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend + 1; ++i) { aux(i, j) = z_r(i, j, k) - z_r(i - 1, j, k); FC(i, j) = rho(i, j, k) - rho(i - 1, j, k); }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff = 2.0 * aux(i, j) * aux(i + 1, j);
        if (cff > eps) { double cff1 = 1.0 / (aux(i, j) + aux(i + 1, j)); dZx(i, j) = cff * cff1; } else dZx(i, j) = 0.0;
        double cff1 = 2.0 * FC(i, j) * FC(i + 1, j);
        if (cff1 > eps) { double cff2 = 1.0 / (FC(i, j) + FC(i + 1, j)); dRx(i, j) = cff1 * cff2; } else dRx(i, j) = 0.0;
      }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i)
        ru(i, j, k) = m.on_u(i, j) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k)) *
                      (P(i - 1, j, k) - P(i, j, k) -
                       HalfGRho * ((rho(i, j, k) + rho(i - 1, j, k)) * (z_r(i, j, k) - z_r(i - 1, j, k)) -
                                   OneFifth * ((dRx(i, j) - dRx(i - 1, j)) * (z_r(i, j, k) - z_r(i - 1, j, k) - OneTwelfth * (dZx(i, j) + dZx(i - 1, j))) -
                                               (dZx(i, j) - dZx(i - 1, j)) * (rho(i, j, k) - rho(i - 1, j, k) - OneTwelfth * (dRx(i, j) + dRx(i - 1, j))))));
  }
  // :360-418  ETA-component
  for (int k = N; k >= 1; --k) {
    for (int j = JstrV - 1; j <= Jend + 1; ++j)
      for (int i = Istr; i <= Iend; ++i) { aux(i, j) = z_r(i, j, k) - z_r(i, j - 1, k); FC(i, j) = rho(i, j, k) - rho(i, j - 1, k); }
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = 2.0 * aux(i, j) * aux(i, j + 1);
        if (cff > eps) { double cff1 = 1.0 / (aux(i, j) + aux(i, j + 1)); dZx(i, j) = cff * cff1; } else dZx(i, j) = 0.0;
        double cff1 = 2.0 * FC(i, j) * FC(i, j + 1);
        if (cff1 > eps) { double cff2 = 1.0 / (FC(i, j) + FC(i, j + 1)); dRx(i, j) = cff1 * cff2; } else dRx(i, j) = 0.0;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i)
        rv(i, j, k) = m.om_v(i, j) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k)) *
                      (P(i, j - 1, k) - P(i, j, k) -
                       HalfGRho * ((rho(i, j, k) + rho(i, j - 1, k)) * (z_r(i, j, k) - z_r(i, j - 1, k)) -
                                   OneFifth * ((dRx(i, j) - dRx(i, j - 1)) * (z_r(i, j, k) - z_r(i, j - 1, k) - OneTwelfth * (dZx(i, j) + dZx(i, j - 1))) -
                                               (dZx(i, j) - dZx(i, j - 1)) * (rho(i, j, k) - rho(i, j - 1, k) - OneTwelfth * (dRx(i, j) + dRx(i, j - 1))))));
  }
}

static void prsgrd31(Model& m, const Bnd& b, bool wj) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double g = c.g, rho0 = c.rho0;
  F3 &rho = m.rho, &z_r = m.z_r, &z_w = m.z_w, &Hz = m.Hz; F3 ru = m.ru[m.nrhs], rv = m.rv[m.nrhs];
  const double fac1 = 0.5 * g / rho0, fac2 = 1000.0 * g / rho0, fac3 = 0.25 * g / rho0;
  std::vector<double> phie_v(ImaxS - IminS + 1), phix_v(ImaxS - IminS + 1);
  double* phie = phie_v.data() - IminS; double* phix = phix_v.data() - IminS;
  for (int j = Jstr; j <= Jend; ++j) {
    for (int i = IstrU; i <= Iend; ++i) {
      double cff1 = z_w(i, j, N) - z_r(i, j, N) + z_w(i - 1, j, N) - z_r(i - 1, j, N);
      phix[i] = fac1 * (rho(i, j, N) - rho(i - 1, j, N)) * cff1;
      if (c.atm_press) phix[i] = phix[i] + (100.0 / rho0) * (m.Pair(i, j) - m.Pair(i - 1, j));        // ATM_PRESS prsgrd31.h:196-198, :213-215
      phix[i] = phix[i] + (fac2 + fac1 * (rho(i, j, N) + rho(i - 1, j, N))) * (z_w(i, j, N) - z_w(i - 1, j, N));
      ru(i, j, N) = -0.5 * (Hz(i, j, N) + Hz(i - 1, j, N)) * phix[i] * m.on_u(i, j);
    }
    for (int k = N - 1; k >= 1; --k)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff1, cff2, cff3, cff4;
        if (wj) {                                                                   // WJ_GRADP :236-254
          cff1 = 1.0 / ((z_r(i, j, k + 1) - z_r(i, j, k)) * (z_r(i - 1, j, k + 1) - z_r(i - 1, j, k)));
          cff2 = z_r(i, j, k) - z_r(i - 1, j, k) + z_r(i, j, k + 1) - z_r(i - 1, j, k + 1);
          cff3 = z_r(i, j, k + 1) - z_r(i, j, k) - z_r(i - 1, j, k + 1) + z_r(i - 1, j, k);
          const double gamma = 0.125 * cff1 * cff2 * cff3;
          cff1 = (1.0 + gamma) * (rho(i, j, k + 1) - rho(i - 1, j, k + 1)) + (1.0 - gamma) * (rho(i, j, k) - rho(i - 1, j, k));
          cff2 = rho(i, j, k + 1) + rho(i - 1, j, k + 1) - rho(i, j, k) - rho(i - 1, j, k);
          cff3 = z_r(i, j, k + 1) + z_r(i - 1, j, k + 1) - z_r(i, j, k) - z_r(i - 1, j, k);
          cff4 = (1.0 + gamma) * (z_r(i, j, k + 1) - z_r(i - 1, j, k + 1)) + (1.0 - gamma) * (z_r(i, j, k) - z_r(i - 1, j, k));
        } else {
        cff1 = rho(i, j, k + 1) - rho(i - 1, j, k + 1) + rho(i, j, k) - rho(i - 1, j, k);
        cff2 = rho(i, j, k + 1) + rho(i - 1, j, k + 1) - rho(i, j, k) - rho(i - 1, j, k);
        cff3 = z_r(i, j, k + 1) + z_r(i - 1, j, k + 1) - z_r(i, j, k) - z_r(i - 1, j, k);
        cff4 = z_r(i, j, k + 1) - z_r(i - 1, j, k + 1) + z_r(i, j, k) - z_r(i - 1, j, k);
        }
        phix[i] = phix[i] + fac3 * (cff1 * cff3 - cff2 * cff4);
        ru(i, j, k) = -0.5 * (Hz(i, j, k) + Hz(i - 1, j, k)) * phix[i] * m.on_u(i, j);
      }
    if (j >= JstrV) {
      for (int i = Istr; i <= Iend; ++i) {
        double cff1 = z_w(i, j, N) - z_r(i, j, N) + z_w(i, j - 1, N) - z_r(i, j - 1, N);
        phie[i] = fac1 * (rho(i, j, N) - rho(i, j - 1, N)) * cff1;
        if (c.atm_press) phie[i] = phie[i] + (100.0 / rho0) * (m.Pair(i, j) - m.Pair(i, j - 1));
        phie[i] = phie[i] + (fac2 + fac1 * (rho(i, j, N) + rho(i, j - 1, N))) * (z_w(i, j, N) - z_w(i, j - 1, N));
        rv(i, j, N) = -0.5 * (Hz(i, j, N) + Hz(i, j - 1, N)) * phie[i] * m.om_v(i, j);
      }
      for (int k = N - 1; k >= 1; --k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff1, cff2, cff3, cff4;
          if (wj) {                                                                 // WJ_GRADP :317-335
            cff1 = 1.0 / ((z_r(i, j, k + 1) - z_r(i, j, k)) * (z_r(i, j - 1, k + 1) - z_r(i, j - 1, k)));
            cff2 = z_r(i, j, k) - z_r(i, j - 1, k) + z_r(i, j, k + 1) - z_r(i, j - 1, k + 1);
            cff3 = z_r(i, j, k + 1) - z_r(i, j, k) - z_r(i, j - 1, k + 1) + z_r(i, j - 1, k);
            const double gamma = 0.125 * cff1 * cff2 * cff3;
            cff1 = (1.0 + gamma) * (rho(i, j, k + 1) - rho(i, j - 1, k + 1)) + (1.0 - gamma) * (rho(i, j, k) - rho(i, j - 1, k));
            cff2 = rho(i, j, k + 1) + rho(i, j - 1, k + 1) - rho(i, j, k) - rho(i, j - 1, k);
            cff3 = z_r(i, j, k + 1) + z_r(i, j - 1, k + 1) - z_r(i, j, k) - z_r(i, j - 1, k);
            cff4 = (1.0 + gamma) * (z_r(i, j, k + 1) - z_r(i, j - 1, k + 1)) + (1.0 - gamma) * (z_r(i, j, k) - z_r(i, j - 1, k));
          } else {
          cff1 = rho(i, j, k + 1) - rho(i, j - 1, k + 1) + rho(i, j, k) - rho(i, j - 1, k);
          cff2 = rho(i, j, k + 1) + rho(i, j - 1, k + 1) - rho(i, j, k) - rho(i, j - 1, k);
          cff3 = z_r(i, j, k + 1) + z_r(i, j - 1, k + 1) - z_r(i, j, k) - z_r(i, j - 1, k);
          cff4 = z_r(i, j, k + 1) - z_r(i, j - 1, k + 1) + z_r(i, j, k) - z_r(i, j - 1, k);
          }
          phie[i] = phie[i] + fac3 * (cff1 * cff3 - cff2 * cff4);
          rv(i, j, k) = -0.5 * (Hz(i, j, k) + Hz(i, j - 1, k)) * phie[i] * m.om_v(i, j);
        }
    }
  }
}

// prsgrd40_tile (ROMS/Nonlinear/prsgrd40.h:176-270; PJ_GRADP): finite-volume pressure Jacobian (Lin, 1997)
static void prsgrd40(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double g = c.g, rho0 = c.rho0;
  F3 &rho = m.rho, &z_w = m.z_w, &Hz = m.Hz; F3 ru = m.ru[m.nrhs], rv = m.rv[m.nrhs];
  SK FC(IminS, ImaxS, 0, N);
  S3 FX(IminS, ImaxS, JminS, JmaxS, 1, N), P(IminS, ImaxS, JminS, JmaxS, 0, N);
  for (int j = JstrV - 1; j <= Jend; ++j) {
    for (int i = IstrU - 1; i <= Iend; ++i) { P(i, j, N) = 0.0; if (c.atm_press) P(i, j, N) = P(i, j, N) + (100.0 / g) * (m.Pair(i, j) - 1013.25); }   // prsgrd40.h:169-172, :194-196
    for (int k = N; k >= 1; --k)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        P(i, j, k - 1) = P(i, j, k) + Hz(i, j, k) * rho(i, j, k);
        FX(i, j, k) = 0.5 * Hz(i, j, k) * (P(i, j, k) + P(i, j, k - 1));
      }
    if (j >= Jstr) {
      for (int i = IstrU; i <= Iend; ++i) FC(i, N) = 0.0;
      const double cff = 0.5 * g, cff1 = g / rho0;
      for (int k = N; k >= 1; --k)
        for (int i = IstrU; i <= Iend; ++i) {
          const double dh = z_w(i, j, k - 1) - z_w(i - 1, j, k - 1);
          FC(i, k - 1) = 0.5 * dh * (P(i, j, k - 1) + P(i - 1, j, k - 1));
          ru(i, j, k) = (cff * (Hz(i - 1, j, k) + Hz(i, j, k)) * (z_w(i - 1, j, N) - z_w(i, j, N)) +
                         cff1 * (FX(i - 1, j, k) - FX(i, j, k) + FC(i, k) - FC(i, k - 1))) * m.on_u(i, j);
        }
    }
    if (j >= JstrV) {
      for (int i = Istr; i <= Iend; ++i) FC(i, N) = 0.0;
      const double cff = 0.5 * g, cff1 = g / rho0;
      for (int k = N; k >= 1; --k)
        for (int i = Istr; i <= Iend; ++i) {
          const double dh = z_w(i, j, k - 1) - z_w(i, j - 1, k - 1);
          FC(i, k - 1) = 0.5 * dh * (P(i, j, k - 1) + P(i, j - 1, k - 1));
          rv(i, j, k) = (cff * (Hz(i, j - 1, k) + Hz(i, j, k)) * (z_w(i, j - 1, N) - z_w(i, j, N)) +
                         cff1 * (FX(i, j - 1, k) - FX(i, j, k) + FC(i, k) - FC(i, k - 1))) * m.om_v(i, j);
        }
    }
  }
}

// prsgrd.F:16-26.  dj_gradps: 0 prsgrd31, 1 DJ_GRADPS prsgrd32, 2 PJ_GRADP prsgrd40, 3 WJ_GRADP (prsgrd31 with the weighted Jacobian)
void prsgrd(Model& m, const Bnd& b) {
  switch (m.c.dj_gradps) {
    case 1: prsgrd32(m, b); break;
    case 2: prsgrd40(m, b); break;
    case 3: prsgrd31(m, b, true); break;
    default: prsgrd31(m, b, false);
  }
}

}  // namespace orc

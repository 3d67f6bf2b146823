// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/step3d_uv.F:111-1480 (step3d_uv_tile): corrector for u,v; SPLINES_VVISC implicit vertical viscosity
// (:344-396, :677-729); 2-D/3-D coupling (:469-605, :802-938); boundary conditions (:956-965); coupled mass fluxes
// and ubar/vbar reset (:1002-1432); periodic exchanges (:1438-1461).
#include "roms_oracle.hpp"

namespace orc {

void step3d_uv(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, Mm = c.Mm; const double dt = c.dt;
  const int nnew = m.nnew, nrhs = m.nrhs;
  const bool nospl = c.nospl_vvisc != 0;                          // SPLINES_VVISC undefined: the centred tridiagonal system (:397-462, :730-795)
  F3 &Hz = m.Hz, &Akv = m.Akv, &Huon = m.Huon, &Hvom = m.Hvom; F2 &pm = m.pm, &pn = m.pn;
  F3 u = m.u[nnew], v = m.v[nnew], ru = m.ru[nrhs], rv = m.rv[nrhs];
  SK AK(IminS, ImaxS, 0, N), BC(IminS, ImaxS, 0, N), CF(IminS, ImaxS, 0, N), DC(IminS, ImaxS, 0, N), FC(IminS, ImaxS, 0, N),
      Hzk(IminS, ImaxS, 1, N), oHz(IminS, ImaxS, 1, N);
  double cffAB;
  if (m.iic == m.ntfirst) cffAB = 0.25 * dt;
  else if (m.iic == m.ntfirst + 1) cffAB = 0.25 * dt * 3.0 / 2.0;
  else cffAB = 0.25 * dt * 23.0 / 12.0;

  for (int j = Jstr; j <= Jend; ++j) {
    // ---- u
    for (int i = IstrU; i <= Iend; ++i) {
      AK(i, 0) = 0.5 * (Akv(i - 1, j, 0) + Akv(i, j, 0));
      for (int k = 1; k <= N; ++k) {
        AK(i, k) = 0.5 * (Akv(i - 1, j, k) + Akv(i, j, k));
        Hzk(i, k) = 0.5 * (Hz(i - 1, j, k) + Hz(i, j, k));
        oHz(i, k) = 1.0 / Hzk(i, k);
      }
    }
    for (int i = IstrU; i <= Iend; ++i) DC(i, 0) = cffAB * (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
    for (int k = 1; k <= N; ++k)
      for (int i = IstrU; i <= Iend; ++i) {
        u(i, j, k) = u(i, j, k) + DC(i, 0) * ru(i, j, k);
        if (!nospl) u(i, j, k) = u(i, j, k) * oHz(i, k);                       // # ifdef SPLINES_VVISC (:319-321)
      }
    if (nospl) {                                                                // # else of SPLINES_VVISC, :397-462: centred tridiagonal system
      const double cff = -c.lambda * dt / 0.5;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          const double cff1 = 1.0 / (m.z_r(i, j, k + 1) + m.z_r(i - 1, j, k + 1) - m.z_r(i, j, k) - m.z_r(i - 1, j, k));
          FC(i, k) = cff * cff1 * AK(i, k);
        }
      for (int i = IstrU; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) { DC(i, k) = u(i, j, k); BC(i, k) = Hzk(i, k) - FC(i, k) - FC(i, k - 1); }
      for (int i = IstrU; i <= Iend; ++i) { const double cf = 1.0 / BC(i, 1); CF(i, 1) = cf * FC(i, 1); DC(i, 1) = cf * DC(i, 1); }
      for (int k = 2; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          const double cf = 1.0 / (BC(i, k) - FC(i, k - 1) * CF(i, k - 1));
          CF(i, k) = cf * FC(i, k);
          DC(i, k) = cf * (DC(i, k) - FC(i, k - 1) * DC(i, k - 1));
        }
      for (int i = IstrU; i <= Iend; ++i) {
        DC(i, N) = (DC(i, N) - FC(i, N - 1) * DC(i, N - 1)) / (BC(i, N) - FC(i, N - 1) * CF(i, N - 1));
        u(i, j, N) = DC(i, N);
      }
      for (int k = N - 1; k >= 1; --k)
        for (int i = IstrU; i <= Iend; ++i) { DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1); u(i, j, k) = DC(i, k); }
    } else {
      double cff1 = 1.0 / 6.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          FC(i, k) = cff1 * Hzk(i, k) - dt * AK(i, k - 1) * oHz(i, k);
          CF(i, k) = cff1 * Hzk(i, k + 1) - dt * AK(i, k + 1) * oHz(i, k + 1);
        }
      for (int i = IstrU; i <= Iend; ++i) { CF(i, 0) = 0.0; DC(i, 0) = 0.0; }
      cff1 = 1.0 / 3.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          BC(i, k) = cff1 * (Hzk(i, k) + Hzk(i, k + 1)) + dt * AK(i, k) * (oHz(i, k) + oHz(i, k + 1));
          double cff = 1.0 / (BC(i, k) - FC(i, k) * CF(i, k - 1));
          CF(i, k) = cff * CF(i, k);
          DC(i, k) = cff * (u(i, j, k + 1) - u(i, j, k) - FC(i, k) * DC(i, k - 1));
        }
      for (int i = IstrU; i <= Iend; ++i) DC(i, N) = 0.0;
      for (int k = N - 1; k >= 1; --k) for (int i = IstrU; i <= Iend; ++i) DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1);
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          DC(i, k) = DC(i, k) * AK(i, k);
          double cff = dt * oHz(i, k) * (DC(i, k) - DC(i, k - 1));
          u(i, j, k) = u(i, j, k) + cff;
        }
    }
    for (int i = IstrU; i <= Iend; ++i) { CF(i, 0) = Hzk(i, 1); DC(i, 0) = u(i, j, 1) * Hzk(i, 1); }
    for (int k = 2; k <= N; ++k)
      for (int i = IstrU; i <= Iend; ++i) { CF(i, 0) = CF(i, 0) + Hzk(i, k); DC(i, 0) = DC(i, 0) + u(i, j, k) * Hzk(i, k); }
    for (int i = IstrU; i <= Iend; ++i) {
      double cff1 = 1.0 / (CF(i, 0) * m.on_u(i, j));
      DC(i, 0) = (DC(i, 0) * m.on_u(i, j) - m.DU_avg1(i, j)) * cff1;
    }
    for (int k = 1; k <= N; ++k) for (int i = IstrU; i <= Iend; ++i) u(i, j, k) = u(i, j, k) - DC(i, 0);
    // ---- v
    if (j >= JstrV) {
      for (int i = Istr; i <= Iend; ++i) {
        AK(i, 0) = 0.5 * (Akv(i, j - 1, 0) + Akv(i, j, 0));
        for (int k = 1; k <= N; ++k) {
          AK(i, k) = 0.5 * (Akv(i, j - 1, k) + Akv(i, j, k));
          Hzk(i, k) = 0.5 * (Hz(i, j - 1, k) + Hz(i, j, k));
          oHz(i, k) = 1.0 / Hzk(i, k);
        }
      }
      for (int i = Istr; i <= Iend; ++i) DC(i, 0) = cffAB * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          v(i, j, k) = v(i, j, k) + DC(i, 0) * rv(i, j, k);
          if (!nospl) v(i, j, k) = v(i, j, k) * oHz(i, k);
        }
      if (nospl) {                                                              // :730-795
        const double cff = -c.lambda * dt / 0.5;
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cff1 = 1.0 / (m.z_r(i, j, k + 1) + m.z_r(i, j - 1, k + 1) - m.z_r(i, j, k) - m.z_r(i, j - 1, k));
            FC(i, k) = cff * cff1 * AK(i, k);
          }
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) { DC(i, k) = v(i, j, k); BC(i, k) = Hzk(i, k) - FC(i, k) - FC(i, k - 1); }
        for (int i = Istr; i <= Iend; ++i) { const double cf = 1.0 / BC(i, 1); CF(i, 1) = cf * FC(i, 1); DC(i, 1) = cf * DC(i, 1); }
        for (int k = 2; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cf = 1.0 / (BC(i, k) - FC(i, k - 1) * CF(i, k - 1));
            CF(i, k) = cf * FC(i, k);
            DC(i, k) = cf * (DC(i, k) - FC(i, k - 1) * DC(i, k - 1));
          }
        for (int i = Istr; i <= Iend; ++i) {
          DC(i, N) = (DC(i, N) - FC(i, N - 1) * DC(i, N - 1)) / (BC(i, N) - FC(i, N - 1) * CF(i, N - 1));
          v(i, j, N) = DC(i, N);
        }
        for (int k = N - 1; k >= 1; --k)
          for (int i = Istr; i <= Iend; ++i) { DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1); v(i, j, k) = DC(i, k); }
      } else {
      double cff1 = 1.0 / 6.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          FC(i, k) = cff1 * Hzk(i, k) - dt * AK(i, k - 1) * oHz(i, k);
          CF(i, k) = cff1 * Hzk(i, k + 1) - dt * AK(i, k + 1) * oHz(i, k + 1);
        }
      for (int i = Istr; i <= Iend; ++i) { CF(i, 0) = 0.0; DC(i, 0) = 0.0; }
      cff1 = 1.0 / 3.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          BC(i, k) = cff1 * (Hzk(i, k) + Hzk(i, k + 1)) + dt * AK(i, k) * (oHz(i, k) + oHz(i, k + 1));
          double cff = 1.0 / (BC(i, k) - FC(i, k) * CF(i, k - 1));
          CF(i, k) = cff * CF(i, k);
          DC(i, k) = cff * (v(i, j, k + 1) - v(i, j, k) - FC(i, k) * DC(i, k - 1));
        }
      for (int i = Istr; i <= Iend; ++i) DC(i, N) = 0.0;
      for (int k = N - 1; k >= 1; --k) for (int i = Istr; i <= Iend; ++i) DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1);
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          DC(i, k) = DC(i, k) * AK(i, k);
          double cff = dt * oHz(i, k) * (DC(i, k) - DC(i, k - 1));
          v(i, j, k) = v(i, j, k) + cff;
        }
      }
      for (int i = Istr; i <= Iend; ++i) { CF(i, 0) = Hzk(i, 1); DC(i, 0) = v(i, j, 1) * Hzk(i, 1); }
      for (int k = 2; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) { CF(i, 0) = CF(i, 0) + Hzk(i, k); DC(i, 0) = DC(i, 0) + v(i, j, k) * Hzk(i, k); }
      for (int i = Istr; i <= Iend; ++i) {
        double cff1_ = 1.0 / (CF(i, 0) * m.om_v(i, j));
        DC(i, 0) = (DC(i, 0) * m.om_v(i, j) - m.DV_avg1(i, j)) * cff1_;
      }
      for (int k = 1; k <= N; ++k) for (int i = Istr; i <= Iend; ++i) v(i, j, k) = v(i, j, k) - DC(i, 0);
    }
  }

  // ---- :956-965
  u3dbc(m, b, nnew); v3dbc(m, b, nnew);

  // ---- :1002-1432  couple 2-D and 3-D momentum, boundary rows included
  for (int j = JstrT; j <= JendT; ++j) {
    for (int i = IstrP; i <= IendT; ++i) { DC(i, 0) = 0.0; CF(i, 0) = 0.0; FC(i, 0) = 0.0; }
    for (int k = 1; k <= N; ++k)
      for (int i = IstrP; i <= IendT; ++i) {
        double cff = 0.5 * m.on_u(i, j);
        DC(i, k) = cff * (Hz(i, j, k) + Hz(i - 1, j, k));
        DC(i, 0) = DC(i, 0) + DC(i, k);
        CF(i, 0) = CF(i, 0) + DC(i, k) * u(i, j, k);
      }
    for (int i = IstrP; i <= IendT; ++i) {
      DC(i, 0) = 1.0 / DC(i, 0);
      CF(i, 0) = DC(i, 0) * (CF(i, 0) - m.DU_avg1(i, j));
      m.ubar[1](i, j) = DC(i, 0) * m.DU_avg1(i, j);
      m.ubar[2](i, j) = m.ubar[1](i, j);
    }
    // (:1078-1130 closed E/W wall corrections: not live, EW periodic)
    if (j == 0)
      for (int k = 1; k <= N; ++k) for (int i = IstrU; i <= Iend; ++i) u(i, j, k) = u(i, j, k) - CF(i, 0);
    if (j == Mm + 1)
      for (int k = 1; k <= N; ++k) for (int i = IstrU; i <= Iend; ++i) u(i, j, k) = u(i, j, k) - CF(i, 0);
    for (int k = N; k >= 1; --k)
      for (int i = IstrP; i <= IendT; ++i) {
        Huon(i, j, k) = 0.5 * (Huon(i, j, k) + u(i, j, k) * DC(i, k));
        FC(i, 0) = FC(i, 0) + Huon(i, j, k);
      }
    for (int i = IstrP; i <= IendT; ++i) FC(i, 0) = DC(i, 0) * (FC(i, 0) - m.DU_avg2(i, j));
    for (int k = 1; k <= N; ++k) for (int i = IstrP; i <= IendT; ++i) Huon(i, j, k) = Huon(i, j, k) - DC(i, k) * FC(i, 0);
    if (j >= Jstr) {
      for (int i = IstrT; i <= IendT; ++i) { DC(i, 0) = 0.0; CF(i, 0) = 0.0; FC(i, 0) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrT; i <= IendT; ++i) {
          double cff = 0.5 * m.om_v(i, j);
          DC(i, k) = cff * (Hz(i, j, k) + Hz(i, j - 1, k));
          DC(i, 0) = DC(i, 0) + DC(i, k);
          CF(i, 0) = CF(i, 0) + DC(i, k) * v(i, j, k);
        }
      for (int i = IstrT; i <= IendT; ++i) {
        DC(i, 0) = 1.0 / DC(i, 0);
        CF(i, 0) = DC(i, 0) * (CF(i, 0) - m.DV_avg1(i, j));
        m.vbar[1](i, j) = DC(i, 0) * m.DV_avg1(i, j);
        m.vbar[2](i, j) = m.vbar[1](i, j);
      }
      if (j == 1)
        for (int k = 1; k <= N; ++k) for (int i = Istr; i <= Iend; ++i) v(i, j, k) = v(i, j, k) - CF(i, 0);
      if (j == Mm + 1)
        for (int k = 1; k <= N; ++k) for (int i = Istr; i <= Iend; ++i) v(i, j, k) = v(i, j, k) - CF(i, 0);
      for (int k = N; k >= 1; --k)
        for (int i = IstrT; i <= IendT; ++i) {
          Hvom(i, j, k) = 0.5 * (Hvom(i, j, k) + v(i, j, k) * DC(i, k));
          FC(i, 0) = FC(i, 0) + Hvom(i, j, k);
        }
      for (int i = IstrT; i <= IendT; ++i) FC(i, 0) = DC(i, 0) * (FC(i, 0) - m.DV_avg2(i, j));
      for (int k = 1; k <= N; ++k) for (int i = IstrT; i <= IendT; ++i) Hvom(i, j, k) = Hvom(i, j, k) - DC(i, k) * FC(i, 0);
    }
  }
  // ---- :1438-1461
  exchange_u3d(m, b, u); exchange_v3d(m, b, v); exchange_u3d(m, b, Huon); exchange_v3d(m, b, Hvom);
  for (int k = 1; k <= 2; ++k) { exchange_u2d(m, b, m.ubar[k]); exchange_v2d(m, b, m.vbar[k]); }
}

}  // namespace orc

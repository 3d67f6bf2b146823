// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/pre_step3d.F:123-1154 (pre_step3d_tile): predictor for tracers at n+1/2, explicit part of the
// vertical diffusion / viscosity, AB3 loading of u,v(nnew).
#include "roms_oracle.hpp"

namespace orc {

void pre_step3d(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, NT = c.NT, nstp = m.nstp, nnew = m.nnew, nrhs = m.nrhs;
  const double dt = c.dt, lambda = c.lambda;
  const double eps = 1.0e-16;                                  // pre_step3d.F:287
  const bool first = (m.iic == m.ntfirst), second = (m.iic == m.ntfirst + 1);
  F3 &Hz = m.Hz, &Huon = m.Huon, &Hvom = m.Hvom, &W = m.W, &z_r = m.z_r, &Akv = m.Akv;
  F2 &pm = m.pm, &pn = m.pn;
  SK CF(IminS, ImaxS, 0, N), DC(IminS, ImaxS, 0, N), FC(IminS, ImaxS, 0, N);
  S2 FE(IminS, ImaxS, JminS, JmaxS), FX(IminS, ImaxS, JminS, JmaxS), curv(IminS, ImaxS, JminS, JmaxS), grad(IminS, ImaxS, JminS, JmaxS);

  // ---- :342-582  horizontal advection -> t(:,:,:,3,itrc)
  for (int itrc = 0; itrc < NT; ++itrc) {
    F3 tst = m.t[nstp][itrc], tnw = m.t[nnew][itrc], t3 = m.t[3][itrc];
    for (int k = 1; k <= N; ++k) {
      if (c.hadv == HADV_C2) {                                  // :345-362
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr; i <= Iend + 1; ++i) FX(i, j) = Huon(i, j, k) * 0.5 * (tst(i - 1, j, k) + tst(i, j, k));
        for (int j = Jstr; j <= Jend + 1; ++j)
          for (int i = Istr; i <= Iend; ++i) FE(i, j) = Hvom(i, j, k) * 0.5 * (tst(i, j - 1, k) + tst(i, j, k));
      } else {                                                  // :386-521 (AKIMA4 / CENTERED4 / UPSTREAM3)
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istrm1; i <= Iendp2; ++i) FX(i, j) = tst(i, j, k) - tst(i - 1, j, k);
        // closed-wall copies :403-416 only if .not.EWperiodic -> not live
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr - 1; i <= Iend + 1; ++i) {
            if (c.hadv == HADV_U3) curv(i, j) = FX(i + 1, j) - FX(i, j);
            else if (c.hadv == HADV_A4) {
              double cff = 2.0 * FX(i + 1, j) * FX(i, j);
              if (cff > eps) grad(i, j) = cff / (FX(i + 1, j) + FX(i, j)); else grad(i, j) = 0.0;
            } else grad(i, j) = 0.5 * (FX(i + 1, j) + FX(i, j));
          }
        {
          const double cff1 = 1.0 / 6.0, cff2 = 1.0 / 3.0;
          for (int j = Jstr; j <= Jend; ++j)
            for (int i = Istr; i <= Iend + 1; ++i) {
              if (c.hadv == HADV_U3)
                FX(i, j) = Huon(i, j, k) * 0.5 * (tst(i - 1, j, k) + tst(i, j, k)) -
                           cff1 * (curv(i - 1, j) * std::max(Huon(i, j, k), 0.0) + curv(i, j) * std::min(Huon(i, j, k), 0.0));
              else
                FX(i, j) = Huon(i, j, k) * 0.5 * (tst(i - 1, j, k) + tst(i, j, k) - cff2 * (grad(i, j) - grad(i - 1, j)));
            }
        }
        for (int j = Jstrm1; j <= Jendp2; ++j)
          for (int i = Istr; i <= Iend; ++i) FE(i, j) = tst(i, j, k) - tst(i, j - 1, k);
        if (b.Southern_Edge) for (int i = Istr; i <= Iend; ++i) FE(i, Jstr - 1) = FE(i, Jstr);      // :468-474
        if (b.Northern_Edge) for (int i = Istr; i <= Iend; ++i) FE(i, Jend + 2) = FE(i, Jend + 1);  // :475-481
        for (int j = Jstr - 1; j <= Jend + 1; ++j)
          for (int i = Istr; i <= Iend; ++i) {
            if (c.hadv == HADV_U3) curv(i, j) = FE(i, j + 1) - FE(i, j);
            else if (c.hadv == HADV_A4) {
              double cff = 2.0 * FE(i, j + 1) * FE(i, j);
              if (cff > eps) grad(i, j) = cff / (FE(i, j + 1) + FE(i, j)); else grad(i, j) = 0.0;
            } else grad(i, j) = 0.5 * (FE(i, j + 1) + FE(i, j));
          }
        {
          const double cff1 = 1.0 / 6.0, cff2 = 1.0 / 3.0;
          for (int j = Jstr; j <= Jend + 1; ++j)
            for (int i = Istr; i <= Iend; ++i) {
              if (c.hadv == HADV_U3)
                FE(i, j) = Hvom(i, j, k) * 0.5 * (tst(i, j - 1, k) + tst(i, j, k)) -
                           cff1 * (curv(i, j - 1) * std::max(Hvom(i, j, k), 0.0) + curv(i, j) * std::min(Hvom(i, j, k), 0.0));
              else
                FE(i, j) = Hvom(i, j, k) * 0.5 * (tst(i, j - 1, k) + tst(i, j, k) - cff2 * (grad(i, j) - grad(i, j - 1)));
            }
        }
      }
      // :557-580
      const double Gamma = 1.0 / 6.0;
      double cff, cff1, cff2;
      if (first) { cff = 0.5 * dt; cff1 = 1.0; cff2 = 0.0; }
      else { cff = (1.0 - Gamma) * dt; cff1 = 0.5 + Gamma; cff2 = 0.5 - Gamma; }
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i)
          t3(i, j, k) = Hz(i, j, k) * (cff1 * tst(i, j, k) + cff2 * tnw(i, j, k)) -
                        cff * pm(i, j) * pn(i, j) * (FX(i + 1, j) - FX(i, j) + FE(i, j + 1) - FE(i, j));
    }
  }

  // ---- :619-826  vertical advection + pseudo-compressible divide
  for (int j = Jstr; j <= Jend; ++j) {
    for (int itrc = 0; itrc < NT; ++itrc) {
      F3 tst = m.t[nstp][itrc], t3 = m.t[3][itrc];
      if (c.vadv == VADV_SPLINES) {                              // :622-665: conservative parabolic splines (NEUMANN, pre_step3d.F:4)
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 1.5 * tst(i, j, 1); CF(i, 1) = 0.5; }
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cff = 1.0 / (2.0 * Hz(i, j, k) + Hz(i, j, k + 1) * (2.0 - CF(i, k)));
            CF(i, k + 1) = cff * Hz(i, j, k);
            FC(i, k) = cff * (3.0 * (Hz(i, j, k) * tst(i, j, k + 1) + Hz(i, j, k + 1) * tst(i, j, k)) - Hz(i, j, k + 1) * FC(i, k - 1));
          }
        for (int i = Istr; i <= Iend; ++i) FC(i, N) = (3.0 * tst(i, j, N) - FC(i, N - 1)) / (2.0 - CF(i, N));
        for (int k = N - 1; k >= 0; --k)
          for (int i = Istr; i <= Iend; ++i) {
            FC(i, k) = FC(i, k) - CF(i, k + 1) * FC(i, k + 1);
            FC(i, k + 1) = W(i, j, k + 1) * FC(i, k + 1);
          }
        for (int i = Istr; i <= Iend; ++i) { FC(i, N) = 0.0; FC(i, 0) = 0.0; }
      } else if (c.vadv == VADV_A4) {                                  // :667-707
        for (int k = 1; k <= N - 1; ++k) for (int i = Istr; i <= Iend; ++i) FC(i, k) = tst(i, j, k + 1) - tst(i, j, k);
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = FC(i, 1); FC(i, N) = FC(i, N - 1); }
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            double cff = 2.0 * FC(i, k) * FC(i, k - 1);
            if (cff > eps) CF(i, k) = cff / (FC(i, k) + FC(i, k - 1)); else CF(i, k) = 0.0;
          }
        const double cff1 = 1.0 / 3.0;
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i)
            FC(i, k) = W(i, j, k) * 0.5 * (tst(i, j, k) + tst(i, j, k + 1) - cff1 * (CF(i, k + 1) - CF(i, k)));
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      } else if (c.vadv == VADV_C2) {                           // :709-727
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) FC(i, k) = W(i, j, k) * 0.5 * (tst(i, j, k) + tst(i, j, k + 1));
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      } else {                                                  // CENTERED4 :751-785
        const double cff1 = 0.5, cff2 = 7.0 / 12.0, cff3 = 1.0 / 12.0;
        for (int k = 2; k <= N - 2; ++k)
          for (int i = Istr; i <= Iend; ++i)
            FC(i, k) = W(i, j, k) * (cff2 * (tst(i, j, k) + tst(i, j, k + 1)) - cff3 * (tst(i, j, k - 1) + tst(i, j, k + 2)));
        for (int i = Istr; i <= Iend; ++i) {
          FC(i, 0) = 0.0;
          FC(i, 1) = W(i, j, 1) * (cff1 * tst(i, j, 1) + cff2 * tst(i, j, 2) - cff3 * tst(i, j, 3));
          FC(i, N - 1) = W(i, j, N - 1) * (cff1 * tst(i, j, N) + cff2 * tst(i, j, N - 1) - cff3 * tst(i, j, N - 2));
          FC(i, N) = 0.0;
        }
      }
      const double Gamma = 1.0 / 6.0;
      double cff = first ? 0.5 * dt : (1.0 - Gamma) * dt;       // :799-803
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i)
          DC(i, k) = 1.0 / (Hz(i, j, k) - cff * pm(i, j) * pn(i, j) *
                                              (Huon(i + 1, j, k) - Huon(i, j, k) + Hvom(i, j + 1, k) - Hvom(i, j, k) + (W(i, j, k) - W(i, j, k - 1))));
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff1 = cff * pm(i, j) * pn(i, j);
          t3(i, j, k) = DC(i, k) * (t3(i, j, k) - cff1 * (FC(i, k) - FC(i, k - 1)));
        }
    }
  }

  // ---- :837-906  t(nnew) = Hz*t(nstp) + explicit vertical diffusion + surface/bottom fluxes
  for (int j = Jstr; j <= Jend; ++j) {
    double cff3 = dt * (1.0 - lambda);
    for (int itrc = 0; itrc < NT; ++itrc) {
      F3 tst = m.t[nstp][itrc], tnw = m.t[nnew][itrc]; F3 Akt = m.Akt[std::min(c.NT, itrc + 1) - 1];
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff = 1.0 / (z_r(i, j, k + 1) - z_r(i, j, k));
          FC(i, k) = cff3 * cff * Akt(i, j, k) * (tst(i, j, k + 1) - tst(i, j, k));
        }
      if (c.lmd_nonlocal && itrc < (c.salinity ? 2 : 1)) {                 // :850-865 (active tracers, itrc <= NAT)
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) FC(i, k) = FC(i, k) - dt * Akt(i, j, k) * m.ghats[itrc](i, j, k);
      }
      if (c.solar_source && itrc == c.itemp - 1) {                         // :312-333 + lmd_swfrac.F:66-80 (Zscale = -1), :866-883
        static const double lmd_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};      // mod_scalars.F:1502-1512
        static const double lmd_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};
        static const double lmd_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double Z = m.z_w(i, j, N) - m.z_w(i, j, k);
            const int Jindex = (int)m.Jwtype(i, j);
            const double fac1 = -1.0 / lmd_mu1[Jindex - 1], fac2 = -1.0 / lmd_mu2[Jindex - 1], fac3 = lmd_r1[Jindex - 1];
            const double swdk = std::exp(Z * fac1) * fac3 + std::exp(Z * fac2) * (1.0 - fac3);
            FC(i, k) = FC(i, k) + dt * m.srflx(i, j) * swdk;
          }
      }
      for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = dt * m.btflx[itrc](i, j); FC(i, N) = dt * m.stflx[itrc](i, j); }
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff1 = Hz(i, j, k) * tst(i, j, k);
          double cff2 = FC(i, k) - FC(i, k - 1);
          tnw(i, j, k) = cff1 + cff2;
        }
    }
  }

  // ---- :917-1118  momentum predictor
  F3 ust = m.u[nstp], unw = m.u[nnew], vst = m.v[nstp], vnw = m.v[nnew];
  const int indx = 3 - nrhs;
  F3 ru_r = m.ru[nrhs], ru_i = m.ru[indx], rv_r = m.rv[nrhs], rv_i = m.rv[indx];
  for (int j = Jstr; j <= Jend; ++j) {
    double cff3 = dt * (1.0 - lambda);
    for (int k = 1; k <= N - 1; ++k)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff = 1.0 / (z_r(i, j, k + 1) + z_r(i - 1, j, k + 1) - z_r(i, j, k) - z_r(i - 1, j, k));
        FC(i, k) = cff3 * cff * (ust(i, j, k + 1) - ust(i, j, k)) * (Akv(i, j, k) + Akv(i - 1, j, k));
      }
    for (int i = IstrU; i <= Iend; ++i) { FC(i, 0) = dt * m.bustr(i, j); FC(i, N) = dt * m.sustr(i, j); if (c.bodyforce) { FC(i, 0) = 0.0; FC(i, N) = 0.0; } }   // :931-937
    double cff = dt * 0.25;
    for (int i = IstrU; i <= Iend; ++i) DC(i, 0) = cff * (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
    if (first) {
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          double cff1 = ust(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
          double cff2 = FC(i, k) - FC(i, k - 1);
          unw(i, j, k) = cff1 + cff2;
        }
    } else if (second) {
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          double cff1 = ust(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
          double cff2 = FC(i, k) - FC(i, k - 1);
          double cff3_ = 0.5 * DC(i, 0);
          unw(i, j, k) = cff1 - cff3_ * ru_i(i, j, k) + cff2;
        }
    } else {
      const double cff1 = 5.0 / 12.0, cff2 = 16.0 / 12.0;
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          double cff3_ = ust(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
          double cff4 = FC(i, k) - FC(i, k - 1);
          unw(i, j, k) = cff3_ + DC(i, 0) * (cff1 * ru_r(i, j, k) - cff2 * ru_i(i, j, k)) + cff4;
        }
    }
    if (j >= JstrV) {
      cff3 = dt * (1.0 - lambda);
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cf = 1.0 / (z_r(i, j, k + 1) + z_r(i, j - 1, k + 1) - z_r(i, j, k) - z_r(i, j - 1, k));
          FC(i, k) = cff3 * cf * (vst(i, j, k + 1) - vst(i, j, k)) * (Akv(i, j, k) + Akv(i, j - 1, k));
        }
      for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = dt * m.bvstr(i, j); FC(i, N) = dt * m.svstr(i, j); if (c.bodyforce) { FC(i, 0) = 0.0; FC(i, N) = 0.0; } }   // :1036-1042
      cff = dt * 0.25;
      for (int i = Istr; i <= Iend; ++i) DC(i, 0) = cff * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      if (first) {
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            double cff1 = vst(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
            double cff2 = FC(i, k) - FC(i, k - 1);
            vnw(i, j, k) = cff1 + cff2;
          }
      } else if (second) {
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            double cff1 = vst(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
            double cff2 = FC(i, k) - FC(i, k - 1);
            double cff3_ = 0.5 * DC(i, 0);
            vnw(i, j, k) = cff1 - cff3_ * rv_i(i, j, k) + cff2;
          }
      } else {
        const double cff1 = 5.0 / 12.0, cff2 = 16.0 / 12.0;
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            double cff3_ = vst(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
            double cff4 = FC(i, k) - FC(i, k - 1);
            vnw(i, j, k) = cff3_ + DC(i, 0) * (cff1 * rv_r(i, j, k) - cff2 * rv_i(i, j, k)) + cff4;
          }
      }
    }
  }

  // ---- :1126-1142  boundary conditions + periodic exchange of the n+1/2 tracers
  for (int itrc = 0; itrc < NT; ++itrc) { t3dbc(m, b, 3, itrc); exchange_r3d(m, b, m.t[3][itrc]); }
}

}  // namespace orc

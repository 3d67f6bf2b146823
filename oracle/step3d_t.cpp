// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/step3d_t.F:108-1680 (step3d_t_tile): tracer corrector.  Horizontal advection of t(:,:,:,3,:)
// (CENTERED2 :390-407; AKIMA4/CENTERED4/UPSTREAM3 :591-725), stepping :861-873; vertical advection (AKIMA4 :938-977,
// CENTERED2 :978-994, CENTERED4 :1091-1126), stepping :1189-1207; SPLINES_VDIFF implicit diffusion :1370-1427;
// t3dbc + periodic exchange :1551-1621.  MPDATA/HSIMT/SPLINES advection are not selectable here.
#include "roms_oracle.hpp"

namespace orc {

void step3d_t(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const bool nospl = c.nospl_vdiff != 0;                          // SPLINES_VDIFF undefined (:1196-1198, :1430-1499)
  const int N = c.N, NT = c.NT, nnew = m.nnew; const double dt = c.dt;
  const double eps = 1.0e-16;
  F3 &Hz = m.Hz, &Huon = m.Huon, &Hvom = m.Hvom, &W = m.W; F2 &pm = m.pm, &pn = m.pn;
  SK CF(IminS, ImaxS, 0, N), BC(IminS, ImaxS, 0, N), DC(IminS, ImaxS, 0, N), FC(IminS, ImaxS, 0, N);
  S2 FE(IminS, ImaxS, JminS, JmaxS), FX(IminS, ImaxS, JminS, JmaxS), curv(IminS, ImaxS, JminS, JmaxS), grad(IminS, ImaxS, JminS, JmaxS);
  S3 oHz(IminS, ImaxS, JminS, JmaxS, 1, N);
  for (int k = 1; k <= N; ++k)
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) oHz(i, j, k) = 1.0 / Hz(i, j, k);

  for (int itrc = 0; itrc < NT; ++itrc) {
    F3 t3 = m.t[3][itrc], tn = m.t[nnew][itrc];
    for (int k = 1; k <= N; ++k) {
      if (c.hadv == HADV_C2) {
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr; i <= Iend + 1; ++i) FX(i, j) = Huon(i, j, k) * 0.5 * (t3(i - 1, j, k) + t3(i, j, k));
        for (int j = Jstr; j <= Jend + 1; ++j)
          for (int i = Istr; i <= Iend; ++i) FE(i, j) = Hvom(i, j, k) * 0.5 * (t3(i, j - 1, k) + t3(i, j, k));
      } else {
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istrm1; i <= Iendp2; ++i) FX(i, j) = t3(i, j, k) - t3(i - 1, j, k);
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
                FX(i, j) = Huon(i, j, k) * 0.5 * (t3(i - 1, j, k) + t3(i, j, k)) -
                           cff1 * (curv(i - 1, j) * std::max(Huon(i, j, k), 0.0) + curv(i, j) * std::min(Huon(i, j, k), 0.0));
              else
                FX(i, j) = Huon(i, j, k) * 0.5 * (t3(i - 1, j, k) + t3(i, j, k) - cff2 * (grad(i, j) - grad(i - 1, j)));
            }
        }
        for (int j = Jstrm1; j <= Jendp2; ++j)
          for (int i = Istr; i <= Iend; ++i) FE(i, j) = t3(i, j, k) - t3(i, j - 1, k);
        if (b.Southern_Edge) for (int i = Istr; i <= Iend; ++i) FE(i, Jstr - 1) = FE(i, Jstr);
        if (b.Northern_Edge) for (int i = Istr; i <= Iend; ++i) FE(i, Jend + 2) = FE(i, Jend + 1);
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
                FE(i, j) = Hvom(i, j, k) * 0.5 * (t3(i, j - 1, k) + t3(i, j, k)) -
                           cff1 * (curv(i, j - 1) * std::max(Hvom(i, j, k), 0.0) + curv(i, j) * std::min(Hvom(i, j, k), 0.0));
              else
                FE(i, j) = Hvom(i, j, k) * 0.5 * (t3(i, j - 1, k) + t3(i, j, k) - cff2 * (grad(i, j) - grad(i, j - 1)));
            }
        }
      }
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          double cff = dt * pm(i, j) * pn(i, j);
          double cff1 = cff * (FX(i + 1, j) - FX(i, j));
          double cff2 = cff * (FE(i, j + 1) - FE(i, j));
          double cff3 = cff1 + cff2;
          tn(i, j, k) = tn(i, j, k) - cff3;
        }
    }
  }

  for (int itrc = 0; itrc < NT; ++itrc) {
    F3 t3 = m.t[3][itrc], tn = m.t[nnew][itrc];
    for (int j = Jstr; j <= Jend; ++j) {
      if (c.vadv == VADV_SPLINES) {                              // step3d_t.F:894-937: conservative parabolic splines (not NEUMANN here)
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 2.0 * t3(i, j, 1); CF(i, 1) = 1.0; }
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cff = 1.0 / (2.0 * Hz(i, j, k) + Hz(i, j, k + 1) * (2.0 - CF(i, k)));
            CF(i, k + 1) = cff * Hz(i, j, k);
            FC(i, k) = cff * (3.0 * (Hz(i, j, k) * t3(i, j, k + 1) + Hz(i, j, k + 1) * t3(i, j, k)) - Hz(i, j, k + 1) * FC(i, k - 1));
          }
        for (int i = Istr; i <= Iend; ++i) FC(i, N) = (2.0 * t3(i, j, N) - FC(i, N - 1)) / (1.0 - CF(i, N));
        for (int k = N - 1; k >= 0; --k)
          for (int i = Istr; i <= Iend; ++i) {
            FC(i, k) = FC(i, k) - CF(i, k + 1) * FC(i, k + 1);
            FC(i, k + 1) = W(i, j, k + 1) * FC(i, k + 1);
          }
        for (int i = Istr; i <= Iend; ++i) { FC(i, N) = 0.0; FC(i, 0) = 0.0; }
      } else if (c.vadv == VADV_A4) {
        for (int k = 1; k <= N - 1; ++k) for (int i = Istr; i <= Iend; ++i) FC(i, k) = t3(i, j, k + 1) - t3(i, j, k);
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = FC(i, 1); FC(i, N) = FC(i, N - 1); }
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            double cff = 2.0 * FC(i, k) * FC(i, k - 1);
            if (cff > eps) CF(i, k) = cff / (FC(i, k) + FC(i, k - 1)); else CF(i, k) = 0.0;
          }
        const double cff1 = 1.0 / 3.0;
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i)
            FC(i, k) = W(i, j, k) * 0.5 * (t3(i, j, k) + t3(i, j, k + 1) - cff1 * (CF(i, k + 1) - CF(i, k)));
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      } else if (c.vadv == VADV_C2) {
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) FC(i, k) = W(i, j, k) * 0.5 * (t3(i, j, k) + t3(i, j, k + 1));
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      } else {
        const double cff1 = 0.5, cff2 = 7.0 / 12.0, cff3 = 1.0 / 12.0;
        for (int k = 2; k <= N - 2; ++k)
          for (int i = Istr; i <= Iend; ++i)
            FC(i, k) = W(i, j, k) * (cff2 * (t3(i, j, k) + t3(i, j, k + 1)) - cff3 * (t3(i, j, k - 1) + t3(i, j, k + 2)));
        for (int i = Istr; i <= Iend; ++i) {
          FC(i, 0) = 0.0;
          FC(i, 1) = W(i, j, 1) * (cff1 * t3(i, j, 1) + cff2 * t3(i, j, 2) - cff3 * t3(i, j, 3));
          FC(i, N - 1) = W(i, j, N - 1) * (cff1 * t3(i, j, N) + cff2 * t3(i, j, N - 1) - cff3 * t3(i, j, N - 2));
          FC(i, N) = 0.0;
        }
      }
      for (int i = Istr; i <= Iend; ++i) CF(i, 0) = dt * pm(i, j) * pn(i, j);
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff1 = CF(i, 0) * (FC(i, k) - FC(i, k - 1));
          tn(i, j, k) = tn(i, j, k) - cff1;
          if (!nospl) tn(i, j, k) = tn(i, j, k) * oHz(i, j, k);           // # ifdef SPLINES_VDIFF (:1196-1198)
        }
    }
  }

  // ---- :1366-1427  implicit vertical diffusion, parabolic splines
  for (int j = Jstr; j <= Jend; ++j) {
    for (int itrc = 0; itrc < NT; ++itrc) {
      F3 tn = m.t[nnew][itrc]; F3 Akt = m.Akt[std::min(NT, itrc + 1) - 1];
      if (nospl) {                                                               // :1430-1499: centred tridiagonal system
        const double cff = -dt * c.lambda;
        for (int k = 1; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cff1 = 1.0 / (m.z_r(i, j, k + 1) - m.z_r(i, j, k));
            FC(i, k) = cff * cff1 * Akt(i, j, k);
          }
        for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
        for (int k = 1; k <= N; ++k)
          for (int i = Istr; i <= Iend; ++i) { BC(i, k) = Hz(i, j, k) - FC(i, k) - FC(i, k - 1); DC(i, k) = tn(i, j, k); }
        for (int i = Istr; i <= Iend; ++i) { const double cf = 1.0 / BC(i, 1); CF(i, 1) = cf * FC(i, 1); DC(i, 1) = cf * DC(i, 1); }
        for (int k = 2; k <= N - 1; ++k)
          for (int i = Istr; i <= Iend; ++i) {
            const double cf = 1.0 / (BC(i, k) - FC(i, k - 1) * CF(i, k - 1));
            CF(i, k) = cf * FC(i, k);
            DC(i, k) = cf * (DC(i, k) - FC(i, k - 1) * DC(i, k - 1));
          }
        for (int i = Istr; i <= Iend; ++i) {
          DC(i, N) = (DC(i, N) - FC(i, N - 1) * DC(i, N - 1)) / (BC(i, N) - FC(i, N - 1) * CF(i, N - 1));
          tn(i, j, N) = DC(i, N);
        }
        for (int k = N - 1; k >= 1; --k)
          for (int i = Istr; i <= Iend; ++i) { DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1); tn(i, j, k) = DC(i, k); }
        continue;
      }
      double cff1 = 1.0 / 6.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          FC(i, k) = cff1 * Hz(i, j, k) - dt * Akt(i, j, k - 1) * oHz(i, j, k);
          CF(i, k) = cff1 * Hz(i, j, k + 1) - dt * Akt(i, j, k + 1) * oHz(i, j, k + 1);
        }
      for (int i = Istr; i <= Iend; ++i) { CF(i, 0) = 0.0; DC(i, 0) = 0.0; }
      cff1 = 1.0 / 3.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          BC(i, k) = cff1 * (Hz(i, j, k) + Hz(i, j, k + 1)) + dt * Akt(i, j, k) * (oHz(i, j, k) + oHz(i, j, k + 1));
          double cff = 1.0 / (BC(i, k) - FC(i, k) * CF(i, k - 1));
          CF(i, k) = cff * CF(i, k);
          DC(i, k) = cff * (tn(i, j, k + 1) - tn(i, j, k) - FC(i, k) * DC(i, k - 1));
        }
      for (int i = Istr; i <= Iend; ++i) DC(i, N) = 0.0;
      for (int k = N - 1; k >= 1; --k) for (int i = Istr; i <= Iend; ++i) DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1);
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          DC(i, k) = DC(i, k) * Akt(i, j, k);
          double c1 = dt * oHz(i, j, k) * (DC(i, k) - DC(i, k - 1));
          tn(i, j, k) = tn(i, j, k) + c1;
        }
    }
  }
  // ---- :1551-1621
  for (int itrc = 0; itrc < NT; ++itrc) { t3dbc(m, b, nnew, itrc); exchange_r3d(m, b, m.t[nnew][itrc]); }
}

}  // namespace orc

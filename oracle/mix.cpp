// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// Harmonic horizontal mixing: ROMS/Nonlinear/t3dmix2_s.h:198-301, t3dmix2_geo.h:219-419 (dispatch t3dmix.F),
// ROMS/Nonlinear/uv3dmix2_s.h:239-330 (dispatch uv3dmix.F); biharmonic tracer mixing along s-surfaces t3dmix4_s.h:215-476.
#include "roms_oracle.hpp"

namespace orc {

static void t3dmix2_s(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double dt = c.dt;
  F3& Hz = m.Hz; F2 &pm = m.pm, &pn = m.pn;
  S2 FE(IminS, ImaxS, JminS, JmaxS), FX(IminS, ImaxS, JminS, JmaxS);
  for (int itrc = 0; itrc < c.NT; ++itrc) {
    F3 tr = m.t[m.nrhs][itrc], tn = m.t[m.nnew][itrc]; F2 diff2 = m.diff2[itrc];
    for (int k = 1; k <= N; ++k) {
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = Istr; i <= Iend + 1; ++i) {
          double cff = 0.25 * (diff2(i, j) + diff2(i - 1, j)) * m.pmon_u(i, j);
          FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) * (tr(i, j, k) - tr(i - 1, j, k));
        }
      for (int j = Jstr; j <= Jend + 1; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          double cff = 0.25 * (diff2(i, j) + diff2(i, j - 1)) * m.pnom_v(i, j);
          FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) * (tr(i, j, k) - tr(i, j - 1, k));
        }
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          double cff = dt * pm(i, j) * pn(i, j);
          double cff1 = cff * (FX(i + 1, j) - FX(i, j));
          double cff2 = cff * (FE(i, j + 1) - FE(i, j));
          double cff3 = cff1 + cff2;
          tn(i, j, k) = tn(i, j, k) + cff3;
        }
    }
  }
}

static void t3dmix2_geo(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double dt = c.dt;
  F3 &Hz = m.Hz, &z_r = m.z_r; F2 &pm = m.pm, &pn = m.pn;
  S2 FE(IminS, ImaxS, JminS, JmaxS), FX(IminS, ImaxS, JminS, JmaxS);
  S3 dTdz(IminS, ImaxS, JminS, JmaxS, 1, 2), dTdx(IminS, ImaxS, JminS, JmaxS, 1, 2), dTde(IminS, ImaxS, JminS, JmaxS, 1, 2),
      dZdx(IminS, ImaxS, JminS, JmaxS, 1, 2), dZde(IminS, ImaxS, JminS, JmaxS, 1, 2), FS(IminS, ImaxS, JminS, JmaxS, 1, 2);
  for (int itrc = 0; itrc < c.NT; ++itrc) {
    F3 tr = m.t[m.nrhs][itrc], tn = m.t[m.nnew][itrc]; F2 diff2 = m.diff2[itrc];
    int k2 = 1, k1;
    for (int k = 0; k <= N; ++k) {
      k1 = k2; k2 = 3 - k1;
      if (k < N) {
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr; i <= Iend + 1; ++i) {
            double cff = 0.5 * (pm(i, j) + pm(i - 1, j));
            dZdx(i, j, k2) = cff * (z_r(i, j, k + 1) - z_r(i - 1, j, k + 1));
            dTdx(i, j, k2) = cff * (tr(i, j, k + 1) - tr(i - 1, j, k + 1));
          }
        for (int j = Jstr; j <= Jend + 1; ++j)
          for (int i = Istr; i <= Iend; ++i) {
            double cff = 0.5 * (pn(i, j) + pn(i, j - 1));
            dZde(i, j, k2) = cff * (z_r(i, j, k + 1) - z_r(i, j - 1, k + 1));
            dTde(i, j, k2) = cff * (tr(i, j, k + 1) - tr(i, j - 1, k + 1));
          }
      }
      if (k == 0 || k == N) {
        for (int j = Jstr - 1; j <= Jend + 1; ++j)
          for (int i = Istr - 1; i <= Iend + 1; ++i) { dTdz(i, j, k2) = 0.0; FS(i, j, k2) = 0.0; }
      } else {
        for (int j = Jstr - 1; j <= Jend + 1; ++j)
          for (int i = Istr - 1; i <= Iend + 1; ++i) {
            double cff = 1.0 / (z_r(i, j, k + 1) - z_r(i, j, k));
            dTdz(i, j, k2) = cff * (tr(i, j, k + 1) - tr(i, j, k));
          }
      }
      if (k > 0) {
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr; i <= Iend + 1; ++i) {
            double cff = 0.25 * (diff2(i, j) + diff2(i - 1, j)) * m.on_u(i, j);
            FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) *
                       (dTdx(i, j, k1) - 0.5 * (std::min(dZdx(i, j, k1), 0.0) * (dTdz(i - 1, j, k1) + dTdz(i, j, k2)) +
                                                std::max(dZdx(i, j, k1), 0.0) * (dTdz(i - 1, j, k2) + dTdz(i, j, k1))));
          }
        for (int j = Jstr; j <= Jend + 1; ++j)
          for (int i = Istr; i <= Iend; ++i) {
            double cff = 0.25 * (diff2(i, j) + diff2(i, j - 1)) * m.om_v(i, j);
            FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) *
                       (dTde(i, j, k1) - 0.5 * (std::min(dZde(i, j, k1), 0.0) * (dTdz(i, j - 1, k1) + dTdz(i, j, k2)) +
                                                std::max(dZde(i, j, k1), 0.0) * (dTdz(i, j - 1, k2) + dTdz(i, j, k1))));
          }
        if (k < N) {
          for (int j = Jstr; j <= Jend; ++j)
            for (int i = Istr; i <= Iend; ++i) {
              double cff = 0.5 * diff2(i, j);
              double cff1 = std::min(dZdx(i, j, k1), 0.0), cff2 = std::min(dZdx(i + 1, j, k2), 0.0);
              double cff3 = std::max(dZdx(i, j, k2), 0.0), cff4 = std::max(dZdx(i + 1, j, k1), 0.0);
              FS(i, j, k2) = cff * (cff1 * (cff1 * dTdz(i, j, k2) - dTdx(i, j, k1)) + cff2 * (cff2 * dTdz(i, j, k2) - dTdx(i + 1, j, k2)) +
                                    cff3 * (cff3 * dTdz(i, j, k2) - dTdx(i, j, k2)) + cff4 * (cff4 * dTdz(i, j, k2) - dTdx(i + 1, j, k1)));
              cff1 = std::min(dZde(i, j, k1), 0.0); cff2 = std::min(dZde(i, j + 1, k2), 0.0);
              cff3 = std::max(dZde(i, j, k2), 0.0); cff4 = std::max(dZde(i, j + 1, k1), 0.0);
              FS(i, j, k2) = FS(i, j, k2) +
                             cff * (cff1 * (cff1 * dTdz(i, j, k2) - dTde(i, j, k1)) + cff2 * (cff2 * dTdz(i, j, k2) - dTde(i, j + 1, k2)) +
                                    cff3 * (cff3 * dTdz(i, j, k2) - dTde(i, j, k2)) + cff4 * (cff4 * dTdz(i, j, k2) - dTde(i, j + 1, k1)));
            }
        }
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = Istr; i <= Iend; ++i) {
            double cff = dt * pm(i, j) * pn(i, j);
            double cff1 = cff * (FX(i + 1, j) - FX(i, j));
            double cff2 = cff * (FE(i, j + 1) - FE(i, j));
            double cff3 = dt * (FS(i, j, k2) - FS(i, j, k1));
            double cff4 = cff1 + cff2 + cff3;
            tn(i, j, k) = tn(i, j, k) + cff4;
          }
      }
    }
  }
}

// t3dmix4_s_tile (ROMS/Nonlinear/t3dmix4_s.h:215-476; TS_DIF4 + MIX_S_TS; diff4 = SQRT(ABS(tnu4)), read_phypar.F:6905): the
// harmonic operator applied twice.  EW periodic, NS closed walls (LapT = 0 outside a closed wall, :352-376).
static void t3dmix4_s(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double dt = c.dt;
  F3& Hz = m.Hz; F2 &pm = m.pm, &pn = m.pn;
  S2 FE(IminS, ImaxS, JminS, JmaxS), FX(IminS, ImaxS, JminS, JmaxS), LapT(IminS, ImaxS, JminS, JmaxS);
  const int Imin = Istr - 1, Imax = Iend + 1;                                   // :221-227 (EWperiodic)
  const int Jmin = std::max(Jstr - 1, 1), Jmax = std::min(Jend + 1, c.Mm);      // :228-234
  for (int itrc = 0; itrc < c.NT; ++itrc) {
    F3 tr = m.t[m.nrhs][itrc], tn = m.t[m.nnew][itrc]; F2 diff4 = m.diff4[itrc];
    for (int k = 1; k <= N; ++k) {
      for (int j = Jmin; j <= Jmax; ++j)                                        // :243-291
        for (int i = Imin; i <= Imax + 1; ++i) {
          double cff = 0.25 * (diff4(i, j) + diff4(i - 1, j)) * m.pmon_u(i, j);
          FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) * (tr(i, j, k) - tr(i - 1, j, k));
        }
      for (int j = Jmin; j <= Jmax + 1; ++j)                                    // :292-340
        for (int i = Imin; i <= Imax; ++i) {
          double cff = 0.25 * (diff4(i, j) + diff4(i, j - 1)) * m.pnom_v(i, j);
          FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) * (tr(i, j, k) - tr(i, j - 1, k));
        }
      for (int j = Jmin; j <= Jmax; ++j)                                        // :341-348
        for (int i = Imin; i <= Imax; ++i) {
          double cff = 1.0 / Hz(i, j, k);
          LapT(i, j) = pm(i, j) * pn(i, j) * cff * (FX(i + 1, j) - FX(i, j) + FE(i, j + 1) - FE(i, j));
        }
      if (b.Southern_Edge) for (int i = Imin; i <= Imax; ++i) LapT(i, Jstr - 1) = 0.0;   // :378-390 (closed)
      if (b.Northern_Edge) for (int i = Imin; i <= Imax; ++i) LapT(i, Jend + 1) = 0.0;   // :392-404
      for (int j = Jstr; j <= Jend; ++j)                                        // :410-435
        for (int i = Istr; i <= Iend + 1; ++i) {
          double cff = 0.25 * (diff4(i, j) + diff4(i - 1, j)) * m.pmon_u(i, j);
          FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) * (LapT(i, j) - LapT(i - 1, j));
        }
      for (int j = Jstr; j <= Jend + 1; ++j)                                    // :436-461
        for (int i = Istr; i <= Iend; ++i) {
          double cff = 0.25 * (diff4(i, j) + diff4(i, j - 1)) * m.pnom_v(i, j);
          FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) * (LapT(i, j) - LapT(i, j - 1));
        }
      for (int j = Jstr; j <= Jend; ++j)                                        // :465-471
        for (int i = Istr; i <= Iend; ++i) {
          double cff = dt * pm(i, j) * pn(i, j);
          double cff1 = cff * (FX(i + 1, j) - FX(i, j));
          double cff2 = cff * (FE(i, j + 1) - FE(i, j));
          double cff3 = cff1 + cff2;
          tn(i, j, k) = tn(i, j, k) - cff3;
        }
    }
  }
}

void t3dmix4(Model& m, const Bnd& b) {
  if (m.c.mix_geo_ts) { std::fprintf(stderr, "oracle: t3dmix4_geo.h is not restated (TS_DIF4 needs MIX_S_TS)\n"); std::abort(); }
  t3dmix4_s(m, b);
}

void t3dmix2(Model& m, const Bnd& b) {
  if (m.c.mix_geo_ts) t3dmix2_geo(m, b); else t3dmix2_s(m, b);
}

void uv3dmix2(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double dt = c.dt;
  F3& Hz = m.Hz; F2 &pm = m.pm, &pn = m.pn;
  F3 u = m.u[m.nrhs], v = m.v[m.nrhs], un = m.u[m.nnew], vn = m.v[m.nnew];
  S2 UFe(IminS, ImaxS, JminS, JmaxS), VFe(IminS, ImaxS, JminS, JmaxS), UFx(IminS, ImaxS, JminS, JmaxS), VFx(IminS, ImaxS, JminS, JmaxS);
  for (int k = 1; k <= N; ++k) {
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff = Hz(i, j, k) * 0.5 *
                     (m.pmon_r(i, j) * ((pn(i, j) + pn(i + 1, j)) * u(i + 1, j, k) - (pn(i - 1, j) + pn(i, j)) * u(i, j, k)) -
                      m.pnom_r(i, j) * ((pm(i, j) + pm(i, j + 1)) * v(i, j + 1, k) - (pm(i, j - 1) + pm(i, j)) * v(i, j, k)));
        UFx(i, j) = m.on_r(i, j) * m.on_r(i, j) * m.visc2_r(i, j) * cff;
        VFe(i, j) = m.om_r(i, j) * m.om_r(i, j) * m.visc2_r(i, j) * cff;
      }
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = Istr; i <= Iend + 1; ++i) {
        double cff = 0.125 * (Hz(i - 1, j, k) + Hz(i, j, k) + Hz(i - 1, j - 1, k) + Hz(i, j - 1, k)) *
                     (m.pmon_p(i, j) * ((pn(i, j - 1) + pn(i, j)) * v(i, j, k) - (pn(i - 1, j - 1) + pn(i - 1, j)) * v(i - 1, j, k)) +
                      m.pnom_p(i, j) * ((pm(i - 1, j) + pm(i, j)) * u(i, j, k) - (pm(i - 1, j - 1) + pm(i, j - 1)) * u(i, j - 1, k)));
        UFe(i, j) = m.om_p(i, j) * m.om_p(i, j) * m.visc2_p(i, j) * cff;
        VFx(i, j) = m.on_p(i, j) * m.on_p(i, j) * m.visc2_p(i, j) * cff;
      }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff = dt * 0.25 * (pm(i - 1, j) + pm(i, j)) * (pn(i - 1, j) + pn(i, j));
        double cff1 = 0.5 * (pn(i - 1, j) + pn(i, j)) * (UFx(i, j) - UFx(i - 1, j));
        double cff2 = 0.5 * (pm(i - 1, j) + pm(i, j)) * (UFe(i, j + 1) - UFe(i, j));
        double cff3 = cff * (cff1 + cff2);
        m.rufrc(i, j) = m.rufrc(i, j) + cff1 + cff2;
        un(i, j, k) = un(i, j, k) + cff3;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = dt * 0.25 * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
        double cff1 = 0.5 * (pn(i, j - 1) + pn(i, j)) * (VFx(i + 1, j) - VFx(i, j));
        double cff2 = 0.5 * (pm(i, j - 1) + pm(i, j)) * (VFe(i, j) - VFe(i, j - 1));
        double cff3 = cff * (cff1 - cff2);
        m.rvfrc(i, j) = m.rvfrc(i, j) + cff1 - cff2;
        vn(i, j, k) = vn(i, j, k) + cff3;
      }
  }
}

}  // namespace orc

// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).  Glue routines of the main3d chain:
// set_massflux, rho_eos, set_vbc, omega, wvelocity, set_zeta, set_depth.
#include "roms_oracle.hpp"

namespace orc {

// ROMS/Nonlinear/set_massflux.F:140-174
void set_massflux(Model& m, const Bnd& b) {
  ORC_UNPACK_BOUNDS(b);
  const int N = m.c.N; F3 u = m.u[m.nrhs], v = m.v[m.nrhs];
  for (int k = 1; k <= N; ++k) {
    for (int j = JstrT; j <= JendT; ++j)
      for (int i = IstrP; i <= IendT; ++i)
        m.Huon(i, j, k) = 0.5 * (m.Hz(i, j, k) + m.Hz(i - 1, j, k)) * u(i, j, k) * m.on_u(i, j);
    for (int j = JstrP; j <= JendT; ++j)
      for (int i = IstrT; i <= IendT; ++i)
        m.Hvom(i, j, k) = 0.5 * (m.Hz(i, j, k) + m.Hz(i, j - 1, k)) * v(i, j, k) * m.om_v(i, j);
  }
  exchange_u3d(m, b, m.Huon); exchange_v3d(m, b, m.Hvom);
}

// ROMS/Modules/mod_eoscoef.F:24-64
namespace eos {
const double A00 = +1.909256e+04, A01 = +2.098925e+02, A02 = -3.041638e+00, A03 = -1.852732e-03, A04 = -1.361629e-05;
const double B00 = +1.044077e+02, B01 = -6.500517e+00, B02 = +1.553190e-01, B03 = +2.326469e-04;
const double D00 = -5.587545e+00, D01 = +7.390729e-01, D02 = -1.909078e-02;
const double E00 = +4.721788e-01, E01 = +1.028859e-02, E02 = -2.512549e-04, E03 = -5.939910e-07;
const double F00 = -1.571896e-02, F01 = -2.598241e-04, F02 = +7.267926e-06;
const double G00 = +2.042967e-03, G01 = +1.045941e-05, G02 = -5.782165e-10, G03 = +1.296821e-07;
const double H00 = -2.595994e-07, H01 = -1.248266e-09, H02 = -3.508914e-09;
const double Q00 = +9.99842594e+02, Q01 = +6.793952e-02, Q02 = -9.095290e-03, Q03 = +1.001685e-04, Q04 = -1.120083e-06, Q05 = +6.536332e-09;
const double U00 = +8.24493e-01, U01 = -4.08990e-03, U02 = +7.64380e-05, U03 = -8.24670e-07, U04 = +5.38750e-09;
const double V00 = -5.72466e-03, V01 = +1.02270e-04, V02 = -1.65460e-06;
const double W00 = +4.8314e-04;
}  // namespace eos

// ROMS/Nonlinear/rho_eos.F:259-343 for one point (Tt,Ts already clipped by the caller as in :259-261).
// den = in-situ density (kg/m3, NOT anomaly), den1 = density at the surface pressure, bulk = secant bulk modulus.
// extra: [0..2] bulk0, bulk1, bulk2 (BV_FREQUENCY :402-418), [3..6] Dden1DS, Dden1DT, DbulkDS, DbulkDT (:290-294, :330-335)
void eos_point_x(double Tt, double Ts, double Tp, double* den, double* den1, double* bulk, double* extra);
void eos_point(double Tt, double Ts, double Tp, double* den, double* den1, double* bulk) { eos_point_x(Tt, Ts, Tp, den, den1, bulk, nullptr); }
void eos_point_x(double Tt, double Ts, double Tp, double* den, double* den1, double* bulk, double* extra) {
  using namespace eos;
  double C[10];
  double sqrtTs = std::sqrt(Ts);
  double Tpr10 = 0.1 * Tp;
  C[0] = Q00 + Tt * (Q01 + Tt * (Q02 + Tt * (Q03 + Tt * (Q04 + Tt * Q05))));
  C[1] = U00 + Tt * (U01 + Tt * (U02 + Tt * (U03 + Tt * U04)));
  C[2] = V00 + Tt * (V01 + Tt * V02);
  double d1 = C[0] + Ts * (C[1] + sqrtTs * C[2] + Ts * W00);
  C[3] = A00 + Tt * (A01 + Tt * (A02 + Tt * (A03 + Tt * A04)));
  C[4] = B00 + Tt * (B01 + Tt * (B02 + Tt * B03));
  C[5] = D00 + Tt * (D01 + Tt * D02);
  C[6] = E00 + Tt * (E01 + Tt * (E02 + Tt * E03));
  C[7] = F00 + Tt * (F01 + Tt * F02);
  C[8] = G01 + Tt * (G02 + Tt * G03);
  C[9] = H00 + Tt * (H01 + Tt * H02);
  double bulk0 = C[3] + Ts * (C[4] + sqrtTs * C[5]);
  double bulk1 = C[6] + Ts * (C[7] + sqrtTs * G00);
  double bulk2 = C[8] + Ts * C[9];
  double bk = bulk0 - Tp * (bulk1 - Tp * bulk2);
  double cff = 1.0 / (bk + Tpr10);
  *den1 = d1; *bulk = bk; *den = d1 * bk * cff;
  if (extra) {
    double dCdT[10];
    dCdT[0] = Q01 + Tt * (2.0 * Q02 + Tt * (3.0 * Q03 + Tt * (4.0 * Q04 + Tt * 5.0 * Q05)));      // :279-282
    dCdT[1] = U01 + Tt * (2.0 * U02 + Tt * (3.0 * U03 + Tt * 4.0 * U04));
    dCdT[2] = V01 + Tt * 2.0 * V02;
    dCdT[3] = A01 + Tt * (2.0 * A02 + Tt * (3.0 * A03 + Tt * 4.0 * A04));                         // :313-319
    dCdT[4] = B01 + Tt * (2.0 * B02 + Tt * 3.0 * B03);
    dCdT[5] = D01 + Tt * 2.0 * D02;
    dCdT[6] = E01 + Tt * (2.0 * E02 + Tt * 3.0 * E03);
    dCdT[7] = F01 + Tt * 2.0 * F02;
    dCdT[8] = G02 + Tt * 2.0 * G03;
    dCdT[9] = H01 + Tt * 2.0 * H02;
    extra[0] = bulk0; extra[1] = bulk1; extra[2] = bulk2;
    extra[3] = C[1] + 1.5 * C[2] * sqrtTs + 2.0 * W00 * Ts;                                       // Dden1DS :293
    extra[4] = dCdT[0] + Ts * (dCdT[1] + sqrtTs * dCdT[2]);                                       // Dden1DT :294
    extra[5] = C[4] + sqrtTs * 1.5 * C[5] - Tp * (C[7] + sqrtTs * 1.5 * G00 - Tp * C[9]);        // DbulkDS :335-336
    extra[6] = dCdT[3] + Ts * (dCdT[4] + sqrtTs * dCdT[5]) - Tp * (dCdT[6] + Ts * dCdT[7] - Tp * (dCdT[8] + Ts * dCdT[9]));   // DbulkDT :337-339
  }
}

// ROMS/Nonlinear/rho_eos.F: nonlinear :252-483 (+ exchanges :489-526), linear :696-799 (+ :805-842)
void rho_eos(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const int nrhs = m.nrhs;
  F3 T = m.t[nrhs][c.itemp - 1];
  F3 S = (c.salinity && c.NT >= 2) ? m.t[nrhs][c.isalt - 1] : F3();
  SK den(IminS, ImaxS, 1, N), den1(IminS, ImaxS, 1, N);
  SK bulk(IminS, ImaxS, 1, N), bulk0(IminS, ImaxS, 1, N), bulk1(IminS, ImaxS, 1, N), bulk2(IminS, ImaxS, 1, N);
  SK Dden1DS(IminS, ImaxS, 1, N), Dden1DT(IminS, ImaxS, 1, N), DbulkDS(IminS, ImaxS, 1, N), DbulkDT(IminS, ImaxS, 1, N);
  const bool extras = c.bv_frequency || c.eos_tderivative;
  for (int j = JstrT; j <= JendT; ++j) {
    if (c.nonlin_eos) {
      for (int k = 1; k <= N; ++k)
        for (int i = IstrT; i <= IendT; ++i) {
          double Tt = std::max(-2.0, T(i, j, k));
          double Ts = c.salinity ? std::max(0.0, S(i, j, k)) : 0.0;
          double Tp = m.z_r(i, j, k);
          double d, d1, bk, ex[7];
          eos_point_x(Tt, Ts, Tp, &d, &d1, &bk, extras ? ex : nullptr);
          den1(i, k) = d1;
          den(i, k) = d - 1000.0;
          if (extras) {
            bulk(i, k) = bk; bulk0(i, k) = ex[0]; bulk1(i, k) = ex[1]; bulk2(i, k) = ex[2];
            Dden1DS(i, k) = ex[3]; Dden1DT(i, k) = ex[4]; DbulkDS(i, k) = ex[5]; DbulkDT(i, k) = ex[6];
          }
        }
    } else {
      for (int k = 1; k <= N; ++k)
        for (int i = IstrT; i <= IendT; ++i) {
          double r = c.R0 - c.R0 * c.Tcoef * (T(i, j, k) - c.T0);
          if (c.salinity) r = r + c.R0 * c.Scoef * (S(i, j, k) - c.S0);
          r = r - 1000.0;
          den(i, k) = r;
        }
    }
    // VAR_RHO_2D (globaldefs.h:491-495): vertical averages for the barotropic pressure gradient
    for (int i = IstrT; i <= IendT; ++i) {
      double cff1 = den(i, N) * m.Hz(i, j, N);
      m.rhoS(i, j) = 0.5 * cff1 * m.Hz(i, j, N);
      m.rhoA(i, j) = cff1;
    }
    for (int k = N - 1; k >= 1; --k)
      for (int i = IstrT; i <= IendT; ++i) {
        double cff1 = den(i, k) * m.Hz(i, j, k);
        m.rhoS(i, j) = m.rhoS(i, j) + m.Hz(i, j, k) * (m.rhoA(i, j) + 0.5 * cff1);
        m.rhoA(i, j) = m.rhoA(i, j) + cff1;
      }
    double cff2 = 1.0 / c.rho0;
    for (int i = IstrT; i <= IendT; ++i) {
      double cff1 = 1.0 / (m.z_w(i, j, N) - m.z_w(i, j, 0));
      m.rhoA(i, j) = cff2 * cff1 * m.rhoA(i, j);
      m.rhoS(i, j) = 2.0 * cff1 * cff1 * cff2 * m.rhoS(i, j);
    }
    if (c.bv_frequency) {
      if (c.nonlin_eos) {                                                               // :402-418
        for (int k = 1; k <= N - 1; ++k)
          for (int i = IstrT; i <= IendT; ++i) {
            const double zw = m.z_w(i, j, k);
            double bulk_up = bulk0(i, k + 1) - zw * (bulk1(i, k + 1) - bulk2(i, k + 1) * zw);
            double bulk_dn = bulk0(i, k) - zw * (bulk1(i, k) - bulk2(i, k) * zw);
            double cff1 = 1.0 / (bulk_up + 0.1 * zw);
            double cff2 = 1.0 / (bulk_dn + 0.1 * zw);
            double den_up = cff1 * (den1(i, k + 1) * bulk_up);
            double den_dn = cff2 * (den1(i, k) * bulk_dn);
            m.bvf(i, j, k) = -c.g * (den_up - den_dn) / (0.5 * (den_up + den_dn) * (m.z_r(i, j, k + 1) - m.z_r(i, j, k)));
          }
        for (int i = IstrT; i <= IendT; ++i) { m.bvf(i, j, 0) = 0.0; m.bvf(i, j, N) = 0.0; }
      } else {                                                                          // :751-758
        const double gorho0 = c.g / c.rho0;
        for (int k = 1; k <= N - 1; ++k)
          for (int i = IstrT; i <= IendT; ++i)
            m.bvf(i, j, k) = -gorho0 * (den(i, k + 1) - den(i, k)) / (m.z_r(i, j, k + 1) - m.z_r(i, j, k));
      }
    }
    if (c.eos_tderivative) {
      if (c.nonlin_eos) {                                                               // :440-462 (no LMD_DDMIX: k = N only)
        for (int i = IstrT; i <= IendT; ++i) {
          const int k = N;
          double Tpr10 = 0.1 * m.z_r(i, j, k);
          double cff = bulk(i, k) + Tpr10;
          double cff1 = Tpr10 * den1(i, k);
          double cff2 = bulk(i, k) * cff;
          double wrk = (den(i, k) + 1000.0) * cff * cff;
          double Tcof = -(DbulkDT(i, k) * cff1 + Dden1DT(i, k) * cff2);
          double Scof = (DbulkDS(i, k) * cff1 + Dden1DS(i, k) * cff2);
          double cf = 1.0 / wrk;
          m.alpha(i, j) = cf * Tcof;
          m.beta(i, j) = cf * Scof;
        }
      } else {                                                                          // :766-773
        for (int i = IstrT; i <= IendT; ++i) { m.alpha(i, j) = std::fabs(c.Tcoef); m.beta(i, j) = c.salinity ? std::fabs(c.Scoef) : 0.0; }
      }
    }
    for (int k = 1; k <= N; ++k)
      for (int i = IstrT; i <= IendT; ++i) {
        m.rho(i, j, k) = den(i, k);
        m.pden(i, j, k) = c.nonlin_eos ? (den1(i, k) - 1000.0) : den(i, k);
      }
  }
  exchange_r3d(m, b, m.rho); exchange_r3d(m, b, m.pden);
  exchange_r2d(m, b, m.rhoA); exchange_r2d(m, b, m.rhoS);
  if (c.bv_frequency) exchange_w3d(m, b, m.bvf);                                        // :499-503
  if (c.eos_tderivative) { exchange_r2d(m, b, m.alpha); exchange_r2d(m, b, m.beta); }    // :512-517
}

// ROMS/Nonlinear/set_vbc.F: tracer fluxes :278-283, :340-355; quadratic drag :591-624; linear drag :629-652; BCs :657-662
void set_vbc(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N, nrhs = m.nrhs;
  F3 u = m.u[nrhs], v = m.v[nrhs];
  const int it = c.itemp - 1;
  for (int j = JstrR; j <= JendR; ++j)
    for (int i = IstrR; i <= IendR; ++i) { m.stflx[it](i, j) = m.stflux[it](i, j); m.btflx[it](i, j) = m.btflux[it](i, j); }
  if (c.qcorrection) {                                            // QCORRECTION :285-299
    F3 T = m.t[nrhs][it];
    for (int j = JstrR; j <= JendR; ++j)
      for (int i = IstrR; i <= IendR; ++i) m.stflx[it](i, j) = m.stflx[it](i, j) + m.dqdt(i, j) * (T(i, j, N) - m.sst(i, j));
  }
  if (c.limit_stflx_cooling) {                                    // LIMIT_STFLX_COOLING :301-328
    F3 T = m.t[nrhs][it];
    const double cff1 = -2.0;
    for (int j = JstrR; j <= JendR; ++j)
      for (int i = IstrR; i <= IendR; ++i) {
        const double cff2 = m.stflx[it](i, j);
        const double cff3 = 0.5 * (1.0 + std::copysign(1.0, cff1 - T(i, j, N)));
        m.stflx[it](i, j) = cff2 - cff3 * 0.5 * (cff2 - std::fabs(cff2));
      }
  }
  if (c.salinity && c.NT >= 2) {
    const int is = c.isalt - 1; F3 S = m.t[nrhs][is];
    for (int j = JstrR; j <= JendR; ++j)
      for (int i = IstrR; i <= IendR; ++i) {
        double EmP = m.stflux[is](i, j);
        if (c.scorrection == 1)                                   // SCORRECTION :344-347
          m.stflx[is](i, j) = EmP * S(i, j, N) - c.Tnudg_salt * m.Hz(i, j, N) * (S(i, j, N) - m.sss(i, j));
        else if (c.scorrection == 2)                              // SRELAXATION :348-350
          m.stflx[is](i, j) = -c.Tnudg_salt * m.Hz(i, j, N) * (S(i, j, N) - m.sss(i, j));
        else
        m.stflx[is](i, j) = EmP * S(i, j, N);
        m.btflx[is](i, j) = m.btflx[is](i, j) * S(i, j, 1);
      }
  }
  if (c.uv_qdrag == 2) {                                          // UV_LOGDRAG, set_vbc.F:541-586 (no LIMIT_BSTRESS)
    const double vonKar = 0.41, Cdb_min = 0.000001, Cdb_max = 0.5;  // mod_scalars.F:444, :747-748
    S2 wrk(IminS, ImaxS, JminS, JmaxS);
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        const double cff1 = 1.0 / std::log((m.z_r(i, j, 1) - m.z_w(i, j, 0)) / m.ZoBot(i, j));
        const double cff2 = vonKar * vonKar * cff1 * cff1;
        wrk(i, j) = std::min(Cdb_max, std::max(Cdb_min, cff2));
      }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        const double cff1 = 0.25 * (v(i, j, 1) + v(i, j + 1, 1) + v(i - 1, j, 1) + v(i - 1, j + 1, 1));
        const double cff2 = std::sqrt(u(i, j, 1) * u(i, j, 1) + cff1 * cff1);
        m.bustr(i, j) = 0.5 * (wrk(i - 1, j) + wrk(i, j)) * u(i, j, 1) * cff2;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        const double cff1 = 0.25 * (u(i, j, 1) + u(i + 1, j, 1) + u(i, j - 1, 1) + u(i + 1, j - 1, 1));
        const double cff2 = std::sqrt(cff1 * cff1 + v(i, j, 1) * v(i, j, 1));
        m.bvstr(i, j) = 0.5 * (wrk(i, j - 1) + wrk(i, j)) * v(i, j, 1) * cff2;
      }
  } else if (c.uv_qdrag) {
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff1 = 0.25 * (v(i, j, 1) + v(i, j + 1, 1) + v(i - 1, j, 1) + v(i - 1, j + 1, 1));
        double cff2 = std::sqrt(u(i, j, 1) * u(i, j, 1) + cff1 * cff1);
        m.bustr(i, j) = 0.5 * (m.rdrag2(i - 1, j) + m.rdrag2(i, j)) * u(i, j, 1) * cff2;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff1 = 0.25 * (u(i, j, 1) + u(i + 1, j, 1) + u(i, j - 1, 1) + u(i + 1, j - 1, 1));
        double cff2 = std::sqrt(cff1 * cff1 + v(i, j, 1) * v(i, j, 1));
        m.bvstr(i, j) = 0.5 * (m.rdrag2(i, j - 1) + m.rdrag2(i, j)) * v(i, j, 1) * cff2;
      }
  } else {
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) m.bustr(i, j) = 0.5 * (m.rdrag(i - 1, j) + m.rdrag(i, j)) * u(i, j, 1);
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) m.bvstr(i, j) = 0.5 * (m.rdrag(i, j - 1) + m.rdrag(i, j)) * v(i, j, 1);
  }
  if (c.limit_bstress) {                                          // LIMIT_BSTRESS (set_vbc.F:533-540, :562-567, :579-584, :600-605, ...):
    const double cff = 0.75 / c.dt;                               // the stress may slow the bottom layer down, not reverse it
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        const double cff3 = cff * 0.5 * (m.Hz(i - 1, j, 1) + m.Hz(i, j, 1));
        m.bustr(i, j) = std::copysign(1.0, m.bustr(i, j)) * std::min(std::fabs(m.bustr(i, j)), std::fabs(u(i, j, 1)) * cff3);
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        const double cff3 = cff * 0.5 * (m.Hz(i, j - 1, 1) + m.Hz(i, j, 1));
        m.bvstr(i, j) = std::copysign(1.0, m.bvstr(i, j)) * std::min(std::fabs(m.bvstr(i, j)), std::fabs(v(i, j, 1)) * cff3);
      }
  }
  bc_u2d(m, b, m.bustr); bc_v2d(m, b, m.bvstr);
}

// ROMS/Nonlinear/omega.F:147-218
void omega(Model& m, const Bnd& b) {
  ORC_UNPACK_BOUNDS(b);
  const int N = m.c.N;
  std::vector<double> wrkv(ImaxS - IminS + 1);
  double* wrk = wrkv.data() - IminS;
  for (int j = Jstr; j <= Jend; ++j) {
    for (int i = Istr; i <= Iend; ++i) m.W(i, j, 0) = 0.0;
    for (int k = 1; k <= N; ++k)
      for (int i = Istr; i <= Iend; ++i)
        m.W(i, j, k) = m.W(i, j, k - 1) - (m.Huon(i + 1, j, k) - m.Huon(i, j, k) + m.Hvom(i, j + 1, k) - m.Hvom(i, j, k));
    for (int i = Istr; i <= Iend; ++i) wrk[i] = m.W(i, j, N) / (m.z_w(i, j, N) - m.z_w(i, j, 0));
    for (int k = N - 1; k >= 1; --k)
      for (int i = Istr; i <= Iend; ++i) m.W(i, j, k) = m.W(i, j, k) - wrk[i] * (m.z_w(i, j, k) - m.z_w(i, j, 0));
    for (int i = Istr; i <= Iend; ++i) m.W(i, j, N) = 0.0;
  }
  bc_w3d(m, b, m.W);
}

// ROMS/Nonlinear/wvelocity.F:142-256
void wvelocity(Model& m, const Bnd& b, int Ninp) {
  ORC_UNPACK_BOUNDS(b);
  const int N = m.c.N;
  F3 u = m.u[Ninp], v = m.v[Ninp];
  F3 &z_r = m.z_r, &z_w = m.z_w, &W = m.W, &wvel = m.wvel; F2 &pm = m.pm, &pn = m.pn;
  exchange_u2d(m, b, m.DU_avg1); exchange_v2d(m, b, m.DV_avg1);
  S2 wrk(IminS, ImaxS, JminS, JmaxS);
  S3 vert(IminS, ImaxS, JminS, JmaxS, 1, N);
  for (int k = 1; k <= N; ++k) {
    for (int j = Jstr; j <= Jend; ++j) {
      for (int i = Istr; i <= Iend + 1; ++i) wrk(i, j) = u(i, j, k) * (z_r(i, j, k) - z_r(i - 1, j, k)) * (pm(i - 1, j) + pm(i, j));
      for (int i = Istr; i <= Iend; ++i) vert(i, j, k) = 0.25 * (wrk(i, j) + wrk(i + 1, j));
    }
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = Istr; i <= Iend; ++i) wrk(i, j) = v(i, j, k) * (z_r(i, j, k) - z_r(i, j - 1, k)) * (pn(i, j - 1) + pn(i, j));
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) vert(i, j, k) = vert(i, j, k) + 0.25 * (wrk(i, j) + wrk(i, j + 1));
  }
  const double cff1 = 3.0 / 8.0, cff2 = 3.0 / 4.0, cff3 = 1.0 / 8.0, cff4 = 9.0 / 16.0, cff5 = 1.0 / 16.0;
  for (int j = Jstr; j <= Jend; ++j) {
    for (int i = Istr; i <= Iend; ++i)
      wrk(i, j) = (m.DU_avg1(i, j) - m.DU_avg1(i + 1, j) + m.DV_avg1(i, j) - m.DV_avg1(i, j + 1)) / (z_w(i, j, N) - z_w(i, j, 0));
    for (int i = Istr; i <= Iend; ++i) {
      double slope = (z_r(i, j, 1) - z_w(i, j, 0)) / (z_r(i, j, 2) - z_r(i, j, 1));
      wvel(i, j, 0) = cff1 * (vert(i, j, 1) - slope * (vert(i, j, 2) - vert(i, j, 1))) + cff2 * vert(i, j, 1) - cff3 * vert(i, j, 2);
      wvel(i, j, 1) = pm(i, j) * pn(i, j) * (W(i, j, 1) + wrk(i, j) * (z_w(i, j, 1) - z_w(i, j, 0))) + cff1 * vert(i, j, 1) +
                      cff2 * vert(i, j, 2) - cff3 * vert(i, j, 3);
    }
    for (int k = 2; k <= N - 2; ++k)
      for (int i = Istr; i <= Iend; ++i)
        wvel(i, j, k) = pm(i, j) * pn(i, j) * (W(i, j, k) + wrk(i, j) * (z_w(i, j, k) - z_w(i, j, 0))) +
                        cff4 * (vert(i, j, k) + vert(i, j, k + 1)) - cff5 * (vert(i, j, k - 1) + vert(i, j, k + 2));
    for (int i = Istr; i <= Iend; ++i) {
      double slope = (z_w(i, j, N) - z_r(i, j, N)) / (z_r(i, j, N) - z_r(i, j, N - 1));
      wvel(i, j, N) = pm(i, j) * pn(i, j) * wrk(i, j) * (z_w(i, j, N) - z_w(i, j, 0)) +
                      cff1 * (vert(i, j, N) + slope * (vert(i, j, N) - vert(i, j, N - 1))) + cff2 * vert(i, j, N) - cff3 * vert(i, j, N - 1);
      wvel(i, j, N - 1) = pm(i, j) * pn(i, j) * (W(i, j, N - 1) + wrk(i, j) * (z_w(i, j, N - 1) - z_w(i, j, 0))) +
                          cff1 * vert(i, j, N) + cff2 * vert(i, j, N - 1) - cff3 * vert(i, j, N - 2);
    }
  }
  bc_w3d(m, b, wvel);
}

// ROMS/Nonlinear/set_zeta.F:95-109
void set_zeta(Model& m, const Bnd& b) {
  ORC_UNPACK_BOUNDS(b);
  for (int j = JstrR; j <= JendR; ++j)
    for (int i = IstrR; i <= IendR; ++i) { m.zeta[1](i, j) = m.Zt_avg1(i, j); m.zeta[2](i, j) = m.Zt_avg1(i, j); }
  exchange_r2d(m, b, m.zeta[1]); exchange_r2d(m, b, m.zeta[2]);
}

// ROMS/Nonlinear/set_depth.F:160-208 (Vtransform = 1: the original transformation), :210-262 (Vtransform = 2)
void set_depth(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N; const double hc = m.hc;
  if (c.Vtransform == 1) {
    for (int j = JstrT; j <= JendT; ++j) {
      for (int i = IstrT; i <= IendT; ++i) m.z_w(i, j, 0) = -m.h(i, j);
      for (int k = 1; k <= N; ++k) {
        const double cff_r = hc * (m.sc_r[k] - m.Cs_r[k]), cff_w = hc * (m.sc_w[k] - m.Cs_w[k]);
        const double cff1_r = m.Cs_r[k], cff1_w = m.Cs_w[k];
        for (int i = IstrT; i <= IendT; ++i) {
          const double hwater = m.h(i, j);
          const double hinv = 1.0 / hwater;
          const double z_w0 = cff_w + cff1_w * hwater;
          m.z_w(i, j, k) = z_w0 + m.Zt_avg1(i, j) * (1.0 + z_w0 * hinv);
          const double z_r0 = cff_r + cff1_r * hwater;
          m.z_r(i, j, k) = z_r0 + m.Zt_avg1(i, j) * (1.0 + z_r0 * hinv);
          m.Hz(i, j, k) = m.z_w(i, j, k) - m.z_w(i, j, k - 1);
        }
      }
    }
    exchange_r2d(m, b, m.h); exchange_w3d(m, b, m.z_w); exchange_r3d(m, b, m.z_r); exchange_r3d(m, b, m.Hz);
    return;
  }
  for (int j = JstrT; j <= JendT; ++j) {
    for (int i = IstrT; i <= IendT; ++i) m.z_w(i, j, 0) = -m.h(i, j);
    for (int k = 1; k <= N; ++k) {
      double cff_r = hc * m.sc_r[k], cff_w = hc * m.sc_w[k];
      double cff1_r = m.Cs_r[k], cff1_w = m.Cs_w[k];
      for (int i = IstrT; i <= IendT; ++i) {
        double hwater = m.h(i, j);
        double hinv = 1.0 / (hc + hwater);
        double cff2_r = (cff_r + cff1_r * hwater) * hinv;
        double cff2_w = (cff_w + cff1_w * hwater) * hinv;
        m.z_w(i, j, k) = m.Zt_avg1(i, j) + (m.Zt_avg1(i, j) + hwater) * cff2_w;
        m.z_r(i, j, k) = m.Zt_avg1(i, j) + (m.Zt_avg1(i, j) + hwater) * cff2_r;
        m.Hz(i, j, k) = m.z_w(i, j, k) - m.z_w(i, j, k - 1);
      }
    }
  }
  exchange_r2d(m, b, m.h); exchange_w3d(m, b, m.z_w); exchange_r3d(m, b, m.z_r); exchange_r3d(m, b, m.Hz);
}

}  // namespace orc

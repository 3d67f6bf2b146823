// Glue kernels of the main3d chain: set_massflux, rho_eos, set_vbc, omega, wvelocity, set_zeta, set_depth, ana_vmix.
// All are HBM-bound streaming or thread-per-column kernels; xi (i) is the coalesced axis.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

// ---------------------------------------------------------------------------------------------------------------
// set_massflux_tile (ROMS/Nonlinear/set_massflux.F:140-174).  One thread per (i,j,k).
__global__ void __launch_bounds__(256) k_set_massflux(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P, o = o2 + k * p.PL;
  const double* __restrict__ Hz = f.Hz;
  const double hz = Hz[o + i];
  {
    const double x = 0.5 * (hz + Hz[o + i - 1]) * f.u[p.nrhs][o + i] * f.on_u[o2 + i];
    st_w(f.Huon, o, i, x, p);
  }
  if (j >= 1) {
    const double x = 0.5 * (hz + Hz[o - p.P + i]) * f.v[p.nrhs][o + i] * f.om_v[o2 + i];
    st_w(f.Hvom, o, i, x, p);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// rho_eos_tile: nonlinear (ROMS/Nonlinear/rho_eos.F:252-483, coefficients mod_eoscoef.F:24-64) and linear (:696-799).
// One thread per column, top-down so that the VAR_RHO_2D integrals rhoA/rhoS accumulate in registers.
// X: also return what BV_FREQUENCY (:402-418) and the expansion coefficients (:290-294, :330-339, :440-462) need
// X = 1: the bulk-modulus terms bvf needs; X = 2: also the temperature / salinity derivatives (needed at level N only)
struct EosX { double bulk, bulk0, bulk1, bulk2, Dden1DS, Dden1DT, DbulkDS, DbulkDT; };
template <int X>
__device__ __forceinline__ void eos_nl(double Tt, double Ts, double Tp, double& den, double& den1, EosX* x = nullptr) {
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
  const double sqrtTs = sqrt(Ts);
  const double Tpr10 = 0.1 * Tp;
  const double C0 = Q00 + Tt * (Q01 + Tt * (Q02 + Tt * (Q03 + Tt * (Q04 + Tt * Q05))));
  const double C1 = U00 + Tt * (U01 + Tt * (U02 + Tt * (U03 + Tt * U04)));
  const double C2 = V00 + Tt * (V01 + Tt * V02);
  den1 = C0 + Ts * (C1 + sqrtTs * C2 + Ts * W00);
  const double C3 = A00 + Tt * (A01 + Tt * (A02 + Tt * (A03 + Tt * A04)));
  const double C4 = B00 + Tt * (B01 + Tt * (B02 + Tt * B03));
  const double C5 = D00 + Tt * (D01 + Tt * D02);
  const double C6 = E00 + Tt * (E01 + Tt * (E02 + Tt * E03));
  const double C7 = F00 + Tt * (F01 + Tt * F02);
  const double C8 = G01 + Tt * (G02 + Tt * G03);
  const double C9 = H00 + Tt * (H01 + Tt * H02);
  const double bulk0 = C3 + Ts * (C4 + sqrtTs * C5);
  const double bulk1 = C6 + Ts * (C7 + sqrtTs * G00);
  const double bulk2 = C8 + Ts * C9;
  const double bulk = bulk0 - Tp * (bulk1 - Tp * bulk2);
  const double cff = 1.0 / (bulk + Tpr10);
  den = den1 * bulk * cff;
  if (X >= 1) { x->bulk = bulk; x->bulk0 = bulk0; x->bulk1 = bulk1; x->bulk2 = bulk2; }
  if (X >= 2) {
    const double dC0 = Q01 + Tt * (2.0 * Q02 + Tt * (3.0 * Q03 + Tt * (4.0 * Q04 + Tt * 5.0 * Q05)));
    const double dC1 = U01 + Tt * (2.0 * U02 + Tt * (3.0 * U03 + Tt * 4.0 * U04));
    const double dC2 = V01 + Tt * 2.0 * V02;
    const double dC3 = A01 + Tt * (2.0 * A02 + Tt * (3.0 * A03 + Tt * 4.0 * A04));
    const double dC4 = B01 + Tt * (2.0 * B02 + Tt * 3.0 * B03);
    const double dC5 = D01 + Tt * 2.0 * D02;
    const double dC6 = E01 + Tt * (2.0 * E02 + Tt * 3.0 * E03);
    const double dC7 = F01 + Tt * 2.0 * F02;
    const double dC8 = G02 + Tt * 2.0 * G03;
    const double dC9 = H01 + Tt * 2.0 * H02;
    x->Dden1DS = C1 + 1.5 * C2 * sqrtTs + 2.0 * W00 * Ts;
    x->Dden1DT = dC0 + Ts * (dC1 + sqrtTs * dC2);
    x->DbulkDS = C4 + sqrtTs * 1.5 * C5 - Tp * (C7 + sqrtTs * 1.5 * G00 - Tp * C9);
    x->DbulkDT = dC3 + Ts * (dC4 + sqrtTs * dC5) - Tp * (dC6 + Ts * dC7 - Tp * (dC8 + Ts * dC9));
  }
}

#ifndef EOS_PF
#define EOS_PF 4          // L2 prefetch distance (levels) of k_rho_eos<true>
#endif
#ifndef EOS_MINB
#define EOS_MINB 4          // 64 registers: rho_eos<true> 0.35 -> 0.32 ms (3: 0.35, unbounded: 0.45); the plain kernel keeps its 40 (6 CTAs per SM)
#endif
template <bool X>     // X: with the optional outputs bvf / alpha, beta (BV_FREQUENCY; LMD_SKPP || BULK_FLUXES)
__global__ void __launch_bounds__(256, X ? EOS_MINB : 6) k_rho_eos(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // JstrT..JendT = 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P;
  const double* __restrict__ T = f.t[p.nrhs][p.itemp - 1];
  const double* __restrict__ S = (p.salinity && p.NT >= 2) ? f.t[p.nrhs][p.isalt - 1] : nullptr;
  const double* __restrict__ Hz = f.Hz;
  double rhoA = 0.0, rhoS = 0.0;
  EosX up; double den1_up = 0.0, den_up_lin = 0.0, zr_up = 0.0;      // level k+1 of the downward march (bvf at W-point k)
  for (int k = p.N; k >= 1; --k) {
    const int o = o2 + k * p.PL;
    if (X) {               // the variant with bvf runs at half the occupancy of the plain one: start the fetch of lower levels early
      pf_dn<EOS_PF>(T, o + i, k, p.PL); pf_dn<EOS_PF>(f.z_r, o + i, k, p.PL); pf_dn<EOS_PF>(Hz, o + i, k, p.PL); pf_dn<EOS_PF>(f.z_w, o + i, k, p.PL);
      if (S) pf_dn<EOS_PF>(S, o + i, k, p.PL);
    }
    double den, pd;
    EosX cur; double d1c = 0.0;
    const double zr = f.z_r[o + i];
    if (p.nonlin_eos) {
      const double Tt = dmax(-2.0, T[o + i]);
      const double Ts = S ? dmax(0.0, S[o + i]) : 0.0;
      double d, d1;
      eos_nl<X ? 2 : 0>(Tt, Ts, zr, d, d1, &cur);
      den = d - 1000.0;
      pd = d1 - 1000.0;
      d1c = d1;
      if (X && p.eos_tderivative && k == p.N) {                            // :440-462 (no LMD_DDMIX: level N only)
        const double Tpr10 = 0.1 * zr;
        const double cff = cur.bulk + Tpr10;
        const double cff1 = Tpr10 * d1;
        const double cff2 = cur.bulk * cff;
        const double wrk = (den + 1000.0) * cff * cff;
        const double Tcof = -(cur.DbulkDT * cff1 + cur.Dden1DT * cff2);
        const double Scof = (cur.DbulkDS * cff1 + cur.Dden1DS * cff2);
        const double cf = 1.0 / wrk;
        st_w(f.alpha, o2, i, cf * Tcof, p);
        st_w(f.beta, o2, i, cf * Scof, p);
      }
    } else {
      double r = p.R0 - p.R0 * p.Tcoef * (T[o + i] - p.T0);
      if (S) r = r + p.R0 * p.Scoef * (S[o + i] - p.S0);
      r = r - 1000.0;
      den = r; pd = r;
    }
    if (X && p.bv_frequency && k < p.N) {                                    // bvf at W-point k from levels k+1 (up) and k (dn)
      double bv;
      if (p.nonlin_eos) {                                                    // :402-418
        const double zw = f.z_w[o + i];
        const double bulk_up = up.bulk0 - zw * (up.bulk1 - up.bulk2 * zw);
        const double bulk_dn = cur.bulk0 - zw * (cur.bulk1 - cur.bulk2 * zw);
        const double cff1 = 1.0 / (bulk_up + 0.1 * zw);
        const double cff2 = 1.0 / (bulk_dn + 0.1 * zw);
        const double den_up = cff1 * (den1_up * bulk_up);
        const double den_dn = cff2 * (d1c * bulk_dn);
        bv = -p.g * (den_up - den_dn) / (0.5 * (den_up + den_dn) * (zr_up - zr));
      } else {                                                               // :751-758
        const double gorho0 = p.g / p.rho0;
        bv = -gorho0 * (den_up_lin - den) / (zr_up - zr);
      }
      st_w(f.bvf, o, i, bv, p);
    }
    if (X) { up = cur; den1_up = d1c; den_up_lin = den; zr_up = zr; }
    st_w(f.rho, o, i, den, p);
    st_w(f.pden, o, i, pd, p);
    const double hz = Hz[o + i];
    const double cff1 = den * hz;
    if (k == p.N) { rhoS = 0.5 * cff1 * hz; rhoA = cff1; }
    else { rhoS = rhoS + hz * (rhoA + 0.5 * cff1); rhoA = rhoA + cff1; }
  }
  const double cff2 = 1.0 / p.rho0;
  const double cff1 = 1.0 / (f.z_w[o2 + p.N * p.PL + i] - f.z_w[o2 + i]);
  st_w(f.rhoA, o2, i, cff2 * cff1 * rhoA, p);
  st_w(f.rhoS, o2, i, 2.0 * cff1 * cff1 * cff2 * rhoS, p);
  if (X) {
    if (p.bv_frequency && p.nonlin_eos) { st_w(f.bvf, o2, i, 0.0, p); st_w(f.bvf, o2 + p.N * p.PL, i, 0.0, p); }   // :414-417
    if (p.eos_tderivative && !p.nonlin_eos) {                                // :766-773
      st_w(f.alpha, o2, i, fabs(p.Tcoef), p);
      st_w(f.beta, o2, i, S ? fabs(p.Scoef) : 0.0, p);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// UV_LOGDRAG drag coefficient of a cell (set_vbc.F:545-551; vonKar, Cdb_min, Cdb_max: mod_scalars.F:444, :747-748)
__device__ __forceinline__ double logdrag_cd(const Flds& f, int q1, int q0, int q2d) {
  const double cff1 = 1.0 / log((f.z_r[q1] - f.z_w[q0]) / f.ZoBot[q2d]);
  const double cff2 = 0.41 * 0.41 * cff1 * cff1;
  return dmin(0.5, dmax(0.000001, cff2));
}
// set_vbc_tile (ROMS/Nonlinear/set_vbc.F:278-283, :340-355, :541-586 logarithmic, :591-624 quadratic, :629-652 linear, BCs :657-662)
__global__ void __launch_bounds__(256) k_set_vbc(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P, o1 = o2 + p.PL;                       // k = 1
  {
    const int it = p.itemp - 1;
    double sf = f.stflux[it][o2 + i];
    if (p.qcorrection) sf = sf + f.dqdt[o2 + i] * (f.t[p.nrhs][it][o2 + p.N * p.PL + i] - f.sst[o2 + i]);       // QCORRECTION :285-299
    if (p.limit_stflx_cooling) {                                                                             // LIMIT_STFLX_COOLING :301-328
      const double cff3 = 0.5 * (1.0 + copysign(1.0, -2.0 - f.t[p.nrhs][it][o2 + p.N * p.PL + i]));
      sf = sf - cff3 * 0.5 * (sf - fabs(sf));
    }
    f.stflx[it][o2 + i] = sf;
    f.btflx[it][o2 + i] = f.btflux[it][o2 + i];
    if (p.salinity && p.NT >= 2) {
      const int is = p.isalt - 1;
      const double* __restrict__ S = f.t[p.nrhs][is];
      const double EmP = f.stflux[is][o2 + i];
      const int oN = o2 + p.N * p.PL + i;
      if (p.scorrection == 1) f.stflx[is][o2 + i] = EmP * S[oN] - p.Tnudg_salt * f.Hz[oN] * (S[oN] - f.sss[o2 + i]);   // SCORRECTION :344-347
      else if (p.scorrection == 2) f.stflx[is][o2 + i] = -p.Tnudg_salt * f.Hz[oN] * (S[oN] - f.sss[o2 + i]);           // SRELAXATION :348-350
      else
      f.stflx[is][o2 + i] = EmP * S[o2 + p.N * p.PL + i];
      f.btflx[is][o2 + i] = f.btflx[is][o2 + i] * S[o1 + i];
    }
  }
  const double* __restrict__ u = f.u[p.nrhs];
  const double* __restrict__ v = f.v[p.nrhs];
  if (j >= 1 && j <= p.Mm) {
    double bu;
    if (p.uv_qdrag == 2) {
      const double cff1 = 0.25 * (v[o1 + i] + v[o1 + p.P + i] + v[o1 + i - 1] + v[o1 + p.P + i - 1]);
      const double uu = u[o1 + i];
      const double cff2 = sqrt(uu * uu + cff1 * cff1);
      bu = 0.5 * (logdrag_cd(f, o1 + i - 1, o2 + i - 1, o2 + i - 1) + logdrag_cd(f, o1 + i, o2 + i, o2 + i)) * uu * cff2;
    } else if (p.uv_qdrag) {
      const double cff1 = 0.25 * (v[o1 + i] + v[o1 + p.P + i] + v[o1 + i - 1] + v[o1 + p.P + i - 1]);
      const double uu = u[o1 + i];
      const double cff2 = sqrt(uu * uu + cff1 * cff1);
      bu = 0.5 * (f.rdrag2[o2 + i - 1] + f.rdrag2[o2 + i]) * uu * cff2;
    } else {
      bu = 0.5 * (f.rdrag[o2 + i - 1] + f.rdrag[o2 + i]) * u[o1 + i];
    }
    if (p.limit_bstress) {                                       // set_vbc.F:562-567 / :600-605 / :633-638
      const double cff3 = (0.75 / p.dt) * 0.5 * (f.Hz[o1 + i - 1] + f.Hz[o1 + i]);
      bu = copysign(1.0, bu) * dmin(fabs(bu), fabs(u[o1 + i]) * cff3);
    }
    st_u_closed(f.bustr, o2, i, j, bu, p);
  }
  if (j >= 2 && j <= p.Mm) {
    double bv;
    if (p.uv_qdrag == 2) {
      const double cff1 = 0.25 * (u[o1 + i] + u[o1 + i + 1] + u[o1 - p.P + i] + u[o1 - p.P + i + 1]);
      const double vv = v[o1 + i];
      const double cff2 = sqrt(cff1 * cff1 + vv * vv);
      bv = 0.5 * (logdrag_cd(f, o1 - p.P + i, o2 - p.P + i, o2 - p.P + i) + logdrag_cd(f, o1 + i, o2 + i, o2 + i)) * vv * cff2;
    } else if (p.uv_qdrag) {
      const double cff1 = 0.25 * (u[o1 + i] + u[o1 + i + 1] + u[o1 - p.P + i] + u[o1 - p.P + i + 1]);
      const double vv = v[o1 + i];
      const double cff2 = sqrt(cff1 * cff1 + vv * vv);
      bv = 0.5 * (f.rdrag2[o2 - p.P + i] + f.rdrag2[o2 + i]) * vv * cff2;
    } else {
      bv = 0.5 * (f.rdrag[o2 - p.P + i] + f.rdrag[o2 + i]) * v[o1 + i];
    }
    if (p.limit_bstress) {
      const double cff3 = (0.75 / p.dt) * 0.5 * (f.Hz[o1 - p.P + i] + f.Hz[o1 + i]);
      bv = copysign(1.0, bv) * dmin(fabs(bv), fabs(v[o1 + i]) * cff3);
    }
    st_v_closed(f.bvstr, o2, i, j, bv, p);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// omega_tile (ROMS/Nonlinear/omega.F:147-218).  Thread per column: upward prefix sum, then removal of the part
// proportional to the barotropic divergence, then bc_w3d (gradient) + periodic images.
#ifndef GLUE_PF
#define GLUE_PF 0          // L2 prefetch distance (levels) of the streaming column kernels k_omega, k_wvelocity
#endif
// NC > 0: the number of levels is the compile-time constant NC (all BENCHMARK grids have N = 30), so the column of partial sums
// stays in registers; NC = 0 keeps it in a thread-local array (local memory: ~0.5 KB per column of extra L1/L2 traffic).
template <int NC>
__global__ void __launch_bounds__(128) k_omega(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int o2 = j * p.P;
  const int N = NC > 0 ? NC : p.N;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ z_w = f.z_w;
  double Wl[(NC > 0 ? NC : MAXN) + 1];
  double w = 0.0;
  Wl[0] = 0.0;
#pragma unroll
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL;
    pf_up<GLUE_PF>(Huon, o + i, k, N, p.PL); pf_up<GLUE_PF>(Hvom, o + i, k, N, p.PL); pf_up<GLUE_PF>(z_w, o + i, k, N, p.PL);
    w = w - (Huon[o + i + 1] - Huon[o + i] + Hvom[o + p.P + i] - Hvom[o + i]);
    Wl[k] = w;
  }
  const double zw0 = z_w[o2 + i];
  const double wrk = w / (z_w[o2 + N * p.PL + i] - zw0);
  st_r_grad(f.W, o2, i, j, 0.0, p);
#pragma unroll
  for (int k = N - 1; k >= 1; --k) {
    const int o = o2 + k * p.PL;
    const double x = Wl[k] - wrk * (z_w[o + i] - zw0);
    st_r_grad(f.W, o, i, j, x, p);
  }
  st_r_grad(f.W, o2 + N * p.PL, i, j, 0.0, p);
}

// ---------------------------------------------------------------------------------------------------------------
// wvelocity_tile (ROMS/Nonlinear/wvelocity.F:156-256): diagnostic true vertical velocity at W-points.
template <int NC>    // NC > 0: compile-time number of levels, vert() in registers (see k_omega)
__global__ void __launch_bounds__(128) k_wvelocity(Par p, Flds f, int Ninp) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int o2 = j * p.P, P = p.P, N = NC > 0 ? NC : p.N;
  const double* __restrict__ u = f.u[Ninp];
  const double* __restrict__ v = f.v[Ninp];
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ z_w = f.z_w;
  const double* __restrict__ W = f.W;
  const double pmi = f.pm[o2 + i], pni = f.pn[o2 + i];
  const double pmU0 = f.pm[o2 + i - 1] + pmi, pmU1 = pmi + f.pm[o2 + i + 1];
  const double pnV0 = f.pn[o2 - P + i] + pni, pnV1 = pni + f.pn[o2 + P + i];
  double vert[(NC > 0 ? NC : MAXN) + 1];
#pragma unroll
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL;
    pf_up<GLUE_PF>(z_r, o + i, k, N, p.PL); pf_up<GLUE_PF>(u, o + i, k, N, p.PL); pf_up<GLUE_PF>(v, o + i, k, N, p.PL);
    pf_up<GLUE_PF>(W, o + i, k, N, p.PL); pf_up<GLUE_PF>(z_w, o + i, k, N, p.PL);
    const double zr = z_r[o + i];
    const double wu0 = u[o + i] * (zr - z_r[o + i - 1]) * pmU0;
    const double wu1 = u[o + i + 1] * (z_r[o + i + 1] - zr) * pmU1;
    double vt = 0.25 * (wu0 + wu1);
    const double wv0 = v[o + i] * (zr - z_r[o - P + i]) * pnV0;
    const double wv1 = v[o + P + i] * (z_r[o + P + i] - zr) * pnV1;
    vt = vt + 0.25 * (wv0 + wv1);
    vert[k] = vt;
  }
  const double cff1 = 3.0 / 8.0, cff2 = 3.0 / 4.0, cff3 = 1.0 / 8.0, cff4 = 9.0 / 16.0, cff5 = 1.0 / 16.0;
  const double zw0 = z_w[o2 + i], zwN = z_w[o2 + N * p.PL + i];
  const double wrk = (f.DU_avg1[o2 + i] - f.DU_avg1[o2 + i + 1] + f.DV_avg1[o2 + i] - f.DV_avg1[o2 + P + i]) / (zwN - zw0);
  const double pmn = pmi * pni;
  {
    const double slope = (z_r[o2 + p.PL + i] - zw0) / (z_r[o2 + 2 * p.PL + i] - z_r[o2 + p.PL + i]);
    const double w0 = cff1 * (vert[1] - slope * (vert[2] - vert[1])) + cff2 * vert[1] - cff3 * vert[2];
    st_r_grad(f.wvel, o2, i, j, w0, p);
    const int o = o2 + p.PL;
    const double w1 = pmn * (W[o + i] + wrk * (z_w[o + i] - zw0)) + cff1 * vert[1] + cff2 * vert[2] - cff3 * vert[3];
    st_r_grad(f.wvel, o, i, j, w1, p);
  }
#pragma unroll
  for (int k = 2; k <= N - 2; ++k) {
    const int o = o2 + k * p.PL;
    const double x = pmn * (W[o + i] + wrk * (z_w[o + i] - zw0)) + cff4 * (vert[k] + vert[k + 1]) - cff5 * (vert[k - 1] + vert[k + 2]);
    st_r_grad(f.wvel, o, i, j, x, p);
  }
  {
    const int oN = o2 + N * p.PL, oM = oN - p.PL;
    const double slope = (zwN - z_r[oN + i]) / (z_r[oN + i] - z_r[oM + i]);
    const double wN = pmn * wrk * (zwN - zw0) + cff1 * (vert[N] + slope * (vert[N] - vert[N - 1])) + cff2 * vert[N] - cff3 * vert[N - 1];
    st_r_grad(f.wvel, oN, i, j, wN, p);
    const double wM = pmn * (W[oM + i] + wrk * (z_w[oM + i] - zw0)) + cff1 * vert[N] + cff2 * vert[N - 1] - cff3 * vert[N - 2];
    st_r_grad(f.wvel, oM, i, j, wM, p);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// set_zeta_tile (ROMS/Nonlinear/set_zeta.F:95-109)
__global__ void __launch_bounds__(256) k_set_zeta(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P;
  const double z = f.Zt_avg1[o2 + i];
  st_w(f.zeta[1], o2, i, z, p);
  st_w(f.zeta[2], o2, i, z, p);
}

// ---------------------------------------------------------------------------------------------------------------
// set_depth_tile (ROMS/Nonlinear/set_depth.F:210-262, Vtransform = 2)
// VT = 1: the original transformation (set_depth.F:160-208): z = z0 + zeta (1 + z0 / h), z0 = hc (s - C) + C h
template <int VT>
__global__ void __launch_bounds__(256) k_set_depth(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P;
  const double hwater = f.h[o2 + i];
  const double zt = f.Zt_avg1[o2 + i];
  double zw_prev = -hwater;
  st_w(f.z_w, o2, i, zw_prev, p);
  const double hinv = (VT == 1) ? 1.0 / hwater : 1.0 / (p.hc + hwater);
  for (int k = 1; k <= p.N; ++k) {
    const int o = o2 + k * p.PL;
    const double cff1_r = f.Cs_r[k], cff1_w = f.Cs_w[k];
    double zw, zr;
    if (VT == 1) {
      const double cff_r = p.hc * (f.sc_r[k] - cff1_r), cff_w = p.hc * (f.sc_w[k] - cff1_w);
      const double z_w0 = cff_w + cff1_w * hwater;
      zw = z_w0 + zt * (1.0 + z_w0 * hinv);
      const double z_r0 = cff_r + cff1_r * hwater;
      zr = z_r0 + zt * (1.0 + z_r0 * hinv);
    } else {
    const double cff_r = p.hc * f.sc_r[k], cff_w = p.hc * f.sc_w[k];
    const double cff2_r = (cff_r + cff1_r * hwater) * hinv;
    const double cff2_w = (cff_w + cff1_w * hwater) * hinv;
    zw = zt + (zt + hwater) * cff2_w;
    zr = zt + (zt + hwater) * cff2_r;
    }
    st_w(f.z_w, o, i, zw, p);
    st_w(f.z_r, o, i, zr, p);
    st_w(f.Hz, o, i, zw - zw_prev, p);
    zw_prev = zw;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ana_vmix (ROMS/Functionals/ana_vmix.h:200-208 Akv, :327-337 Akt; UPWELLING), k = 1..N-1
__global__ void __launch_bounds__(256) k_ana_vmix(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o = j * p.P + k * p.PL;
  st_w(f.Akv, o, i, 2.0e-3 + 8.0e-3 * exp(f.z_w[o + i] / 150.0), p);
  st_w(f.Akt[p.itemp - 1], o, i, p.Akt_bak[p.itemp - 1], p);
  if (p.salinity && p.NT >= 2) st_w(f.Akt[p.isalt - 1], o, i, p.Akt_bak[p.isalt - 1], p);
}

// bvf_mix_tile (ROMS/Nonlinear/bvf_mix.F:92-127; constants mod_scalars.F:1793-1796) + periodic images (:133-143)
__global__ void __launch_bounds__(256) k_bvf_mix(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int o = j * p.P + k * p.PL;
  const double bvf_numax = 4.0e-4, bvf_numin = 3.0e-5, bvf_nu0 = 1.0e-7, bvf_nu0c = 1.0;
  const double bv = f.bvf[o + i];
  const int it = p.itemp - 1, is = p.isalt - 1;
  double av, at, as;
  if (bv < 0.0) { av = bvf_nu0c; at = bvf_nu0c; as = bvf_nu0c; }
  else if (bv == 0.0) { av = p.Akv_bak; at = p.Akt_bak[it]; as = p.salinity ? p.Akt_bak[is] : 0.0; }
  else {
    const double cff = bvf_nu0 / sqrt(bv);
    at = dmin(bvf_numax, dmax(bvf_numin, cff));
    av = at; as = at;
  }
  st_w(f.Akv, o, i, av, p);
  st_w(f.Akt[it], o, i, at, p);
  if (p.salinity && p.NT >= 2) st_w(f.Akt[is], o, i, as, p);
}

// ---------------------------------------------------------------------------------------------------------------
// set_avg_tile (ROMS/Nonlinear/set_avg.F:237-2600, AVERAGES) for the state variables of the chain: the window's first call
// copies (:280-426), the others add (:1308-1454), and the call that closes the window scales the sums by 1/nAVG and refreshes
// the periodic images (:2327-2598).  One thread per (i,j,k); level k = 0 also handles the 2-D fields.
__global__ void __launch_bounds__(256) k_set_avg(Par p, Flds f, int mode, int norm, double fac, int Kout, int Nout) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // JstrR..JendR = 0..Mm+1
  const int k = blockIdx.z;                                     // 0..N
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P, o = o2 + k * p.PL;
  auto put = [&](double* A, int off, double x) {
    double v = (mode == 0) ? x : A[off + i] + x;
    if (norm) { v = fac * v; st_w(A, off, i, v, p); } else A[off + i] = v;
  };
  if (k == 0) {
    put(f.avgzeta, o2, f.zeta[Kout][o2 + i]);
    put(f.avgu2d, o2, f.ubar[Kout][o2 + i]);
    if (j >= 1) put(f.avgv2d, o2, f.vbar[Kout][o2 + i]);
  } else {
    put(f.avgu3d, o, f.u[Nout][o + i]);
    if (j >= 1) put(f.avgv3d, o, f.v[Nout][o + i]);
    put(f.avgrho, o, f.rho[o + i]);
    for (int it = 0; it < p.NT; ++it) put(f.avgt[it], o, f.t[Nout][it][o + i]);
  }
  put(f.avgw3d, o, f.W[o + i] * f.pm[o2 + i] * f.pn[o2 + i]);
  put(f.avgwvel, o, f.wvel[o + i]);
}

// ---------------------------------------------------------------------------------------------------------------
static inline dim3 g2(const Par& p, dim3 b, int nj, int nz = 1) {
  return dim3((xspan(p) + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz);
}

void launch_set_massflux(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 4); k_set_massflux<<<g2(p, b, p.Mm + 2, p.N), b, 0, s>>>(p, f); }
void launch_rho_eos(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 4);
  if (p.bv_frequency || p.eos_tderivative) k_rho_eos<true><<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f);
  else k_rho_eos<false><<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f);
}
void launch_set_vbc(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 4); k_set_vbc<<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f); }
void launch_omega(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2);
  if (p.N == 30) k_omega<30><<<g2(p, b, p.Mm), b, 0, s>>>(p, f);
  else k_omega<0><<<g2(p, b, p.Mm), b, 0, s>>>(p, f);
}
void launch_wvelocity(const Par& p, const Flds& f, int Ninp, cudaStream_t s) {
  dim3 b(64, 2);
  if (p.N == 30) k_wvelocity<30><<<g2(p, b, p.Mm), b, 0, s>>>(p, f, Ninp);
  else k_wvelocity<0><<<g2(p, b, p.Mm), b, 0, s>>>(p, f, Ninp);
}
void launch_set_avg(const Par& p, const Flds& f, int mode, int norm, double fac, int Kout, int Nout, cudaStream_t s) {
  dim3 b(64, 4);
  k_set_avg<<<g2(p, b, p.Mm + 2, p.N + 1), b, 0, s>>>(p, f, mode, norm, fac, Kout, Nout);
}
void launch_set_zeta(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 4); k_set_zeta<<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f); }
void launch_set_depth(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 4);
  if (p.vtransform == 1) k_set_depth<1><<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f);
  else k_set_depth<2><<<g2(p, b, p.Mm + 2), b, 0, s>>>(p, f);
}
void launch_bvf_mix(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 4); k_bvf_mix<<<g2(p, b, p.Mm, p.N - 1), b, 0, s>>>(p, f); }
void launch_ana_vmix(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 4); k_ana_vmix<<<g2(p, b, p.Mm + 2, p.N - 1), b, 0, s>>>(p, f); }

}  // namespace rb

// Tracer advection flux functions shared by pre_step3d (advects t(nstp)) and step3d_t (advects t(:,:,:,3,:)).
#pragma once
#include "dev.cuh"

namespace rb {

// Face flux of a tracer between cells (m-1) and m along one horizontal axis.
//   d0 = t(m)-t(m-1), dm1 = t(m-1)-t(m-2), dp1 = t(m+1)-t(m), H = Huon/Hvom at the face.
// UPSTREAM3 / AKIMA4 / CENTERED4 / CENTERED2 as in ROMS/Nonlinear/pre_step3d.F:345-457 and step3d_t.F:390-661.
template <int HADV>
__device__ __forceinline__ double hflux(double dm1, double d0, double dp1, double tm1, double t0, double H) {
  if (HADV == 0) {                                      // UPSTREAM3
    const double curv_lo = d0 - dm1, curv_hi = dp1 - d0;
    return H * 0.5 * (tm1 + t0) - (1.0 / 6.0) * (curv_lo * dmax(H, 0.0) + curv_hi * dmin(H, 0.0));
  } else if (HADV == 1) {                               // AKIMA4
    const double eps = 1.0e-16;
    double c = 2.0 * d0 * dm1;
    const double g_lo = (c > eps) ? c / (d0 + dm1) : 0.0;
    c = 2.0 * dp1 * d0;
    const double g_hi = (c > eps) ? c / (dp1 + d0) : 0.0;
    return H * 0.5 * (tm1 + t0 - (1.0 / 3.0) * (g_hi - g_lo));
  } else if (HADV == 2) {                               // CENTERED4
    const double g_lo = 0.5 * (d0 + dm1), g_hi = 0.5 * (dp1 + d0);
    return H * 0.5 * (tm1 + t0 - (1.0 / 3.0) * (g_hi - g_lo));
  } else {                                              // CENTERED2
    return H * 0.5 * (tm1 + t0);
  }
}

// The four face fluxes around cell (i,j) at level offset o (= j*P + k*PL): FX(i), FX(i+1), FE(j), FE(j+1), with the
// closed-wall copies FE(Jstr-1)=FE(Jstr), FE(Jend+2)=FE(Jend+1) of the first differences (pre_step3d.F:468-481,
// step3d_t.F:672-685).
template <int HADV>
__device__ __forceinline__ void hadv_fluxes(const double* __restrict__ T, const double* __restrict__ Huon,
                                            const double* __restrict__ Hvom, int o, int i, int j, const Par& p,
                                            double& FXi, double& FXip, double& FEj, double& FEjp) {
  const int P = p.P;
  const double tm2 = T[o + i - 2], tm1 = T[o + i - 1], t0 = T[o + i], tp1 = T[o + i + 1], tp2 = T[o + i + 2];
  const double dxm1 = tm1 - tm2, dx0 = t0 - tm1, dxp1 = tp1 - t0, dxp2 = tp2 - tp1;
  FXi = hflux<HADV>(dxm1, dx0, dxp1, tm1, t0, Huon[o + i]);
  FXip = hflux<HADV>(dx0, dxp1, dxp2, t0, tp1, Huon[o + i + 1]);
  // eta: differences d(m) = t(m)-t(m-1) exist for m = 1..Mm+1; d(0) = d(1), d(Mm+2) = d(Mm+1)
  const double sm1 = T[o - P + i], sp1 = T[o + P + i];
  const double dy0 = t0 - sm1, dyp1 = sp1 - t0;
  const double dym1 = (j > 1) ? (sm1 - T[o - 2 * P + i]) : dy0;
  const double dyp2 = (j < p.Mm) ? (T[o + 2 * P + i] - sp1) : dyp1;
  FEj = hflux<HADV>(dym1, dy0, dyp1, sm1, t0, Hvom[o + i]);
  FEjp = hflux<HADV>(dy0, dyp1, dyp2, t0, sp1, Hvom[o + P + i]);
}

// Same fluxes from operands that were loaded up front (all global loads of a k-iteration are issued back to back so
// that the thread waits for memory once per level instead of once per dependent group).
struct AdvIn { double tm2, tm1, t0, tp1, tp2, sm2, sm1, sp1, sp2, hu0, hu1, hv0, hv1; };
__device__ __forceinline__ AdvIn adv_load(const double* __restrict__ T, const double* __restrict__ Huon, const double* __restrict__ Hvom,
                                          int o, int i, int j, const Par& p) {
  const int P = p.P;
  AdvIn a;
  const int om2 = (j > 1) ? -2 * P : -P, op2 = (j < p.Mm) ? 2 * P : P;     // clamped rows keep the loads unconditional
  a.tm2 = T[o + i - 2]; a.tm1 = T[o + i - 1]; a.t0 = T[o + i]; a.tp1 = T[o + i + 1]; a.tp2 = T[o + i + 2];
  a.sm2 = T[o + om2 + i]; a.sm1 = T[o - P + i]; a.sp1 = T[o + P + i]; a.sp2 = T[o + op2 + i];
  a.hu0 = Huon[o + i]; a.hu1 = Huon[o + i + 1]; a.hv0 = Hvom[o + i]; a.hv1 = Hvom[o + P + i];
  return a;
}
template <int HADV>
__device__ __forceinline__ void hadv_fluxes_v(const AdvIn& a, int j, int Mm, double& FXi, double& FXip, double& FEj, double& FEjp) {
  const double dxm1 = a.tm1 - a.tm2, dx0 = a.t0 - a.tm1, dxp1 = a.tp1 - a.t0, dxp2 = a.tp2 - a.tp1;
  FXi = hflux<HADV>(dxm1, dx0, dxp1, a.tm1, a.t0, a.hu0);
  FXip = hflux<HADV>(dx0, dxp1, dxp2, a.t0, a.tp1, a.hu1);
  const double dy0 = a.t0 - a.sm1, dyp1 = a.sp1 - a.t0;
  const double dym1 = (j > 1) ? (a.sm1 - a.sm2) : dy0;
  const double dyp2 = (j < Mm) ? (a.sp2 - a.sp1) : dyp1;
  FEj = hflux<HADV>(dym1, dy0, dyp1, a.sm1, a.t0, a.hv0);
  FEjp = hflux<HADV>(dy0, dyp1, dyp2, a.t0, a.sp1, a.hv1);
}

// Vertical advective flux through the top of level k (k = 1..N-1) from a column tc[0..N+1].
// CENTERED4 (pre_step3d.F:751-785), AKIMA4 (:667-707), CENTERED2 (:709-727); same in step3d_t.F:938-1126.
template <int VADV>
__device__ __forceinline__ double vflux(const double* tc, int k, int N, double Wk) {
  if (VADV == 0) {
    if (k == 1) return Wk * (0.5 * tc[1] + (7.0 / 12.0) * tc[2] - (1.0 / 12.0) * tc[3]);
    if (k == N - 1) return Wk * (0.5 * tc[N] + (7.0 / 12.0) * tc[N - 1] - (1.0 / 12.0) * tc[N - 2]);
    return Wk * ((7.0 / 12.0) * (tc[k] + tc[k + 1]) - (1.0 / 12.0) * (tc[k - 1] + tc[k + 2]));
  } else if (VADV == 1) {
    const double eps = 1.0e-16;
    // FC(k) = t(k+1)-t(k) for 1..N-1, FC(0)=FC(1), FC(N)=FC(N-1); CF(k) = harmonic(FC(k),FC(k-1)), k = 1..N
    const double dk = tc[k + 1] - tc[k];
    const double dkm1 = (k > 1) ? (tc[k] - tc[k - 1]) : dk;
    const double dkp1 = (k + 1 <= N - 1) ? (tc[k + 2] - tc[k + 1]) : dk;
    double c = 2.0 * dk * dkm1;
    const double CFk = (c > eps) ? c / (dk + dkm1) : 0.0;
    c = 2.0 * dkp1 * dk;
    const double CFkp = (c > eps) ? c / (dkp1 + dk) : 0.0;
    return Wk * 0.5 * (tc[k] + tc[k + 1] - (1.0 / 3.0) * (CFkp - CFk));
  } else {
    return Wk * 0.5 * (tc[k] + tc[k + 1]);
  }
}

// Same flux from a rolling four-value window (tkm1 = t(k-1), tk, tkp1, tkp2 = t(k+2); out-of-column values are never
// used by the end formulas).
template <int VADV>
__device__ __forceinline__ double vflux4(double tkm1, double tk, double tkp1, double tkp2, int k, int N, double Wk) {
  if (VADV == 0) {
    if (k == 1) return Wk * (0.5 * tk + (7.0 / 12.0) * tkp1 - (1.0 / 12.0) * tkp2);
    if (k == N - 1) return Wk * (0.5 * tkp1 + (7.0 / 12.0) * tk - (1.0 / 12.0) * tkm1);
    return Wk * ((7.0 / 12.0) * (tk + tkp1) - (1.0 / 12.0) * (tkm1 + tkp2));
  } else if (VADV == 3) {
    return 0.0;                                   // SPLINES: the flux comes from vspline_flux, not from the rolling window
  } else if (VADV == 1) {
    const double eps = 1.0e-16;
    const double dk = tkp1 - tk;
    const double dkm1 = (k > 1) ? (tk - tkm1) : dk;
    const double dkp1 = (k + 1 <= N - 1) ? (tkp2 - tkp1) : dk;
    double c = 2.0 * dk * dkm1;
    const double CFk = (c > eps) ? c / (dk + dkm1) : 0.0;
    c = 2.0 * dkp1 * dk;
    const double CFkp = (c > eps) ? c / (dkp1 + dk) : 0.0;
    return Wk * 0.5 * (tk + tkp1 - (1.0 / 3.0) * (CFkp - CFk));
  } else {
    return Wk * 0.5 * (tk + tkp1);
  }
}

// SPLINES vertical advection (pre_step3d.F:622-665 with NEUMANN, step3d_t.F:894-937 without): interfacial tracer values from
// conservative parabolic splines -- a tridiagonal solve along the column -- times W.  FCs / CFs: thread-local columns (0:N);
// on return FCs(k) is the vertical advective flux through W level k.  q = offset of (i,j) in level 0.
template <bool NEUMANN>
__device__ __forceinline__ void vspline_flux(const double* __restrict__ T, const double* __restrict__ Hz, const double* __restrict__ W,
                                             int q, int N, int PL, double* FCs, double* CFs) {
  double tk = T[q + PL], hk = Hz[q + PL];
  FCs[0] = (NEUMANN ? 1.5 : 2.0) * tk;
  CFs[1] = NEUMANN ? 0.5 : 1.0;
  for (int k = 1; k <= N - 1; ++k) {
    const double t1 = T[q + (k + 1) * PL], h1 = Hz[q + (k + 1) * PL];
    const double cff = 1.0 / (2.0 * hk + h1 * (2.0 - CFs[k]));
    CFs[k + 1] = cff * hk;
    FCs[k] = cff * (3.0 * (hk * t1 + h1 * tk) - h1 * FCs[k - 1]);
    tk = t1; hk = h1;
  }
  FCs[N] = ((NEUMANN ? 3.0 : 2.0) * tk - FCs[N - 1]) / ((NEUMANN ? 2.0 : 1.0) - CFs[N]);
  for (int k = N - 1; k >= 0; --k) {
    FCs[k] = FCs[k] - CFs[k + 1] * FCs[k + 1];
    FCs[k + 1] = W[q + (k + 1) * PL] * FCs[k + 1];
  }
  FCs[N] = 0.0; FCs[0] = 0.0;
}

}  // namespace rb

// pre_step3d (tracer predictor + momentum AB3 loading), prsgrd31/32, t3dmix2_s.
#include "dev.cuh"
#include "kernels.h"
#include "k_adv.cuh"

namespace rb {

// ---------------------------------------------------------------------------------------------------------------
// pre_step3d_tile, tracer part (ROMS/Nonlinear/pre_step3d.F:342-582 horizontal, :619-826 vertical + artificial
// continuity, :837-906 t(nnew) loading, :1126-1142 t3dbc + periodic images).  One thread per column and tracer; every
// 3-D input is read once per level and both outputs are written once.
template <int HADV, int VADV>
__global__ void __launch_bounds__(128) k_pre_step3d_t(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int itrc = blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P;
  const double* __restrict__ tst = f.t[p.nstp][itrc];
  double* __restrict__ tnw = f.t[p.nnew][itrc];
  double* __restrict__ t3 = f.t[3][itrc];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Akt = f.Akt[itrc];
  const double pm = f.pm[o2 + i], pn = f.pn[o2 + i];
  const double Gamma = 1.0 / 6.0;
  double cff, cff1h, cff2h;
  if (p.istart == 0) { cff = 0.5 * p.dt; cff1h = 1.0; cff2h = 0.0; }
  else { cff = (1.0 - Gamma) * p.dt; cff1h = 0.5 + Gamma; cff2h = 0.5 - Gamma; }
  // Rolling vertical window of t(nstp) (tkm1 = t(k-1) ... tkp2 = t(k+2)) and software-pipelined operand loads: the loads
  // of level k+1 are issued, back to back, before level k is computed, so a thread waits for memory once per level.
  struct Lvl { AdvIn a; double tnw, hz, W, zr1, akt, tk3; };
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * p.PL;
    Lvl L;
    L.a = adv_load(tst, Huon, Hvom, o, i, j, p);
    L.tnw = tnw[o + i]; L.hz = Hz[o + i]; L.W = W[o + i]; L.akt = Akt[o + i];
    L.zr1 = z_r[o + ((k < N) ? p.PL : 0) + i];                   // z_r(k+1) (k = N: unused)
    L.tk3 = tst[o2 + ((k + 2 <= N) ? (k + 2) : N) * p.PL + i];   // t(k+2), clamped
    return L;
  };
  const double cff3 = p.dt * (1.0 - p.lambda);
  double FCm = 0.0;                                   // advective FC(k-1)
  double FDm = p.dt * f.btflx[itrc][o2 + i];          // diffusive FC(k-1), FC(0) = dt*btflx
  double Wm = W[o2 + i];                              // W(k-1)
  double zrk = z_r[o2 + p.PL + i];                    // z_r(k)
  Lvl cur = load_level(1);
  double tkm1 = cur.a.t0, tk = cur.a.t0, tkp1 = tst[o2 + 2 * p.PL + i], tkp2 = cur.tk3;
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL;
    Lvl nxt = cur;
    if (k < N) nxt = load_level(k + 1);
    const double hz = cur.hz;
    double FXi, FXip, FEj, FEjp;
    hadv_fluxes_v<HADV>(cur.a, j, p.Mm, FXi, FXip, FEj, FEjp);
    const double div = FXip - FXi + FEjp - FEj;
    double t3v = hz * (cff1h * tk + cff2h * cur.tnw) - cff * pm * pn * div;
    const double Wk = cur.W;
    const double FCk = (k < N) ? vflux4<VADV>(tkm1, tk, tkp1, tkp2, k, N, Wk) : 0.0;
    const double DC = 1.0 / (hz - cff * pm * pn * (cur.a.hu1 - cur.a.hu0 + cur.a.hv1 - cur.a.hv0 + (Wk - Wm)));
    const double cff1 = cff * pm * pn;
    t3v = DC * (t3v - cff1 * (FCk - FCm));
    st_r_grad(t3, o, i, j, t3v, p);
    // t(nnew) = Hz*t(nstp) + explicit part of the vertical diffusion (zero for lambda = 1) + surface/bottom fluxes
    double FDk;
    if (k < N) {
      const double c = 1.0 / (cur.zr1 - zrk);
      FDk = cff3 * c * cur.akt * (tkp1 - tk);
    } else {
      FDk = p.dt * f.stflx[itrc][o2 + i];
    }
    const double c1 = hz * tk;
    const double c2 = FDk - FDm;
    tnw[o + i] = c1 + c2;
    FCm = FCk; FDm = FDk; Wm = Wk; zrk = cur.zr1;
    tkm1 = tk; tk = tkp1; tkp1 = tkp2; tkp2 = nxt.tk3;
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pre_step3d_tile, momentum part (ROMS/Nonlinear/pre_step3d.F:917-1118): u,v(nnew) = Hz*u(nstp) + AB3 rhs + explicit
// vertical viscosity flux divergence.  One thread per column, handles the u-point and the v-point of cell (i,j).
__global__ void __launch_bounds__(128) k_pre_step3d_uv(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P;
  const int indx = 3 - p.nrhs;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Akv = f.Akv;
  const double pm0 = f.pm[o2 + i], pn0 = f.pn[o2 + i];
  const double cff3 = p.dt * (1.0 - p.lambda);
  {
    const double* __restrict__ ust = f.u[p.nstp];
    double* __restrict__ unw = f.u[p.nnew];
    const double* __restrict__ ru_r = f.ru[p.nrhs];
    const double* __restrict__ ru_i = f.ru[indx];
    const double cff = p.dt * 0.25;
    const double DC0 = cff * (pm0 + f.pm[o2 + i - 1]) * (pn0 + f.pn[o2 + i - 1]);
    double FCm = p.dt * f.bustr[o2 + i];
    double uk = ust[o2 + p.PL + i];
    for (int k = 1; k <= N; ++k) {
      const int o = o2 + k * p.PL;
      double FCk, ukp = 0.0;
      if (k < N) {
        ukp = ust[o + p.PL + i];
        const double c = 1.0 / (z_r[o + p.PL + i] + z_r[o + p.PL + i - 1] - z_r[o + i] - z_r[o + i - 1]);
        FCk = cff3 * c * (ukp - uk) * (Akv[o + i] + Akv[o + i - 1]);
      } else {
        FCk = p.dt * f.sustr[o2 + i];
      }
      const double a = uk * 0.5 * (Hz[o + i] + Hz[o + i - 1]);
      const double d = FCk - FCm;
      double x;
      if (p.istart == 0) x = a + d;
      else if (p.istart == 1) { const double c3 = 0.5 * DC0; x = a - c3 * ru_i[o + i] + d; }
      else x = a + DC0 * ((5.0 / 12.0) * ru_r[o + i] - (16.0 / 12.0) * ru_i[o + i]) + d;
      unw[o + i] = x;
      FCm = FCk; uk = ukp;
    }
  }
  if (j >= p.JstrV) {
    const double* __restrict__ vst = f.v[p.nstp];
    double* __restrict__ vnw = f.v[p.nnew];
    const double* __restrict__ rv_r = f.rv[p.nrhs];
    const double* __restrict__ rv_i = f.rv[indx];
    const double cff = p.dt * 0.25;
    const double DC0 = cff * (pm0 + f.pm[o2 - P + i]) * (pn0 + f.pn[o2 - P + i]);
    double FCm = p.dt * f.bvstr[o2 + i];
    double vk = vst[o2 + p.PL + i];
    for (int k = 1; k <= N; ++k) {
      const int o = o2 + k * p.PL;
      double FCk, vkp = 0.0;
      if (k < N) {
        vkp = vst[o + p.PL + i];
        const double c = 1.0 / (z_r[o + p.PL + i] + z_r[o + p.PL - P + i] - z_r[o + i] - z_r[o - P + i]);
        FCk = cff3 * c * (vkp - vk) * (Akv[o + i] + Akv[o - P + i]);
      } else {
        FCk = p.dt * f.svstr[o2 + i];
      }
      const double a = vk * 0.5 * (Hz[o + i] + Hz[o - P + i]);
      const double d = FCk - FCm;
      double x;
      if (p.istart == 0) x = a + d;
      else if (p.istart == 1) { const double c3 = 0.5 * DC0; x = a - c3 * rv_i[o + i] + d; }
      else x = a + DC0 * ((5.0 / 12.0) * rv_r[o + i] - (16.0 / 12.0) * rv_i[o + i]) + d;
      vnw[o + i] = x;
      FCm = FCk; vk = vkp;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prsgrd32_tile, phase 1 (ROMS/Nonlinear/prsgrd32.h:236-290): per-column harmonic-mean slopes and downward
// integration of the pressure P over IstrU-1:Iend x JstrV-1:Jend.
__global__ void __launch_bounds__(128) k_prsgrd32_P(Par p, Flds f) {
  const int i = p.Istr - 1 + blockIdx.x * blockDim.x + threadIdx.x;   // IstrU-1 .. Iend
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;            // JstrV-1 .. Jend
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, o2 = j * p.P;
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  const double GRho = p.g / p.rho0, HalfGRho = 0.5 * GRho;
  double rc[MAXN + 1], zc[MAXN + 1];
  for (int k = 1; k <= N; ++k) { rc[k] = f.rho[o2 + k * p.PL + i]; zc[k] = f.z_r[o2 + k * p.PL + i]; }
  const double zwN = f.z_w[o2 + N * p.PL + i];
  // harmonic means at level N: raw(N) = raw(N-1)
  double rawR_k = rc[N] - rc[N - 1], rawZ_k = zc[N] - zc[N - 1];     // raw(N) := raw(N-1)
  double rawR_km = rawR_k, rawZ_km = rawZ_k;                          // raw(N-1)
  double c = 2.0 * rawR_k * rawR_km;
  double dR_kp = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;            // dR(N)
  double dZ_kp = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);         // dZ(N)
  const double cff1 = 1.0 / (zc[N] - zc[N - 1]);
  const double cff2 = 0.5 * (rc[N] - rc[N - 1]) * (zwN - zc[N]) * cff1;
  double Pk = p.g * zwN + GRho * (rc[N] + cff2) * (zwN - zc[N]);
  f.P3[o2 + N * p.PL + i] = Pk;
  for (int k = N - 1; k >= 1; --k) {
    // dR(k) = harmonic(raw(k), raw(k-1)); raw(0) = raw(1)
    rawR_k = rc[k + 1] - rc[k]; rawZ_k = zc[k + 1] - zc[k];
    if (k > 1) { rawR_km = rc[k] - rc[k - 1]; rawZ_km = zc[k] - zc[k - 1]; } else { rawR_km = rawR_k; rawZ_km = rawZ_k; }
    c = 2.0 * rawR_k * rawR_km;
    const double dR_k = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;
    const double dZ_k = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);
    Pk = Pk + HalfGRho * ((rc[k + 1] + rc[k]) * (zc[k + 1] - zc[k]) -
                          OneFifth * ((dR_kp - dR_k) * (zc[k + 1] - zc[k] - OneTwelfth * (dZ_kp + dZ_k)) -
                                      (dZ_kp - dZ_k) * (rc[k + 1] - rc[k] - OneTwelfth * (dR_kp + dR_k))));
    f.P3[o2 + k * p.PL + i] = Pk;
    dR_kp = dR_k; dZ_kp = dZ_k;
  }
}

__device__ __forceinline__ void harm_pair(double a0, double a1, double f0, double f1, double& dZx, double& dRx) {
  const double eps = 1.0e-10;
  const double cff = 2.0 * a0 * a1;
  if (cff > eps) { const double c1 = 1.0 / (a0 + a1); dZx = cff * c1; } else dZx = 0.0;
  const double cff1 = 2.0 * f0 * f1;
  if (cff1 > eps) { const double c2 = 1.0 / (f0 + f1); dRx = cff1 * c2; } else dRx = 0.0;
}

// prsgrd32_tile, phase 2 (prsgrd32.h:296-418): ru,rv(i,j,k,nrhs) from P and the horizontal harmonic-mean slopes.
__global__ void __launch_bounds__(256) k_prsgrd32_R(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int P = p.P, o2 = j * P, o = o2 + k * p.PL;
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0;
  const double HalfGRho = 0.5 * (p.g / p.rho0);
  const double* __restrict__ rho = f.rho;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ PP = f.P3;
  const double r0 = rho[o + i], z0 = z_r[o + i], hz0 = Hz[o + i], P0 = PP[o + i];
  {
    const double rm2 = rho[o + i - 2], rm1 = rho[o + i - 1], rp1 = rho[o + i + 1];
    const double zm2 = z_r[o + i - 2], zm1 = z_r[o + i - 1], zp1 = z_r[o + i + 1];
    // aux(m) = z_r(m)-z_r(m-1), FC(m) = rho(m)-rho(m-1) for m = i-1, i, i+1
    const double am1 = zm1 - zm2, a0 = z0 - zm1, ap1 = zp1 - z0;
    const double fm1 = rm1 - rm2, f0 = r0 - rm1, fp1 = rp1 - r0;
    double dZx_m, dRx_m, dZx_0, dRx_0;
    harm_pair(am1, a0, fm1, f0, dZx_m, dRx_m);        // (i-1)
    harm_pair(a0, ap1, f0, fp1, dZx_0, dRx_0);        // (i)
    const double x = f.on_u[o2 + i] * 0.5 * (hz0 + Hz[o + i - 1]) *
                     (PP[o + i - 1] - P0 -
                      HalfGRho * ((r0 + rm1) * (z0 - zm1) -
                                  OneFifth * ((dRx_0 - dRx_m) * (z0 - zm1 - OneTwelfth * (dZx_0 + dZx_m)) -
                                              (dZx_0 - dZx_m) * (r0 - rm1 - OneTwelfth * (dRx_0 + dRx_m)))));
    f.ru[p.nrhs][o + i] = x;
  }
  if (j >= p.JstrV) {
    const double rm2 = rho[o - 2 * P + i], rm1 = rho[o - P + i], rp1 = rho[o + P + i];
    const double zm2 = z_r[o - 2 * P + i], zm1 = z_r[o - P + i], zp1 = z_r[o + P + i];
    const double am1 = zm1 - zm2, a0 = z0 - zm1, ap1 = zp1 - z0;
    const double fm1 = rm1 - rm2, f0 = r0 - rm1, fp1 = rp1 - r0;
    double dZx_m, dRx_m, dZx_0, dRx_0;
    harm_pair(am1, a0, fm1, f0, dZx_m, dRx_m);        // (j-1)
    harm_pair(a0, ap1, f0, fp1, dZx_0, dRx_0);        // (j)
    const double x = f.om_v[o2 + i] * 0.5 * (hz0 + Hz[o - P + i]) *
                     (PP[o - P + i] - P0 -
                      HalfGRho * ((r0 + rm1) * (z0 - zm1) -
                                  OneFifth * ((dRx_0 - dRx_m) * (z0 - zm1 - OneTwelfth * (dZx_0 + dZx_m)) -
                                              (dZx_0 - dZx_m) * (r0 - rm1 - OneTwelfth * (dRx_0 + dRx_m)))));
    f.rv[p.nrhs][o + i] = x;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prsgrd31_tile (ROMS/Nonlinear/prsgrd31.h:203-359): standard density Jacobian, RHO_SURF on.  Thread per column.
__global__ void __launch_bounds__(128) k_prsgrd31(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P;
  const double fac1 = 0.5 * p.g / p.rho0, fac2 = 1000.0 * p.g / p.rho0, fac3 = 0.25 * p.g / p.rho0;
  const double* __restrict__ rho = f.rho;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ z_w = f.z_w;
  const double* __restrict__ Hz = f.Hz;
  for (int dir = 0; dir < 2; ++dir) {
    if (dir == 1 && j < p.JstrV) break;
    const int s = dir ? P : 1;                        // neighbour stride (i-1 or j-1)
    double* __restrict__ R = dir ? f.rv[p.nrhs] : f.ru[p.nrhs];
    const double met = dir ? f.om_v[o2 + i] : f.on_u[o2 + i];
    const int oN = o2 + N * p.PL + i;
    double cff1 = z_w[oN] - z_r[oN] + z_w[oN - s] - z_r[oN - s];
    double phi = fac1 * (rho[oN] - rho[oN - s]) * cff1;
    phi = phi + (fac2 + fac1 * (rho[oN] + rho[oN - s])) * (z_w[oN] - z_w[oN - s]);
    R[oN] = -0.5 * (Hz[oN] + Hz[oN - s]) * phi * met;
    for (int k = N - 1; k >= 1; --k) {
      const int o = o2 + k * p.PL + i, ou = o + p.PL;
      const double c1 = rho[ou] - rho[ou - s] + rho[o] - rho[o - s];
      const double c2 = rho[ou] + rho[ou - s] - rho[o] - rho[o - s];
      const double c3 = z_r[ou] + z_r[ou - s] - z_r[o] - z_r[o - s];
      const double c4 = z_r[ou] - z_r[ou - s] + z_r[o] - z_r[o - s];
      phi = phi + fac3 * (c1 * c3 - c2 * c4);
      R[o] = -0.5 * (Hz[o] + Hz[o - s]) * phi * met;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// t3dmix2_s_tile (ROMS/Nonlinear/t3dmix2_s.h:198-301): harmonic mixing of tracers along s-surfaces.
__global__ void __launch_bounds__(256) k_t3dmix2_s(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int P = p.P, o2 = j * P, o = o2 + k * p.PL;
  const double* __restrict__ Hz = f.Hz;
  const double hz0 = Hz[o + i], hzW = Hz[o + i - 1], hzE = Hz[o + i + 1], hzS = Hz[o - P + i], hzN = Hz[o + P + i];
  const double cff = p.dt * f.pm[o2 + i] * f.pn[o2 + i];
  for (int it = 0; it < p.NT; ++it) {
    const double* __restrict__ tr = f.t[p.nrhs][it];
    double* __restrict__ tn = f.t[p.nnew][it];
    const double* __restrict__ d2 = f.diff2[it];
    const double t0 = tr[o + i];
    const double d0 = d2[o2 + i];
    const double FXi = 0.25 * (d0 + d2[o2 + i - 1]) * f.pmon_u[o2 + i] * (hz0 + hzW) * (t0 - tr[o + i - 1]);
    const double FXip = 0.25 * (d2[o2 + i + 1] + d0) * f.pmon_u[o2 + i + 1] * (hzE + hz0) * (tr[o + i + 1] - t0);
    const double FEj = 0.25 * (d0 + d2[o2 - P + i]) * f.pnom_v[o2 + i] * (hz0 + hzS) * (t0 - tr[o - P + i]);
    const double FEjp = 0.25 * (d2[o2 + P + i] + d0) * f.pnom_v[o2 + P + i] * (hzN + hz0) * (tr[o + P + i] - t0);
    const double cff1 = cff * (FXip - FXi);
    const double cff2 = cff * (FEjp - FEj);
    const double cff3 = cff1 + cff2;
    tn[o + i] = tn[o + i] + cff3;
  }
}

// ---------------------------------------------------------------------------------------------------------------
static inline dim3 g2(const Par& p, dim3 b, int ni, int nj, int nz = 1) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz); }

template <int H>
static void launch_pre_t_v(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2); dim3 g = g2(p, b, p.Iend - p.Istr + 1, p.Mm, p.NT);
  if (p.vadv == 0) k_pre_step3d_t<H, 0><<<g, b, 0, s>>>(p, f);
  else if (p.vadv == 1) k_pre_step3d_t<H, 1><<<g, b, 0, s>>>(p, f);
  else k_pre_step3d_t<H, 2><<<g, b, 0, s>>>(p, f);
}
void launch_pre_step3d(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.hadv == 0) launch_pre_t_v<0>(p, f, s);
  else if (p.hadv == 1) launch_pre_t_v<1>(p, f, s);
  else if (p.hadv == 2) launch_pre_t_v<2>(p, f, s);
  else launch_pre_t_v<3>(p, f, s);
  dim3 b(64, 2);
  k_pre_step3d_uv<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
}
void launch_prsgrd(const Par& p, const Flds& f, int dj_gradps, cudaStream_t s) {
  if (dj_gradps) {
    dim3 b(64, 2);
    k_prsgrd32_P<<<g2(p, b, p.Iend - p.Istr + 2, p.Mm), b, 0, s>>>(p, f);
    dim3 b2(64, 4);
    k_prsgrd32_R<<<g2(p, b2, p.Iend - p.Istr + 1, p.Mm, p.N), b2, 0, s>>>(p, f);
  } else {
    dim3 b(64, 2);
    k_prsgrd31<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  }
}
void launch_t3dmix2_s(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 4);
  k_t3dmix2_s<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm, p.N), b, 0, s>>>(p, f);
}

}  // namespace rb

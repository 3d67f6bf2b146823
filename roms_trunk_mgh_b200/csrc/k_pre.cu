// pre_step3d (tracer predictor + momentum AB3 loading), prsgrd31/32, t3dmix2_s.
#include "dev.cuh"
#include "kernels.h"
#include "k_adv.cuh"

namespace rb {

// ---------------------------------------------------------------------------------------------------------------
// pre_step3d_tile, tracer part (ROMS/Nonlinear/pre_step3d.F:342-582 horizontal, :619-826 vertical + artificial
// continuity, :837-906 t(nnew) loading, :1126-1142 t3dbc + periodic images).  One thread per column and tracer; every
// 3-D input is read once per level and both outputs are written once.
// MIXS: also apply t3dmix2_s (ROMS/Nonlinear/t3dmix2_s.h:198-301) to t(nnew) in the same pass -- its operands t(nrhs = nstp) at
// i+-1, j+-1 are already in registers for the advective fluxes, only the four Hz neighbours are extra loads -- so the separate
// kernel (7 units of traffic) disappears from the time step.  The additions happen in the reference's order.
// Jerlov water types (mod_scalars.F:1502-1512): reciprocal absorption coefficients and band-1 fraction
__constant__ double c_lmd_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};
__constant__ double c_lmd_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};
__constant__ double c_lmd_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};
template <int HADV, int VADV, bool MIXS, bool SRC>     // SRC: with the KPP nonlocal / shortwave source terms of the vertical flux
#ifndef PRT_PP
#define PRT_PP false
#endif
#ifndef PRT_MINB
#define PRT_MINB 3
#endif
#ifndef PRU_PP
#define PRU_PP false
#endif
#ifndef PRU_MINB
#define PRU_MINB 3
#endif
#ifndef PRT_PF
#define PRT_PF 4          // L2 prefetch distance (levels) in k_pre_step3d_t
#endif
#ifndef PRU_PF
#define PRU_PF 4          // same for k_pre_step3d_uv
#endif
// without the fused t3dmix2_s the kernel fits 128 registers: 4 CTAs per SM (0.86 -> 0.80 ms for pre_step3d on BENCHMARK3 with the
// shipped cpp set); with it (156 registers) the cap costs more in spills than the extra CTA brings (profiles/README.md)
__global__ void __launch_bounds__(128, MIXS ? PRT_MINB : PRT_MINB + 1) k_pre_step3d_t(Par p, Flds f) {
  // the tracer index is the fastest grid dimension: the CTAs of all tracers of one tile run back to back, so the shared
  // operands (Huon, Hvom, W, Hz, z_r) of the second tracer come from L2
  const int itrc = blockIdx.x % p.NT;
  const int i = xcol0(p, (blockIdx.x / p.NT) * blockDim.x) + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
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
  double mW = 0.0, mE = 0.0, mS = 0.0, mN = 0.0, mcff = 0.0;
  if (MIXS) {
    const double* __restrict__ d2 = f.diff2[itrc];
    const int q = o2 + i;
    const double d0 = d2[q];
    mW = 0.25 * (d0 + d2[q - 1]) * f.pmon_u[q];
    mE = 0.25 * (d2[q + 1] + d0) * f.pmon_u[q + 1];
    mS = 0.25 * (d0 + d2[q - P]) * f.pnom_v[q];
    mN = 0.25 * (d2[q + P] + d0) * f.pnom_v[q + P];
    mcff = p.dt * pm * pn;
  }
  // optional terms of the vertical flux FC (benchmark.h): KPP nonlocal transport of the active tracers (:850-865) and, for
  // temperature, the penetrating shortwave radiation (:312-333 with lmd_swfrac.F:66-80, Zscale = -1; :866-883)
  const bool nonloc = SRC && p.lmd_nonlocal && itrc < (p.salinity ? 2 : 1);
  const bool solar = SRC && p.solar_source && itrc == p.itemp - 1;
  const double* __restrict__ ghats = nonloc ? f.ghats[itrc] : nullptr;
  struct Lvl { AdvIn a; double tnw, hz, W, zr1, akt, tk3, hzW, hzE, hzS, hzN, gh, zw; };
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * p.PL;
    Lvl L;
    pf_up<PRT_PF>(tst, o + i, k, N, p.PL); pf_up<PRT_PF>(Huon, o + i, k, N, p.PL); pf_up<PRT_PF>(Hvom, o + i, k, N, p.PL); pf_up<PRT_PF>(tnw, o + i, k, N, p.PL);
    pf_up<PRT_PF>(Hz, o + i, k, N, p.PL); pf_up<PRT_PF>(W, o + i, k, N, p.PL); pf_up<PRT_PF>(Akt, o + i, k, N, p.PL); pf_up<PRT_PF>(z_r, o + i, k, N, p.PL);
    L.a = adv_load(tst, Huon, Hvom, o, i, j, p);
    L.tnw = tnw[o + i]; L.hz = Hz[o + i]; L.W = W[o + i]; L.akt = Akt[o + i];
    if (MIXS) { L.hzW = Hz[o + i - 1]; L.hzE = Hz[o + i + 1]; L.hzS = Hz[o - P + i]; L.hzN = Hz[o + P + i]; }
    L.zr1 = z_r[o + ((k < N) ? p.PL : 0) + i];                   // z_r(k+1) (k = N: unused)
    L.tk3 = tst[o2 + ((k + 2 <= N) ? (k + 2) : N) * p.PL + i];   // t(k+2), clamped
    if (SRC) {                                                   // operands of the optional flux terms, requested with the rest
      L.gh = nonloc ? ghats[o + i] : 0.0;
      L.zw = solar ? f.z_w[o + i] : 0.0;
    }
    return L;
  };
  const double cff3 = p.dt * (1.0 - p.lambda);
  double sw_fac1 = 0.0, sw_fac2 = 0.0, sw_fac3 = 0.0, sw_srflx = 0.0, sw_zwN = 0.0;
  if (solar) {
    int J = (int)f.Jwtype[o2 + i];
    J = (J < 1) ? 1 : (J > 9 ? 9 : J);
    sw_fac1 = -1.0 / c_lmd_mu1[J - 1]; sw_fac2 = -1.0 / c_lmd_mu2[J - 1]; sw_fac3 = c_lmd_r1[J - 1];
    sw_srflx = f.srflx[o2 + i]; sw_zwN = f.z_w[o2 + N * p.PL + i];
  }
  double FCs[VADV == 3 ? MAXN + 1 : 1], CFs[VADV == 3 ? MAXN + 1 : 1];
  if (VADV == 3) vspline_flux<true>(tst, Hz, W, o2 + i, N, p.PL, FCs, CFs);       // SPLINES (NEUMANN: pre_step3d.F:4)
  double FCm = 0.0;                                   // advective FC(k-1)
  double FDm = p.dt * f.btflx[itrc][o2 + i];          // diffusive FC(k-1), FC(0) = dt*btflx
  double Wm = W[o2 + i];                              // W(k-1)
  double zrk = z_r[o2 + p.PL + i];                    // z_r(k)
  double tkm1 = tst[o2 + p.PL + i], tk = tkm1, tkp1 = tst[o2 + 2 * p.PL + i], tkp2 = tst[o2 + 3 * p.PL + i];
  auto level = [&](const Lvl& cur, const Lvl& nxt, int k) {
    const int o = o2 + k * p.PL;
    const double hz = cur.hz;
    double FXi, FXip, FEj, FEjp;
    hadv_fluxes_v<HADV>(cur.a, j, p.Mm, FXi, FXip, FEj, FEjp);
    const double div = FXip - FXi + FEjp - FEj;
    double t3v = hz * (cff1h * tk + cff2h * cur.tnw) - cff * pm * pn * div;
    const double Wk = cur.W;
    const double FCk = (VADV == 3) ? FCs[VADV == 3 ? k : 0] : ((k < N) ? vflux4<VADV>(tkm1, tk, tkp1, tkp2, k, N, Wk) : 0.0);
    const double DC = 1.0 / (hz - cff * pm * pn * (cur.a.hu1 - cur.a.hu0 + cur.a.hv1 - cur.a.hv0 + (Wk - Wm)));
    const double cff1 = cff * pm * pn;
    t3v = DC * (t3v - cff1 * (FCk - FCm));
    st_r_grad(t3, o, i, j, t3v, p);
    // t(nnew) = Hz*t(nstp) + explicit part of the vertical diffusion (zero for lambda = 1) + surface/bottom fluxes
    double FDk;
    if (k < N) {
      const double c = 1.0 / (cur.zr1 - zrk);
      FDk = cff3 * c * cur.akt * (tkp1 - tk);
      if (SRC && nonloc) FDk = FDk - p.dt * cur.akt * cur.gh;
      if (SRC && solar) {
        const double Z = sw_zwN - cur.zw;
        const double swdk = exp(Z * sw_fac1) * sw_fac3 + exp(Z * sw_fac2) * (1.0 - sw_fac3);
        FDk = FDk + p.dt * sw_srflx * swdk;
      }
    } else {
      FDk = p.dt * f.stflx[itrc][o2 + i];
    }
    const double c1 = hz * tk;
    const double c2 = FDk - FDm;
    double tnv = c1 + c2;
    if (MIXS) {
      const double t0 = cur.a.t0;
      const double MXi = mW * (hz + cur.hzW) * (t0 - cur.a.tm1);
      const double MXip = mE * (cur.hzE + hz) * (cur.a.tp1 - t0);
      const double MEj = mS * (hz + cur.hzS) * (t0 - cur.a.sm1);
      const double MEjp = mN * (cur.hzN + hz) * (cur.a.sp1 - t0);
      const double m1 = mcff * (MXip - MXi);
      const double m2 = mcff * (MEjp - MEj);
      const double m3 = m1 + m2;
      tnv = tnv + m3;
    }
    tnw[o + i] = tnv;
    FCm = FCk; FDm = FDk; Wm = Wk; zrk = cur.zr1;
    tkm1 = tk; tk = tkp1; tkp1 = tkp2; tkp2 = nxt.tk3;
  };
  {
    sweep_levels<PRT_PP>(N, load_level, level);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pre_step3d_tile, momentum part (ROMS/Nonlinear/pre_step3d.F:917-1118): u,v(nnew) = Hz*u(nstp) + AB3 rhs + explicit
// vertical viscosity flux divergence.  One thread per column, handles the u-point and the v-point of cell (i,j) in one
// upward march so that Hz, z_r and Akv are read once; the operands of level k+1 are requested before level k is computed.
__global__ void __launch_bounds__(128, PRU_MINB) k_pre_step3d_uv(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, PL = p.PL, o2 = j * P + i;
  const int indx = 3 - p.nrhs;
  const int istart = p.istart;
  const bool dov = (j >= p.JstrV);
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Akv = f.Akv;
  const double* __restrict__ ust = f.u[p.nstp];
  const double* __restrict__ vst = f.v[p.nstp];
  double* __restrict__ unw = f.u[p.nnew];
  double* __restrict__ vnw = f.v[p.nnew];
  const double* __restrict__ ru_r = f.ru[p.nrhs];
  const double* __restrict__ ru_i = f.ru[indx];
  const double* __restrict__ rv_r = f.rv[p.nrhs];
  const double* __restrict__ rv_i = f.rv[indx];
  const double pm0 = f.pm[o2], pn0 = f.pn[o2];
  const double cff3 = p.dt * (1.0 - p.lambda);
  const double cff = p.dt * 0.25;
  const double DCu = cff * (pm0 + f.pm[o2 - 1]) * (pn0 + f.pn[o2 - 1]);
  const double DCv = cff * (pm0 + f.pm[o2 - P]) * (pn0 + f.pn[o2 - P]);
  struct Lvl { double hz0, hzW, hzS, zr0, zrW, zrS, ak0, akW, akS, up, vp, rur, rui, rvr, rvi; };
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * PL;
    const int ou = (k < N) ? o + PL : o;                 // level k+1 operands (unused at k = N)
    Lvl L;
    pf_up<PRU_PF>(Hz, o, k, N, PL); pf_up<PRU_PF>(z_r, o, k, N, PL); pf_up<PRU_PF>(Akv, o, k, N, PL); pf_up<PRU_PF>(ust, o, k, N, PL); pf_up<PRU_PF>(vst, o, k, N, PL);
    if (istart >= 1) { pf_up<PRU_PF>(ru_i, o, k, N, PL); pf_up<PRU_PF>(rv_i, o, k, N, PL); }
    if (istart == 2) { pf_up<PRU_PF>(ru_r, o, k, N, PL); pf_up<PRU_PF>(rv_r, o, k, N, PL); }
    L.hz0 = Hz[o]; L.hzW = Hz[o - 1]; L.hzS = Hz[o - P];
    L.zr0 = z_r[ou]; L.zrW = z_r[ou - 1]; L.zrS = z_r[ou - P];
    L.ak0 = Akv[o]; L.akW = Akv[o - 1]; L.akS = Akv[o - P];
    L.up = ust[ou]; L.vp = vst[ou];
    L.rur = L.rui = L.rvr = L.rvi = 0.0;
    if (istart >= 1) { L.rui = ru_i[o]; L.rvi = rv_i[o]; }
    if (istart == 2) { L.rur = ru_r[o]; L.rvr = rv_r[o]; }
    return L;
  };
  // BODYFORCE (pre_step3d.F:931-937, :1036-1042): the stresses do not enter as boundary fluxes (dt * 0 = 0, the reference's FC = 0)
  double FCum = p.dt * (p.bodyforce ? 0.0 : f.bustr[o2]), FCvm = p.dt * (p.bodyforce ? 0.0 : f.bvstr[o2]);
  const double sus = p.bodyforce ? 0.0 : f.sustr[o2], svs = p.bodyforce ? 0.0 : f.svstr[o2];
  double uk = ust[o2 + PL], vk = vst[o2 + PL];
  double zk0 = z_r[o2 + PL], zkW = z_r[o2 + PL - 1], zkS = z_r[o2 + PL - P];
  auto level = [&](const Lvl& cur, const Lvl&, int k) {
    const int o = o2 + k * PL;
    {
      double FCk;
      if (k < N) {
        const double c = 1.0 / (cur.zr0 + cur.zrW - zk0 - zkW);
        FCk = cff3 * c * (cur.up - uk) * (cur.ak0 + cur.akW);
      } else {
        FCk = p.dt * sus;
      }
      const double a = uk * 0.5 * (cur.hz0 + cur.hzW);
      const double d = FCk - FCum;
      double x;
      if (istart == 0) x = a + d;
      else if (istart == 1) { const double c3 = 0.5 * DCu; x = a - c3 * cur.rui + d; }
      else x = a + DCu * ((5.0 / 12.0) * cur.rur - (16.0 / 12.0) * cur.rui) + d;
      unw[o] = x;
      FCum = FCk;
    }
    if (dov) {
      double FCk;
      if (k < N) {
        const double c = 1.0 / (cur.zr0 + cur.zrS - zk0 - zkS);
        FCk = cff3 * c * (cur.vp - vk) * (cur.ak0 + cur.akS);
      } else {
        FCk = p.dt * svs;
      }
      const double a = vk * 0.5 * (cur.hz0 + cur.hzS);
      const double d = FCk - FCvm;
      double x;
      if (istart == 0) x = a + d;
      else if (istart == 1) { const double c3 = 0.5 * DCv; x = a - c3 * cur.rvi + d; }
      else x = a + DCv * ((5.0 / 12.0) * cur.rvr - (16.0 / 12.0) * cur.rvi) + d;
      vnw[o] = x;
      FCvm = FCk;
    }
    uk = cur.up; vk = cur.vp; zk0 = cur.zr0; zkW = cur.zrW; zkS = cur.zrS;
  };
  {
    sweep_levels<PRU_PP>(N, load_level, level);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prsgrd32_tile, phase 1 (ROMS/Nonlinear/prsgrd32.h:236-290): per-column harmonic-mean slopes and downward
// integration of the pressure P over IstrU-1:Iend x JstrV-1:Jend.
__global__ void __launch_bounds__(128) k_prsgrd32_P(Par p, Flds f) {
  const int i = p.Istr - 1 + blockIdx.x * blockDim.x + threadIdx.x;   // IstrU-1 .. Iend
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;            // JstrV-1 .. Jend
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, PL = p.PL, o2 = j * p.P + i;
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  const double GRho = p.g / p.rho0, HalfGRho = 0.5 * GRho;
  const double* __restrict__ rho = f.rho;
  const double* __restrict__ z_r = f.z_r;
  double* __restrict__ P3 = f.P3;
  // rolling window (k+1, k, k-1) of rho and z_r, marching downward; the loads of a batch of levels are issued together
  constexpr int CH = 5;
  const int oN = o2 + N * PL;
  double r_kp = rho[oN], z_kp = z_r[oN];                               // level k+1 (starts at N)
  double r_k = rho[oN - PL], z_k = z_r[oN - PL];                       // level k   (starts at N-1)
  const double zwN = f.z_w[oN];
  // harmonic means at level N: raw(N) = raw(N-1)
  double rawR_k = r_kp - r_k, rawZ_k = z_kp - z_k;                     // raw(N) := raw(N-1)
  double rawR_km = rawR_k, rawZ_km = rawZ_k;                           // raw(N-1)
  double c = 2.0 * rawR_k * rawR_km;
  double dR_kp = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;             // dR(N)
  double dZ_kp = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);          // dZ(N)
  const double cff1 = 1.0 / (z_kp - z_k);
  const double cff2 = 0.5 * (r_kp - r_k) * (zwN - z_kp) * cff1;
  double Pk = p.atm_press ? p.g * zwN + (100.0 / p.rho0) * (f.Pair[o2] - 1013.25) + GRho * (r_kp + cff2) * (zwN - z_kp)   // ATM_PRESS :265-267
                          : p.g * zwN + GRho * (r_kp + cff2) * (zwN - z_kp);
  P3[oN] = Pk;
  for (int kt = N - 1; kt >= 1; kt -= CH) {
    double lr[CH], lz[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int km = (kt - q - 1 >= 1) ? kt - q - 1 : 1;               // level k-1 (clamped; unused when k = 1)
      lr[q] = rho[o2 + km * PL]; lz[q] = z_r[o2 + km * PL];
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int k = kt - q;
      if (k >= 1) {
        // dR(k) = harmonic(raw(k), raw(k-1)); raw(0) = raw(1)
        rawR_k = r_kp - r_k; rawZ_k = z_kp - z_k;
        if (k > 1) { rawR_km = r_k - lr[q]; rawZ_km = z_k - lz[q]; } else { rawR_km = rawR_k; rawZ_km = rawZ_k; }
        c = 2.0 * rawR_k * rawR_km;
        const double dR_k = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;
        const double dZ_k = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);
        Pk = Pk + HalfGRho * ((r_kp + r_k) * (z_kp - z_k) -
                              OneFifth * ((dR_kp - dR_k) * (z_kp - z_k - OneTwelfth * (dZ_kp + dZ_k)) -
                                          (dZ_kp - dZ_k) * (r_kp - r_k - OneTwelfth * (dR_kp + dR_k))));
        P3[o2 + k * PL] = Pk;
        dR_kp = dR_k; dZ_kp = dZ_k;
        r_kp = r_k; z_kp = z_k; r_k = lr[q]; z_k = lz[q];
      }
    }
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
// WJ: the weighted Jacobian of Song & Wright (WJ_GRADP, :236-254, :317-335).
template <bool WJ>
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
    if (p.atm_press) phi = phi + (100.0 / p.rho0) * (f.Pair[o2 + i] - f.Pair[o2 + i - s]);                  // ATM_PRESS
    phi = phi + (fac2 + fac1 * (rho[oN] + rho[oN - s])) * (z_w[oN] - z_w[oN - s]);
    R[oN] = -0.5 * (Hz[oN] + Hz[oN - s]) * phi * met;
    for (int k = N - 1; k >= 1; --k) {
      const int o = o2 + k * p.PL + i, ou = o + p.PL;
      double c1, c4;
      if (WJ) {
        const double w1 = 1.0 / ((z_r[ou] - z_r[o]) * (z_r[ou - s] - z_r[o - s]));
        const double w2 = z_r[o] - z_r[o - s] + z_r[ou] - z_r[ou - s];
        const double w3 = z_r[ou] - z_r[o] - z_r[ou - s] + z_r[o - s];
        const double gamma = 0.125 * w1 * w2 * w3;
        c1 = (1.0 + gamma) * (rho[ou] - rho[ou - s]) + (1.0 - gamma) * (rho[o] - rho[o - s]);
        c4 = (1.0 + gamma) * (z_r[ou] - z_r[ou - s]) + (1.0 - gamma) * (z_r[o] - z_r[o - s]);
      } else {
        c1 = rho[ou] - rho[ou - s] + rho[o] - rho[o - s];
        c4 = z_r[ou] - z_r[ou - s] + z_r[o] - z_r[o - s];
      }
      const double c2 = rho[ou] + rho[ou - s] - rho[o] - rho[o - s];
      const double c3 = z_r[ou] + z_r[ou - s] - z_r[o] - z_r[o - s];
      phi = phi + fac3 * (c1 * c3 - c2 * c4);
      R[o] = -0.5 * (Hz[o] + Hz[o - s]) * phi * met;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prsgrd40_tile (ROMS/Nonlinear/prsgrd40.h:176-270; PJ_GRADP): finite-volume pressure Jacobian (Lin, 1997).  Thread per column,
// marching downward: the hydrostatic sums P of the column and of its western and southern neighbours are carried in registers
// (the reference keeps a private 3-D P; each neighbour sum is re-accumulated by the same expression, so it is bit-identical),
// with the face integrals FC(k) of the two directions.  No divisions; rho, Hz, z_w are each read once per column they belong to.
__global__ void __launch_bounds__(128) k_prsgrd40(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P + i;
  const double* __restrict__ rho = f.rho;
  const double* __restrict__ z_w = f.z_w;
  const double* __restrict__ Hz = f.Hz;
  double* __restrict__ ru = f.ru[p.nrhs];
  double* __restrict__ rv = f.rv[p.nrhs];
  const bool dov = (j >= p.JstrV);
  const double cff = 0.5 * p.g, cff1 = p.g / p.rho0;
  const double on_u = f.on_u[o2], om_v = f.om_v[o2];
  const int oN = o2 + N * p.PL;
  const double dzu = z_w[oN - 1] - z_w[oN], dzv = z_w[oN - P] - z_w[oN];          // z_w(i-1,j,N) - z_w(i,j,N), z_w(i,j-1,N) - z_w(i,j,N)
  double P0 = 0.0, PW = 0.0, PS = 0.0, FCu = 0.0, FCv = 0.0;
  if (p.atm_press) {                                                   // prsgrd40.h:194-196
    const double fac = 100.0 / p.g;
    P0 = P0 + fac * (f.Pair[o2] - 1013.25); PW = PW + fac * (f.Pair[o2 - 1] - 1013.25); PS = PS + fac * (f.Pair[o2 - P] - 1013.25);
  }
  for (int k = N; k >= 1; --k) {
    const int o = o2 + k * p.PL, od = o - p.PL;
    const double hz0 = Hz[o], hzW = Hz[o - 1];
    const double P0n = P0 + hz0 * rho[o], PWn = PW + hzW * rho[o - 1];
    const double FX0 = 0.5 * hz0 * (P0 + P0n), FXW = 0.5 * hzW * (PW + PWn);
    const double zw0 = z_w[od];
    {
      const double dh = zw0 - z_w[od - 1];
      const double FCn = 0.5 * dh * (P0n + PWn);
      ru[o] = (cff * (hzW + hz0) * dzu + cff1 * (FXW - FX0 + FCu - FCn)) * on_u;
      FCu = FCn;
    }
    if (dov) {
      const double hzS = Hz[o - P];
      const double PSn = PS + hzS * rho[o - P];
      const double FXS = 0.5 * hzS * (PS + PSn);
      const double dh = zw0 - z_w[od - P];
      const double FCn = 0.5 * dh * (P0n + PSn);
      rv[o] = (cff * (hzS + hz0) * dzv + cff1 * (FXS - FX0 + FCv - FCn)) * om_v;
      FCv = FCn; PS = PSn;
    }
    P0 = P0n; PW = PWn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// t3dmix2_s_tile (ROMS/Nonlinear/t3dmix2_s.h:198-301): harmonic mixing of tracers along s-surfaces.  One thread per
// column; the level-independent leading factors 0.25*(diff2+diff2)*pmon_u / pnom_v of the four face fluxes are formed
// once, and the operands of level k+1 are requested before level k is computed.
template <int NTR>
__global__ void __launch_bounds__(128) k_t3dmix2_s(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, PL = p.PL, o2 = j * P + i;
  const double* __restrict__ Hz = f.Hz;
  const double cff = p.dt * f.pm[o2] * f.pn[o2];
  const double* __restrict__ tr[NTR];
  double* __restrict__ tn[NTR];
  double cW[NTR], cE[NTR], cS[NTR], cN[NTR];
#pragma unroll
  for (int it = 0; it < NTR; ++it) {
    tr[it] = f.t[p.nrhs][it]; tn[it] = f.t[p.nnew][it];
    const double* __restrict__ d2 = f.diff2[it];
    const double d0 = d2[o2];
    cW[it] = 0.25 * (d0 + d2[o2 - 1]) * f.pmon_u[o2];
    cE[it] = 0.25 * (d2[o2 + 1] + d0) * f.pmon_u[o2 + 1];
    cS[it] = 0.25 * (d0 + d2[o2 - P]) * f.pnom_v[o2];
    cN[it] = 0.25 * (d2[o2 + P] + d0) * f.pnom_v[o2 + P];
  }
  struct Lvl { double hz0, hzW, hzE, hzS, hzN, t0[NTR], tW[NTR], tE[NTR], tS[NTR], tN[NTR], tn[NTR]; };
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * PL;
    Lvl L;
    L.hz0 = Hz[o]; L.hzW = Hz[o - 1]; L.hzE = Hz[o + 1]; L.hzS = Hz[o - P]; L.hzN = Hz[o + P];
#pragma unroll
    for (int it = 0; it < NTR; ++it) {
      L.t0[it] = tr[it][o]; L.tW[it] = tr[it][o - 1]; L.tE[it] = tr[it][o + 1]; L.tS[it] = tr[it][o - P]; L.tN[it] = tr[it][o + P];
      L.tn[it] = tn[it][o];
    }
    return L;
  };
  auto level = [&](const Lvl& cur, const Lvl&, int k) {
    const int o = o2 + k * PL;
#pragma unroll
    for (int it = 0; it < NTR; ++it) {
      const double t0 = cur.t0[it];
      const double FXi = cW[it] * (cur.hz0 + cur.hzW) * (t0 - cur.tW[it]);
      const double FXip = cE[it] * (cur.hzE + cur.hz0) * (cur.tE[it] - t0);
      const double FEj = cS[it] * (cur.hz0 + cur.hzS) * (t0 - cur.tS[it]);
      const double FEjp = cN[it] * (cur.hzN + cur.hz0) * (cur.tN[it] - t0);
      const double cff1 = cff * (FXip - FXi);
      const double cff2 = cff * (FEjp - FEj);
      const double cff3 = cff1 + cff2;
      tn[it][o] = cur.tn[it] + cff3;
    }
  };
  {
    sweep_levels<true>(N, load_level, level);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// t3dmix4_s_tile (ROMS/Nonlinear/t3dmix4_s.h:215-476; TS_DIF4 + MIX_S_TS): biharmonic mixing of tracers along s-surfaces, the
// harmonic operator applied twice with diff4 = SQRT(ABS(tnu4)).  No vertical coupling: one thread per (i,j,k).  The reference
// keeps LapT in a private 2-D scratch array; here a thread evaluates LapT at its own cell and at its four neighbours (each from
// that cell's four face fluxes, the expressions of :243-348 verbatim, so the five values are bit-identical to the scratch array's)
// and applies the second operator (:410-471) in registers: t(nrhs) and Hz are read on a 13-point diamond that the L1 serves, no
// scratch traffic.  LapT = 0 outside the closed southern / northern walls (:378-404); EW periodic (ghost columns of t(nrhs), Hz).
__device__ __forceinline__ double lapT4(const Par& p, const Flds& f, const double* __restrict__ tr, const double* __restrict__ d4,
                                        int i, int j, int ok) {
  if (j < 1 || j > p.Mm) return 0.0;
  const int P = p.P, o2 = j * P + i, o = o2 + ok;
  const double* __restrict__ Hz = f.Hz;
  const double h0 = Hz[o], t0 = tr[o], d0 = d4[o2];
  double cff = 0.25 * (d0 + d4[o2 - 1]) * f.pmon_u[o2];
  const double FX0 = cff * (h0 + Hz[o - 1]) * (t0 - tr[o - 1]);
  cff = 0.25 * (d4[o2 + 1] + d0) * f.pmon_u[o2 + 1];
  const double FX1 = cff * (Hz[o + 1] + h0) * (tr[o + 1] - t0);
  cff = 0.25 * (d0 + d4[o2 - P]) * f.pnom_v[o2];
  const double FE0 = cff * (h0 + Hz[o - P]) * (t0 - tr[o - P]);
  cff = 0.25 * (d4[o2 + P] + d0) * f.pnom_v[o2 + P];
  const double FE1 = cff * (Hz[o + P] + h0) * (tr[o + P] - t0);
  cff = 1.0 / h0;
  return f.pm[o2] * f.pn[o2] * cff * (FX1 - FX0 + FE1 - FE0);
}
__global__ void __launch_bounds__(128) k_t3dmix4_s(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int P = p.P, ok = k * p.PL, o2 = j * P + i, o = o2 + ok;
  const double* __restrict__ Hz = f.Hz;
  const double h0 = Hz[o], hW = Hz[o - 1], hE = Hz[o + 1], hS = Hz[o - P], hN = Hz[o + P];
  const double dtmn = p.dt * f.pm[o2] * f.pn[o2];
  for (int it = 0; it < p.NT; ++it) {
    const double* __restrict__ tr = f.t[p.nrhs][it];
    const double* __restrict__ d4 = f.diff4[it];
    double* __restrict__ tn = f.t[p.nnew][it];
    const double L0 = lapT4(p, f, tr, d4, i, j, ok), LW = lapT4(p, f, tr, d4, i - 1, j, ok), LE = lapT4(p, f, tr, d4, i + 1, j, ok);
    const double LS = lapT4(p, f, tr, d4, i, j - 1, ok), LN = lapT4(p, f, tr, d4, i, j + 1, ok);
    const double d0 = d4[o2];
    double cff = 0.25 * (d0 + d4[o2 - 1]) * f.pmon_u[o2];
    const double FX0 = cff * (h0 + hW) * (L0 - LW);
    cff = 0.25 * (d4[o2 + 1] + d0) * f.pmon_u[o2 + 1];
    const double FX1 = cff * (hE + h0) * (LE - L0);
    cff = 0.25 * (d0 + d4[o2 - P]) * f.pnom_v[o2];
    const double FE0 = cff * (h0 + hS) * (L0 - LS);
    cff = 0.25 * (d4[o2 + P] + d0) * f.pnom_v[o2 + P];
    const double FE1 = cff * (hN + h0) * (LN - L0);
    const double cff1 = dtmn * (FX1 - FX0);
    const double cff2 = dtmn * (FE1 - FE0);
    const double cff3 = cff1 + cff2;
    tn[o] = tn[o] - cff3;
  }
}

// ---------------------------------------------------------------------------------------------------------------
static inline dim3 g2(const Par& p, dim3 b, int ni, int nj, int nz = 1) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz); }

template <int H, bool SRC>
static void launch_pre_t_v(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2); dim3 g = g2(p, b, xspan(p), p.Mm);
  g.x *= p.NT;
  if (p.fuse_tmix) {
    if (p.vadv == 0) k_pre_step3d_t<H, 0, true, SRC><<<g, b, 0, s>>>(p, f);
    else if (p.vadv == 1) k_pre_step3d_t<H, 1, true, SRC><<<g, b, 0, s>>>(p, f);
    else if (p.vadv == 2) k_pre_step3d_t<H, 2, true, SRC><<<g, b, 0, s>>>(p, f);
    else k_pre_step3d_t<H, 3, true, SRC><<<g, b, 0, s>>>(p, f);
    return;
  }
  if (p.vadv == 0) k_pre_step3d_t<H, 0, false, SRC><<<g, b, 0, s>>>(p, f);
  else if (p.vadv == 1) k_pre_step3d_t<H, 1, false, SRC><<<g, b, 0, s>>>(p, f);
  else if (p.vadv == 2) k_pre_step3d_t<H, 2, false, SRC><<<g, b, 0, s>>>(p, f);
  else k_pre_step3d_t<H, 3, false, SRC><<<g, b, 0, s>>>(p, f);
}
template <int H>
static void launch_pre_t_h(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.solar_source || p.lmd_nonlocal) launch_pre_t_v<H, true>(p, f, s); else launch_pre_t_v<H, false>(p, f, s);
}
void launch_pre_step3d_t(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.hadv == 0) launch_pre_t_h<0>(p, f, s);
  else if (p.hadv == 1) launch_pre_t_h<1>(p, f, s);
  else if (p.hadv == 2) launch_pre_t_h<2>(p, f, s);
  else launch_pre_t_h<3>(p, f, s);
}
void launch_pre_step3d_uv(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2);
  k_pre_step3d_uv<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
}
void launch_prsgrd(const Par& p, const Flds& f, int dj_gradps, cudaStream_t s) {
  if (dj_gradps == 2) {                                                  // PJ_GRADP
    dim3 b(64, 2);
    k_prsgrd40<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  } else if (dj_gradps == 1) {
    dim3 b(64, 2);
    k_prsgrd32_P<<<g2(p, b, p.Iend - p.Istr + 2, p.Mm), b, 0, s>>>(p, f);
    dim3 b2(64, 4);
    k_prsgrd32_R<<<g2(p, b2, p.Iend - p.Istr + 1, p.Mm, p.N), b2, 0, s>>>(p, f);
  } else {
    dim3 b(64, 2);
    if (dj_gradps == 3) k_prsgrd31<true><<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);     // WJ_GRADP
    else k_prsgrd31<false><<<g2(p, b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  }
}
void launch_t3dmix4_s(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2);
  k_t3dmix4_s<<<g2(p, b, p.Iend - p.Istr + 1, p.Mm, p.N), b, 0, s>>>(p, f);
}
void launch_t3dmix2_s(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2); dim3 g = g2(p, b, p.Iend - p.Istr + 1, p.Mm);
  if (p.NT == 1) k_t3dmix2_s<1><<<g, b, 0, s>>>(p, f);
  else k_t3dmix2_s<2><<<g, b, 0, s>>>(p, f);
}

}  // namespace rb

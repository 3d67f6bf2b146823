// Surface forcing and vertical mixing of the shipped BENCHMARK cpp set (ROMS/Include/benchmark.h): bulk_flux (BULK_FLUXES with
// LONGWAVE; COARE 3.0) and lmd_vmix (LMD_MIXING with LMD_RIMIX, LMD_CONVEC, LMD_SKPP, LMD_NONLOCAL, RI_SPLINES).
#include "dev.cuh"
#include "kernels.h"

namespace rb {

namespace {
// mod_scalars.F:431-444, :1415-1422
constexpr double kPi = 3.14159265358979323846;
constexpr double Cp = 3985.0, StefBo = 5.67e-8, emmiss = 0.97, vonKar = 0.41;
constexpr double blk_Cpa = 1004.67, blk_Cpw = 4000.0, blk_Rgas = 287.1, blk_Zabl = 600.0, blk_beta = 1.2;
// mod_scalars.F:1552-1629
constexpr double lmd_Ri0 = 0.7, lmd_bvfcon = -2.0e-5, lmd_nu0c = 0.01, lmd_nu0m = 10.0e-4, lmd_nu0s = 10.0e-4;
constexpr double lmd_Cstar = 10.0, lmd_Cv = 1.25, lmd_Ric = 0.3, lmd_am = 1.257, lmd_as = -28.86, lmd_betaT = -0.2, lmd_cekman = 0.7,
                 lmd_cmonob = 1.0, lmd_cm = 8.36, lmd_cs = 98.96, lmd_epsilon = 0.1, lmd_zetam = -0.2, lmd_zetas = -1.0;
}  // namespace
// Jerlov water types (mod_scalars.F:1502-1512)
__constant__ double k_lmd_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};
__constant__ double k_lmd_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};
__constant__ double k_lmd_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};

// ---------------------------------------------------------------------------------------------------------------
// bulk_psiu / bulk_psit (bulk_flux.F:950-1066)
__device__ __forceinline__ double bulk_psiu(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = pow(1.0 - 15.0 * ZoL, 0.25);
    const double psik = 2.0 * log(0.5 * (1.0 + x)) + log(0.5 * (1.0 + x * x)) - 2.0 * atan(x) + 0.5 * kPi;
    double cff = sqrt(3.0);
    const double y = pow(1.0 - 10.15 * ZoL, r3);
    const double psic = 1.5 * log(r3 * (1.0 + y + y * y)) - cff * atan((1.0 + 2.0 * y) / cff) + kPi / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = dmin(50.0, 0.35 * ZoL);
  return -((1.0 + ZoL) + 0.6667 * (ZoL - 14.28) / exp(cff) + 8.525);
}
__device__ __forceinline__ double bulk_psit(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = pow(1.0 - 15.0 * ZoL, 0.5);
    const double psik = 2.0 * log(0.5 * (1.0 + x));
    double cff = sqrt(3.0);
    const double y = pow(1.0 - 34.15 * ZoL, r3);
    const double psic = 1.5 * log(r3 * (1.0 + y + y * y)) - cff * atan((1.0 + 2.0 * y) / cff) + kPi / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = dmin(50.0, 0.35 * ZoL);
  return -(pow(1.0 + 2.0 * ZoL, 1.5) + 0.6667 * (ZoL - 14.28) / exp(cff) + 8.525);
}

// bulk_flux_tile, the rho-point part (bulk_flux.F:396-855, :868-877): one thread per rho point of Istr-1:Iend x 0:Mm+1.  The wind
// stress components at the rho points go to the scratch planes Taux / Tauy (column Istr-1 included: sustr(Istr) needs it).
__global__ void __launch_bounds__(128) k_bulk_flux(Par p, Flds f) {
  const int i = p.Istr - 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int q = j * p.P + i;
  const double g = p.g, eps = 1.0e-20, r3 = 1.0 / 3.0;
  const double blk_ZW = p.blk_ZW, blk_ZT = p.blk_ZT, blk_ZQ = p.blk_ZQ;
  const double Uair = f.Uwind[q], Vair = f.Vwind[q];
  const double Wmag = sqrt(Uair * Uair + Vair * Vair);
  const double PairM = f.Pair[q];
  const double TairC = f.Tair[q], TairK = TairC + 273.16;
  const double TseaC = f.t[p.nrhs][p.itemp - 1][q + p.N * p.PL], TseaK = TseaC + 273.16;
  const double RH = f.Hair[q];
  const double rain = f.rain[q], cloud = f.cloud[q];
  // LONGWAVE (Berliand), :448-458
  double cff = (0.7859 + 0.03477 * TairC) / (1.0 + 0.00412 * TairC);
  const double e_sat = pow(10.0, cff);
  const double vap_p = e_sat * RH;
  const double cff2 = TairK * TairK * TairK;
  const double cff1 = cff2 * TairK;
  const double LRad = -emmiss * StefBo * (cff1 * (0.39 - 0.05 * sqrt(vap_p)) * (1.0 - 0.6823 * cloud * cloud) + cff2 * 4.0 * (TseaK - TairK));
  // specific humidities, :514-544
  cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * exp(17.502 * TairC / (240.97 + TairC));
  const double Qair = 0.62197 * (cff / (PairM - 0.378 * cff));
  double Q;
  if (RH < 2.0) { cff = cff * RH; Q = 0.62197 * (cff / (PairM - 0.378 * cff)); }
  else Q = RH / 1000.0;
  cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * exp(17.502 * TseaC / (240.97 + TseaC));
  cff = cff * 0.98;
  const double Qsea = 0.62197 * (cff / (PairM - 0.378 * cff));
  // :554-578
  const double rhoAir = PairM * 100.0 / (blk_Rgas * TairK * (1.0 + 0.61 * Q));
  const double VisAir = 1.326e-5 * (1.0 + TairC * (6.542e-3 + TairC * (8.301e-6 - 4.84e-9 * TairC)));
  const double Hlv = (2.501 - 0.00237 * TseaC) * 1.0e+6;
  double Wgus = 0.5;
  double delW = sqrt(Wmag * Wmag + Wgus * Wgus);
  const double delQ = Qsea - Q, delT = TseaC - TairC;
  // neutral coefficients, :582-592
  double ZoW = 0.0001;
  const double u10 = delW * log(10.0 / ZoW) / log(blk_ZW / ZoW);
  double Wstar = 0.035 * u10;
  const double Zo10 = 0.011 * Wstar * Wstar / g + 0.11 * VisAir / Wstar;
  double t_ = vonKar / log(10.0 / Zo10);
  const double Cd10 = t_ * t_;
  const double Ch10 = 0.00115;
  const double Ct10 = Ch10 / sqrt(Cd10);
  const double ZoT10 = 10.0 / exp(vonKar / Ct10);
  t_ = vonKar / log(blk_ZW / Zo10);
  double Cd = t_ * t_;
  // Richardson number, :596-612
  const double Ct = vonKar / log(blk_ZT / ZoT10);
  const double CC = vonKar * Ct / Cd;
  const double delTc = 0.0, delQc = 0.0;
  const double Ribcu = -blk_ZW / (blk_Zabl * 0.004 * (blk_beta * blk_beta * blk_beta));
  const double Ri = -g * blk_ZW * ((delT - delTc) + 0.61 * TairK * delQ) / (TairK * delW * delW);
  double Zetu;
  if (Ri < 0.0) Zetu = CC * Ri / (1.0 + Ri / Ribcu);
  else Zetu = CC * Ri / (1.0 + 3.0 * Ri / CC);
  const double L10 = blk_ZW / Zetu;
  // first guesses, :616-624
  Wstar = delW * vonKar / (log(blk_ZW / Zo10) - bulk_psiu(blk_ZW / L10));
  double Tstar = -(delT - delTc) * vonKar / (log(blk_ZT / ZoT10) - bulk_psit(blk_ZT / L10));
  // blk_ZQ == blk_ZT (the usual case, roms_benchmark1.in): the humidity terms repeat the temperature ones bit for bit
  const bool zq_is_zt = (blk_ZQ == blk_ZT);
  double Qstar = -(delQ - delQc) * vonKar / (log(blk_ZQ / ZoT10) - bulk_psit(blk_ZQ / L10));
  // Charnock, :629-637
  double charn;
  if (delW > 18.0) charn = 0.018;
  else if (10.0 < delW && delW <= 18.0) charn = 0.011 + 0.125 * (0.018 - 0.011) * (delW - 10.0);
  else charn = 0.011;
  // iteration, :655-715
#pragma unroll 1
  for (int Iter = 1; Iter <= 3; ++Iter) {
    ZoW = charn * Wstar * Wstar / g + 0.11 * VisAir / (Wstar + eps);
    const double Rr = ZoW * Wstar / VisAir;
    const double ZoQ = dmin(1.15e-4, 5.5e-5 / pow(Rr, 0.6));
    const double ZoT = ZoQ;
    const double ZoL = vonKar * g * blk_ZW * (Tstar * (1.0 + 0.61 * Q) + 0.61 * TairK * Qstar) / (TairK * Wstar * Wstar * (1.0 + 0.61 * Q) + eps);
    const double L = blk_ZW / (ZoL + eps);
    const double Wpsi = bulk_psiu(ZoL);
    const double Tpsi = bulk_psit(blk_ZT / L);
    const double Qpsi = zq_is_zt ? Tpsi : bulk_psit(blk_ZQ / L);
    Wstar = dmax(eps, delW * vonKar / (log(blk_ZW / ZoW) - Wpsi));
    const double lT = log(blk_ZT / ZoT);
    const double lQ = zq_is_zt ? lT : log(blk_ZQ / ZoQ);                  // ZoT == ZoQ (:668)
    Tstar = -(delT - delTc) * vonKar / (lT - Tpsi);
    Qstar = -(delQ - delQc) * vonKar / (lQ - Qpsi);
    const double Bf = -g / TairK * Wstar * (Tstar + 0.61 * TairK * Qstar);
    if (Bf > 0.0) Wgus = blk_beta * pow(Bf * blk_Zabl, r3);
    else Wgus = 0.2;
    delW = sqrt(Wmag * Wmag + Wgus * Wgus);
  }
  // fluxes, :775-855
  const double Wspeed = sqrt(Wmag * Wmag + Wgus * Wgus);
  Cd = Wstar * Wstar / (Wspeed * Wspeed + eps);
  const double Hs = -blk_Cpa * rhoAir * Wstar * Tstar;
  const double diffw = 2.11e-5 * pow(TairK / 273.16, 1.94);
  const double diffh = 0.02411 * (1.0 + TairC * (3.309e-3 - 1.44e-6 * TairC)) / (rhoAir * blk_Cpa);
  cff = Qair * Hlv / (blk_Rgas * TairK * TairK);
  const double wet_bulb = 1.0 / (1.0 + 0.622 * (cff * Hlv * diffw) / (blk_Cpa * diffh));
  const double Hsr = rain * wet_bulb * blk_Cpw * ((TseaC - TairC) + (Qsea - Q) * Hlv / blk_Cpa);
  const double SHeat = (Hs + Hsr);
  const double Hl = -Hlv * rhoAir * Wstar * Qstar;
  const double upvel = -1.61 * Wstar * Qstar - (1.0 + 1.61 * Q) * Wstar * Tstar / TairK;
  const double Hlw = rhoAir * Hlv * upvel * Q;
  const double LHeat = (Hl + Hlw);
  const double Taur = 0.85 * rain * Wmag;
  cff = rhoAir * Cd * Wspeed;
  f.Taux[q] = (cff * Uair + Taur * copysign(1.0, Uair));
  f.Tauy[q] = (cff * Vair + Taur * copysign(1.0, Vair));
  if (i >= p.Istr) {
    // kinematic fluxes, :868-877, and the periodic images (:930-941)
    const double Hscale = 1.0 / (p.rho0 * Cp);
    const double lr = LRad * Hscale, lh = -LHeat * Hscale, sh = -SHeat * Hscale;
    const int o2 = j * p.P;
    st_w(f.lrflx, o2, i, lr, p);
    st_w(f.lhflx, o2, i, lh, p);
    st_w(f.shflx, o2, i, sh, p);
    st_w(f.stflux[p.itemp - 1], o2, i, (f.srflx[q] + lr + lh + sh), p);
  }
}

// bulk_flux_tile, the kinematic wind stress at u and v points (bulk_flux.F:904-924) + periodic images (:942-947)
__global__ void __launch_bounds__(256) k_bulk_stress(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int o2 = j * p.P;
  const double cff = 0.5 / p.rho0;
  st_w(f.sustr, o2, i, cff * (f.Taux[o2 + i - 1] + f.Taux[o2 + i]), p);
  if (j >= 1) st_w(f.svstr, o2, i, cff * (f.Tauy[o2 - p.P + i] + f.Tauy[o2 + i]), p);
}

// ---------------------------------------------------------------------------------------------------------------
// turbulent velocity scales (lmd_skpp.F:454-476, :700-722, :835-857)
__device__ __forceinline__ void lmd_wscale(double Ustar, double sigma, double Bf, double& wm, double& ws) {
  const double r3 = 1.0 / 3.0, small = 1.0e-20;
  const double Ustar3 = Ustar * Ustar * Ustar;
  const double zetahat = vonKar * sigma * Bf;
  const double zetapar = zetahat / (Ustar3 + small);
  if (zetahat >= 0.0) {
    wm = vonKar * Ustar / (1.0 + 5.0 * zetapar);
    ws = wm;
  } else {
    if (zetapar > lmd_zetam) wm = vonKar * Ustar * pow(1.0 - 16.0 * zetapar, 0.25);
    else wm = vonKar * pow(lmd_am * Ustar3 - lmd_cm * zetahat, r3);
    if (zetapar > lmd_zetas) ws = vonKar * Ustar * pow(1.0 - 16.0 * zetapar, 0.5);
    else ws = vonKar * pow(lmd_as * Ustar3 - lmd_cs * zetahat, r3);
  }
}

// lmd_vmix: lmd_vmix_tile (lmd_vmix.F:182-347), lmd_skpp_tile (lmd_skpp.F:246-923) and lmd_finish_tile (lmd_vmix.F:508-659) in one
// pass, one thread per column.  The parabolic splines of the shear (dU, dV) are the same in lmd_vmix_tile and lmd_skpp_tile and are
// built once.  Only the four spline columns live in thread-local memory: the shortwave fraction / buoyancy flux of a level and the
// interior coefficients are evaluated where they are used (the boundary layer needs them at a few levels near the surface), so
// Akv / Akt are written once and nothing else is staged.  The column Iend-1 copy of the eastern edge (lmd_vmix.F:568-575) is
// k_lmd_east.  NC > 0: N is the compile-time constant NC (thread-local columns sized for it).
#ifndef LMD_MINB
#define LMD_MINB 5          // 0.64 ms (3: 0.77, 4: 0.67, 6: 0.66)
#endif
#ifndef LMD_PF
#define LMD_PF 6          // L2 prefetch distance (levels) of the upward spline sweep
#endif
struct LmdCol {           // per-column constants of the shortwave / buoyancy profile
  double zwN, Bo, Bosol, fac1, fac2, fac3;
};
// lmd_swfrac.F:66-80 (Zscale = -1) and the total buoyancy flux (lmd_skpp.F:312-316) at depth Z below the surface
__device__ __forceinline__ double lmd_swdk(const LmdCol& c, double Z) { return exp(Z * c.fac1) * c.fac3 + exp(Z * c.fac2) * (1.0 - c.fac3); }
__device__ __forceinline__ double lmd_bflux(const LmdCol& c, double swdk) { return (c.Bo + c.Bosol * (1.0 - swdk)); }
// interior coefficients of one level (lmd_vmix.F:230-234, :309-347): Richardson-number mixing + internal waves
__device__ __forceinline__ void lmd_interior(double bv, double dUk, double dVk, double& av, double& at) {
  const double e14 = 1.0e-14;
  double shear2 = dUk * dUk + dVk * dVk;
  const double Rig = bv / (shear2 + e14);
  double cff = dmin(1.0, dmax(0.0, Rig) / lmd_Ri0);
  double nu_sx = 1.0 - cff * cff;
  nu_sx = nu_sx * nu_sx * nu_sx;
  shear2 = bv / (Rig + e14);
  cff = shear2 * shear2 / (shear2 * shear2 + 16.0e-10);
  nu_sx = cff * nu_sx;
  cff = 1.0 / sqrt(dmax(bv, 1.0e-7));
  av = 1.0e-6 * cff + lmd_nu0m * nu_sx;
  at = 1.0e-7 * cff + lmd_nu0s * nu_sx;
}

template <int NC>
__global__ void __launch_bounds__(128, LMD_MINB) k_lmd_vmix(Par p, Flds f) {
  constexpr int NA = (NC > 0 ? NC : MAXN) + 1;
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = NC > 0 ? NC : p.N;
  const int P = p.P, PL = p.PL, o2 = j * P, q = o2 + i;
  const int it_T = p.itemp - 1, it_S = p.isalt - 1;
  const bool salt = p.salinity != 0;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ z_w = f.z_w;
  const double* __restrict__ pden = f.pden;
  const double* __restrict__ bvf = f.bvf;
  const double* __restrict__ u = f.u[p.nstp];
  const double* __restrict__ v = f.v[p.nstp];
  double FC[NA], dR[NA], dU[NA], dV[NA];
  const double g = p.g, gorho0 = p.g / p.rho0, eps = 1.0e-10;
  const double lmd_Cg = lmd_Cstar * vonKar * pow(lmd_cs * vonKar * lmd_epsilon, 1.0 / 3.0);      // mod_scalars.F:4330
  const double Vtc = lmd_Cv * sqrt(-lmd_betaT) / (sqrt(lmd_cs * lmd_epsilon) * lmd_Ric * vonKar * vonKar);
  LmdCol c;
  c.zwN = z_w[q + N * PL];
  const double zwN = c.zwN;
  const double stT = f.stflx[it_T][q], stS = salt ? f.stflx[it_S][q] : 0.0, sr = f.srflx[q], al = f.alpha[q];
  // lmd_skpp.F:256-290
  double sl_dpth = lmd_epsilon * (zwN - f.hsbl[q]);
  double Ustar;
  {
    const double a = 0.5 * (f.sustr[q] + f.sustr[q + 1]), b = 0.5 * (f.svstr[q] + f.svstr[q + P]);
    Ustar = sqrt(sqrt(a * a + b * b));
  }
  if (salt) c.Bo = g * (al * (stT - sr) - f.beta[q] * stS);
  else c.Bo = g * al * (stT - sr);
  c.Bosol = g * al * sr;
  {
    int J = (int)f.Jwtype[q];
    J = (J < 1) ? 1 : (J > 9 ? 9 : J);
    c.fac1 = -1.0 / k_lmd_mu1[J - 1]; c.fac2 = -1.0 / k_lmd_mu2[J - 1]; c.fac3 = k_lmd_r1[J - 1];
  }
  // parabolic splines of pden, u, v at W points, upward sweep (lmd_skpp.F:342-364 == lmd_vmix.F:196-216 for dU, dV); the operands
  // of level k+2 are requested while level k is computed
  FC[0] = 0.0; dR[0] = 0.0; dU[0] = 0.0; dV[0] = 0.0;
  {
    struct Lv { double hz, pd, u0, u1, v0, v1; };
    auto ld = [&](int k) -> Lv {
      const int o = q + k * PL;
      pf_up<LMD_PF>(Hz, o, k, N, PL); pf_up<LMD_PF>(pden, o, k, N, PL); pf_up<LMD_PF>(u, o, k, N, PL); pf_up<LMD_PF>(v, o, k, N, PL);
      return Lv{Hz[o], pden[o], u[o], u[o + 1], v[o], v[o + P]};
    };
    Lv a = ld(1), b = ld(2), n2 = b;
#pragma unroll 2
    for (int k = 1; k <= N - 1; ++k) {
      if (k + 2 <= N) n2 = ld(k + 2);
      const double cff = 1.0 / (2.0 * b.hz + a.hz * (2.0 - FC[k - 1]));
      FC[k] = cff * b.hz;
      dR[k] = cff * (6.0 * (b.pd - a.pd) - a.hz * dR[k - 1]);
      dU[k] = cff * (3.0 * (b.u0 - a.u0 + b.u1 - a.u1) - a.hz * dU[k - 1]);
      dV[k] = cff * (3.0 * (b.v0 - a.v0 + b.v1 - a.v1) - a.hz * dV[k - 1]);
      a = b; b = n2;
    }
  }
  // downward sweep (lmd_skpp.F:365-377) ...
  dR[N] = 0.0; dU[N] = 0.0; dV[N] = 0.0;
  for (int k = N - 1; k >= 1; --k) {
    dR[k] = dR[k] - FC[k] * dR[k + 1];
    dU[k] = dU[k] - FC[k] * dU[k + 1];
    dV[k] = dV[k] - FC[k] * dV[k + 1];
  }
  // ... and the bulk Richardson criterion, top down, until FC changes sign (lmd_skpp.F:435-508, SASHA)
  int ksbl = 1;
  double hsbl = z_w[q + PL];
  {
    const double c13 = 1.0 / 3.0, c16 = 1.0 / 6.0;
    const double hzN = Hz[q + N * PL];
    const double Rref = pden[q + N * PL] + hzN * (c13 * dR[N] + c16 * dR[N - 1]);
    const double Uref = 0.5 * (u[q + N * PL] + u[q + 1 + N * PL]) + hzN * (c13 * dU[N] + c16 * dU[N - 1]);
    const double Vref = 0.5 * (v[q + N * PL] + v[q + P + N * PL]) + hzN * (c13 * dV[N] + c16 * dV[N - 1]);
    double FCk = 0.0;
    for (int k = N; k >= 2; --k) {
      const int o = q + k * PL;
      const double zwm = z_w[o - PL];
      const double depth = zwN - zwm;
      const double Bf = lmd_bflux(c, lmd_swdk(c, depth));
      const double sigma = (Bf < 0.0) ? dmin(sl_dpth, depth) : depth;
      double wm, ws;
      lmd_wscale(Ustar, sigma, Bf, wm, ws);
      const double hz = Hz[o];
      const double Rk = pden[o] - hz * (c13 * dR[k - 1] + c16 * dR[k]);
      const double Uk = 0.5 * (u[o] + u[o + 1]) - hz * (c13 * dU[k - 1] + c16 * dU[k]);
      const double Vk = 0.5 * (v[o] + v[o + P]) - hz * (c13 * dV[k - 1] + c16 * dV[k]);
      const double Ritop = -gorho0 * (Rref - Rk) * depth;
      const double du = Uref - Uk, dv = Vref - Vk;
      const double Ribot = du * du + dv * dv + Vtc * depth * ws * sqrt(fabs(bvf[o - PL]));
      const double FCm = Ritop - lmd_Ric * Ribot;
      if (FCm > 0.0) {
        hsbl = (z_w[o] * FCm - zwm * FCk) / (FCm - FCk);
        ksbl = k;
        break;
      }
      FCk = FCm;
    }
  }
  // limits under stable forcing (lmd_skpp.F:551-589)
  const double zw0 = z_w[q];
  {
    const double Bfsfc = lmd_bflux(c, lmd_swdk(c, zwN - hsbl));
    if (Ustar > 0.0 && Bfsfc > 0.0) {
      const double hekman = lmd_cekman * Ustar / dmax(fabs(f.f[q]), eps);
      const double hmonob = lmd_cmonob * Ustar * Ustar * Ustar / dmax(vonKar * Bfsfc, eps);
      hsbl = (zwN - dmin(dmin(hekman, hmonob), zwN - hsbl));
    }
    hsbl = dmin(hsbl, zwN);
    hsbl = dmax(hsbl, zw0);
  }
  st_r_grad(f.hsbl, o2, i, j, hsbl, p);                       // bc_r2d (lmd_skpp.F:636-647)
  // new boundary-layer index (:651-660)
  ksbl = 1;
  for (int k = N; k >= 2; --k)
    if (z_w[q + (k - 1) * PL] < hsbl) { ksbl = k; break; }
  // buoyancy flux and velocity scales at hsbl (:666-737)
  const double Bfsfc = lmd_bflux(c, lmd_swdk(c, zwN - hsbl));
  sl_dpth = lmd_epsilon * (zwN - hsbl);
  double wm, ws;
  {
    const double cff = (Bfsfc > 0.0) ? 1.0 : lmd_epsilon;
    const double sigma = cff * (zwN - hsbl);
    lmd_wscale(Ustar, sigma, Bfsfc, wm, ws);
  }
  const double f1 = 5.0 * dmax(0.0, Bfsfc) * vonKar / (Ustar * Ustar * Ustar * Ustar + eps);
  // shape functions at hsbl (:739-818)
  const double zbl = zwN - hsbl;
  double Gm1, Gt1, Gs1 = 0.0, dGm1dS, dGt1dS, dGs1dS = 0.0;
  if (hsbl > z_w[q + PL]) {
    const int k = ksbl;
    const double zk = z_w[q + k * PL], zkm = z_w[q + (k - 1) * PL];
    const double cff = 1.0 / (zk - zkm);
    const double cff_dn = cff * (hsbl - zkm);
    const double cff_up = cff * (zk - hsbl);
    // levels 1..N-1 hold this step's interior values; level N is whatever the array holds (never set by lmd_vmix)
    double avk, atk, ask, avm, atm;
    lmd_interior(bvf[q + (k - 1) * PL], dU[k - 1], dV[k - 1], avm, atm);
    if (k <= N - 1) { lmd_interior(bvf[q + k * PL], dU[k], dV[k], avk, atk); ask = atk; }
    else { avk = f.Akv[q + k * PL]; atk = f.Akt[it_T][q + k * PL]; ask = salt ? f.Akt[it_S][q + k * PL] : 0.0; }
    double K_bl = cff_dn * avk + cff_up * avm;
    double dK_bl = cff * (avk - avm);
    Gm1 = K_bl / (zbl * wm + eps);
    dGm1dS = dmin(0.0, -dK_bl / (wm + eps) - K_bl * f1);
    K_bl = cff_dn * atk + cff_up * atm;
    dK_bl = cff * (atk - atm);
    Gt1 = K_bl / (zbl * ws + eps);
    dGt1dS = dmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
    if (salt) {
      K_bl = cff_dn * ask + cff_up * atm;
      dK_bl = cff * (ask - atm);
      Gs1 = K_bl / (zbl * ws + eps);
      dGs1dS = dmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
    }
  } else {
    ksbl = 0;
    const double a = 0.5 * (f.bustr[q] + f.bustr[q + 1]), b = 0.5 * (f.bvstr[q] + f.bvstr[q + P]);
    const double Ustarb = sqrt(sqrt(a * a + b * b));
    const double dK_bl = vonKar * Ustarb;
    const double K_bl = dK_bl * (hsbl - zw0);
    Gm1 = K_bl / (zbl * wm + eps);
    dGm1dS = dmin(0.0, -dK_bl / (wm + eps) - K_bl * f1);
    Gt1 = K_bl / (zbl * ws + eps);
    dGt1dS = dmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
    Gs1 = Gt1; dGs1dS = dGt1dS;
  }
  f.ksbl[q] = (double)ksbl;
  // boundary-layer coefficients above ksbl (:826-923), LMD_CONVEC on every level (lmd_vmix.F:508-532), boundary copies + bc_w3d
  double* __restrict__ Akv = f.Akv;
  double* __restrict__ AkT = f.Akt[it_T];
  double* __restrict__ AkS = salt ? f.Akt[it_S] : nullptr;
  double* __restrict__ ghT = f.ghats[it_T];
  double* __restrict__ ghS = salt ? f.ghats[it_S] : nullptr;
  for (int k = 0; k <= N; k += N) {                           // levels 0 and N: ghats keeps its first value, Akv / Akt only get the copies
    const int o = o2 + k * PL;
    const double swdk = lmd_swdk(c, zwN - z_w[o + i]);
    const double c0 = 1.0 - (0.5 + copysign(0.5, lmd_bflux(c, swdk)));
    ghT[o + i] = -c0 * (stT - sr + sr * (1.0 - swdk));
    if (salt) ghS[o + i] = c0 * stS;
    st_r_grad(Akv, o, i, j, Akv[o + i], p);
    st_r_grad(AkT, o, i, j, AkT[o + i], p);
    if (salt) st_r_grad(AkS, o, i, j, AkS[o + i], p);
  }
#pragma unroll 2
  for (int k = 1; k <= N - 1; ++k) {
    const int o = o2 + k * PL;
    const double bv = bvf[o + i];
    double av, at;
    lmd_interior(bv, dU[k], dV[k], av, at);
    double as = at;
    double gT = 0.0, gS = 0.0;
    if (k > ksbl) {
      const double depth = zwN - z_w[o + i];
      const double swdk = lmd_swdk(c, depth);
      const double Bf = lmd_bflux(c, swdk);
      double sigma = (Bf < 0.0) ? dmin(sl_dpth, depth) : depth;
      lmd_wscale(Ustar, sigma, Bf, wm, ws);
      sigma = depth / (zbl + eps);
      const double a1 = sigma - 2.0, a2 = 3.0 - 2.0 * sigma, a3 = sigma - 1.0;
      const double Gm = a1 + a2 * Gm1 + a3 * dGm1dS;
      const double Gt = a1 + a2 * Gt1 + a3 * dGt1dS;
      av = depth * wm * (1.0 + sigma * Gm);
      at = depth * ws * (1.0 + sigma * Gt);
      if (salt) {
        const double Gs = a1 + a2 * Gs1 + a3 * dGs1dS;
        as = depth * ws * (1.0 + sigma * Gs);
      }
      const double c0 = 1.0 - (0.5 + copysign(0.5, Bf));
      const double cff = lmd_Cg * c0 / (zbl * ws + eps);
      gT = cff * (-c0 * (stT - sr + sr * (1.0 - swdk)));
      gS = cff * (c0 * stS);
    }
    ghT[o + i] = gT;
    if (salt) ghS[o + i] = gS;
    double cff = dmax(bv, lmd_bvfcon);
    cff = dmin(1.0, (lmd_bvfcon - cff) / lmd_bvfcon);
    double nu_sxc = 1.0 - cff * cff;
    nu_sxc = nu_sxc * nu_sxc * nu_sxc;
    st_r_grad(Akv, o, i, j, av + lmd_nu0c * nu_sxc, p);
    st_r_grad(AkT, o, i, j, at + lmd_nu0c * nu_sxc, p);
    if (salt) st_r_grad(AkS, o, i, j, as + lmd_nu0c * nu_sxc, p);
  }
}

// lmd_finish_tile on the tile that owns the eastern edge (lmd_vmix.F:568-575): column Iend-1 of Akv / Akt takes the values of
// column Iend (every level, rows 0..Mm+1 after the southern / northern copies), and so does its periodic image.
__global__ void __launch_bounds__(128) k_lmd_east(Par p, Flds f, int NAT) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;          // 0..Mm+1
  const int k = blockIdx.y;                                     // 0..N
  if (j > p.Mm + 1) return;
  const int o = j * p.P + k * p.PL;
  for (int a = 0; a <= NAT; ++a) {
    double* __restrict__ A = (a == 0) ? f.Akv : f.Akt[a - 1];
    const double x = A[o + p.Lm];
    A[o + p.Lm - 1] = x;
    if (p.ew_wrap) A[o - 1] = x;
  }
}

void launch_bulk_flux(const Par& p, const Flds& f, cudaStream_t s) {
  const int ni = p.Iend - p.Istr + 1;
  { dim3 b(64, 2), g((ni + 1 + b.x - 1) / b.x, (p.Mm + 2 + b.y - 1) / b.y); k_bulk_flux<<<g, b, 0, s>>>(p, f); }
  { dim3 b(64, 4), g((ni + b.x - 1) / b.x, (p.Mm + 2 + b.y - 1) / b.y); k_bulk_stress<<<g, b, 0, s>>>(p, f); }
}

void launch_lmd_vmix(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2), g((xspan(p) + b.x - 1) / b.x, (p.Mm + b.y - 1) / b.y);
  if (p.N == 30) k_lmd_vmix<30><<<g, b, 0, s>>>(p, f);
  else k_lmd_vmix<0><<<g, b, 0, s>>>(p, f);
  if (p.Iend == p.Lm) {                                       // this launch covers the eastern edge of the domain
    dim3 bb(128), gg((p.Mm + 2 + 127) / 128, p.N + 1);
    k_lmd_east<<<gg, bb, 0, s>>>(p, f, p.salinity ? 2 : 1);
  }
}

}  // namespace rb

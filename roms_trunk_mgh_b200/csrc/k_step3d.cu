// step3d_uv_tile and step3d_t_tile: correctors with implicit vertical mixing (parabolic-spline tridiagonal systems,
// SPLINES_VVISC / SPLINES_VDIFF).
//
// Thread-per-column Thomas solver, xi is the coalesced axis.  The solver is a dependent chain (one FP64 division per
// level), so what matters on B200 is how many columns an SM keeps in flight: the kernels below hold per column only the
// two forward-sweep arrays CF(k), DC(k) in shared memory ([k][thread], conflict free, 2*N*8 bytes per column) and stream
// everything else level by level with software-pipelined loads, which keeps them near 100 registers and 14 warps per
// SM for N = 30.  Values needed again by the back substitution (the pre-solve right-hand side, Hz, Ak) are re-read
// while they are still in L2 (the columns in flight are ~20 MB) instead of being kept in registers.  Arithmetic and
// summation order are exactly those of the reference loops.
#include "dev.cuh"
#include "kernels.h"
#include "k_adv.cuh"

namespace rb {

#ifndef S3D_TS
#define S3D_TS 64
#endif
#ifndef S3D_CH
#define S3D_CH 6
#endif
#ifndef S3T_CH
#define S3T_CH 10
#endif
#ifndef S3D_MINB
#define S3D_MINB 7
#endif
#ifndef S3U_DEPTH
#define S3U_DEPTH 0        // levels in flight in pass 1 of k_step3d_uv (0: sweep_levels<S3U_PP>)
#endif
#ifndef S3U_PP
#define S3U_PP true
#endif
#ifndef S3T_PP
#define S3T_PP true      // ping-pong level buffers in k_step3d_t (see sweep_levels): pays once the register budget allows ~164
#endif
#ifndef S3T_MINB
#define S3T_MINB 5
#endif
#ifndef S3U_RING
#define S3U_RING 0        // > 0: pass 1 of k_step3d_uv streams its operands through a per-thread shared-memory ring of this many
#endif                    // levels filled with cp.async (LDGSTS): levels in flight without a register cost
#ifndef S3T_RING
#define S3T_RING 0        // same for pass 1 of k_step3d_t (18 operands per level)
#endif
#ifndef S3U_PF
#define S3U_PF 4          // L2 prefetch distance (levels) of pass 1 in k_step3d_uv
#endif
#ifndef S3T_PF
#define S3T_PF 8          // same for k_step3d_t
#endif
#ifndef S3U_PF2
#define S3U_PF2 0         // L2 prefetch distance (levels, downward) of pass 2 in k_step3d_uv
#endif
#ifndef S3T_PF2
#define S3T_PF2 0         // same for k_step3d_t
#endif
constexpr int TS = S3D_TS;   // threads (columns) per block
constexpr int CH = S3D_CH;   // levels per batch of independent loads in the downward / coupling passes of k_step3d_uv
constexpr int CHT = S3T_CH;  // ... and in the back substitution of k_step3d_t (profiles/r01_sweep_batch_depth.log)
static_assert(EDGE_W % TS == 0, "split launches need CTA widths that divide EDGE_W");

// One forward-elimination step of the spline system for row m = k-1 once level k is known
// (step3d_uv.F:344-375, step3d_t.F:1370-1405): h/o/AK suffix m = level k-1, k = level k, AKmm = AK(k-2).
__device__ __forceinline__ void spline_forward(double hm, double om, double hk, double ok, double AKmm, double AKm, double AKk,
                                               double dx, double dt, double& CFp, double& DCp) {
  const double FCv = (1.0 / 6.0) * hm - dt * AKmm * om;
  const double CFv = (1.0 / 6.0) * hk - dt * AKk * ok;
  const double BCv = (1.0 / 3.0) * (hm + hk) + dt * AKm * (om + ok);
  const double cff = 1.0 / (BCv - FCv * CFp);
  CFp = cff * CFv;
  DCp = cff * (dx - FCv * DCp);
}

// ---------------------------------------------------------------------------------------------------------------
// step3d_uv_tile (ROMS/Nonlinear/step3d_uv.F:288-950 time step + implicit viscosity + vertical-mean replacement;
// :956-965 closed-wall BCs; :1002-1432 coupling with DU_avg1/DU_avg2, ubar/vbar reset, corrected Huon/Hvom;
// :1438-1461 periodic images).  DIR = 0: u-points, DIR = 1: v-points.
// SPL = false: SPLINES_VVISC not defined -- the centred tridiagonal system of :397-462 / :730-795 (u is NOT divided by Hz first; the
// matrix carries Hz): the same delayed forward elimination with FC(k) = -lambda dt / 0.5 * AK(k) / (sum of the two columns' dz_r).
template <int DIR, bool SPL = true>
__global__ void __launch_bounds__(TS, S3D_MINB) k_step3d_uv(Par p, Flds f) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const int N = p.N;
  double* sA = sm + tid;                 // CF(k) -> x(k) -> hv(k)      slot (k-1)*TS
  double* sB = sm + N * TS + tid;        // DC(k) -> Hzk(k) -> d(k)
  const int i = xcol0(p, blockIdx.x * TS) + tid;
  const int j = (DIR ? 2 : 1) + blockIdx.y;                                // u: Jstr..Jend, v: JstrV..Jend
  if (i > p.Iend) return;
  const int P = p.P, Mm = p.Mm, PL = p.PL, o2 = j * P + i;
  const int s = DIR ? P : 1;                                               // stride to the (i-1) / (j-1) neighbour
  double* __restrict__ X = DIR ? f.v[p.nnew] : f.u[p.nnew];
  const double* __restrict__ R = DIR ? f.rv[p.nrhs] : f.ru[p.nrhs];
  double* __restrict__ HUV = DIR ? f.Hvom : f.Huon;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Akv = f.Akv;
  const double* __restrict__ Davg1 = DIR ? f.DV_avg1 : f.DU_avg1;
  const double* __restrict__ Davg2 = DIR ? f.DV_avg2 : f.DU_avg2;
  const double* __restrict__ met = DIR ? f.om_v : f.on_u;                  // on_u (u) / om_v (v)
  double* __restrict__ bar1 = DIR ? f.vbar[1] : f.ubar[1];
  double* __restrict__ bar2 = DIR ? f.vbar[2] : f.ubar[2];
  double cffAB;
  if (p.istart == 0) cffAB = 0.25 * p.dt;
  else if (p.istart == 1) cffAB = 0.25 * p.dt * 3.0 / 2.0;
  else cffAB = 0.25 * p.dt * 23.0 / 12.0;
  const double DC0 = cffAB * (f.pm[o2] + f.pm[o2 - s]) * (f.pn[o2] + f.pn[o2 - s]);
  const double dt = p.dt;

  // ---- pass 1 (upward): right-hand side x(k) and the forward elimination; loads of level k+1 are issued before level k
  // is computed
  struct Lvl { double ak0, ak1, h0, h1, x, r, z0, z1; };
  const double* __restrict__ z_r = f.z_r;
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * PL;
    pf_l2(HUV + o);                                                        // first touched by the coupling pass
    pf_up<S3U_PF>(Akv, o, k, N, PL); pf_up<S3U_PF>(Hz, o, k, N, PL); pf_up<S3U_PF>(X, o, k, N, PL); pf_up<S3U_PF>(R, o, k, N, PL);
    return Lvl{Akv[o - s], Akv[o], Hz[o - s], Hz[o], X[o], R[o], SPL ? 0.0 : z_r[o], SPL ? 0.0 : z_r[o - s]};
  };
  double AKm = 0.5 * (Akv[o2 - s] + Akv[o2]), AKmm = 0.0, AKN;
  double topDC = 0.0;                                                      // !SPL: DC(N) of the centred system
  {
    double hm = 0.0, om = 0.0, xm = 0.0, CFp = 0.0, DCp = 0.0;
    double zm0 = 0.0, zm1 = 0.0, FCmm = 0.0;                               // !SPL: z_r of level k-1 in the two columns, FC(k-2)
    const double cffd = -p.lambda * dt / 0.5;
    auto level = [&](const Lvl& cur, const Lvl&, int k) {
      const double AKk = 0.5 * (cur.ak0 + cur.ak1);
      const double hk = 0.5 * (cur.h0 + cur.h1);
      const double ok = 1.0 / hk;
      double xv = cur.x + DC0 * cur.r;
      if (SPL) xv = xv * ok;
      if (k >= 2) {
        if (SPL) spline_forward(hm, om, hk, ok, AKmm, AKm, AKk, xv - xm, dt, CFp, DCp);
        else {                                                             // row m = k-1 of the centred system
          const double cff1 = 1.0 / (cur.z0 + cur.z1 - zm0 - zm1);
          const double FCm = cffd * cff1 * AKm;
          const double BCm = hm - FCm - FCmm;
          const double cf = 1.0 / (BCm - FCmm * CFp);
          CFp = cf * FCm;
          DCp = cf * (xm - FCmm * DCp);
          FCmm = FCm;
        }
        sA[(k - 2) * TS] = CFp; sB[(k - 2) * TS] = DCp;
      }
      AKmm = AKm; AKm = AKk; hm = hk; om = ok; xm = xv; zm0 = cur.z0; zm1 = cur.z1;
    };
#if S3U_RING > 0
    // Register-free deep prefetch: every thread copies the six operands of level k + S3U_RING straight into its own slots of a
    // shared-memory ring (one commit group per level) and only ever reads back what it copied itself, so no barrier is needed;
    // S3U_RING levels (x 48 bytes per thread) are in flight while the dependent FP64 chain of level k runs.
    {
      constexpr int D = S3U_RING;
      double* ring = sm + 2 * N * TS + tid;                                // [D][6][TS]
      auto issue = [&](int k) {
        if (k <= N) {
          const int o = o2 + k * PL;
          double* r = ring + ((k - 1) % D) * 6 * TS;
          cp_async8(r, Akv + o - s); cp_async8(r + TS, Akv + o); cp_async8(r + 2 * TS, Hz + o - s); cp_async8(r + 3 * TS, Hz + o);
          cp_async8(r + 4 * TS, X + o); cp_async8(r + 5 * TS, R + o);
          pf_l2(HUV + o);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int k = 1; k <= D; ++k) issue(k);
      for (int k = 1; k <= N; ++k) {
        cp_async_wait<D - 1>();
        const double* r = ring + ((k - 1) % D) * 6 * TS;
        const Lvl cur{r[0], r[TS], r[2 * TS], r[3 * TS], r[4 * TS], r[5 * TS], 0.0, 0.0};
        issue(k + D);
        level(cur, cur, k);
      }
    }
#elif S3U_DEPTH >= 2
    sweep_levels_deep<S3U_DEPTH>(N, load_level, level);
#else
    sweep_levels<S3U_PP>(N, load_level, level);
#endif
    AKN = AKm;
    if (!SPL) topDC = (xm - FCmm * DCp) / (hm - FCmm - FCmm * CFp);         // :436-441: BC(N) = Hzk(N) - FC(N) - FC(N-1), FC(N) = 0
  }
  // ---- pass 2 (downward): back substitution fused with the update x(k) += dt*oHz(k)*(AK(k)*DC(k) - AK(k-1)*DC(k-1));
  // x(k) and Hzk(k) replace CF(k), DC(k) in shared memory (both already consumed at level k+1)
  if (!SPL) {                                                              // centred system: x(k) = DC(k) - CF(k) x(k+1) (:442-462)
    double dk = topDC;
    for (int k = N; k >= 1; --k) {
      const int o = o2 + k * PL;
      if (k < N) dk = sB[(k - 1) * TS] - sA[(k - 1) * TS] * dk;
      sA[(k - 1) * TS] = dk;
      sB[(k - 1) * TS] = 0.5 * (Hz[o - s] + Hz[o]);
    }
  } else {
    double dk = 0.0;                       // DC(N) = 0
    double ak = dk * AKN;
    for (int kt = N; kt >= 1; kt -= CH) {
      double lx[CH], lr[CH], lh0[CH], lh1[CH], la0[CH], la1[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int k = (kt - q >= 1) ? kt - q : 1;
        const int o = o2 + k * PL;
        pf_dn<S3U_PF2>(X, o, k, PL); pf_dn<S3U_PF2>(R, o, k, PL); pf_dn<S3U_PF2>(Hz, o, k, PL); pf_dn<S3U_PF2>(Akv, o - PL, k, PL);
        lx[q] = X[o]; lr[q] = R[o]; lh0[q] = Hz[o - s]; lh1[q] = Hz[o]; la0[q] = Akv[o - PL - s]; la1[q] = Akv[o - PL];
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int k = kt - q;
        if (k >= 1) {
          const double hk = 0.5 * (lh0[q] + lh1[q]);
          const double ok = 1.0 / hk;
          double xv = lx[q] + DC0 * lr[q];
          xv = xv * ok;
          double dkm1 = 0.0, akm1 = 0.0;
          if (k > 1) {
            dkm1 = sB[(k - 2) * TS] - sA[(k - 2) * TS] * dk;
            akm1 = dkm1 * (0.5 * (la0[q] + la1[q]));
          }
          const double cff = dt * ok * (ak - akm1);
          sA[(k - 1) * TS] = xv + cff;
          sB[(k - 1) * TS] = hk;
          dk = dkm1; ak = akm1;
        }
      }
    }
  }
  // ---- replace the vertical mean with the one from the barotropic sub-cycle (:469-605)
  double dcm;
  {
    double cf0 = sB[0], dc0 = sA[0] * sB[0];
    for (int k = 2; k <= N; ++k) { const double hk = sB[(k - 1) * TS]; cf0 = cf0 + hk; dc0 = dc0 + sA[(k - 1) * TS] * hk; }
    const double m = met[o2];
    const double cff1 = 1.0 / (cf0 * m);
    dcm = (dc0 * m - Davg1[o2]) * cff1;
  }
  // ---- coupling (:1002-1432) for this row, then for the wall rows owned by the edge threads
  //   u: rows 0 and Mm+1 carry u = gamma2*u(wall-adjacent row) (u3dbc) and get their own coupling pass
  //   v: row 1 (the wall itself, v = 0 from v3dbc) is handled by the j = 2 thread, row Mm+1 by the j = Mm thread
  // own row: x(k) - dcm from shared memory; wall rows: scale * (the own-row values just stored to X)
  auto couple = [&](int jj, double scale, bool own) {
    const int q2 = jj * P + i;
    double dc0 = 0.0, cf0 = 0.0, fc0 = 0.0;
    const double mq = met[q2];
    const double cff = 0.5 * mq;
    for (int k = 1; k <= N; ++k) {
      double d, xs;
      if (own) {
        d = mq * sB[(k - 1) * TS];                       // (0.5*m)*(a+b) == m*(0.5*(a+b)) bitwise
        xs = sA[(k - 1) * TS] - dcm;
      } else {
        const int o = q2 + k * PL;
        d = cff * (Hz[o] + Hz[o - s]);
        xs = (scale == 0.0) ? 0.0 : scale * X[o2 + k * PL];
      }
      sA[(k - 1) * TS] = xs; sB[(k - 1) * TS] = d;
      dc0 = dc0 + d;
      cf0 = cf0 + d * xs;
    }
    dc0 = 1.0 / dc0;
    cf0 = dc0 * (cf0 - Davg1[q2]);
    const double b = dc0 * Davg1[q2];
    st_w(bar1, q2 - i, i, b, p);
    st_w(bar2, q2 - i, i, b, p);
    // boundary rows only: remove the mismatch of the vertical mean (:1132-1188, :1350-1406)
    const bool wall = DIR ? (jj == 1 || jj == Mm + 1) : (jj == 0 || jj == Mm + 1);
    for (int kt = N; kt >= 1; kt -= CH) {
      double lh[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) { const int k = (kt - q >= 1) ? kt - q : 1; lh[q] = HUV[q2 + k * PL]; }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int k = kt - q;
        if (k >= 1) {
          const double xs = sA[(k - 1) * TS];
          const double xk = wall ? (xs - cf0) : xs;
          st_w(X, q2 + k * PL - i, i, xk, p);
          const double hv = 0.5 * (lh[q] + xk * sB[(k - 1) * TS]);
          sA[(k - 1) * TS] = hv;
          fc0 = fc0 + hv;
        }
      }
    }
    fc0 = dc0 * (fc0 - Davg2[q2]);
    for (int k = 1; k <= N; ++k) st_w(HUV, q2 + k * PL - i, i, sA[(k - 1) * TS] - sB[(k - 1) * TS] * fc0, p);
  };
  couple(j, 1.0, true);
  if (DIR == 0) {
    if (j == 1) couple(0, p.gamma2, false);
    if (j == Mm) couple(Mm + 1, p.gamma2, false);
  } else {
    if (j == 2) couple(1, 0.0, false);
    if (j == Mm) couple(Mm + 1, 0.0, false);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// step3d_t_tile (ROMS/Nonlinear/step3d_t.F:388-876 horizontal advection of t(:,:,:,3,:), :883-1210 vertical advection,
// :1366-1427 implicit diffusion, :1551-1621 t3dbc + periodic images).  One thread per column and tracer.
// SPL = false: SPLINES_VDIFF not defined -- t(nnew) stays in Hz*t units after the advection (:1196-1198) and the centred tridiagonal
// system of :1430-1499 (matrix Hz(k) - FC(k) - FC(k-1), FC(k) = -lambda dt Akt(k) / (z_r(k+1) - z_r(k))) returns the tracer.
template <int HADV, int VADV, bool SPL = true>
__global__ void __launch_bounds__(TS, S3T_MINB) k_step3d_t(Par p, Flds f) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const int N = p.N;
  double* sCF = sm + tid;
  double* sDC = sm + N * TS + tid;
  // tracer index fastest: the CTAs of all tracers of one tile run back to back and share Huon, Hvom, W, Hz through L2
  const int itrc = blockIdx.x % p.NT;
  const int i = xcol0(p, (blockIdx.x / p.NT) * TS) + tid;
  const int j = 1 + blockIdx.y;
  if (i > p.Iend) return;
  const int P = p.P, PL = p.PL, o2 = j * P;
  const double* __restrict__ t3 = f.t[3][itrc];
  double* tn = f.t[p.nnew][itrc];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  const double* __restrict__ Akt = f.Akt[itrc];
  const double pm = f.pm[o2 + i], pn = f.pn[o2 + i];
  const double cffh = p.dt * pm * pn;
  const double dt = p.dt;
  // ---- pass 1 (upward): t(nnew) - advection, /Hz -> x(k), parked in t(nnew); forward elimination
  struct Lvl { AdvIn a; double tn, hz, W, akt, tk3, zr; };
  const double* __restrict__ z_r = f.z_r;
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * PL;
    Lvl L;
    pf_up<S3T_PF>(t3, o + i, k, N, PL); pf_up<S3T_PF>(Huon, o + i, k, N, PL); pf_up<S3T_PF>(Hvom, o + i, k, N, PL); pf_up<S3T_PF>(tn, o + i, k, N, PL);
    pf_up<S3T_PF>(Hz, o + i, k, N, PL); pf_up<S3T_PF>(W, o + i, k, N, PL); pf_up<S3T_PF>(Akt, o + i, k, N, PL);
    L.a = adv_load(t3, Huon, Hvom, o, i, j, p);
    L.tn = tn[o + i]; L.hz = Hz[o + i]; L.W = W[o + i]; L.akt = Akt[o + i];
    L.tk3 = t3[o2 + ((k + 2 <= N) ? (k + 2) : N) * PL + i];                // t(k+2), clamped
    L.zr = SPL ? 0.0 : z_r[o + i];
    return L;
  };
  double FCs[VADV == 3 ? MAXN + 1 : 1], CFs[VADV == 3 ? MAXN + 1 : 1];
  if (VADV == 3) vspline_flux<false>(t3, Hz, W, o2 + i, N, PL, FCs, CFs);         // SPLINES (step3d_t.F:894-937)
  double AKm = Akt[o2 + i], AKmm = 0.0, AKN;
  double topDC = 0.0;                                                      // !SPL: DC(N) of the centred system
  {
    double tkm1 = t3[o2 + PL + i], tk = tkm1, tkp1 = t3[o2 + 2 * PL + i], tkp2 = t3[o2 + 3 * PL + i];
    double FCm = 0.0, hm = 0.0, om = 0.0, xm = 0.0, CFp = 0.0, DCp = 0.0;
    double zm = 0.0, FDmm = 0.0;                                           // !SPL: z_r(k-1), FC(k-2) of the diffusion matrix
    const double cffd = -dt * p.lambda;
    auto level = [&](const Lvl& cur, const Lvl& nxt, int k) {
      const double hk = cur.hz;
      const double ok = 1.0 / hk;
      const double AKk = cur.akt;
      // horizontal: t(nnew) -= dt*pm*pn*(div)  (:861-873)
      double FXi, FXip, FEj, FEjp;
      hadv_fluxes_v<HADV>(cur.a, j, p.Mm, FXi, FXip, FEj, FEjp);
      const double c1 = cffh * (FXip - FXi);
      const double c2 = cffh * (FEjp - FEj);
      const double c3 = c1 + c2;
      double tv = cur.tn - c3;
      // vertical (:1189-1207)
      const double FCk = (VADV == 3) ? FCs[VADV == 3 ? k : 0] : ((k < N) ? vflux4<VADV>(tkm1, tk, tkp1, tkp2, k, N, cur.W) : 0.0);
      const double cv = cffh * (FCk - FCm);
      tv = tv - cv;
      if (SPL) tv = tv * ok;
      if (SPL) tn[o2 + k * PL + i] = tv;                                   // (the centred system needs no parked copy)
      if (k >= 2) {
        if (SPL) spline_forward(hm, om, hk, ok, AKmm, AKm, AKk, tv - xm, dt, CFp, DCp);
        else {                                                             // row m = k-1 (:1431-1461)
          const double cff1 = 1.0 / (cur.zr - zm);
          const double FDm = cffd * cff1 * AKm;
          const double BCm = hm - FDm - FDmm;
          const double cf = 1.0 / (BCm - FDmm * CFp);
          CFp = cf * FDm;
          DCp = cf * (xm - FDmm * DCp);
          FDmm = FDm;
        }
        sCF[(k - 2) * TS] = CFp; sDC[(k - 2) * TS] = DCp;
      }
      AKmm = AKm; AKm = AKk; hm = hk; om = ok; xm = tv; FCm = FCk; zm = cur.zr;
      tkm1 = tk; tk = tkp1; tkp1 = tkp2; tkp2 = nxt.tk3;
    };
#if S3T_RING > 0
    {
      constexpr int D = S3T_RING, NF = 18;
      double* ring = sm + 2 * N * TS + tid;                                // [D][NF][TS]
      const int om2 = (j > 1) ? -2 * P : -P, op2 = (j < p.Mm) ? 2 * P : P; // clamped rows, as in adv_load
      auto issue = [&](int k) {
        if (k <= N) {
          const int o = o2 + k * PL + i;
          double* r = ring + ((k - 1) % D) * NF * TS;
          cp_async8(r, t3 + o - 2); cp_async8(r + TS, t3 + o - 1); cp_async8(r + 2 * TS, t3 + o); cp_async8(r + 3 * TS, t3 + o + 1);
          cp_async8(r + 4 * TS, t3 + o + 2); cp_async8(r + 5 * TS, t3 + o + om2); cp_async8(r + 6 * TS, t3 + o - P);
          cp_async8(r + 7 * TS, t3 + o + P); cp_async8(r + 8 * TS, t3 + o + op2);
          cp_async8(r + 9 * TS, Huon + o); cp_async8(r + 10 * TS, Huon + o + 1); cp_async8(r + 11 * TS, Hvom + o); cp_async8(r + 12 * TS, Hvom + o + P);
          cp_async8(r + 13 * TS, tn + o); cp_async8(r + 14 * TS, Hz + o); cp_async8(r + 15 * TS, W + o); cp_async8(r + 16 * TS, Akt + o);
          cp_async8(r + 17 * TS, t3 + o2 + ((k + 3 <= N) ? (k + 3) : N) * PL + i);   // tk3 of level k+1: what `level` takes from nxt
        }
        cp_async_commit();
      };
#pragma unroll
      for (int k = 1; k <= D; ++k) issue(k);
      for (int k = 1; k <= N; ++k) {
        cp_async_wait<D - 1>();
        const double* r = ring + ((k - 1) % D) * NF * TS;
        Lvl cur;
        cur.a = AdvIn{r[0], r[TS], r[2 * TS], r[3 * TS], r[4 * TS], r[5 * TS], r[6 * TS], r[7 * TS], r[8 * TS], r[9 * TS], r[10 * TS], r[11 * TS], r[12 * TS]};
        cur.tn = r[13 * TS]; cur.hz = r[14 * TS]; cur.W = r[15 * TS]; cur.akt = r[16 * TS]; cur.tk3 = r[17 * TS]; cur.zr = 0.0;
        issue(k + D);
        level(cur, cur, k);                                                // only nxt.tk3 is read from the second argument
      }
    }
#else
    sweep_levels<S3T_PP>(N, load_level, level);
#endif
    AKN = AKm;
    if (!SPL) topDC = (xm - FDmm * DCp) / (hm - FDmm - FDmm * CFp);         // :1462-1470
  }
  // ---- pass 2 (downward): back substitution + update + t3dbc / periodic images
  if (!SPL) {                                                              // :1471-1499: t(k) = DC(k) - CF(k) t(k+1)
    double dk = topDC;
    for (int k = N; k >= 1; --k) {
      if (k < N) dk = sDC[(k - 1) * TS] - sCF[(k - 1) * TS] * dk;
      st_r_grad(tn, o2 + k * PL, i, j, dk, p);
    }
  } else {
    double dk = 0.0;
    double ak = dk * AKN;
    for (int kt = N; kt >= 1; kt -= CHT) {
      double lx[CHT], lh[CHT], la[CHT];
#pragma unroll
      for (int q = 0; q < CHT; ++q) {
        const int k = (kt - q >= 1) ? kt - q : 1;
        const int o = o2 + k * PL + i;
        pf_dn<S3T_PF2>(tn, o, k, PL); pf_dn<S3T_PF2>(Hz, o, k, PL); pf_dn<S3T_PF2>(Akt, o - PL, k, PL);
        lx[q] = tn[o]; lh[q] = Hz[o]; la[q] = Akt[o - PL];
      }
#pragma unroll
      for (int q = 0; q < CHT; ++q) {
        const int k = kt - q;
        if (k >= 1) {
          const double ok = 1.0 / lh[q];
          double dkm1 = 0.0, akm1 = 0.0;
          if (k > 1) {
            dkm1 = sDC[(k - 2) * TS] - sCF[(k - 2) * TS] * dk;
            akm1 = dkm1 * la[q];
          }
          const double cff = dt * ok * (ak - akm1);
          st_r_grad(tn, o2 + k * PL, i, j, lx[q] + cff, p);
          dk = dkm1; ak = akm1;
        }
      }
    }
  }
}

static size_t smem_cols(int N) { return (size_t)2 * N * TS * sizeof(double); }
static size_t smem_cols_t(int N) { return (size_t)(2 * N + 18 * S3T_RING) * TS * sizeof(double); }
static size_t smem_cols_uv(int N) { return (size_t)(2 * N + 6 * S3U_RING) * TS * sizeof(double); }
template <typename K>
static void allow_smem(K kern, size_t bytes) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }

void launch_step3d_uv(const Par& p, const Flds& f, cudaStream_t s) {
  const size_t sm = smem_cols_uv(p.N);
  static size_t allowed[MAXDEV] = {0};
  size_t& al = allowed[cur_dev()];
  if (sm > al) { allow_smem(k_step3d_uv<0>, sm); allow_smem(k_step3d_uv<1>, sm); al = sm; }
  const int nbx = (xspan(p) + TS - 1) / TS;
  if (p.nospl_vvisc) {                                                   // SPLINES_VVISC not defined
    static size_t allowed2[MAXDEV] = {0};
    size_t& a2 = allowed2[cur_dev()];
    if (sm > a2) { allow_smem(k_step3d_uv<0, false>, sm); allow_smem(k_step3d_uv<1, false>, sm); a2 = sm; }
    k_step3d_uv<0, false><<<dim3(nbx, p.Mm), TS, sm, s>>>(p, f);
    k_step3d_uv<1, false><<<dim3(nbx, p.Mm - 1), TS, sm, s>>>(p, f);
    return;
  }
  k_step3d_uv<0><<<dim3(nbx, p.Mm), TS, sm, s>>>(p, f);
  k_step3d_uv<1><<<dim3(nbx, p.Mm - 1), TS, sm, s>>>(p, f);
}

template <int H, int V>
static void launch_s3t(const Par& p, const Flds& f, cudaStream_t s) {
  const size_t sm = smem_cols_t(p.N);
  static size_t allowed[MAXDEV] = {0};
  size_t& al = allowed[cur_dev()];
  if (sm > al) { allow_smem(k_step3d_t<H, V>, sm); allow_smem(k_step3d_t<H, V, false>, sm); al = sm; }
  const int nbx = (xspan(p) + TS - 1) / TS;
  if (p.nospl_vdiff) k_step3d_t<H, V, false><<<dim3(nbx * p.NT, p.Mm), TS, sm, s>>>(p, f);       // SPLINES_VDIFF not defined
  else k_step3d_t<H, V><<<dim3(nbx * p.NT, p.Mm), TS, sm, s>>>(p, f);
}
template <int H>
static void launch_s3t_v(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.vadv == 0) launch_s3t<H, 0>(p, f, s);
  else if (p.vadv == 1) launch_s3t<H, 1>(p, f, s);
  else if (p.vadv == 2) launch_s3t<H, 2>(p, f, s);
  else launch_s3t<H, 3>(p, f, s);
}
void launch_step3d_t(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.hadv == 0) launch_s3t_v<0>(p, f, s);
  else if (p.hadv == 1) launch_s3t_v<1>(p, f, s);
  else if (p.hadv == 2) launch_s3t_v<2>(p, f, s);
  else launch_s3t_v<3>(p, f, s);
}

}  // namespace rb

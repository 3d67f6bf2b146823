// step3d_uv_tile and step3d_t_tile: correctors with implicit vertical mixing (parabolic-spline tridiagonal systems,
// SPLINES_VVISC / SPLINES_VDIFF).  Thread-per-column Thomas solver: the forward sweep keeps CF(k), DC(k) in
// thread-private arrays, the back substitution runs in the same thread; xi stays the coalesced axis.
#include "dev.cuh"
#include "kernels.h"
#include "k_adv.cuh"

namespace rb {

// Parabolic-spline implicit vertical mixing for one column (step3d_uv.F:344-396, step3d_t.F:1370-1427).
// x[1..N] in/out, Hzk[1..N], oHz[1..N], AK[0..N].
__device__ __forceinline__ void spline_implicit(double* x, const double* Hzk, const double* oHz, const double* AK, int N, double dt,
                                                double* CF, double* DC) {
  CF[0] = 0.0; DC[0] = 0.0;
  for (int k = 1; k <= N - 1; ++k) {
    const double FCk = (1.0 / 6.0) * Hzk[k] - dt * AK[k - 1] * oHz[k];
    const double CFk = (1.0 / 6.0) * Hzk[k + 1] - dt * AK[k + 1] * oHz[k + 1];
    const double BCk = (1.0 / 3.0) * (Hzk[k] + Hzk[k + 1]) + dt * AK[k] * (oHz[k] + oHz[k + 1]);
    const double cff = 1.0 / (BCk - FCk * CF[k - 1]);
    CF[k] = cff * CFk;
    DC[k] = cff * (x[k + 1] - x[k] - FCk * DC[k - 1]);
  }
  DC[N] = 0.0;
  for (int k = N - 1; k >= 1; --k) DC[k] = DC[k] - CF[k] * DC[k + 1];
  for (int k = 1; k <= N; ++k) {
    DC[k] = DC[k] * AK[k];
    const double cff = dt * oHz[k] * (DC[k] - DC[k - 1]);
    x[k] = x[k] + cff;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// step3d_uv_tile (ROMS/Nonlinear/step3d_uv.F:288-950 time step + implicit viscosity + vertical-mean replacement;
// :956-965 closed-wall BCs; :1002-1432 coupling with DU_avg1/DU_avg2, ubar/vbar reset, corrected Huon/Hvom;
// :1438-1461 periodic images).  DIR = 0: u-points, DIR = 1: v-points.
template <int DIR>
__global__ void __launch_bounds__(128) k_step3d_uv(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR ? 2 : 1) + blockIdx.y * blockDim.y + threadIdx.y;     // u: Jstr..Jend, v: JstrV..Jend
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, Mm = p.Mm, o2 = j * P + i;
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
  double x[MAXN + 1], Hzk[MAXN + 1], oHz[MAXN + 1], AK[MAXN + 1], CF[MAXN + 1], DC[MAXN + 1];
  double cffAB;
  if (p.istart == 0) cffAB = 0.25 * p.dt;
  else if (p.istart == 1) cffAB = 0.25 * p.dt * 3.0 / 2.0;
  else cffAB = 0.25 * p.dt * 23.0 / 12.0;
  AK[0] = 0.5 * (Akv[o2 - s] + Akv[o2]);
  const double DC0 = cffAB * (f.pm[o2] + f.pm[o2 - s]) * (f.pn[o2] + f.pn[o2 - s]);
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL;
    AK[k] = 0.5 * (Akv[o - s] + Akv[o]);
    Hzk[k] = 0.5 * (Hz[o - s] + Hz[o]);
    oHz[k] = 1.0 / Hzk[k];
    double xv = X[o] + DC0 * R[o];
    xv = xv * oHz[k];
    x[k] = xv;
  }
  spline_implicit(x, Hzk, oHz, AK, N, p.dt, CF, DC);
  // replace the vertical mean with the one from the barotropic sub-cycle (:469-605)
  {
    double cf0 = Hzk[1], dc0 = x[1] * Hzk[1];
    for (int k = 2; k <= N; ++k) { cf0 = cf0 + Hzk[k]; dc0 = dc0 + x[k] * Hzk[k]; }
    const double m = met[o2];
    const double cff1 = 1.0 / (cf0 * m);
    dc0 = (dc0 * m - Davg1[o2]) * cff1;
    for (int k = 1; k <= N; ++k) x[k] = x[k] - dc0;
  }
  // ---- coupling (:1002-1432) for this row, then for the wall rows owned by the edge threads
  //   u: rows 0 and Mm+1 carry u = gamma2*u(wall-adjacent row) (u3dbc) and get their own coupling pass
  //   v: row 1 (the wall itself, v = 0 from v3dbc) is handled by the j = 2 thread, row Mm+1 by the j = Mm thread
  auto couple = [&](int jj, const double* xx) {
    const int q2 = jj * P + i;
    double* dck = CF;                                  // CF/DC are free after the implicit solve
    double* hvk = DC;
    double dc0 = 0.0, cf0 = 0.0, fc0 = 0.0;
    const double cff = 0.5 * met[q2];
    for (int k = 1; k <= N; ++k) {
      const int o = q2 + k * p.PL;
      dck[k] = cff * (Hz[o] + Hz[o - s]);
      dc0 = dc0 + dck[k];
      cf0 = cf0 + dck[k] * xx[k];
    }
    dc0 = 1.0 / dc0;
    cf0 = dc0 * (cf0 - Davg1[q2]);
    const double b = dc0 * Davg1[q2];
    st_w(bar1, q2 - i, i, b, p);
    st_w(bar2, q2 - i, i, b, p);
    // boundary rows only: remove the mismatch of the vertical mean (:1132-1188, :1350-1406)
    const bool wall = DIR ? (jj == 1 || jj == Mm + 1) : (jj == 0 || jj == Mm + 1);
    for (int k = N; k >= 1; --k) {
      const int o = q2 + k * p.PL;
      const double xk = wall ? (xx[k] - cf0) : xx[k];
      st_w(X, o - i, i, xk, p);
      const double hv = 0.5 * (HUV[o] + xk * dck[k]);
      hvk[k] = hv;
      fc0 = fc0 + hv;
    }
    fc0 = dc0 * (fc0 - Davg2[q2]);
    for (int k = 1; k <= N; ++k) {
      const int o = q2 + k * p.PL;
      st_w(HUV, o - i, i, hvk[k] - dck[k] * fc0, p);
    }
  };
  couple(j, x);
  if (DIR == 0) {
    if (j == 1) { double xb[MAXN + 1]; for (int k = 1; k <= N; ++k) xb[k] = p.gamma2 * x[k]; couple(0, xb); }
    if (j == Mm) { double xb[MAXN + 1]; for (int k = 1; k <= N; ++k) xb[k] = p.gamma2 * x[k]; couple(Mm + 1, xb); }
  } else {
    if (j == 2) { double xb[MAXN + 1]; for (int k = 1; k <= N; ++k) xb[k] = 0.0; couple(1, xb); }
    if (j == Mm) { double xb[MAXN + 1]; for (int k = 1; k <= N; ++k) xb[k] = 0.0; couple(Mm + 1, xb); }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// step3d_t_tile (ROMS/Nonlinear/step3d_t.F:388-876 horizontal advection of t(:,:,:,3,:), :883-1210 vertical advection,
// :1366-1427 implicit diffusion, :1551-1621 t3dbc + periodic images).  One thread per column and tracer.
template <int HADV, int VADV>
__global__ void __launch_bounds__(128) k_step3d_t(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int itrc = blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P;
  const double* __restrict__ t3 = f.t[3][itrc];
  double* __restrict__ tn = f.t[p.nnew][itrc];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  const double* __restrict__ Akt = f.Akt[itrc];
  const double pm = f.pm[o2 + i], pn = f.pn[o2 + i];
  double x[MAXN + 1], hz[MAXN + 1], oHz[MAXN + 1], AK[MAXN + 1], CF[MAXN + 1], DC[MAXN + 1], tc[MAXN + 2];
  for (int k = 1; k <= N; ++k) tc[k] = t3[o2 + k * p.PL + i];
  tc[0] = tc[1]; tc[N + 1] = tc[N];
  AK[0] = Akt[o2 + i];
  const double cffh = p.dt * pm * pn;
  double FCm = 0.0;
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL;
    hz[k] = Hz[o + i];
    oHz[k] = 1.0 / hz[k];
    AK[k] = Akt[o + i];
    // horizontal: t(nnew) -= dt*pm*pn*(div)  (:861-873)
    double FXi, FXip, FEj, FEjp;
    hadv_fluxes<HADV>(t3, Huon, Hvom, o, i, j, p, FXi, FXip, FEj, FEjp);
    const double c1 = cffh * (FXip - FXi);
    const double c2 = cffh * (FEjp - FEj);
    const double c3 = c1 + c2;
    double tv = tn[o + i] - c3;
    // vertical (:1189-1207)
    const double FCk = (k < N) ? vflux<VADV>(tc, k, N, W[o + i]) : 0.0;
    const double cv = cffh * (FCk - FCm);
    tv = tv - cv;
    tv = tv * oHz[k];
    x[k] = tv;
    FCm = FCk;
  }
  spline_implicit(x, hz, oHz, AK, N, p.dt, CF, DC);
  for (int k = 1; k <= N; ++k) st_r_grad(tn, o2 + k * p.PL, i, j, x[k], p);
}

static inline dim3 g2(dim3 b, int ni, int nj, int nz = 1) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz); }

void launch_step3d_uv(const Par& p, const Flds& f, cudaStream_t s) {
  if (launch_step3d_uv_n(p, f, s)) return;          // compile-time-N fast path (k_step3d_n.cu)
  dim3 b(64, 2);
  k_step3d_uv<0><<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  k_step3d_uv<1><<<g2(b, p.Iend - p.Istr + 1, p.Mm - 1), b, 0, s>>>(p, f);
}

template <int H>
static void launch_s3t_v(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2); dim3 g = g2(b, p.Iend - p.Istr + 1, p.Mm, p.NT);
  if (p.vadv == 0) k_step3d_t<H, 0><<<g, b, 0, s>>>(p, f);
  else if (p.vadv == 1) k_step3d_t<H, 1><<<g, b, 0, s>>>(p, f);
  else k_step3d_t<H, 2><<<g, b, 0, s>>>(p, f);
}
void launch_step3d_t(const Par& p, const Flds& f, cudaStream_t s) {
  if (launch_step3d_t_n(p, f, s)) return;           // compile-time-N fast path (k_step3d_n.cu)
  if (p.hadv == 0) launch_s3t_v<0>(p, f, s);
  else if (p.hadv == 1) launch_s3t_v<1>(p, f, s);
  else if (p.hadv == 2) launch_s3t_v<2>(p, f, s);
  else launch_s3t_v<3>(p, f, s);
}

}  // namespace rb

// Compile-time-N (N = 30, the BENCHMARK grids) versions of the thread-per-column correctors step3d_uv / step3d_t.
// The generic kernels in k_step3d.cu keep six thread-private k-arrays in local memory; at BENCHMARK3 size that working
// set (~230 MB for the resident threads) overflows L2 and costs ~5x the algorithmic DRAM traffic.  Here the k loops are
// fully unrolled so that the solution x(k) and the Thomas coefficients CF(k), DC(k) live in registers, while the column
// geometry (Hzk, 1/Hzk, AK) sits in shared memory laid out [k][thread] (conflict free).  Same arithmetic, same order.
#include "dev.cuh"
#include "kernels.h"
#include "k_adv.cuh"

namespace rb {

constexpr int TN = 64;       // threads per block (one column each)

// parabolic-spline implicit solve (step3d_uv.F:344-396, step3d_t.F:1370-1427); geometry in shared memory
template <int NN>
__device__ __forceinline__ void spline_implicit_n(double (&x)[NN + 1], const double* sHzk, const double* sOHz, const double* sAK, double dt,
                                                  double (&CF)[NN + 1], double (&DC)[NN + 1]) {
  CF[0] = 0.0; DC[0] = 0.0;
#pragma unroll
  for (int k = 1; k <= NN - 1; ++k) {
    const double hk = sHzk[k * TN], hk1 = sHzk[(k + 1) * TN], ok = sOHz[k * TN], ok1 = sOHz[(k + 1) * TN];
    const double FCk = (1.0 / 6.0) * hk - dt * sAK[(k - 1) * TN] * ok;
    const double CFk = (1.0 / 6.0) * hk1 - dt * sAK[(k + 1) * TN] * ok1;
    const double BCk = (1.0 / 3.0) * (hk + hk1) + dt * sAK[k * TN] * (ok + ok1);
    const double cff = 1.0 / (BCk - FCk * CF[k - 1]);
    CF[k] = cff * CFk;
    DC[k] = cff * (x[k + 1] - x[k] - FCk * DC[k - 1]);
  }
  DC[NN] = 0.0;
#pragma unroll
  for (int k = NN - 1; k >= 1; --k) DC[k] = DC[k] - CF[k] * DC[k + 1];
#pragma unroll
  for (int k = 1; k <= NN; ++k) {
    DC[k] = DC[k] * sAK[k * TN];
    const double cff = dt * sOHz[k * TN] * (DC[k] - DC[k - 1]);
    x[k] = x[k] + cff;
  }
}

// step3d_uv_tile (ROMS/Nonlinear/step3d_uv.F:288-950, :956-965, :1002-1432, :1438-1461); see k_step3d.cu for the map
template <int DIR, int NN>
__global__ void __launch_bounds__(TN) k_step3d_uv_n(Par p, Flds f) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  double* sHzk = sm + tid; double* sOHz = sm + (NN + 1) * TN + tid; double* sAK = sm + 2 * (NN + 1) * TN + tid;
  const int i = p.Istr + blockIdx.x * TN + tid;
  const int j = (DIR ? 2 : 1) + blockIdx.y;
  if (i > p.Iend) return;
  const int P = p.P, Mm = p.Mm, PL = p.PL, o2 = j * P + i;
  const int s = DIR ? P : 1;
  double* __restrict__ X = DIR ? f.v[p.nnew] : f.u[p.nnew];
  const double* __restrict__ R = DIR ? f.rv[p.nrhs] : f.ru[p.nrhs];
  double* __restrict__ HUV = DIR ? f.Hvom : f.Huon;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Akv = f.Akv;
  const double* __restrict__ Davg1 = DIR ? f.DV_avg1 : f.DU_avg1;
  const double* __restrict__ Davg2 = DIR ? f.DV_avg2 : f.DU_avg2;
  const double* __restrict__ met = DIR ? f.om_v : f.on_u;
  double* __restrict__ bar1 = DIR ? f.vbar[1] : f.ubar[1];
  double* __restrict__ bar2 = DIR ? f.vbar[2] : f.ubar[2];
  double x[NN + 1], CF[NN + 1], DC[NN + 1];
  double cffAB;
  if (p.istart == 0) cffAB = 0.25 * p.dt;
  else if (p.istart == 1) cffAB = 0.25 * p.dt * 3.0 / 2.0;
  else cffAB = 0.25 * p.dt * 23.0 / 12.0;
  // Stage the whole column of every 3-D operand with cp.async (global -> shared, no register cost): ~210 independent 8-byte
  // copies per thread are in flight at once, which is what saturates HBM for this otherwise latency-bound solver.
  double* sX = sm + 3 * (NN + 1) * TN + tid; double* sR = sm + 4 * (NN + 1) * TN + tid;
  double* sKb = sm + 5 * (NN + 1) * TN + tid; double* sHUV = sm + 6 * (NN + 1) * TN + tid;
  cp_async8(sAK, Akv + o2); cp_async8(sKb, Akv + o2 - s);
#pragma unroll
  for (int k = 1; k <= NN; ++k) {
    const int o = o2 + k * PL;
    cp_async8(sAK + k * TN, Akv + o); cp_async8(sKb + k * TN, Akv + o - s);
    cp_async8(sHzk + k * TN, Hz + o); cp_async8(sOHz + k * TN, Hz + o - s);
    cp_async8(sX + k * TN, X + o); cp_async8(sR + k * TN, R + o); cp_async8(sHUV + k * TN, HUV + o);
  }
  cp_async_wait_all();
  sAK[0] = 0.5 * (sKb[0] + sAK[0]);
  const double DC0 = cffAB * (f.pm[o2] + f.pm[o2 - s]) * (f.pn[o2] + f.pn[o2 - s]);
#pragma unroll
  for (int k = 1; k <= NN; ++k) {
    sAK[k * TN] = 0.5 * (sKb[k * TN] + sAK[k * TN]);
    const double hk = 0.5 * (sOHz[k * TN] + sHzk[k * TN]);
    const double ok = 1.0 / hk;
    sHzk[k * TN] = hk; sOHz[k * TN] = ok;
    double xv = sX[k * TN] + DC0 * sR[k * TN];
    xv = xv * ok;
    x[k] = xv;
  }
  spline_implicit_n<NN>(x, sHzk, sOHz, sAK, p.dt, CF, DC);
  {
    double cf0 = sHzk[TN], dc0 = x[1] * sHzk[TN];
#pragma unroll
    for (int k = 2; k <= NN; ++k) { const double hk = sHzk[k * TN]; cf0 = cf0 + hk; dc0 = dc0 + x[k] * hk; }
    const double m = met[o2];
    const double cff1 = 1.0 / (cf0 * m);
    dc0 = (dc0 * m - Davg1[o2]) * cff1;
#pragma unroll
    for (int k = 1; k <= NN; ++k) x[k] = x[k] - dc0;
  }
  // coupling (:1002-1432); own row first, then the wall rows owned by the edge threads (see k_step3d.cu)
  auto couple = [&](int jj, double scale, bool own) {
    const int q2 = jj * P + i;
    double dc0 = 0.0, cf0 = 0.0, fc0 = 0.0;
    const double mq = met[q2];
    const double cff = 0.5 * mq;
#pragma unroll
    for (int k = 1; k <= NN; ++k) {
      const int o = q2 + k * PL;
      const double d = own ? (mq * sHzk[k * TN]) : (cff * (Hz[o] + Hz[o - s]));   // (0.5*m)*(a+b) == m*(0.5*(a+b)) bitwise
      CF[k] = d;
      dc0 = dc0 + d;
      cf0 = cf0 + d * ((scale == 0.0) ? 0.0 : scale * x[k]);
    }
    dc0 = 1.0 / dc0;
    cf0 = dc0 * (cf0 - Davg1[q2]);
    const double b = dc0 * Davg1[q2];
    st_w(bar1, q2 - i, i, b, p);
    st_w(bar2, q2 - i, i, b, p);
    const bool wall = DIR ? (jj == 1 || jj == Mm + 1) : (jj == 0 || jj == Mm + 1);
#pragma unroll
    for (int k = NN; k >= 1; --k) {
      const int o = q2 + k * PL;
      const double xs = (scale == 0.0) ? 0.0 : scale * x[k];
      const double xk = wall ? (xs - cf0) : xs;
      st_w(X, o - i, i, xk, p);
      const double hv = 0.5 * ((own ? sHUV[k * TN] : HUV[o]) + xk * CF[k]);
      DC[k] = hv;
      fc0 = fc0 + hv;
    }
    fc0 = dc0 * (fc0 - Davg2[q2]);
#pragma unroll
    for (int k = 1; k <= NN; ++k) st_w(HUV, q2 + k * PL - i, i, DC[k] - CF[k] * fc0, p);
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

// step3d_t_tile (ROMS/Nonlinear/step3d_t.F:388-876, :883-1210, :1366-1427, :1551-1621)
template <int HADV, int VADV, int NN>
__global__ void __launch_bounds__(TN) k_step3d_t_n(Par p, Flds f) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  double* sHz = sm + tid; double* sOHz = sm + (NN + 1) * TN + tid; double* sAK = sm + 2 * (NN + 1) * TN + tid;
  const int i = p.Istr + blockIdx.x * TN + tid;
  const int j = 1 + blockIdx.y;
  const int itrc = blockIdx.z;
  if (i > p.Iend) return;
  const int P = p.P, PL = p.PL, o2 = j * P;
  const double* __restrict__ t3 = f.t[3][itrc];
  double* __restrict__ tn = f.t[p.nnew][itrc];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  const double* __restrict__ Akt = f.Akt[itrc];
  const double pm = f.pm[o2 + i], pn = f.pn[o2 + i];
  double x[NN + 1], CF[NN + 1], DC[NN + 1];
  sAK[0] = Akt[o2 + i];
  const double cffh = p.dt * pm * pn;
  double FCm = 0.0;
  double tkm1, tk = t3[o2 + PL + i], tkp1 = t3[o2 + 2 * PL + i], tkp2 = t3[o2 + 3 * PL + i];
  tkm1 = tk;
#pragma unroll
  for (int k = 1; k <= NN; ++k) {
    const int o = o2 + k * PL;
    const double hk = Hz[o + i];
    const double ok = 1.0 / hk;
    sHz[k * TN] = hk; sOHz[k * TN] = ok; sAK[k * TN] = Akt[o + i];
    double FXi, FXip, FEj, FEjp;
    hadv_fluxes<HADV>(t3, Huon, Hvom, o, i, j, p, FXi, FXip, FEj, FEjp);
    const double c1 = cffh * (FXip - FXi);
    const double c2 = cffh * (FEjp - FEj);
    const double c3 = c1 + c2;
    double tv = tn[o + i] - c3;
    const double FCk = (k < NN) ? vflux4<VADV>(tkm1, tk, tkp1, tkp2, k, NN, W[o + i]) : 0.0;
    const double cv = cffh * (FCk - FCm);
    tv = tv - cv;
    tv = tv * ok;
    x[k] = tv;
    FCm = FCk;
    tkm1 = tk; tk = tkp1; tkp1 = tkp2;
    if (k + 3 <= NN) tkp2 = t3[o + 3 * PL + i];
  }
  spline_implicit_n<NN>(x, sHz, sOHz, sAK, p.dt, CF, DC);
#pragma unroll
  for (int k = 1; k <= NN; ++k) st_r_grad(tn, o2 + k * PL, i, j, x[k], p);
}

static size_t smem_n(int NN) { return (size_t)3 * (NN + 1) * TN * sizeof(double); }

bool launch_step3d_uv_n(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.N != 30) return false;
  static bool once = false;
  const size_t sm = (size_t)7 * 31 * TN * sizeof(double);      // Hzk, oHz, AK, X, R, Akv(i-1), Huon/Hvom columns
  if (!once) {
    cudaFuncSetAttribute(k_step3d_uv_n<0, 30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_step3d_uv_n<1, 30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    once = true;
  }
  const int nbx = (p.Iend - p.Istr + 1 + TN - 1) / TN;
  k_step3d_uv_n<0, 30><<<dim3(nbx, p.Mm), TN, sm, s>>>(p, f);
  k_step3d_uv_n<1, 30><<<dim3(nbx, p.Mm - 1), TN, sm, s>>>(p, f);
  return true;
}

bool launch_step3d_t_n(const Par& p, const Flds& f, cudaStream_t s) {
  if (p.N != 30 || p.hadv != 0 || p.vadv != 0) return false;
  static bool once = false;
  const size_t sm = smem_n(30);
  if (!once) { cudaFuncSetAttribute(k_step3d_t_n<0, 0, 30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); once = true; }
  const int nbx = (p.Iend - p.Istr + 1 + TN - 1) / TN;
  k_step3d_t_n<0, 0, 30><<<dim3(nbx, p.Mm, p.NT), TN, sm, s>>>(p, f);
  return true;
}

}  // namespace rb

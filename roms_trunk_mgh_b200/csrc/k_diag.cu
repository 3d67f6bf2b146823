// diag_tile (ROMS/Nonlinear/diag.F:207-437) and the SEAMOUNT ana_diag maxima (ROMS/Functionals/ana_diag.h:116-142).
// Three small kernels keep the reference's summation order (j-collapse per i, then the sum over i), so single-tile
// results are reproducible run to run and comparable with the CPU restatement.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

constexpr int NDV = 14;   // per-column values: ke, pe, vol, C, Cu, Cv, Cw, speed, rho, umax, vmax, ubarmax, vbarmax, (pad)

// stage 1: one thread per column (rows 0..Mm+1; sums only on rows 1..Mm)
__global__ void __launch_bounds__(128) k_diag_col(Par p, Flds f, double* __restrict__ S, int knew) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm + 1) return;
  const int N = p.N, P = p.P, o2 = j * P + i;
  const int ncol = P * (p.Mm + 2);
  double* __restrict__ s = S + j * P + (i - p.Istr);
  const double* __restrict__ un = f.u[p.nnew];
  const double* __restrict__ vn = f.v[p.nnew];
  double umax = 0.0, vmax = 0.0;
  for (int k = 1; k <= N; ++k) {
    umax = dmax(umax, un[o2 + k * p.PL]);
    if (j >= 1) vmax = dmax(vmax, vn[o2 + k * p.PL]);
  }
  s[9 * ncol] = umax; s[10 * ncol] = vmax;
  s[11 * ncol] = dmax(0.0, f.ubar[knew][o2]);
  s[12 * ncol] = (j >= 1) ? dmax(0.0, f.vbar[knew][o2]) : 0.0;
  if (j < 1 || j > p.Mm) {
    for (int q = 0; q < 9; ++q) s[q * ncol] = 0.0;
    s[8 * ncol] = -1.0e37;
    return;
  }
  const double* __restrict__ u = f.u[p.nstp];
  const double* __restrict__ v = f.v[p.nstp];
  const double zw0 = f.z_w[o2], zwN = f.z_w[o2 + N * p.PL];
  double ke = 0.0, pe = 0.5 * p.g * zwN * zwN;
  const double cff = p.g / p.rho0;
  double mC = 0.0, mCu = 0.0, mCv = 0.0, mCw = 0.0, mspeed = 0.0, mrho = -1.0e37;
  const double pm = f.pm[o2], pn = f.pn[o2];
  for (int k = N; k >= 1; --k) {
    const int o = o2 + k * p.PL;
    const double hz = f.Hz[o];
    const double u2v2 = u[o] * u[o] + u[o + 1] * u[o + 1] + v[o] * v[o] + v[o + P] * v[o + P];
    ke = ke + hz * 0.25 * u2v2;
    pe = pe + cff * hz * (f.rho[o] + 1000.0) * (f.z_r[o] - zw0);
    const double Cu = 0.5 * fabs(u[o] + u[o + 1]) * p.dt * pm;
    const double Cv = 0.5 * fabs(v[o] + v[o + P]) * p.dt * pn;
    const double Cw = 0.5 * fabs(f.wvel[o - p.PL] + f.wvel[o]) * p.dt / hz;
    const double C = Cu + Cv + Cw;
    if (C > mC) { mC = C; mCu = Cu; mCv = Cv; mCw = Cw; }
    mspeed = dmax(mspeed, sqrt(0.5 * u2v2));
    mrho = dmax(mrho, f.rho[o]);
  }
  const double omn = f.omn[o2];
  s[0] = omn * ke; s[ncol] = omn * pe; s[2 * ncol] = omn * (zwN - zw0);
  s[3 * ncol] = mC; s[4 * ncol] = mCu; s[5 * ncol] = mCv; s[6 * ncol] = mCw; s[7 * ncol] = mspeed; s[8 * ncol] = mrho;
}

// stage 2: one thread per i, collapse j in ascending order (diag.F:298-310).  The three sums keep the reference's order; their
// operands are fetched RB rows at a time so that the loads of a batch are in flight together (the additions are the only
// dependent chain).  The maxima are order-free apart from the first-maximum-wins rule of the Courant number.
constexpr int RB = 8;
__global__ void __launch_bounds__(64) k_diag_rows(Par p, const double* __restrict__ S, double* __restrict__ R) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > p.Iend) return;
  const int P = p.P, ncol = P * (p.Mm + 2);
  const double* __restrict__ s0 = S + (i - p.Istr);
  double ke = 0.0, pe = 0.0, vol = 0.0, mC = 0.0, mCu = 0.0, mCv = 0.0, mCw = 0.0, msp = 0.0, mrho = -1.0e37;
  double umax = 0.0, vmax = 0.0, ubm = 0.0, vbm = 0.0;
  for (int jb = 1; jb <= p.Mm; jb += RB) {
    double a[RB], b[RB], c[RB], d[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int j = min(jb + q, p.Mm);
      a[q] = s0[j * P]; b[q] = s0[j * P + ncol]; c[q] = s0[j * P + 2 * ncol]; d[q] = s0[j * P + 3 * ncol];
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      if (jb + q <= p.Mm) {
        ke = ke + a[q]; pe = pe + b[q]; vol = vol + c[q];
        if (d[q] > mC) { const double* s = s0 + (jb + q) * P; mC = d[q]; mCu = s[4 * ncol]; mCv = s[5 * ncol]; mCw = s[6 * ncol]; }
      }
    }
  }
  for (int jb = 0; jb <= p.Mm + 1; jb += RB) {
    double a[RB], b[RB], c[RB], d[RB], e[RB], g[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int j = min(jb + q, p.Mm + 1);
      a[q] = s0[j * P + 7 * ncol]; b[q] = s0[j * P + 8 * ncol]; c[q] = s0[j * P + 9 * ncol]; d[q] = s0[j * P + 10 * ncol];
      e[q] = s0[j * P + 11 * ncol]; g[q] = s0[j * P + 12 * ncol];
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int j = min(jb + q, p.Mm + 1);
      if (j >= 1 && j <= p.Mm) { msp = dmax(msp, a[q]); mrho = dmax(mrho, b[q]); }
      umax = dmax(umax, c[q]); vmax = dmax(vmax, d[q]); ubm = dmax(ubm, e[q]); vbm = dmax(vbm, g[q]);
    }
  }
  double* r = R + (size_t)(i - p.Istr) * NDV;
  r[0] = ke; r[1] = pe; r[2] = vol; r[3] = mC; r[4] = mCu; r[5] = mCv; r[6] = mCw; r[7] = msp; r[8] = mrho;
  r[9] = umax; r[10] = vmax; r[11] = ubm; r[12] = vbm;
}

// stage 3: sum over i in ascending order (diag.F:311-318).  out16 = avgke*vol, avgpe*vol, volume, maxC, Cu, Cv, Cw,
// maxspeed, maxrho, umax, vmax, ubarmax, vbarmax  (tile-local; the caller finishes the division / cross-tile reduce).
// One CTA: all threads stage a chunk of R in shared memory (coalesced); thread 0 adds the three sums in the reference's
// order, warp 1 reduces the maxima (order-free; the Courant number keeps the first maximum: larger value, then lower i).
constexpr int FCH = 256;
__global__ void __launch_bounds__(256) k_diag_final(Par p, const double* __restrict__ R, double* __restrict__ out) {
  __shared__ double sh[FCH * NDV];
  double ke = 0.0, pe = 0.0, vol = 0.0;
  double mC = 0.0, msp = 0.0, mrho = -1.0e37, umax = 0.0, vmax = 0.0, ubm = 0.0, vbm = 0.0;
  int iC = 0x7fffffff;
  const int ni = p.Iend - p.Istr + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q0 = 0; q0 < ni; q0 += FCH) {
    const int n = min(FCH, ni - q0);
    __syncthreads();
    for (int x = threadIdx.x; x < n * NDV; x += blockDim.x) sh[x] = R[(size_t)q0 * NDV + x];
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 0; q < n; ++q) { const double* r = sh + q * NDV; vol = vol + r[2]; pe = pe + r[1]; ke = ke + r[0]; }
    } else if (warp == 1) {
      for (int q = lane; q < n; q += 32) {
        const double* r = sh + q * NDV;
        if (r[3] > mC) { mC = r[3]; iC = q0 + q; }
        msp = dmax(msp, r[7]); mrho = dmax(mrho, r[8]);
        umax = dmax(umax, r[9]); vmax = dmax(vmax, r[10]); ubm = dmax(ubm, r[11]); vbm = dmax(vbm, r[12]);
      }
    }
  }
  if (warp == 1) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const double oC = __shfl_xor_sync(0xffffffffu, mC, d); const int oi = __shfl_xor_sync(0xffffffffu, iC, d);
      if (oC > mC || (oC == mC && oi < iC)) { mC = oC; iC = oi; }
      msp = dmax(msp, __shfl_xor_sync(0xffffffffu, msp, d)); mrho = dmax(mrho, __shfl_xor_sync(0xffffffffu, mrho, d));
      umax = dmax(umax, __shfl_xor_sync(0xffffffffu, umax, d)); vmax = dmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
      ubm = dmax(ubm, __shfl_xor_sync(0xffffffffu, ubm, d)); vbm = dmax(vbm, __shfl_xor_sync(0xffffffffu, vbm, d));
    }
    if (lane == 0) {
      double mCu = 0.0, mCv = 0.0, mCw = 0.0;
      if (mC > 0.0 && iC < ni) { const double* r = R + (size_t)iC * NDV; mCu = r[4]; mCv = r[5]; mCw = r[6]; }
      out[3] = mC; out[4] = mCu; out[5] = mCv; out[6] = mCw; out[7] = msp; out[8] = mrho;
      out[9] = umax; out[10] = vmax; out[11] = ubm; out[12] = vbm;
    }
  }
  if (threadIdx.x == 0) { out[0] = ke; out[1] = pe; out[2] = vol; }
}

int diag_partial_doubles(const Par& p) { return NDV * p.P * (p.Mm + 2) + NDV * (p.Iend - p.Istr + 1); }

void launch_diag(const Par& p, const Flds& f, double* partial, double* out16, int knew, cudaStream_t s) {
  double* S = partial;
  double* R = partial + (size_t)NDV * p.P * (p.Mm + 2);
  double* Sq = S;                 // S[q][j*P + (i-Istr)]
  dim3 b(64, 2);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + 2 + b.y - 1) / b.y);
  k_diag_col<<<g, b, 0, s>>>(p, f, Sq, knew);
  k_diag_rows<<<(p.Iend - p.Istr + 1 + 63) / 64, 64, 0, s>>>(p, Sq, R);
  k_diag_final<<<1, 256, 0, s>>>(p, R, out16);
}

}  // namespace rb

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

// stage 2: one thread per i, collapse j in ascending order (diag.F:298-310)
__global__ void k_diag_rows(Par p, const double* __restrict__ S, double* __restrict__ R) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > p.Iend) return;
  const int P = p.P, ncol = P * (p.Mm + 2);
  double ke = 0.0, pe = 0.0, vol = 0.0, mC = 0.0, mCu = 0.0, mCv = 0.0, mCw = 0.0, msp = 0.0, mrho = -1.0e37;
  double umax = 0.0, vmax = 0.0, ubm = 0.0, vbm = 0.0;
  for (int j = 0; j <= p.Mm + 1; ++j) {
    const double* s = S + j * P + (i - p.Istr);
    if (j >= 1 && j <= p.Mm) {
      ke = ke + s[0]; pe = pe + s[ncol]; vol = vol + s[2 * ncol];
      if (s[3 * ncol] > mC) { mC = s[3 * ncol]; mCu = s[4 * ncol]; mCv = s[5 * ncol]; mCw = s[6 * ncol]; }
      msp = dmax(msp, s[7 * ncol]); mrho = dmax(mrho, s[8 * ncol]);
    }
    umax = dmax(umax, s[9 * ncol]); vmax = dmax(vmax, s[10 * ncol]); ubm = dmax(ubm, s[11 * ncol]); vbm = dmax(vbm, s[12 * ncol]);
  }
  double* r = R + (size_t)(i - p.Istr) * NDV;
  r[0] = ke; r[1] = pe; r[2] = vol; r[3] = mC; r[4] = mCu; r[5] = mCv; r[6] = mCw; r[7] = msp; r[8] = mrho;
  r[9] = umax; r[10] = vmax; r[11] = ubm; r[12] = vbm;
}

// stage 3: sum over i in ascending order (diag.F:311-318).  out16 = avgke*vol, avgpe*vol, volume, maxC, Cu, Cv, Cw,
// maxspeed, maxrho, umax, vmax, ubarmax, vbarmax  (tile-local; the caller finishes the division / cross-tile reduce)
__global__ void k_diag_final(Par p, const double* __restrict__ R, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ke = 0.0, pe = 0.0, vol = 0.0, mC = 0.0, mCu = 0.0, mCv = 0.0, mCw = 0.0, msp = 0.0, mrho = -1.0e37;
  double umax = 0.0, vmax = 0.0, ubm = 0.0, vbm = 0.0;
  const int ni = p.Iend - p.Istr + 1;
  for (int q = 0; q < ni; ++q) {
    const double* r = R + (size_t)q * NDV;
    vol = vol + r[2]; pe = pe + r[1]; ke = ke + r[0];
    if (r[3] > mC) { mC = r[3]; mCu = r[4]; mCv = r[5]; mCw = r[6]; }
    msp = dmax(msp, r[7]); mrho = dmax(mrho, r[8]);
    umax = dmax(umax, r[9]); vmax = dmax(vmax, r[10]); ubm = dmax(ubm, r[11]); vbm = dmax(vbm, r[12]);
  }
  out[0] = ke; out[1] = pe; out[2] = vol; out[3] = mC; out[4] = mCu; out[5] = mCv; out[6] = mCw; out[7] = msp; out[8] = mrho;
  out[9] = umax; out[10] = vmax; out[11] = ubm; out[12] = vbm;
}

int diag_partial_doubles(const Par& p) { return NDV * p.P * (p.Mm + 2) + NDV * (p.Iend - p.Istr + 1); }

void launch_diag(const Par& p, const Flds& f, double* partial, double* out16, int knew, cudaStream_t s) {
  double* S = partial;
  double* R = partial + (size_t)NDV * p.P * (p.Mm + 2);
  double* Sq = S;                 // S[q][j*P + (i-Istr)]
  dim3 b(64, 2);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + 2 + b.y - 1) / b.y);
  k_diag_col<<<g, b, 0, s>>>(p, f, Sq, knew);
  k_diag_rows<<<(p.Iend - p.Istr + 1 + 127) / 128, 128, 0, s>>>(p, Sq, R);
  k_diag_final<<<1, 32, 0, s>>>(p, R, out16);
}

}  // namespace rb

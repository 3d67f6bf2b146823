// t3dmix2_geo_tile (ROMS/Nonlinear/t3dmix2_geo.h:219-419): harmonic tracer mixing rotated to geopotential surfaces.
// The reference keeps a two-level (k1,k2) rolling buffer of dTdz, dTdx, dTde, dZdx, dZde, FS in private 2-D scratch;
// here every thread owns a column and carries the same two levels in registers for its cell, its four faces and its
// four neighbours, marching k = 0..N once.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

#ifndef GEO_PF
#define GEO_PF 4          // L2 prefetch distance (levels)
#endif
#ifndef GEO_MINB
#define GEO_MINB 4          // 128 registers: the allocation the kernel had before the bound was made explicit; 5 and 6 are slower
#endif
__global__ void __launch_bounds__(128, GEO_MINB) k_t3dmix2_geo(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int itrc = blockIdx.z;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P + i;
  const double* __restrict__ tr = f.t[p.nrhs][itrc];
  double* __restrict__ tn = f.t[p.nnew][itrc];
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ d2 = f.diff2[itrc];
  // neighbour offsets: 0 = C(i,j), 1 = W(i-1,j), 2 = E(i+1,j), 3 = S(i,j-1), 4 = N(i,j+1)
  const int off[5] = {0, -1, +1, -P, +P};
  const double cxa = 0.5 * (f.pm[o2] + f.pm[o2 - 1]), cxb = 0.5 * (f.pm[o2 + 1] + f.pm[o2]);      // faces i, i+1
  const double cya = 0.5 * (f.pn[o2] + f.pn[o2 - P]), cyb = 0.5 * (f.pn[o2 + P] + f.pn[o2]);      // faces j, j+1
  const double kxa = 0.25 * (d2[o2] + d2[o2 - 1]) * f.on_u[o2], kxb = 0.25 * (d2[o2 + 1] + d2[o2]) * f.on_u[o2 + 1];
  const double kya = 0.25 * (d2[o2] + d2[o2 - P]) * f.om_v[o2], kyb = 0.25 * (d2[o2 + P] + d2[o2]) * f.om_v[o2 + P];
  const double ks = 0.5 * d2[o2];
  const double cdt = p.dt * f.pm[o2] * f.pn[o2];
  double tk[5], zk[5], tk1[5], zk1[5];
  double dTdz_p[5], dTdz_c[5];
  double dTdx_p[2] = {0, 0}, dZdx_p[2] = {0, 0}, dTde_p[2] = {0, 0}, dZde_p[2] = {0, 0};
  double dTdx_c[2] = {0, 0}, dZdx_c[2] = {0, 0}, dTde_c[2] = {0, 0}, dZde_c[2] = {0, 0};
  double FS_p = 0.0, FS_c = 0.0;
#pragma unroll
  for (int c = 0; c < 5; ++c) { tk[c] = zk[c] = tk1[c] = zk1[c] = 0.0; dTdz_p[c] = dTdz_c[c] = 0.0; }
  for (int k = 0; k <= N; ++k) {
    if (k < N) {
      const int o = o2 + (k + 1) * p.PL;
      // start the DRAM fetch of the rows a few levels up (the neighbouring rows j-1, j+1 are fetched by their own threads)
      pf_up<GEO_PF>(tr, o, k + 1, N, p.PL); pf_up<GEO_PF>(z_r, o, k + 1, N, p.PL); pf_up<GEO_PF>(Hz, o, k + 1, N, p.PL); pf_up<GEO_PF>(tn, o, k + 1, N, p.PL);
#pragma unroll
      for (int c = 0; c < 5; ++c) { tk1[c] = tr[o + off[c]]; zk1[c] = z_r[o + off[c]]; }
      dZdx_c[0] = cxa * (zk1[0] - zk1[1]); dTdx_c[0] = cxa * (tk1[0] - tk1[1]);
      dZdx_c[1] = cxb * (zk1[2] - zk1[0]); dTdx_c[1] = cxb * (tk1[2] - tk1[0]);
      dZde_c[0] = cya * (zk1[0] - zk1[3]); dTde_c[0] = cya * (tk1[0] - tk1[3]);
      dZde_c[1] = cyb * (zk1[4] - zk1[0]); dTde_c[1] = cyb * (tk1[4] - tk1[0]);
    }
    if (k == 0 || k == N) {
#pragma unroll
      for (int c = 0; c < 5; ++c) dTdz_c[c] = 0.0;
      FS_c = 0.0;
    } else {
#pragma unroll
      for (int c = 0; c < 5; ++c) { const double cff = 1.0 / (zk1[c] - zk[c]); dTdz_c[c] = cff * (tk1[c] - tk[c]); }
    }
    if (k > 0) {
      const int o = o2 + k * p.PL;
      const double hz0 = Hz[o];
      const double FXi = kxa * (hz0 + Hz[o - 1]) *
                         (dTdx_p[0] - 0.5 * (dmin(dZdx_p[0], 0.0) * (dTdz_p[1] + dTdz_c[0]) + dmax(dZdx_p[0], 0.0) * (dTdz_c[1] + dTdz_p[0])));
      const double FXip = kxb * (Hz[o + 1] + hz0) *
                          (dTdx_p[1] - 0.5 * (dmin(dZdx_p[1], 0.0) * (dTdz_p[0] + dTdz_c[2]) + dmax(dZdx_p[1], 0.0) * (dTdz_c[0] + dTdz_p[2])));
      const double FEj = kya * (hz0 + Hz[o - P]) *
                         (dTde_p[0] - 0.5 * (dmin(dZde_p[0], 0.0) * (dTdz_p[3] + dTdz_c[0]) + dmax(dZde_p[0], 0.0) * (dTdz_c[3] + dTdz_p[0])));
      const double FEjp = kyb * (Hz[o + P] + hz0) *
                          (dTde_p[1] - 0.5 * (dmin(dZde_p[1], 0.0) * (dTdz_p[0] + dTdz_c[4]) + dmax(dZde_p[1], 0.0) * (dTdz_c[0] + dTdz_p[4])));
      if (k < N) {
        const double dz = dTdz_c[0];
        double c1 = dmin(dZdx_p[0], 0.0), c2 = dmin(dZdx_c[1], 0.0), c3 = dmax(dZdx_c[0], 0.0), c4 = dmax(dZdx_p[1], 0.0);
        FS_c = ks * (c1 * (c1 * dz - dTdx_p[0]) + c2 * (c2 * dz - dTdx_c[1]) + c3 * (c3 * dz - dTdx_c[0]) + c4 * (c4 * dz - dTdx_p[1]));
        c1 = dmin(dZde_p[0], 0.0); c2 = dmin(dZde_c[1], 0.0); c3 = dmax(dZde_c[0], 0.0); c4 = dmax(dZde_p[1], 0.0);
        FS_c = FS_c + ks * (c1 * (c1 * dz - dTde_p[0]) + c2 * (c2 * dz - dTde_c[1]) + c3 * (c3 * dz - dTde_c[0]) + c4 * (c4 * dz - dTde_p[1]));
      }
      const double a1 = cdt * (FXip - FXi);
      const double a2 = cdt * (FEjp - FEj);
      const double a3 = p.dt * (FS_c - FS_p);
      const double a4 = a1 + a2 + a3;
      tn[o] = tn[o] + a4;
    }
    // roll k2 -> k1
#pragma unroll
    for (int c = 0; c < 5; ++c) { dTdz_p[c] = dTdz_c[c]; tk[c] = tk1[c]; zk[c] = zk1[c]; }
#pragma unroll
    for (int q = 0; q < 2; ++q) { dTdx_p[q] = dTdx_c[q]; dZdx_p[q] = dZdx_c[q]; dTde_p[q] = dTde_c[q]; dZde_p[q] = dZde_c[q]; }
    FS_p = FS_c;
  }
}

void launch_t3dmix2_geo(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 2);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + b.y - 1) / b.y, p.NT);
  k_t3dmix2_geo<<<g, b, 0, s>>>(p, f);
}

}  // namespace rb

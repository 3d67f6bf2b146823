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

// ---------------------------------------------------------------------------------------------------------------
// The same routine with the neighbour quantities shared through shared memory.  A CTA owns GT_X x GT_Y columns and marches k;
// per level every thread loads t(nrhs), z_r of level k+1 and Hz of level k for ITS column only and computes ITS dT/dz (one
// division), the first 2*(GT_X+GT_Y) threads do the same for one cell of the one-cell rim, and the four neighbours' values come
// from the tile.  Against the column kernel above: 1.3 instead of 5 divisions and 4-7 instead of 16 global loads per point and
// level.  Expressions and their order are unchanged (bit-identical results).
#ifndef GEO_TY
#define GEO_TY 4           // 32x4 columns per CTA, 4 CTAs per SM: 0.49 ms (32x8 with 2: 0.53, with 3: 1.21; 32x16: 0.60)
#endif
#ifndef GEO_TMINB
#define GEO_TMINB 4
#endif
constexpr int GT_X = 32, GT_Y = GEO_TY, GT_SW = GT_X + 2, GT_SH = GT_Y + 2, GT_NH = 2 * GT_X + 2 * GT_Y;
static_assert(GT_NH <= GT_X * GT_Y, "one rim cell per thread");
__global__ void __launch_bounds__(GT_X * GT_Y, GEO_TMINB) k_t3dmix2_geo_tiled(Par p, Flds f) {
  __shared__ double sT[GT_SH][GT_SW], sZ[GT_SH][GT_SW], sH[GT_SH][GT_SW], sD[GT_SH][GT_SW];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * GT_X + tx;
  const int i0 = p.Istr + blockIdx.x * GT_X, j0 = 1 + blockIdx.y * GT_Y;
  const int i = i0 + tx, j = j0 + ty;
  const int itrc = blockIdx.z;
  const bool own = (i <= p.Iend && j <= p.Mm);
  const int N = p.N, P = p.P, PL = p.PL;
  // columns beyond the tile's range take part in the barriers with clamped (valid) addresses and store nothing
  const int ic = (i <= p.Iend + 1) ? i : p.Iend + 1, jc = (j <= p.Mm + 1) ? j : p.Mm + 1;
  const int o2 = jc * P + ic;
  const double* __restrict__ tr = f.t[p.nrhs][itrc];
  double* __restrict__ tn = f.t[p.nnew][itrc];
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ d2 = f.diff2[itrc];
  // rim cell of this thread (tid < GT_NH): south row, north row, west column, east column of the (GT_Y+2) x (GT_X+2) tile
  int hx = 0, hy = 0;
  const bool rim = tid < GT_NH;
  if (tid < GT_X) { hx = tid + 1; hy = 0; }
  else if (tid < 2 * GT_X) { hx = tid - GT_X + 1; hy = GT_SH - 1; }
  else if (tid < 2 * GT_X + GT_Y) { hx = 0; hy = tid - 2 * GT_X + 1; }
  else if (rim) { hx = GT_SW - 1; hy = tid - 2 * GT_X - GT_Y + 1; }
  int hi = i0 + hx - 1, hj = j0 + hy - 1;
  hi = (hi <= p.Iend + 1) ? hi : p.Iend + 1; hj = (hj <= p.Mm + 1) ? hj : p.Mm + 1;
  const int oh = hj * P + hi;
  double cxa = 0, cxb = 0, cya = 0, cyb = 0, kxa = 0, kxb = 0, kya = 0, kyb = 0, ks = 0, cdt = 0;
  if (own) {
    cxa = 0.5 * (f.pm[o2] + f.pm[o2 - 1]); cxb = 0.5 * (f.pm[o2 + 1] + f.pm[o2]);
    cya = 0.5 * (f.pn[o2] + f.pn[o2 - P]); cyb = 0.5 * (f.pn[o2 + P] + f.pn[o2]);
    kxa = 0.25 * (d2[o2] + d2[o2 - 1]) * f.on_u[o2]; kxb = 0.25 * (d2[o2 + 1] + d2[o2]) * f.on_u[o2 + 1];
    kya = 0.25 * (d2[o2] + d2[o2 - P]) * f.om_v[o2]; kyb = 0.25 * (d2[o2 + P] + d2[o2]) * f.om_v[o2 + P];
    ks = 0.5 * d2[o2];
    cdt = p.dt * f.pm[o2] * f.pn[o2];
  }
  double t0 = 0.0, z0 = 0.0, ht0 = 0.0, hz0r = 0.0;                 // level k of the own column / of the rim cell
  double dpC = 0.0, dpW = 0.0, dpE = 0.0, dpS = 0.0, dpN = 0.0;      // dT/dz at W level k-1
  double dTdx_p0 = 0, dTdx_p1 = 0, dZdx_p0 = 0, dZdx_p1 = 0, dTde_p0 = 0, dTde_p1 = 0, dZde_p0 = 0, dZde_p1 = 0;
  double dTdx_c0 = 0, dTdx_c1 = 0, dZdx_c0 = 0, dZdx_c1 = 0, dTde_c0 = 0, dTde_c1 = 0, dZde_c0 = 0, dZde_c1 = 0;
  double FS_p = 0.0, FS_c = 0.0;
  // software pipeline: t / z_r of level k+2 and Hz / t(nnew) of level k+1 are requested before level k is staged
  double nt = tr[o2 + PL], nz = z_r[o2 + PL], nH = 0.0, ntn = 0.0;
  double hnt = 0.0, hnz = 0.0, hnH = 0.0;
  if (rim) { hnt = tr[oh + PL]; hnz = z_r[oh + PL]; }
  for (int k = 0; k <= N; ++k) {
    // ---- stage the own column and the rim cell: t, z_r of level k+1, Hz of level k, dT/dz at W level k
    const double t1 = nt, z1 = nz, hcur = nH, ctn = ntn;
    const double ht1 = hnt, hz1 = hnz, hhcur = hnH;
    if (k + 2 <= N) {
      const int o = o2 + (k + 2) * PL;
      pf_up<GEO_PF>(tr, o, k + 2, N, PL); pf_up<GEO_PF>(z_r, o, k + 2, N, PL); pf_up<GEO_PF>(Hz, o - PL, k + 1, N, PL); pf_up<GEO_PF>(tn, o - PL, k + 1, N, PL);
      nt = tr[o]; nz = z_r[o];
      if (rim) { hnt = tr[oh + (k + 2) * PL]; hnz = z_r[oh + (k + 2) * PL]; }
    }
    if (k + 1 <= N) {
      nH = Hz[o2 + (k + 1) * PL]; ntn = tn[o2 + (k + 1) * PL];
      if (rim) hnH = Hz[oh + (k + 1) * PL];
    }
    double dcC = 0.0;
    if (k < N) { sT[ty + 1][tx + 1] = t1; sZ[ty + 1][tx + 1] = z1; }
    if (k > 0) sH[ty + 1][tx + 1] = hcur;
    if (k > 0 && k < N) { const double cff = 1.0 / (z1 - z0); dcC = cff * (t1 - t0); }
    sD[ty + 1][tx + 1] = dcC;
    if (rim) {
      double hd = 0.0;
      if (k < N) { sT[hy][hx] = ht1; sZ[hy][hx] = hz1; }
      if (k > 0) sH[hy][hx] = hhcur;
      if (k > 0 && k < N) { const double cff = 1.0 / (hz1 - hz0r); hd = cff * (ht1 - ht0); }
      sD[hy][hx] = hd;
      ht0 = ht1; hz0r = hz1;
    }
    __syncthreads();
    const double dcW = sD[ty + 1][tx], dcE = sD[ty + 1][tx + 2], dcS = sD[ty][tx + 1], dcN = sD[ty + 2][tx + 1];
    if (k < N) {
      const double zW = sZ[ty + 1][tx], zE = sZ[ty + 1][tx + 2], zS = sZ[ty][tx + 1], zN = sZ[ty + 2][tx + 1];
      const double tW = sT[ty + 1][tx], tE = sT[ty + 1][tx + 2], tS = sT[ty][tx + 1], tN = sT[ty + 2][tx + 1];
      dZdx_c0 = cxa * (z1 - zW); dTdx_c0 = cxa * (t1 - tW);
      dZdx_c1 = cxb * (zE - z1); dTdx_c1 = cxb * (tE - t1);
      dZde_c0 = cya * (z1 - zS); dTde_c0 = cya * (t1 - tS);
      dZde_c1 = cyb * (zN - z1); dTde_c1 = cyb * (tN - t1);
    }
    if (k == 0 || k == N) FS_c = 0.0;
    if (k > 0) {
      const double hz0 = sH[ty + 1][tx + 1];
      const double FXi = kxa * (hz0 + sH[ty + 1][tx]) *
                         (dTdx_p0 - 0.5 * (dmin(dZdx_p0, 0.0) * (dpW + dcC) + dmax(dZdx_p0, 0.0) * (dcW + dpC)));
      const double FXip = kxb * (sH[ty + 1][tx + 2] + hz0) *
                          (dTdx_p1 - 0.5 * (dmin(dZdx_p1, 0.0) * (dpC + dcE) + dmax(dZdx_p1, 0.0) * (dcC + dpE)));
      const double FEj = kya * (hz0 + sH[ty][tx + 1]) *
                         (dTde_p0 - 0.5 * (dmin(dZde_p0, 0.0) * (dpS + dcC) + dmax(dZde_p0, 0.0) * (dcS + dpC)));
      const double FEjp = kyb * (sH[ty + 2][tx + 1] + hz0) *
                          (dTde_p1 - 0.5 * (dmin(dZde_p1, 0.0) * (dpC + dcN) + dmax(dZde_p1, 0.0) * (dcC + dpN)));
      if (k < N) {
        const double dz = dcC;
        double c1 = dmin(dZdx_p0, 0.0), c2 = dmin(dZdx_c1, 0.0), c3 = dmax(dZdx_c0, 0.0), c4 = dmax(dZdx_p1, 0.0);
        FS_c = ks * (c1 * (c1 * dz - dTdx_p0) + c2 * (c2 * dz - dTdx_c1) + c3 * (c3 * dz - dTdx_c0) + c4 * (c4 * dz - dTdx_p1));
        c1 = dmin(dZde_p0, 0.0); c2 = dmin(dZde_c1, 0.0); c3 = dmax(dZde_c0, 0.0); c4 = dmax(dZde_p1, 0.0);
        FS_c = FS_c + ks * (c1 * (c1 * dz - dTde_p0) + c2 * (c2 * dz - dTde_c1) + c3 * (c3 * dz - dTde_c0) + c4 * (c4 * dz - dTde_p1));
      }
      if (own) {
        const int o = o2 + k * PL;
        const double a1 = cdt * (FXip - FXi);
        const double a2 = cdt * (FEjp - FEj);
        const double a3 = p.dt * (FS_c - FS_p);
        const double a4 = a1 + a2 + a3;
        tn[o] = ctn + a4;
      }
    }
    __syncthreads();                                            // the tile is overwritten by the next level
    dpC = dcC; dpW = dcW; dpE = dcE; dpS = dcS; dpN = dcN;
    dTdx_p0 = dTdx_c0; dTdx_p1 = dTdx_c1; dZdx_p0 = dZdx_c0; dZdx_p1 = dZdx_c1;
    dTde_p0 = dTde_c0; dTde_p1 = dTde_c1; dZde_p0 = dZde_c0; dZde_p1 = dZde_c1;
    FS_p = FS_c; t0 = t1; z0 = z1;
  }
}

#ifndef GEO_TILED
#define GEO_TILED 1
#endif
void launch_t3dmix2_geo(const Par& p, const Flds& f, cudaStream_t s) {
#if GEO_TILED
  dim3 b(GT_X, GT_Y);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + b.y - 1) / b.y, p.NT);
  k_t3dmix2_geo_tiled<<<g, b, 0, s>>>(p, f);
#else
  dim3 b(64, 2);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + b.y - 1) / b.y, p.NT);
  k_t3dmix2_geo<<<g, b, 0, s>>>(p, f);
#endif
}

}  // namespace rb

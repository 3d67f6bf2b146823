// REJECTED VARIANT (kept as a tuning record, not built): prsgrd32 as one kernel with P in shared memory.
// Measured on B200, BENCHMARK3: 0.50 ms (6 levels per barrier, 128 registers) and 0.40 ms (1 level per barrier) against 0.30 ms
// for the two-kernel form (k_prsgrd32_P + k_prsgrd32_R, k_pre.cu) although it moves 630 MB instead of 1123 MB: the routine is
// bound by FP64 division latency at the occupancy the fused column kernel allows, not by HBM (profiles/README.md).
// ---------------------------------------------------------------------------------------------------------------
// prsgrd32_tile (ROMS/Nonlinear/prsgrd32.h:236-418) in ONE pass.  The reference integrates the pressure P downward into a
// private 3-D scratch array (:236-290) and then forms ru, rv from P(i-1), P(i) / P(j-1), P(j) and the horizontal
// harmonic-mean slopes (:296-418).  Here a CTA of PG_TX x PG_TY columns marches from k = N to 1: every thread integrates
// the P of its own column in registers (the same recurrence, same operation order), publishes the level's value in shared
// memory, and the threads that own an output point pick up their western / southern neighbour's P there -- the 3-D scratch
// never exists, which removes 3 of the routine's 8 array passes through HBM.  Column (tx = 0) and row (ty = 0) of the CTA only
// supply P to their neighbours; a CTA outputs (PG_TX-1) x (PG_TY-1) points.
constexpr int PG_TX = 32, PG_TY = 8, PG_CH = 6;
__device__ __forceinline__ void harm_pair(double a0, double a1, double f0, double f1, double& dZx, double& dRx) {
  const double eps = 1.0e-10;
  const double cff = 2.0 * a0 * a1;
  if (cff > eps) { const double c1 = 1.0 / (a0 + a1); dZx = cff * c1; } else dZx = 0.0;
  const double cff1 = 2.0 * f0 * f1;
  if (cff1 > eps) { const double c2 = 1.0 / (f0 + f1); dRx = cff1 * c2; } else dRx = 0.0;
}

__global__ void __launch_bounds__(PG_TX * PG_TY, 2) k_prsgrd32(Par p, Flds f) {
  // P of PG_CH consecutive levels of every column of the CTA, double-buffered: one barrier per chunk of levels, so the
  // downward recurrence runs PG_CH levels ahead of the (mutually independent, division-heavy) ru / rv evaluations that consume it
  __shared__ double sP[2][PG_CH][PG_TY][PG_TX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = p.Istr - 1 + blockIdx.x * (PG_TX - 1) + tx;          // IstrU-1 .. Iend
  const int j = blockIdx.y * (PG_TY - 1) + ty;                        // JstrV-1 = 1 .. Jend (row 0 of the first CTA row idles)
  const int N = p.N, P = p.P, PL = p.PL;
  const bool pv = i <= p.Iend && j >= 1 && j <= p.Mm;                 // this thread integrates a P column
  const bool ou = pv && tx >= 1 && ty >= 1;                           // ... and owns ru(i,j,:)
  const bool ov = ou && j >= p.JstrV;                                 // ... and rv(i,j,:)
  const int o2 = pv ? j * P + i : p.P + p.Istr;                       // safe dummy column for idle threads
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  const double GRho = p.g / p.rho0, HalfGRho = 0.5 * GRho;
  const double* __restrict__ rho = f.rho;
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ Hz = f.Hz;
  double* __restrict__ ru = f.ru[p.nrhs];
  double* __restrict__ rv = f.rv[p.nrhs];
  const double onu = ou ? f.on_u[o2] : 0.0, omv = ov ? f.om_v[o2] : 0.0;

  // ru, rv at level k from this column's (r0, z0, P0) and the neighbours' P of the same level (:296-418)
  auto rhs_level = [&](int k, double r0, double z0, double P0, double PW, double PS) {
    const int o = o2 + k * PL;
    const double hz0 = Hz[o];
    {
      const double rm2 = rho[o - 2], rm1 = rho[o - 1], rp1 = rho[o + 1];
      const double zm2 = z_r[o - 2], zm1 = z_r[o - 1], zp1 = z_r[o + 1];
      const double am1 = zm1 - zm2, a0 = z0 - zm1, ap1 = zp1 - z0;
      const double fm1 = rm1 - rm2, f0 = r0 - rm1, fp1 = rp1 - r0;
      double dZx_m, dRx_m, dZx_0, dRx_0;
      harm_pair(am1, a0, fm1, f0, dZx_m, dRx_m);        // (i-1)
      harm_pair(a0, ap1, f0, fp1, dZx_0, dRx_0);        // (i)
      const double x = onu * 0.5 * (hz0 + Hz[o - 1]) *
                       (PW - P0 -
                        HalfGRho * ((r0 + rm1) * (z0 - zm1) -
                                    OneFifth * ((dRx_0 - dRx_m) * (z0 - zm1 - OneTwelfth * (dZx_0 + dZx_m)) -
                                                (dZx_0 - dZx_m) * (r0 - rm1 - OneTwelfth * (dRx_0 + dRx_m)))));
      ru[o] = x;
    }
    if (ov) {
      const double rm2 = rho[o - 2 * P], rm1 = rho[o - P], rp1 = rho[o + P];
      const double zm2 = z_r[o - 2 * P], zm1 = z_r[o - P], zp1 = z_r[o + P];
      const double am1 = zm1 - zm2, a0 = z0 - zm1, ap1 = zp1 - z0;
      const double fm1 = rm1 - rm2, f0 = r0 - rm1, fp1 = rp1 - r0;
      double dZx_m, dRx_m, dZx_0, dRx_0;
      harm_pair(am1, a0, fm1, f0, dZx_m, dRx_m);        // (j-1)
      harm_pair(a0, ap1, f0, fp1, dZx_0, dRx_0);        // (j)
      const double x = omv * 0.5 * (hz0 + Hz[o - P]) *
                       (PS - P0 -
                        HalfGRho * ((r0 + rm1) * (z0 - zm1) -
                                    OneFifth * ((dRx_0 - dRx_m) * (z0 - zm1 - OneTwelfth * (dZx_0 + dZx_m)) -
                                                (dZx_0 - dZx_m) * (r0 - rm1 - OneTwelfth * (dRx_0 + dRx_m)))));
      rv[o] = x;
    }
  };

  // rolling window of the downward recurrence (:236-290) when level k is integrated: (r_kp, z_kp) = level k+1,
  // (r_k, z_k) = level k, (r_km, z_km) = level k-1.  Level N starts it with (r_k, r_km) = levels (N, N-1).
  const int oN = o2 + N * PL;
  double r_kp = 0.0, z_kp = 0.0;
  double r_k = rho[oN], z_k = z_r[oN];
  double r_km = rho[oN - PL], z_km = z_r[oN - PL];
  const double zwN = f.z_w[oN];
  double dR_kp = 0.0, dZ_kp = 0.0, Pk = 0.0;
  int buf = 0;
  for (int kt = N; kt >= 1; kt -= PG_CH) {
    double lr[PG_CH], lz[PG_CH], r0[PG_CH], z0[PG_CH];
    // the levels that enter the window during this chunk, requested together (level k-2 after level k has been integrated)
#pragma unroll
    for (int q = 0; q < PG_CH; ++q) {
      const int km = (kt - q - 2 >= 1) ? kt - q - 2 : 1;
      lr[q] = rho[o2 + km * PL]; lz[q] = z_r[o2 + km * PL];
      pf_dn<PG_CH>(rho, o2 + km * PL, km, PL); pf_dn<PG_CH>(z_r, o2 + km * PL, km, PL);
    }
#pragma unroll
    for (int q = 0; q < PG_CH; ++q) {
      const int k = kt - q;
      if (k >= 1) {
        if (k == N) {
          // level N (:236-265): raw(N) = raw(N-1), surface pressure
          const double rawR_k = r_k - r_km, rawZ_k = z_k - z_km;
          const double rawR_km = rawR_k, rawZ_km = rawZ_k;
          const double c = 2.0 * rawR_k * rawR_km;
          dR_kp = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;             // dR(N)
          dZ_kp = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);          // dZ(N)
          const double cff1 = 1.0 / (z_k - z_km);
          const double cff2 = 0.5 * (r_k - r_km) * (zwN - z_k) * cff1;
          Pk = p.g * zwN + GRho * (r_k + cff2) * (zwN - z_k);
        } else {
          // level k < N (:266-290)
          const double rawR_k = r_kp - r_k, rawZ_k = z_kp - z_k;
          double rawR_km, rawZ_km;
          if (k > 1) { rawR_km = r_k - r_km; rawZ_km = z_k - z_km; } else { rawR_km = rawR_k; rawZ_km = rawZ_k; }   // raw(0) = raw(1)
          const double c = 2.0 * rawR_k * rawR_km;
          const double dR_k = (c > eps) ? c / (rawR_k + rawR_km) : 0.0;
          const double dZ_k = 2.0 * rawZ_k * rawZ_km / (rawZ_k + rawZ_km);
          Pk = Pk + HalfGRho * ((r_kp + r_k) * (z_kp - z_k) -
                                OneFifth * ((dR_kp - dR_k) * (z_kp - z_k - OneTwelfth * (dZ_kp + dZ_k)) -
                                            (dZ_kp - dZ_k) * (r_kp - r_k - OneTwelfth * (dR_kp + dR_k))));
          dR_kp = dR_k; dZ_kp = dZ_k;
        }
        sP[buf][q][ty][tx] = Pk;
        r0[q] = r_k; z0[q] = z_k;
        r_kp = r_k; z_kp = z_k; r_k = r_km; z_k = z_km; r_km = lr[q]; z_km = lz[q];
      }
    }
    __syncthreads();                    // one barrier per chunk: the other buffer may still be read by slower warps
    if (ou) {
#pragma unroll
      for (int q = 0; q < PG_CH; ++q) {
        const int k = kt - q;
        if (k >= 1) rhs_level(k, r0[q], z0[q], sP[buf][q][ty][tx], sP[buf][q][ty][tx - 1], sP[buf][q][ty - 1][tx]);
      }
    }
    buf ^= 1;
  }
}


// REJECTED VARIANT (kept as a tuning record, not built): set_massflux + omega + wvelocity in one column pass.
// Measured on B200, BENCHMARK3: 0.59 ms (wvelocity column in shared memory, 166 registers) and 0.80 ms (both columns in
// registers, 254) against 0.41 ms for the three separate kernels, although it moves 9 instead of 16 whole arrays: the
// (i,j,k)-parallel k_set_massflux runs at 5.2 TB/s and the fused column kernel loses that parallelism (profiles/README.md).
// ---------------------------------------------------------------------------------------------------------------
// Whole-step path only: set_massflux_tile (set_massflux.F:140-174), omega_tile (omega.F:147-218) and wvelocity_tile
// (wvelocity.F:156-256) in one pass over the columns.  Separately the three routines move 16 whole 3-D arrays (Huon and Hvom
// are written and read back, u, v, z_r, z_w, W are read twice); fused, 5 are read and 4 written.  omega needs Huon(i+1,j) and
// Hvom(i,j+1): they are recomputed here from the neighbour's operands with the very expression the neighbour's thread stores,
// so every stored value is bit-identical to the three separate kernels (which roms_b200_run_phase keeps using).  Nothing
// between set_massflux and omega in main3d.F:307-474 reads Huon / Hvom (rho_eos, diag, set_vbc, ana_vmix), so running the
// fused kernel at omega's position leaves every other routine's inputs as they were.
template <int NC, bool WVEL>
__global__ void __launch_bounds__(128) k_flux_omega_wvel(Par p, Flds f) {
  const int i = xcol0(p, blockIdx.x * blockDim.x) + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..Mm+1 (mass fluxes); omega / wvelocity on 1..Mm
  if (i > p.Iend || j > p.Mm + 1) return;
  const int P = p.P, PL = p.PL, o2 = j * P, N = NC > 0 ? NC : p.N;
  const bool inner = j >= 1 && j <= p.Mm;
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ u = f.u[p.nrhs];
  const double* __restrict__ v = f.v[p.nrhs];
  const double* __restrict__ z_r = f.z_r;
  const double* __restrict__ z_w = f.z_w;
  const double onu0 = f.on_u[o2 + i], omv0 = f.om_v[o2 + i];
  double onu1 = 0.0, omv1 = 0.0, pmU0 = 0.0, pmU1 = 0.0, pnV0 = 0.0, pnV1 = 0.0, pmn = 0.0;
  if (inner) {
    onu1 = f.on_u[o2 + i + 1]; omv1 = f.om_v[o2 + P + i];
    if (WVEL) {
      const double pmi = f.pm[o2 + i], pni = f.pn[o2 + i];
      pmU0 = f.pm[o2 + i - 1] + pmi; pmU1 = pmi + f.pm[o2 + i + 1];
      pnV0 = f.pn[o2 - P + i] + pni; pnV1 = pni + f.pn[o2 + P + i];
      pmn = pmi * pni;
    }
  }
  // omega's partial sums stay in registers (compile-time N); wvelocity's column of horizontal terms goes to shared memory,
  // [level][thread], so that the kernel keeps the occupancy of the separate k_omega / k_wvelocity (both columns in registers: 254)
  double Wl[(NC > 0 ? NC : MAXN) + 1];
  __shared__ double sv[(WVEL && NC > 0) ? NC + 1 : 1][(WVEL && NC > 0) ? 128 : 1];
  double vloc[(WVEL && NC == 0) ? MAXN + 1 : 1];
  const int tl = threadIdx.y * blockDim.x + threadIdx.x;
#define vert(k) (*((NC > 0) ? &sv[(WVEL && NC > 0) ? (k) : 0][(WVEL && NC > 0) ? tl : 0] : &vloc[(WVEL && NC == 0) ? (k) : 0]))
  double w = 0.0;
  Wl[0] = 0.0;
#pragma unroll
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * PL;
    pf_up<GLUE_PF>(Hz, o + i, k, N, PL); pf_up<GLUE_PF>(u, o + i, k, N, PL); pf_up<GLUE_PF>(v, o + i, k, N, PL);
    const double hz = Hz[o + i];
    const double ui = u[o + i];
    const double hu0 = 0.5 * (hz + Hz[o + i - 1]) * ui * onu0;                 // set_massflux.F:146-157
    st_w(f.Huon, o, i, hu0, p);
    double vi = 0.0, hv0 = 0.0;
    if (j >= 1) {
      vi = v[o + i];
      hv0 = 0.5 * (hz + Hz[o - P + i]) * vi * omv0;                            // :158-169
      st_w(f.Hvom, o, i, hv0, p);
    }
    if (inner) {
      const double uE = u[o + i + 1], vN = v[o + P + i];
      const double hu1 = 0.5 * (Hz[o + i + 1] + hz) * uE * onu1;               // Huon(i+1,j,k)
      const double hv1 = 0.5 * (Hz[o + P + i] + hz) * vN * omv1;               // Hvom(i,j+1,k)
      w = w - (hu1 - hu0 + hv1 - hv0);                                         // omega.F:147-160
      Wl[k] = w;
      if (WVEL) {                                                              // wvelocity.F:156-190
        const double zr = z_r[o + i];
        const double wu0 = ui * (zr - z_r[o + i - 1]) * pmU0;
        const double wu1 = uE * (z_r[o + i + 1] - zr) * pmU1;
        double vt = 0.25 * (wu0 + wu1);
        const double wv0 = vi * (zr - z_r[o - P + i]) * pnV0;
        const double wv1 = vN * (z_r[o + P + i] - zr) * pnV1;
        vt = vt + 0.25 * (wv0 + wv1);
        vert(k) = vt;
      }
    }
  }
  if (!inner) return;
  // ---- omega: remove the part proportional to the free-surface tendency (omega.F:192-210), bc_w3d
  const double zw0 = z_w[o2 + i], zwN = z_w[o2 + N * PL + i];
  {
    const double wrk = w / (zwN - zw0);
    st_r_grad(f.W, o2, i, j, 0.0, p);
#pragma unroll
    for (int k = N - 1; k >= 1; --k) {
      const int o = o2 + k * PL;
      const double x = Wl[k] - wrk * (z_w[o + i] - zw0);
      Wl[k] = x;
      st_r_grad(f.W, o, i, j, x, p);
    }
    st_r_grad(f.W, o2 + N * PL, i, j, 0.0, p);
  }
  if (!WVEL) return;
  // ---- wvelocity (wvelocity.F:191-256)
  const double cff1 = 3.0 / 8.0, cff2 = 3.0 / 4.0, cff3 = 1.0 / 8.0, cff4 = 9.0 / 16.0, cff5 = 1.0 / 16.0;
  const double wrk = (f.DU_avg1[o2 + i] - f.DU_avg1[o2 + i + 1] + f.DV_avg1[o2 + i] - f.DV_avg1[o2 + P + i]) / (zwN - zw0);
  {
    const double slope = (z_r[o2 + PL + i] - zw0) / (z_r[o2 + 2 * PL + i] - z_r[o2 + PL + i]);
    const double w0 = cff1 * (vert(1) - slope * (vert(2) - vert(1))) + cff2 * vert(1) - cff3 * vert(2);
    st_r_grad(f.wvel, o2, i, j, w0, p);
    const int o = o2 + PL;
    const double w1 = pmn * (Wl[1] + wrk * (z_w[o + i] - zw0)) + cff1 * vert(1) + cff2 * vert(2) - cff3 * vert(3);
    st_r_grad(f.wvel, o, i, j, w1, p);
  }
#pragma unroll
  for (int k = 2; k <= N - 2; ++k) {
    const int o = o2 + k * PL;
    const double x = pmn * (Wl[k] + wrk * (z_w[o + i] - zw0)) + cff4 * (vert(k) + vert(k + 1)) - cff5 * (vert(k - 1) + vert(k + 2));
    st_r_grad(f.wvel, o, i, j, x, p);
  }
  {
    const int oN = o2 + N * PL, oM = oN - PL;
    const double slope = (zwN - z_r[oN + i]) / (z_r[oN + i] - z_r[oM + i]);
    const double wN = pmn * wrk * (zwN - zw0) + cff1 * (vert(N) + slope * (vert(N) - vert(N - 1))) + cff2 * vert(N) - cff3 * vert(N - 1);
    st_r_grad(f.wvel, oN, i, j, wN, p);
    const double wM = pmn * (Wl[N - 1] + wrk * (z_w[oM + i] - zw0)) + cff1 * vert(N) + cff2 * vert(N - 1) - cff3 * vert(N - 2);
    st_r_grad(f.wvel, oM, i, j, wM, p);
  }
}
#undef vert


// step2d_tile (ROMS/Nonlinear/step2d_LF_AM3.h:137-2528): one barotropic LF-AM3 predictor or corrector sub-step as a
// single kernel.  Everything a point needs derives from the read-only old time levels (krhs, kstp), so no grid-wide
// synchronisation is required inside a sub-step.  Each CTA owns a TX x TY tile of rho points and stages in shared
// memory, with a 3/2-point halo, what the reference keeps in private 2-D scratch arrays, so that every flux is
// evaluated once per face instead of once per consumer:
//   stage 0  Drhs, ubar(krhs), vbar(krhs)                                   (:548-552)
//   stage 1  DUon, DVom                                                     (:553-574)
//   stage 2  zeta_new -> Dnew, zwrk, gzeta, gzeta2, gzetaSA                 (:770-851); zeta(knew), rzeta(krhs) (:860-929)
//            rho-point fluxes: advective UFx, VFe (:1104-1112,:1263-1272), Coriolis (:1291-1300), curvilinear
//            (:1333-1347), viscous UFx, VFe (:1400-1414)
//            psi-point fluxes: advective UFe, VFx (:1141-1150,:1213-1222), viscous UFe, VFx (:1415-1430)
//   stage 3  fast-time averages (:614-682), pressure gradient (:944-1019), flux divergences, 2-D/3-D coupling
//            (:1884-2065), LF / AM3 stepping (:2098-2255), rhs history (:2420-2430), closed-wall BCs, periodic images.
// The kernel is bound by the L1 / shared-memory data pipe and by latency, not by HBM bandwidth (profiles/README.md), so the
// global operands of a stage are requested as early as the register budget allows -- those of stages 0, 1 and 2a at the top
// of the kernel, those of stage 3 before the barrier that ends stage 2 -- and the CTA has enough threads (NTH) to cover the
// (TX+1)x(TY+1) flux regions in a single pass.  On a ring of tiles the xi-halo exchange of the sub-step is part of this
// kernel (template XCH, dev.cuh Xchg).
#include <cstdio>
#include <cstring>
#include "dev.cuh"
#include "kernels.h"

namespace rb {

#ifndef S2D_TY
#define S2D_TY 8
#endif
#ifndef S2D_NTH
#define S2D_NTH 320
#endif
constexpr int TX = 32, TY = S2D_TY;      // output tile
constexpr int NTH = S2D_NTH;             // threads per CTA (>= (TX+1)*(TY+1) = 297)
constexpr int HL = 3, HH = 2;            // low / high halo of the staged inputs
constexpr int SW = TX + HL + HH;         // staged width  (37)
constexpr int SH = TY + HL + HH;         // staged height (13)
constexpr int ZW = TX + 1, ZH = TY + 1;  // flux / zeta regions: one extra column and row
constexpr int NS = SW * SH, NZ = ZW * ZH;
constexpr int SMEM_DOUBLES = 5 * NS + 17 * NZ;
static_assert(TX == 32 && TX * ZH + ZH <= NTH && NS <= 2 * NTH && TX * TY + 64 <= NTH, "tile / thread-count mismatch");

#ifndef S2D_MINB
#define S2D_MINB 2
#endif
#ifndef S2D_PF
#define S2D_PF 0        // L2 prefetch of the stage 2b / 3 operands at kernel start: measured slower (3.37 vs 3.09 ms per loop)
#endif
#ifndef S2D_EVICT
#define S2D_EVICT 1      // 1: the static operands (metrics, h, rhoA/rhoS, rufrc/rvfrc) are loaded with an L2 evict-first hint so that
#endif                   // they do not displace the time-varying barotropic state (86 MB at BENCHMARK3) from the 126 MB L2
#if S2D_EVICT
__device__ __forceinline__ unsigned long long evict_first_policy() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double lds_pol(const double* a, unsigned long long pol) {
  double v;
  asm("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
#define lds_(a) lds_pol(a, l2pol)
#else
#define lds_(a) (*(a))
#endif
// Time-varying 2-D fields.  In the exchange instance the ghost columns of these arrays are written by this very kernel (the pull
// below, through x.recvf) before they are read, so the reads must neither be declared __restrict__ / const (eligible for the
// read-only ld.global.nc path) nor be served from a stale L1 line: they go to L2 (ld.global.cg).  The single-tile instance
// keeps the plain cached loads.
template <bool XCH>
__device__ __forceinline__ double ldt(const double* a) { return XCH ? __ldcg(a) : *a; }

// C2: UV_C2ADVECTION, second-order centred advective fluxes (step2d_LF_AM3.h:1026-1080) instead of the fourth-order centred default
template <bool XCH, bool C2 = false>   // XCH: with the fused halo exchange (multi-GPU peer path); the single-tile instance carries none of it
__global__ void __launch_bounds__(NTH, S2D_MINB) k_step2d(Par p, Flds f, Xchg x) {
#if S2D_EVICT
  const unsigned long long l2pol = evict_first_policy();
#endif
  extern __shared__ double smem[];
  double* sD = smem; double* sU = sD + NS; double* sV = sU + NS; double* sDU = sV + NS; double* sDV = sDU + NS;
  // regions with origin (i0-1, j0-1): zeta-stage and rho-point fluxes
  double* sDnew = sDV + NS; double* sZw = sDnew + NZ; double* sG = sZw + NZ; double* sG2 = sG + NZ; double* sGSA = sG2 + NZ;
  double* aUFx = sGSA + NZ; double* aVFe = aUFx + NZ; double* cUFx = aVFe + NZ; double* cVFe = cUFx + NZ;
  double* kUFx = cVFe + NZ; double* kVFe = kUFx + NZ; double* vUFx = kVFe + NZ; double* vVFe = vUFx + NZ;
  // regions with origin (i0, j0): psi-point fluxes
  double* aUFe = vVFe + NZ; double* aVFx = aUFe + NZ; double* vUFe = aVFx + NZ; double* vVFx = vUFe + NZ;
  const int tid = threadIdx.x;
  // Tile origin.  The row blocks cover the interior rows 1..Mm; the two wall rows 0 and Mm+1 (fast-time averages only) are
  // taken by the two warps that idle in stage 3, in the first / last row block -- for Mm = 256 that is 32 instead of 33
  // row blocks, i.e. 2048 CTAs = 6.9 waves of 2 x 148 resident CTAs instead of 7.1 (a whole extra wave).
  // With the fused exchange the CTAs of the two edge column blocks are scheduled first (CTAs start in linear block order):
  // their pushes are then on the wire while the interior of this sub-step is still being computed, and the next sub-step's
  // edge CTAs -- first again -- find them delivered.
  int bx = blockIdx.x, by = blockIdx.y;
  if (XCH && gridDim.x >= 3) {
    const int nbx = gridDim.x, nby = gridDim.y, bid = bx + nbx * by;
    if (bid < 2 * nby) { bx = (bid & 1) ? nbx - 1 : 0; by = bid >> 1; }
    else { const int r = bid - 2 * nby; bx = 1 + r % (nbx - 2); by = r / (nbx - 2); }
  }
  const int i0 = xcol0(p, bx * TX), j0 = 1 + by * TY;
  const int P = p.P, Mm = p.Mm;
  // Periodic images (exchange_2d.F) only exist for columns 1..2 and Lm-2..Lm: the stores of all other column blocks skip the
  // two tests per store (CTA-uniform), which is ~4 % of this kernel's instructions.
  Par ps = p;
  ps.ew_wrap = (p.ew_wrap && (i0 <= 2 || i0 + TX - 1 >= p.Lm - 2)) ? 1 : 0;
  const bool PRED = p.predictor != 0;
  const bool FIRST = (p.iif == 1);
  const bool active = (p.iif <= p.nfast);                              // :755 (the nfast+1-th call only averages)
  const double* __restrict__ h = f.h;
  const double* zr = f.zeta[p.krhs];
  const double* zs = f.zeta[p.kstp];
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;

  // ---- fused halo exchange (dev.cuh Xchg): epochs; pull the ghost columns the previous sub-step's neighbours pushed.
  // One thread per CTA reads the epoch (every thread doing so would serialise ~10^4 requests per launch on one L2 line).
  __shared__ unsigned long long s_epoch;
  unsigned long long xe0 = 0;
  unsigned xtag = 0;
  size_t xbase = 0;
  if (XCH && (x.send | x.recv)) {
    const bool needW = (i0 - HL < x.Istr), needE = (i0 + TX - 1 + HH > x.Iend);
    if (tid == 0) s_epoch = *(volatile unsigned long long*)x.box;      // epoch of the latest completed push (same on every rank)
    if (x.recv && (needW || needE)) {                                  // CTA-uniform
      __syncthreads();
      xe0 = s_epoch;
      const unsigned tag = (unsigned)xe0;
      const double* slot = x.box + XHDR + (xe0 & (XSLOTS - 1)) * xslot_doubles(x.nj);
      const int nW = needW ? x.nrecv * SH * XNW : 0, nE = needE ? x.nrecv * SH * XNE : 0;
      for (int idx = tid; idx < nW + nE; idx += NTH) {
        const bool w = idx < nW;
        const int q = w ? idx : idx - nW, nc = w ? XNW : XNE;
        const int c = q % nc, r = (q / nc) % SH, fld = q / (nc * SH);
        const int j = j0 - HL + r;
        if (j >= 0 && j <= Mm + 1) {
          const double* line = slot + (w ? xline_w(x.nj, fld, j, c) : xline_e(x.nj, fld, j, c));
          double v;
          ll_wait(line, tag, v, x.timeout_ns, x.err);
          x.recvf[fld][j * P + (w ? x.Istr - XNW + c : x.Iend + 1 + c)] = v;
        }
      }
      __syncthreads();      // every ghost value this CTA reads below was written by this CTA above
    }
  }
  auto push = [&](int fld, int i, int j, double v) {
    if (i >= x.Iend - (XNW - 1)) ll_store(x.boxE + xbase + xline_w(x.nj, fld, j, i - (x.Iend - (XNW - 1))), v, xtag);   // -> east neighbour's west ghosts
    if (i <= x.Istr + (XNE - 1)) ll_store(x.boxW + xbase + xline_e(x.nj, fld, j, i - x.Istr), v, xtag);                 // -> west neighbour's east ghosts
  };

  const Xchg& xc = x;     // (`x` is shadowed by the new velocity in stage 3)

  // Stage-2 thread map: warps 0..ZH-1 take one region row each (32 columns, so a warp never straddles a row: conflict-free
  // 64-bit shared-memory accesses), the first ZH lanes of warp ZH take the 33rd column.
  const int za = (tid < TX * ZH) ? (tid & (TX - 1)) : TX;
  const int zb = (tid < TX * ZH) ? (tid / TX) : (tid - TX * ZH);
  const int zi = zb * ZW + za;
  // The global operands of stage 2a do not depend on shared memory: they are requested up front, together with those of
  // stage 0, so that the CTA pays one DRAM latency for both.
  const bool okA = active && tid < TX * ZH + ZH && (j0 - 1 + zb) >= 1 && (j0 - 1 + zb) <= Mm && (i0 - 1 + za) <= p.Iend;
  const int qA = okA ? (j0 - 1 + zb) * P + (i0 - 1 + za) : (j0 * P + i0);
  const double zs_q = ldt<XCH>(zs + qA), zr_q = ldt<XCH>(zr + qA), pm_q = lds_(pm + qA), pn_q = lds_(pn + qA), h_q = lds_(h + qA), rS = lds_(f.rhoS + qA), rA = lds_(f.rhoA + qA);
  const double fomn_q = lds_(f.fomn + qA), visc_q = lds_(f.visc2_r + qA), pmon_q = lds_(f.pmon_r + qA), pnom_q = lds_(f.pnom_r + qA);
  const double pnE_a = lds_(pn + qA + 1), pnW_a = lds_(pn + qA - 1), pmN_a = lds_(pm + qA + P), pmS_a = lds_(pm + qA - P), onr = lds_(f.on_r + qA), omr = lds_(f.om_r + qA);
  double dndx_q = 0.0, dmde_q = 0.0, rz_s = 0.0, rz_p = 0.0;
  if (p.curvgrid) { dndx_q = lds_(f.dndx + qA); dmde_q = lds_(f.dmde + qA); }
  if (!FIRST && !PRED) { rz_s = ldt<XCH>(f.rzeta[p.kstp] + qA); rz_p = ldt<XCH>(f.rzeta[p.ptsk] + qA); }
#if S2D_PF
  // L2 prefetch of the operands of stages 2b and 3 (18 arrays x TY rows x two 128-byte lines, one request per thread): no
  // register cost, and the loads issued two barriers later find their lines in L2.
  {
    const int a = tid >> 4, r = (tid >> 1) & 7, half = tid & 1;
    const double* pa = nullptr;
    switch (a) {
      case 0: pa = f.DU_avg2; break; case 1: pa = f.DV_avg2; break; case 2: pa = f.rufrc; break; case 3: pa = f.rvfrc; break;
      case 4: pa = f.ubar[p.kstp]; break; case 5: pa = f.vbar[p.kstp]; break;
      case 6: pa = f.visc2_p; break; case 7: pa = f.pmon_p; break; case 8: pa = f.pnom_p; break; case 9: pa = f.om_p; break; case 10: pa = f.on_p; break;
      case 11: if (PRED && !FIRST) pa = f.Zt_avg1; break; case 12: if (PRED && !FIRST) pa = f.DU_avg1; break; case 13: if (PRED && !FIRST) pa = f.DV_avg1; break;
      case 14: if (!PRED && !FIRST) pa = f.rubar[p.kstp]; break; case 15: if (!PRED && !FIRST) pa = f.rubar[p.ptsk]; break;
      case 16: if (!PRED && !FIRST) pa = f.rvbar[p.kstp]; break; case 17: if (!PRED && !FIRST) pa = f.rvbar[p.ptsk]; break;
      default: break;
    }
    if (pa && r < TY && j0 + r <= Mm + 1) pf_l2(pa + (j0 + r) * P + i0 + half * 16);
  }
#endif
  // ---- stages 0/1: Drhs, ubar, vbar, DUon, DVom on the staged region (two items per thread, loads first)
  {
    const double* ur = f.ubar[p.krhs];
    const double* vr = f.vbar[p.krhs];
    double zv[2], hv[2], uv[2], vv[2], onu[2], omv[2];
    bool ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tid + r * NTH;
      const int a = s % SW, b = s / SW;
      const int i = i0 - HL + a, j = j0 - HL + b;
      ok[r] = (s < NS) && i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
      const int q = ok[r] ? (j * P + i) : (j0 * P + i0);               // safe dummy address
      zv[r] = ldt<XCH>(zr + q); hv[r] = lds_(h + q); uv[r] = ldt<XCH>(ur + q); vv[r] = ldt<XCH>(vr + q); onu[r] = lds_(f.on_u + q); omv[r] = lds_(f.om_v + q);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tid + r * NTH;
      if (s < NS) { sD[s] = ok[r] ? (zv[r] + hv[r]) : 0.0; sU[s] = ok[r] ? uv[r] : 0.0; sV[s] = ok[r] ? vv[r] : 0.0; }
    }
    __syncthreads();
    if (XCH && x.send) {
      xe0 = s_epoch;
      const unsigned long long e = xe0 + 1;
      xtag = (unsigned)e; xbase = XHDR + (e & (XSLOTS - 1)) * xslot_doubles(x.nj);
      if (tid == 0) {
        // this CTA has read the epoch; the CTA that completes the count publishes the new one (it is only read by later
        // kernels of this stream)
        unsigned long long* hdr = (unsigned long long*)x.box;
        if (atomicAdd(&hdr[1], 1ULL) == (unsigned long long)gridDim.x * gridDim.y - 1) { hdr[1] = 0; __threadfence(); hdr[0] = e; }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tid + r * NTH;
      if (s < NS) {
        const int a = s % SW, b = s / SW;
        const int i = i0 - HL + a, j = j0 - HL + b;
        double du = 0.0, dv = 0.0;
        if (ok[r] && i > p.LBi && a >= 1) {
          const double c = 0.5 * onu[r];
          const double c1 = c * (sD[s] + sD[s - 1]);
          du = sU[s] * c1;
        }
        if (ok[r] && j >= 1 && b >= 1) {
          const double c = 0.5 * omv[r];
          const double c1 = c * (sD[s] + sD[s - SW]);
          dv = sV[s] * c1;
        }
        sDU[s] = du; sDV[s] = dv;
      }
    }
  }
  __syncthreads();

#define D_(di, dj) sD[c0 + (dj) * SW + (di)]
#define U_(di, dj) sU[c0 + (dj) * SW + (di)]
#define V_(di, dj) sV[c0 + (dj) * SW + (di)]
#define DU_(di, dj) sDU[c0 + (dj) * SW + (di)]
#define DV_(di, dj) sDV[c0 + (dj) * SW + (di)]
#define GXU(di, dj) (U_((di)-1, dj) - 2.0 * U_(di, dj) + U_((di) + 1, dj))
#define GXDU(di, dj) (DU_((di)-1, dj) - 2.0 * DU_(di, dj) + DU_((di) + 1, dj))
#define GXDV(di, dj) (DV_((di)-1, dj) - 2.0 * DV_(di, dj) + DV_((di) + 1, dj))
#define GYU(di, dj) (U_(di, (dj)-1) - 2.0 * U_(di, dj) + U_(di, (dj) + 1))
#define GXV(di, dj) (V_((di)-1, dj) - 2.0 * V_(di, dj) + V_((di) + 1, dj))
#define GYDU(di, dj) (DU_(di, (dj)-1) - 2.0 * DU_(di, dj) + DU_(di, (dj) + 1))
#define GYV(di, dj) (V_(di, (dj)-1) - 2.0 * V_(di, dj) + V_(di, (dj) + 1))
#define GYDV(di, dj) (DV_(di, (dj)-1) - 2.0 * DV_(di, dj) + DV_(di, (dj) + 1))

  if (active && tid < TX * ZH + ZH) {
    const double c6 = 1.0 / 6.0;
    // ---- stage 2a: rho-point quantities at (i0-1+za, j0-1+zb)
    {
      const int i = i0 - 1 + za, j = j0 - 1 + zb;
      const int c0 = (zb + HL - 1) * SW + (za + HL - 1);               // staged index of (i,j)
      // Straight-line arithmetic for every thread of the region (invalid points compute on in-bounds dummy operands and are
      // zeroed at the end): one basic block, so the independent flux chains below overlap in the FP64 pipe instead of
      // running one after the other behind a branch.
      double Dnew, zwrk, gz, gz2, gsa;
      double a_ufx, a_vfe, c_ufx, c_vfe, k_ufx = 0.0, k_vfe = 0.0, v_ufx, v_vfe;
      {
        // new free surface (:770-851)
        const double dd = (DU_(0, 0) - DU_(1, 0)) + (DV_(0, 0) - DV_(0, 1));
        double zeta_new;
        const double pmn = pm_q * pn_q;
        if (FIRST) {
          zeta_new = zs_q + pmn * p.dtfast * dd;
          zwrk = 0.5 * (zs_q + zeta_new);
        } else if (PRED) {
          const double cff1 = 2.0 * p.dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
          zeta_new = zs_q + pmn * cff1 * dd;
          zwrk = cff5 * zr_q + cff4 * (zs_q + zeta_new);
        } else {
          const double cff1 = p.dtfast * 5.0 / 12.0, cff2 = p.dtfast * 8.0 / 12.0, cff3 = p.dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
          const double cff = cff1 * dd;
          zeta_new = zs_q + pmn * (cff + cff2 * rz_s - cff3 * rz_p);
          zwrk = cff5 * zeta_new + cff4 * zr_q;
        }
        Dnew = zeta_new + h_q;
        gz = (1000.0 / p.rho0 + rS) * zwrk;
        gz2 = gz * zwrk;
        gsa = zwrk * (rS - rA);
        if (okA && za >= 1 && zb >= 1) {                                 // own points of this tile
          const int q = qA;
          (void)q;
          st_r_grad(f.zeta[p.knew], j * P, i, j, zeta_new, ps);
          if (PRED) st_w(f.rzeta[p.krhs], j * P, i, dd, ps);
          if (XCH && x.send && (i >= x.Iend - (XNW - 1) || i <= x.Istr + (XNE - 1))) {
            push(0, i, j, zeta_new);
            if (j == 1) push(0, i, 0, zeta_new);
            if (j == Mm) push(0, i, Mm + 1, zeta_new);
            if (PRED) {                                                // rzeta has no wall-row values of its own: forward what is there
              push(3, i, j, dd);
              if (j == 1) push(3, i, 0, f.rzeta[p.krhs][i]);
              if (j == Mm) push(3, i, Mm + 1, f.rzeta[p.krhs][(Mm + 1) * P + i]);
            }
          }
        }
        // advective UFx at rho(i,j) (:1104-1112)
        if (C2) a_ufx = 0.25 * (DU_(0, 0) + DU_(1, 0)) * (U_(0, 0) + U_(1, 0));                         // :1027-1038
        else a_ufx = 0.25 * (U_(0, 0) + U_(1, 0) - c6 * (GXU(0, 0) + GXU(1, 0))) * (DU_(0, 0) + DU_(1, 0) - c6 * (GXDU(0, 0) + GXDU(1, 0)));
        // advective VFe at rho(i,j) (:1263-1272): grad/Dgrad rows 2..Mm with wall copies (1)=(2), (Mm+1)=(Mm)
        if (C2) a_vfe = 0.25 * (DV_(0, 0) + DV_(0, 1)) * (V_(0, 0) + V_(0, 1));                         // :1066-1077
        else {
          const int da = (j < 2) ? 1 : 0, db = (j + 1 > Mm) ? 0 : 1;
          a_vfe = 0.25 * (V_(0, 0) + V_(0, 1) - c6 * (GYV(0, da) + GYV(0, db))) * (DV_(0, 0) + DV_(0, 1) - c6 * (GYDV(0, da) + GYDV(0, db)));
        }
        // Coriolis (:1291-1300) and curvilinear (:1333-1347) at rho(i,j)
        const double D0 = D_(0, 0);
        const double vS = V_(0, 0) + V_(0, 1), uS = U_(0, 0) + U_(1, 0);
        {
          const double c = 0.5 * D0 * fomn_q;
          c_ufx = c * vS; c_vfe = c * uS;
        }
        if (p.curvgrid) {
          const double c1 = 0.5 * vS, c2 = 0.5 * uS;
          const double c = D0 * (c1 * dndx_q - c2 * dmde_q);
          k_ufx = c * c1; k_vfe = c * c2;
        }
        // viscous stress at rho(i,j) (:1400-1414)
        {
          const double cr = visc_q * D0 * 0.5 *
                            (pmon_q * ((pn_q + pnE_a) * U_(1, 0) - (pnW_a + pn_q) * U_(0, 0)) -
                             pnom_q * ((pm_q + pmN_a) * V_(0, 1) - (pmS_a + pm_q) * V_(0, 0)));
          v_ufx = onr * onr * cr; v_vfe = omr * omr * cr;
        }
      }
      sDnew[zi] = okA ? Dnew : 0.0; sZw[zi] = okA ? zwrk : 0.0; sG[zi] = okA ? gz : 0.0; sG2[zi] = okA ? gz2 : 0.0; sGSA[zi] = okA ? gsa : 0.0;
      aUFx[zi] = okA ? a_ufx : 0.0; aVFe[zi] = okA ? a_vfe : 0.0; cUFx[zi] = okA ? c_ufx : 0.0; cVFe[zi] = okA ? c_vfe : 0.0;
      kUFx[zi] = okA ? k_ufx : 0.0; kVFe[zi] = okA ? k_vfe : 0.0; vUFx[zi] = okA ? v_ufx : 0.0; vVFe[zi] = okA ? v_vfe : 0.0;
    }
    // ---- stage 2b: psi-point fluxes at (i0+za, j0+zb)
    {
      const int i = i0 + za, j = j0 + zb;
      const int c0 = (zb + HL) * SW + (za + HL);
      const bool okB = j >= 1 && j <= Mm + 1 && i <= p.Iend + 1;
      const int q = okB ? j * P + i : (j0 + 1) * P + i0 + 1;             // dummy: a point of this tile that has all four neighbours
      double a_ufe, a_vfx, v_ufe, v_vfx;
      {
        const double visc_q = lds_(f.visc2_p + q), pmon_q = lds_(f.pmon_p + q), pnom_q = lds_(f.pnom_p + q), omp = lds_(f.om_p + q), onp = lds_(f.on_p + q);
        const double pn_q = lds_(pn + q), pnS = lds_(pn + q - P), pnW = lds_(pn + q - 1), pnSW = lds_(pn + q - P - 1);
        const double pm_q = lds_(pm + q), pmS = lds_(pm + q - P), pmW = lds_(pm + q - 1), pmSW = lds_(pm + q - P - 1);
        // advective UFe at psi(i,j) (:1141-1150): grad = d2y(ubar), rows 1..Mm with wall copies (0)=(1), (Mm+1)=(Mm)
        if (C2) a_ufe = 0.25 * (DV_(0, 0) + DV_(-1, 0)) * (U_(0, 0) + U_(0, -1));                      // :1040-1051
        else {
          const int d0 = (j > Mm) ? -1 : 0, dm = (j - 1 < 1) ? 0 : -1;
          a_ufe = 0.25 * (U_(0, 0) + U_(0, -1) - c6 * (GYU(0, d0) + GYU(0, dm))) * (DV_(0, 0) + DV_(-1, 0) - c6 * (GXDV(0, 0) + GXDV(-1, 0)));
        }
        // advective VFx at psi(i,j), j = 2..Mm (:1213-1222)
        if (C2) a_vfx = 0.25 * (DU_(0, 0) + DU_(0, -1)) * (V_(0, 0) + V_(-1, 0));                      // :1053-1064
        else a_vfx = 0.25 * (V_(0, 0) + V_(-1, 0) - c6 * (GXV(0, 0) + GXV(-1, 0))) * (DU_(0, 0) + DU_(0, -1) - c6 * (GYDU(0, 0) + GYDU(0, -1)));
        // viscous stress at psi(i,j) (:1394-1430)
        {
          const double Dp = 0.25 * (D_(0, 0) + D_(-1, 0) + D_(0, -1) + D_(-1, -1));
          const double cp = visc_q * Dp * 0.5 *
                            (pmon_q * ((pnS + pn_q) * V_(0, 0) - (pnSW + pnW) * V_(-1, 0)) +
                             pnom_q * ((pmW + pm_q) * U_(0, 0) - (pmSW + pmS) * U_(0, -1)));
          v_ufe = omp * omp * cp; v_vfx = onp * onp * cp;
        }
      }
      aUFe[zi] = okB ? a_ufe : 0.0; aVFx[zi] = (okB && j >= 2 && j <= Mm) ? a_vfx : 0.0; vUFe[zi] = okB ? v_ufe : 0.0; vVFx[zi] = okB ? v_vfx : 0.0;
    }
  }
  // ---- stage 3: one thread per rho point of the tile (+ one warp each for wall rows 0 and Mm+1).  None of its global
  // operands depends on what stage 2 stored, so they are requested BEFORE the barrier that ends stage 2: the loads are in
  // flight while the slower warps of the CTA finish their fluxes.
  int tx, ty;
  bool live = true;
  if (tid < TX * TY) {
    tx = tid % TX; ty = tid / TX;
    if (j0 + ty > Mm) live = false;
  } else {
    const int w = (tid - TX * TY) >> 5;
    tx = tid & 31;
    ty = 0;
    if (w == 0 && by == 0) ty = -1;                              // row 0
    else if (w == 1 && j0 <= Mm && j0 + TY - 1 >= Mm) ty = Mm + 1 - j0;  // row Mm+1, in the block that holds row Mm
    else live = false;
  }
  if (i0 + tx > p.Iend) live = false;
  if (!live) { tx = 0; ty = 0; }                                        // dummy point of this tile: keeps the loads in bounds
  const int i = i0 + tx, j = j0 + ty;
  const int o = j * P + i;
  const bool inner = active && j >= 1 && j <= Mm;
  const bool dov = inner && (j >= p.JstrV);
  const int oS = (j >= 1) ? o - P : o;                                  // row j-1 (clamped so the loads stay in bounds)
  // all global operands of this stage, issued back to back
  const double zr_o = ldt<XCH>(zr + o);
  const double av_du2 = f.DU_avg2[o], av_dv2 = f.DV_avg2[o];
  double av_zt = 0.0, av_du1 = 0.0, av_dv1 = 0.0;
  if (PRED && !FIRST) { av_zt = f.Zt_avg1[o]; av_du1 = f.DU_avg1[o]; av_dv1 = f.DV_avg1[o]; }
  const double h0 = lds_(h + o), hW = lds_(h + o - 1), hS = lds_(h + oS);
  const double rA0 = lds_(f.rhoA + o), rAW = lds_(f.rhoA + o - 1), rAS = lds_(f.rhoA + oS);
  const double pm0 = lds_(pm + o), pmW = lds_(pm + o - 1), pmS = lds_(pm + oS), pn0 = lds_(pn + o), pnW = lds_(pn + o - 1), pnS = lds_(pn + oS);
  const double onu = lds_(f.on_u + o), omv = lds_(f.om_v + o);
  const double zs0 = ldt<XCH>(zs + o), zsW = ldt<XCH>(zs + o - 1), zsS = ldt<XCH>(zs + oS);
  const double us = f.ubar[p.kstp][o], vs = f.vbar[p.kstp][o];
  const double rufrc_o = lds_(f.rufrc + o), rvfrc_o = lds_(f.rvfrc + o);
  double rub_s = 0.0, rub_p = 0.0, rvb_s = 0.0, rvb_p = 0.0, ru_n = 0.0, ru_so = 0.0, rv_n = 0.0, rv_so = 0.0;
  if (!FIRST && !PRED) { rub_s = f.rubar[p.kstp][o]; rub_p = f.rubar[p.ptsk][o]; rvb_s = f.rvbar[p.kstp][o]; rvb_p = f.rvbar[p.ptsk][o]; }
  if (FIRST && PRED && p.istart >= 1) { ru_n = f.ru[p.nnew][o]; rv_n = f.rv[p.nnew][o]; ru_so = f.ru[p.nstp][o]; rv_so = f.rv[p.nstp][o]; }
  __syncthreads();
  if (!live) return;

  // fast-time averages (:614-682); rows 0..Mm+1 for Zt/DU, rows 1..Mm+1 for DV
  {
    const int c0 = (ty + HL) * SW + (tx + HL);
    const double DUo = DU_(0, 0), DVo = DV_(0, 0);
    if (PRED) {
      if (FIRST) {
        const double cff2 = (-1.0 / 12.0) * p.w2_p1;
        st_w(f.Zt_avg1, j * P, i, 0.0, ps);
        st_w(f.DU_avg1, j * P, i, 0.0, ps);
        f.DU_avg2[o] = cff2 * DUo;
        if (j >= 1) { st_w(f.DV_avg1, j * P, i, 0.0, ps); f.DV_avg2[o] = cff2 * DVo; }
      } else {
        const double cff1 = p.w1_m1;
        const double cff2 = (8.0 / 12.0) * p.w2_0 - (1.0 / 12.0) * p.w2_p1;
        st_w(f.Zt_avg1, j * P, i, av_zt + cff1 * zr_o, ps);
        st_w(f.DU_avg1, j * P, i, av_du1 + cff1 * DUo, ps);
        f.DU_avg2[o] = av_du2 + cff2 * DUo;
        if (j >= 1) {
          st_w(f.DV_avg1, j * P, i, av_dv1 + cff1 * DVo, ps);
          f.DV_avg2[o] = av_dv2 + cff2 * DVo;
        }
      }
    } else {
      const double cff2 = FIRST ? p.w2_0 : (5.0 / 12.0) * p.w2_0;
      f.DU_avg2[o] = av_du2 + cff2 * DUo;
      if (j >= 1) f.DV_avg2[o] = av_dv2 + cff2 * DVo;
    }
  }
  if (!inner) return;
  const int z0 = (ty + 1) * ZW + (tx + 1), zW = z0 - 1, zS = z0 - ZW;   // rho-region indices of (i,j), (i-1,j), (i,j-1)
  const int p0 = ty * ZW + tx, pE = p0 + 1, pN = p0 + ZW;               // psi-region indices of (i,j), (i+1,j), (i,j+1)

  // ---- u-point (i,j) and v-point (i,j): the arithmetic of both first, as one basic block (two independent dependency chains
  // that overlap in the FP64 pipe), the stores afterwards; the v results of row 1 (the wall) are computed and dropped
  const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
  double rhs_u = cff1 * onu *
                 ((hW + h0) * (sG[zW] - sG[z0]) +
                  (hW - h0) * (sGSA[zW] + sGSA[z0] + cff2 * (rAW - rA0) * (sZw[zW] - sZw[z0])) +
                  (sG2[zW] - sG2[z0]));
  double rhs_v = cff1 * omv *
                 ((hS + h0) * (sG[zS] - sG[z0]) +
                  (hS - h0) * (sGSA[zS] + sGSA[z0] + cff2 * (rAS - rA0) * (sZw[zS] - sZw[z0])) +
                  (sG2[zS] - sG2[z0]));
  {
    const double a1 = aUFx[z0] - aUFx[zW];
    const double a2 = aUFe[pN] - aUFe[p0];
    const double fc = a1 + a2;
    rhs_u = rhs_u - fc;
  }
  {
    const double a1 = aVFx[pE] - aVFx[p0];
    const double a2 = aVFe[z0] - aVFe[zS];
    const double fc = a1 + a2;
    rhs_v = rhs_v - fc;
  }
  rhs_u = rhs_u + 0.5 * (cUFx[z0] + cUFx[zW]);
  rhs_v = rhs_v - 0.5 * (cVFe[z0] + cVFe[zS]);
  if (p.curvgrid) {
    rhs_u = rhs_u + 0.5 * (kUFx[z0] + kUFx[zW]);
    rhs_v = rhs_v - 0.5 * (kVFe[z0] + kVFe[zS]);
  }
  {
    const double a1 = 0.5 * (pnW + pn0) * (vUFx[z0] - vUFx[zW]);
    const double a2 = 0.5 * (pmW + pm0) * (vUFe[pN] - vUFe[p0]);
    const double fc = a1 + a2;
    rhs_u = rhs_u + fc;
  }
  {
    const double a1 = 0.5 * (pnS + pn0) * (vVFx[pE] - vVFx[p0]);
    const double a2 = 0.5 * (pmS + pm0) * (vVFe[z0] - vVFe[zS]);
    const double fc = a1 - a2;
    rhs_v = rhs_v + fc;
  }
  // coupling with the 3-D equations (:1884-2065); level k = 0 planes of ru carry the AB3 history of the 2-D forcing
  double rf_u = 0.0, rf_v = 0.0;
  if (FIRST && PRED) {
    rf_u = rufrc_o - rhs_u;
    rf_v = rvfrc_o - rhs_v;
    if (p.istart == 0) { rhs_u = rhs_u + rf_u; rhs_v = rhs_v + rf_v; }
    else if (p.istart == 1) { rhs_u = rhs_u + 1.5 * rf_u - 0.5 * ru_n; rhs_v = rhs_v + 1.5 * rf_v - 0.5 * rv_n; }
    else {
      rhs_u = rhs_u + (23.0 / 12.0) * rf_u - (16.0 / 12.0) * ru_n + (5.0 / 12.0) * ru_so;
      rhs_v = rhs_v + (23.0 / 12.0) * rf_v - (16.0 / 12.0) * rv_n + (5.0 / 12.0) * rv_so;
    }
  } else {
    rhs_u = rhs_u + rufrc_o;
    rhs_v = rhs_v + rvfrc_o;
  }
  // time stepping (:2098-2255)
  double xu, xv;
  {
    const double Dstp_u = (zs0 + h0) + (zsW + hW), Dstp_v = (zs0 + h0) + (zsS + hS);
    const double cff_u = (pm0 + pmW) * (pn0 + pnW), cff_v = (pm0 + pmS) * (pn0 + pnS);
    const double fc_u = 1.0 / (sDnew[z0] + sDnew[zW]), fc_v = 1.0 / (sDnew[z0] + sDnew[zS]);
    if (FIRST || PRED) {
      const double c1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
      xu = (us * Dstp_u + cff_u * c1 * rhs_u) * fc_u;
      xv = (vs * Dstp_v + cff_v * c1 * rhs_v) * fc_v;
    } else {
      const double c1 = 0.5 * p.dtfast * 5.0 / 12.0, c2 = 0.5 * p.dtfast * 8.0 / 12.0, c3 = 0.5 * p.dtfast * 1.0 / 12.0;
      xu = (us * Dstp_u + cff_u * (c1 * rhs_u + c2 * rub_s - c3 * rub_p)) * fc_u;
      xv = (vs * Dstp_v + cff_v * (c1 * rhs_v + c2 * rvb_s - c3 * rvb_p)) * fc_v;
    }
  }
  // ---- stores: rhs history (:2420-2430), BCs (:2451-2460), periodic images (:2509-2524)
  if (FIRST && PRED) { f.rufrc[o] = rf_u; f.ru[p.nstp][o] = rf_u; }
  st_u_closed(f.ubar[p.knew], j * P, i, j, xu, ps);
  if (PRED) f.rubar[p.krhs][o] = rhs_u;
  if (XCH && xc.send && (i >= xc.Iend - (XNW - 1) || i <= xc.Istr + (XNE - 1))) {
    push(1, i, j, xu);
    if (j == 1) push(1, i, 0, p.gamma2 * xu);
    if (j == Mm) push(1, i, Mm + 1, p.gamma2 * xu);
  }
  if (dov) {
    if (FIRST && PRED) { f.rvfrc[o] = rf_v; f.rv[p.nstp][o] = rf_v; }
    st_v_closed(f.vbar[p.knew], j * P, i, j, xv, ps);
    if (PRED) f.rvbar[p.krhs][o] = rhs_v;
    if (XCH && xc.send && (i >= xc.Iend - (XNW - 1) || i <= xc.Istr + (XNE - 1))) {
      push(2, i, j, xv);
      if (j == 2) { push(2, i, 1, 0.0); push(2, i, 0, f.vbar[p.knew][i]); }   // row 1 is the wall (v = 0); row 0 is never written
      if (j == Mm) push(2, i, Mm + 1, 0.0);
    }
  }
}

void launch_step2d(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x) {
  dim3 g((xspan(p) + TX - 1) / TX, (p.Mm + TY - 1) / TY);
  const size_t smem = (size_t)SMEM_DOUBLES * sizeof(double);
  static bool done[MAXDEV] = {false};
  bool& once = done[cur_dev()];
  if (!once) {
    // 59.6 KB of dynamic shared memory is above the 48 KB default: without the opt-in the launch fails
    cudaError_t e1 = cudaFuncSetAttribute(k_step2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(k_step2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_step2d<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_step2d<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      std::fprintf(stderr, "roms_b200: cannot opt in to %zu bytes of shared memory for k_step2d: %s\n", smem, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return;      // the sticky CUDA error is picked up by the caller's cudaGetLastError (run_phase_async -> exit_flag 8)
    }
    once = true;
  }
  Xchg none;
  std::memset(&none, 0, sizeof(none));
  if (p.uv_adv == 3) {                                                  // UV_C2ADVECTION
    if (x && (x->send || x->recv)) k_step2d<true, true><<<g, NTH, smem, s>>>(p, f, *x);
    else k_step2d<false, true><<<g, NTH, smem, s>>>(p, f, none);
    return;
  }
  if (x && (x->send || x->recv)) k_step2d<true><<<g, NTH, smem, s>>>(p, f, *x);
  else k_step2d<false><<<g, NTH, smem, s>>>(p, f, none);
}

}  // namespace rb

// LOOP_2D (ROMS/Nonlinear/main3d.F:592-700) as ONE persistent kernel: all 2*nfast+1 calls of step2d_tile
// (ROMS/Nonlinear/step2d_LF_AM3.h:137-2528) of a baroclinic step, for tiles small enough that every CTA can own a fixed
// TX x TYL block of rho points for the whole loop (BENCHMARK3 on 8 GPUs: 256 x 256 points per GPU = 128 CTAs on 148 SMs).
//
// Why: with one kernel per sub-step (k_step2d.cu) such a tile is a single wave of CTAs, so a sub-step costs one CTA latency
// (three dependent L2 / DRAM round trips and three barriers, ~8 us) plus a graph-node hand-off, 59 times per step -- the
// latency floor SURVEY.md section 8d names.  Here
//   * the ~26 operands that do not change during the loop (metrics, h, rhoA / rhoS, viscosity, the 3-D forcing rufrc /
//     rvfrc) are staged in shared memory ONCE per baroclinic step instead of being fetched 59 times;
//   * the fast-time averages Zt_avg1, DU_avg1/2, DV_avg1/2 (read-modify-written by the same thread in every call) stay in
//     registers for the whole loop and are stored once;
//   * consecutive sub-steps are ordered by per-CTA completion flags: a CTA starts call c when its (up to) 8 neighbours have
//     published call c-1 -- everything it reads off-tile was written then, and everything it is about to overwrite was last
//     read then -- so there is no grid-wide barrier and no kernel boundary inside the loop;
//   * on a ring of GPUs the xi-halo travels exactly as in k_step2d<true> (LL push at the stores, pull at the start of the next
//     call; dev.cuh Xchg), with the epoch derived from the call number.
// The per-point arithmetic is the same sequence of IEEE operations as k_step2d.cu (and the oracle): the strict build is
// bit-exact with either.  Time-varying fields are read with ld.global.cg (L2): they are written by other SMs during the
// kernel, so the non-coherent L1 must be bypassed.
#include <cstdio>
#include <cstring>
#include "dev.cuh"
#include "kernels.h"

namespace rb {

namespace loopk {
#ifndef LK_TY
#define LK_TY 8
#endif
#ifndef LK_NTH
#define LK_NTH (LK_TY == 8 ? 320 : 640)
#endif
#ifndef LK_MINB
#define LK_MINB (LK_TY == 8 ? 2 : 1)
#endif
constexpr int TX = 32, TY = LK_TY;       // output tile of one CTA
constexpr int NTH = LK_NTH;              // >= (TX+1)*(TY+1) and >= TX*TY + 64
constexpr int HL = 3, HH = 2;
constexpr int SW = TX + HL + HH, SH = TY + HL + HH;      // staged inputs, origin (i0-3, j0-3): 37 x 21
constexpr int ZW = TX + 1, ZH = TY + 1;                  // rho region, origin (i0-1, j0-1), and psi region, origin (i0, j0): 33 x 17
constexpr int MW = TX + 4, MH = TY + 4;                  // pm / pn with one more ring, origin (i0-2, j0-2): 36 x 20
constexpr int NS = SW * SH, NZ = ZW * ZH, NM = MW * MH, NO = TX * TY;
// work: sD sU sV sDU sDV (NS) + 17 flux regions (NZ); static: h on_u om_v (NS), pm pn (NM), 10 rho-point + 5 psi-point (NZ).
// 32 x 8 tiles: 113.7 KB, so that two CTAs share an SM (one computes while the other waits for its neighbours or for L2).
constexpr int SMEM_DOUBLES = 5 * NS + 17 * NZ + 3 * NS + 2 * NM + 15 * NZ;
static_assert(TX == 32 && TX * ZH + ZH <= NTH && NS <= 2 * NTH && TX * TY + 64 <= NTH && NM <= 2 * NTH, "tile / thread-count mismatch");
static_assert(((size_t)SMEM_DOUBLES * 8 + 1024 + 64) * LK_MINB <= 233472, "shared memory budget of one SM");
}  // namespace loopk

using namespace loopk;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// time-varying field written by other CTAs during this kernel: L2 is the point of coherence
__device__ __forceinline__ double ldv(const double* a) { return __ldcg(a); }

#ifdef LK_TRACE
// tuning aid (tools/variant.py -DLK_TRACE): wall-clock stamps of one CTA at the stage boundaries of calls 20..27
#define TRACE(slot) do { if (tid == 0 && bx == (int)gridDim.x / 2 && by == (int)gridDim.y / 2 && c >= 20 && c < 28) ctl.flags[2048 + (c - 20) * 8 + (slot)] = (unsigned long long)gtime_ns(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

template <bool XCH>
__global__ void __launch_bounds__(NTH, LK_MINB) k_step2d_loop(Par p, Flds f, Xchg x, LoopCtl ctl) {
  extern __shared__ double smem[];
  double* sD = smem; double* sU = sD + NS; double* sV = sU + NS; double* sDU = sV + NS; double* sDV = sDU + NS;
  double* sDnew = sDV + NS; double* sZw = sDnew + NZ; double* sG = sZw + NZ; double* sG2 = sG + NZ; double* sGSA = sG2 + NZ;
  double* aUFx = sGSA + NZ; double* aVFe = aUFx + NZ; double* cUFx = aVFe + NZ; double* cVFe = cUFx + NZ;
  double* kUFx = cVFe + NZ; double* kVFe = kUFx + NZ; double* vUFx = kVFe + NZ; double* vVFe = vUFx + NZ;
  double* aUFe = vVFe + NZ; double* aVFx = aUFe + NZ; double* vUFe = aVFx + NZ; double* vVFx = vUFe + NZ;
  // operands that do not change during the loop
  double* cH = vVFx + NZ; double* cONU = cH + NS; double* cOMV = cONU + NS;
  double* cPM = cOMV + NS; double* cPN = cPM + NM;
  double* cRS = cPN + NM; double* cRA = cRS + NZ; double* cFOMN = cRA + NZ; double* cVISR = cFOMN + NZ; double* cPMONR = cVISR + NZ;
  double* cPNOMR = cPMONR + NZ; double* cONR = cPNOMR + NZ; double* cOMR = cONR + NZ; double* cDNDX = cOMR + NZ; double* cDMDE = cDNDX + NZ;
  double* cVISP = cDMDE + NZ; double* cPMONP = cVISP + NZ; double* cPNOMP = cPMONP + NZ; double* cOMP = cPNOMP + NZ; double* cONP = cOMP + NZ;
  __shared__ unsigned long long s_base, s_epoch;

  const int tid = threadIdx.x;
  const int bx = blockIdx.x, by = blockIdx.y, nbx = gridDim.x, nby = gridDim.y;
  const int i0 = p.Istr + bx * TX, j0 = 1 + by * TY;
  const int P = p.P, Mm = p.Mm;
  const double* __restrict__ h = f.h;
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;

  if (tid == 0) {
    s_base = ctl.base[0];
    s_epoch = (XCH && (x.box != nullptr)) ? *(volatile unsigned long long*)x.box : 0ULL;
  }

  // ---- stage the loop-invariant operands (zero where the array has no such element: those entries are never used)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int s = tid + r * NTH;
    if (s < NS) {
      const int a = s % SW, b = s / SW;
      const int i = i0 - HL + a, j = j0 - HL + b;
      const bool ok = i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
      const int q = j * P + i;
      cH[s] = ok ? __ldg(h + q) : 0.0; cONU[s] = ok ? __ldg(f.on_u + q) : 0.0; cOMV[s] = ok ? __ldg(f.om_v + q) : 0.0;
    }
    if (s < NM) {
      const int a = s % MW, b = s / MW;
      const int i = i0 - 2 + a, j = j0 - 2 + b;
      const bool ok = i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
      const int q = j * P + i;
      cPM[s] = ok ? __ldg(pm + q) : 0.0; cPN[s] = ok ? __ldg(pn + q) : 0.0;
    }
  }
  if (tid < NZ) {
    const int za_ = tid % ZW, zb_ = tid / ZW;
    {
      const int i = i0 - 1 + za_, j = j0 - 1 + zb_;
      const bool ok = i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
      const int q = j * P + i;
      cRS[tid] = ok ? __ldg(f.rhoS + q) : 0.0; cRA[tid] = ok ? __ldg(f.rhoA + q) : 0.0; cFOMN[tid] = ok ? __ldg(f.fomn + q) : 0.0;
      cVISR[tid] = ok ? __ldg(f.visc2_r + q) : 0.0; cPMONR[tid] = ok ? __ldg(f.pmon_r + q) : 0.0; cPNOMR[tid] = ok ? __ldg(f.pnom_r + q) : 0.0;
      cONR[tid] = ok ? __ldg(f.on_r + q) : 0.0; cOMR[tid] = ok ? __ldg(f.om_r + q) : 0.0;
      cDNDX[tid] = (ok && p.curvgrid) ? __ldg(f.dndx + q) : 0.0; cDMDE[tid] = (ok && p.curvgrid) ? __ldg(f.dmde + q) : 0.0;
    }
    {
      const int i = i0 + za_, j = j0 + zb_;
      const bool ok = i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
      const int q = j * P + i;
      cVISP[tid] = ok ? __ldg(f.visc2_p + q) : 0.0; cPMONP[tid] = ok ? __ldg(f.pmon_p + q) : 0.0; cPNOMP[tid] = ok ? __ldg(f.pnom_p + q) : 0.0;
      cOMP[tid] = ok ? __ldg(f.om_p + q) : 0.0; cONP[tid] = ok ? __ldg(f.on_p + q) : 0.0;
    }
  }

  // ---- fixed thread maps (the same for every call)
  // stage 2: warps 0..ZH-1 take one region row each, the first ZH lanes of warp ZH take the 33rd column
  const int za = (tid < TX * ZH) ? (tid & (TX - 1)) : TX;
  const int zb = (tid < TX * ZH) ? (tid / TX) : (tid - TX * ZH);
  const int zi = zb * ZW + za;
  const bool st2 = tid < TX * ZH + ZH;
  const bool okA0 = st2 && (j0 - 1 + zb) >= 1 && (j0 - 1 + zb) <= Mm && (i0 - 1 + za) <= p.Iend;
  const int qA = okA0 ? (j0 - 1 + zb) * P + (i0 - 1 + za) : (j0 * P + i0);
  const int mA = (zb + 1) * MW + (za + 1);                 // pm / pn index of rho-region point (za, zb)
  // stage 3: one thread per rho point of the tile, + one warp each for wall rows 0 and Mm+1
  int tx, ty;
  bool live = true;
  if (tid < TX * TY) {
    tx = tid % TX; ty = tid / TX;
    if (j0 + ty > Mm) live = false;
  } else {
    const int w = (tid - TX * TY) >> 5;
    tx = tid & 31;
    ty = 0;
    if (w == 0 && by == 0) ty = -1;
    else if (w == 1 && j0 <= Mm && j0 + TY - 1 >= Mm) ty = Mm + 1 - j0;
    else live = false;
  }
  if (i0 + tx > p.Iend) live = false;
  if (!live) { tx = 0; ty = 0; }
  const int i3 = i0 + tx, j3 = j0 + ty;
  const int o = j3 * P + i3;
  const int oS = (j3 >= 1) ? o - P : o;
  const int a3 = (ty + HL) * SW + (tx + HL);               // staged index of (i3, j3)
  const int a3S = (j3 >= 1) ? a3 - SW : a3;
  const int m3 = (ty + 2) * MW + (tx + 2);                 // pm / pn index of (i3, j3)
  const int m3S = (j3 >= 1) ? m3 - MW : m3;
  const int z0 = (ty + 1) * ZW + (tx + 1), zW = z0 - 1, zS = (j3 >= 1) ? z0 - ZW : z0;
  const int p0 = ty * ZW + tx, pE = p0 + 1, pN = p0 + ZW;
  // fast-time averages of this thread's point: registers for the whole loop (the first call initialises them, :614-682)
  double av_zt = 0.0, av_du1 = 0.0, av_du2 = 0.0, av_dv1 = 0.0, av_dv2 = 0.0;
  // 3-D forcing of this thread's u / v point: read once, replaced by the first call (:1884-2065)
  double rufrc_o = 0.0, rvfrc_o = 0.0;
  if (live && j3 >= 1 && j3 <= Mm) { rufrc_o = __ldg(f.rufrc + o); rvfrc_o = __ldg(f.rvfrc + o); }

  // neighbours whose completion flags order the calls (periodic wrap only when this tile owns the whole xi range)
  int nb_id = -1;
  if (tid < 8) {
    const int dx = (tid < 3) ? tid - 1 : (tid < 5 ? (tid == 3 ? -1 : 1) : tid - 6);
    const int dy = (tid < 3) ? -1 : (tid < 5 ? 0 : 1);
    int nx = bx + dx;
    const int ny = by + dy;
    if (p.ew_wrap) nx = (nx + nbx) % nbx;
    if (nx >= 0 && nx < nbx && ny >= 0 && ny < nby && !(nx == bx && ny == by)) nb_id = nx + nbx * ny;
  }
  __syncthreads();
  const unsigned long long base = s_base;
  const unsigned long long xe0 = s_epoch;
  const Xchg& xc = x;

  for (int c = 1; c <= ctl.ncall; ++c) {
    const LoopStep st = ctl.steps[c - 1];
    const bool PRED = st.predictor != 0;
    const bool FIRST = (st.iif == 1);
    const bool active = (st.iif <= p.nfast);
    const bool LAST = (c == ctl.ncall);
    const double* zr = f.zeta[st.krhs];
    const double* zs = f.zeta[st.kstp];

    TRACE(0);
    // ---- order: every neighbour has finished call c-1
    if (c > 1) {
      if (nb_id >= 0) {
        const unsigned long long want = base + (unsigned long long)(c - 1);
        const unsigned long long* fl = ctl.flags + nb_id;
        if (ld_acquire_u64(fl) < want) {
          const long long t0 = gtime_ns();
          for (;;) {
            bool done = false;
#pragma unroll 1
            for (int q = 0; q < 64 && !done; ++q) done = ld_acquire_u64(fl) >= want;
            if (done) break;
            if (*(volatile unsigned long long*)ctl.err != 0ULL) break;                    // another wait already gave up
            if (ctl.timeout_ns > 0 && gtime_ns() - t0 > ctl.timeout_ns) { *ctl.err = 2ULL; break; }
          }
        }
      }
      __syncthreads();
    }

    // ---- fused halo exchange: pull the ghost columns the neighbours pushed in call c-1 (dev.cuh Xchg)
    unsigned xtag = 0;
    size_t xbase = 0;
    if (XCH) {
      if (st.recv) {
        const bool needW = (i0 - HL < x.Istr), needE = (i0 + TX - 1 + HH > x.Iend);
        if (needW || needE) {
          const unsigned long long er = xe0 + (unsigned long long)(c - 1);
          const unsigned tag = (unsigned)er;
          const double* slot = x.box + XHDR + (er & (XSLOTS - 1)) * xslot_doubles(x.nj);
          double* rf[XF] = {f.zeta[st.rk], f.ubar[st.rk], f.vbar[st.rk], f.rzeta[st.rr]};
          const int nW = needW ? st.nrecv * SH * XNW : 0, nE = needE ? st.nrecv * SH * XNE : 0;
          for (int idx = tid; idx < nW + nE; idx += NTH) {
            const bool w = idx < nW;
            const int q = w ? idx : idx - nW, nc = w ? XNW : XNE;
            const int cc = q % nc, r = (q / nc) % SH, fld = q / (nc * SH);
            const int j = j0 - HL + r;
            if (j >= 0 && j <= Mm + 1) {
              const double* line = slot + (w ? xline_w(x.nj, fld, j, cc) : xline_e(x.nj, fld, j, cc));
              double v;
              ll_wait(line, tag, v, x.timeout_ns, x.err);
              __stcg(&rf[fld][j * P + (w ? x.Istr - XNW + cc : x.Iend + 1 + cc)], v);
            }
          }
          __syncthreads();      // every ghost value this CTA reads below was written by this CTA above
        }
      }
      if (st.send) {
        const unsigned long long e = xe0 + (unsigned long long)c;
        xtag = (unsigned)e; xbase = XHDR + (e & (XSLOTS - 1)) * xslot_doubles(x.nj);
      }
    }
    auto push = [&](int fld, int i, int j, double v) {
      if (i >= xc.Iend - (XNW - 1)) ll_store(xc.boxE + xbase + xline_w(xc.nj, fld, j, i - (xc.Iend - (XNW - 1))), v, xtag);
      if (i <= xc.Istr + (XNE - 1)) ll_store(xc.boxW + xbase + xline_e(xc.nj, fld, j, i - xc.Istr), v, xtag);
    };
    const bool xsend = XCH && st.send;

    TRACE(1);
    // ---- time-varying operands of stages 0 and 2a, requested together
    const bool okA = active && okA0;
    double zs_q = 0.0, zr_q = 0.0, rz_s = 0.0, rz_p = 0.0;
    if (st2) {
      zs_q = ldv(zs + qA); zr_q = ldv(zr + qA);
      if (!FIRST && !PRED) { rz_s = ldv(f.rzeta[st.kstp] + qA); rz_p = ldv(f.rzeta[st.ptsk] + qA); }
    }
    // ---- stages 0/1: Drhs, ubar, vbar, DUon, DVom on the staged region (:548-574)
    {
      const double* ur = f.ubar[st.krhs];
      const double* vr = f.vbar[st.krhs];
      double zv[2], uv[2], vv[2];
      bool ok[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = tid + r * NTH;
        const int a = s % SW, b = s / SW;
        const int i = i0 - HL + a, j = j0 - HL + b;
        ok[r] = (s < NS) && i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1;
        const int q = ok[r] ? (j * P + i) : (j0 * P + i0);
        zv[r] = ldv(zr + q); uv[r] = ldv(ur + q); vv[r] = ldv(vr + q);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = tid + r * NTH;
        if (s < NS) { sD[s] = ok[r] ? (zv[r] + cH[s]) : 0.0; sU[s] = ok[r] ? uv[r] : 0.0; sV[s] = ok[r] ? vv[r] : 0.0; }
      }
      __syncthreads();
      TRACE(2);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = tid + r * NTH;
        if (s < NS) {
          const int a = s % SW, b = s / SW;
          const int i = i0 - HL + a, j = j0 - HL + b;
          double du = 0.0, dv = 0.0;
          if (ok[r] && i > p.LBi && a >= 1) {
            const double cc = 0.5 * cONU[s];
            const double c1 = cc * (sD[s] + sD[s - 1]);
            du = sU[s] * c1;
          }
          if (ok[r] && j >= 1 && b >= 1) {
            const double cc = 0.5 * cOMV[s];
            const double c1 = cc * (sD[s] + sD[s - SW]);
            dv = sV[s] * c1;
          }
          sDU[s] = du; sDV[s] = dv;
        }
      }
    }
    __syncthreads();

    TRACE(3);
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

    if (active && st2) {
      const double c6 = 1.0 / 6.0;
      // ---- stage 2a: rho-point quantities at (i0-1+za, j0-1+zb)
      {
        const int i = i0 - 1 + za, j = j0 - 1 + zb;
        const int c0 = (zb + HL - 1) * SW + (za + HL - 1);
        // straight-line arithmetic for every thread of the region, results zeroed for invalid points (see k_step2d.cu)
        double Dnew, zwrk, gz, gz2, gsa;
        double a_ufx, a_vfe, c_ufx, c_vfe, k_ufx = 0.0, k_vfe = 0.0, v_ufx, v_vfe;
        {
          const double pm_q = cPM[mA], pn_q = cPN[mA], h_q = cH[c0], rS = cRS[zi], rA = cRA[zi];
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
          if (okA && za >= 1 && zb >= 1) {                               // own points of this tile
            st_r_grad(f.zeta[st.knew], j * P, i, j, zeta_new, p);
            if (PRED) st_w(f.rzeta[st.krhs], j * P, i, dd, p);
            if (xsend && (i >= x.Iend - (XNW - 1) || i <= x.Istr + (XNE - 1))) {
              push(0, i, j, zeta_new);
              if (j == 1) push(0, i, 0, zeta_new);
              if (j == Mm) push(0, i, Mm + 1, zeta_new);
              if (PRED) {                                                // rzeta has no wall-row values of its own: forward what is there
                push(3, i, j, dd);
                if (j == 1) push(3, i, 0, ldv(f.rzeta[st.krhs] + i));
                if (j == Mm) push(3, i, Mm + 1, ldv(f.rzeta[st.krhs] + (Mm + 1) * P + i));
              }
            }
          }
          // advective UFx at rho(i,j) (:1104-1112)
          a_ufx = 0.25 * (U_(0, 0) + U_(1, 0) - c6 * (GXU(0, 0) + GXU(1, 0))) * (DU_(0, 0) + DU_(1, 0) - c6 * (GXDU(0, 0) + GXDU(1, 0)));
          // advective VFe at rho(i,j) (:1263-1272)
          {
            const int da = (j < 2) ? 1 : 0, db = (j + 1 > Mm) ? 0 : 1;
            a_vfe = 0.25 * (V_(0, 0) + V_(0, 1) - c6 * (GYV(0, da) + GYV(0, db))) * (DV_(0, 0) + DV_(0, 1) - c6 * (GYDV(0, da) + GYDV(0, db)));
          }
          // Coriolis (:1291-1300) and curvilinear (:1333-1347) at rho(i,j)
          const double D0 = D_(0, 0);
          const double vS = V_(0, 0) + V_(0, 1), uS = U_(0, 0) + U_(1, 0);
          {
            const double cc = 0.5 * D0 * cFOMN[zi];
            c_ufx = cc * vS; c_vfe = cc * uS;
          }
          if (p.curvgrid) {
            const double c1 = 0.5 * vS, c2 = 0.5 * uS;
            const double cc = D0 * (c1 * cDNDX[zi] - c2 * cDMDE[zi]);
            k_ufx = cc * c1; k_vfe = cc * c2;
          }
          // viscous stress at rho(i,j) (:1400-1414)
          {
            const double pnE_a = cPN[mA + 1], pnW_a = cPN[mA - 1], pmN_a = cPM[mA + MW], pmS_a = cPM[mA - MW];
            const double cr = cVISR[zi] * D0 * 0.5 *
                              (cPMONR[zi] * ((pn_q + pnE_a) * U_(1, 0) - (pnW_a + pn_q) * U_(0, 0)) -
                               cPNOMR[zi] * ((pm_q + pmN_a) * V_(0, 1) - (pmS_a + pm_q) * V_(0, 0)));
            const double onr = cONR[zi], omr = cOMR[zi];
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
        const int mP = (zb + 2) * MW + (za + 2);                         // pm / pn index of (i, j)
        const bool okB = j >= 1 && j <= Mm + 1 && i <= p.Iend + 1;
        double a_ufe, a_vfx, v_ufe, v_vfx;
        {
          const double pn_q = cPN[mP], pnS = cPN[mP - MW], pnW = cPN[mP - 1], pnSW = cPN[mP - MW - 1];
          const double pm_q = cPM[mP], pmS = cPM[mP - MW], pmW = cPM[mP - 1], pmSW = cPM[mP - MW - 1];
          // advective UFe at psi(i,j) (:1141-1150)
          {
            const int d0 = (j > Mm) ? -1 : 0, dm = (j - 1 < 1) ? 0 : -1;
            a_ufe = 0.25 * (U_(0, 0) + U_(0, -1) - c6 * (GYU(0, d0) + GYU(0, dm))) * (DV_(0, 0) + DV_(-1, 0) - c6 * (GXDV(0, 0) + GXDV(-1, 0)));
          }
          // advective VFx at psi(i,j), j = 2..Mm (:1213-1222)
          a_vfx = 0.25 * (V_(0, 0) + V_(-1, 0) - c6 * (GXV(0, 0) + GXV(-1, 0))) * (DU_(0, 0) + DU_(0, -1) - c6 * (GYDU(0, 0) + GYDU(0, -1)));
          // viscous stress at psi(i,j) (:1394-1430)
          {
            const double Dp = 0.25 * (D_(0, 0) + D_(-1, 0) + D_(0, -1) + D_(-1, -1));
            const double cp = cVISP[zi] * Dp * 0.5 *
                              (cPMONP[zi] * ((pnS + pn_q) * V_(0, 0) - (pnSW + pnW) * V_(-1, 0)) +
                               cPNOMP[zi] * ((pmW + pm_q) * U_(0, 0) - (pmSW + pmS) * U_(0, -1)));
            const double omp = cOMP[zi], onp = cONP[zi];
            v_ufe = omp * omp * cp; v_vfx = onp * onp * cp;
          }
        }
        aUFe[zi] = okB ? a_ufe : 0.0; aVFx[zi] = (okB && j >= 2 && j <= Mm) ? a_vfx : 0.0; vUFe[zi] = okB ? v_ufe : 0.0; vVFx[zi] = okB ? v_vfx : 0.0;
      }
    }
    // ---- stage 3: its time-varying operands are requested before the barrier that ends stage 2
    const int i = i3, j = j3;
    const bool inner = active && j >= 1 && j <= Mm;
    const bool dov = inner && (j >= p.JstrV);
    double zr_o = 0.0, zs0 = 0.0, zsW = 0.0, zsS = 0.0, us = 0.0, vs = 0.0;
    double rub_s = 0.0, rub_p = 0.0, rvb_s = 0.0, rvb_p = 0.0, ru_n = 0.0, ru_so = 0.0, rv_n = 0.0, rv_so = 0.0;
    if (live) {
      if (PRED && !FIRST) zr_o = ldv(zr + o);
      if (inner) {
        zs0 = ldv(zs + o); zsW = ldv(zs + o - 1); zsS = ldv(zs + oS);
        us = ldv(f.ubar[st.kstp] + o); vs = ldv(f.vbar[st.kstp] + o);
        if (!FIRST && !PRED) { rub_s = ldv(f.rubar[st.kstp] + o); rub_p = ldv(f.rubar[st.ptsk] + o); rvb_s = ldv(f.rvbar[st.kstp] + o); rvb_p = ldv(f.rvbar[st.ptsk] + o); }
        if (FIRST && PRED && p.istart >= 1) { ru_n = ldv(f.ru[p.nnew] + o); rv_n = ldv(f.rv[p.nnew] + o); ru_so = ldv(f.ru[p.nstp] + o); rv_so = ldv(f.rv[p.nstp] + o); }
      }
    }
    __syncthreads();
    TRACE(4);
    if (live) {
      // fast-time averages (:614-682); rows 0..Mm+1 for Zt/DU, rows 1..Mm+1 for DV.  Kept in registers, stored by the last call.
      {
        const double DUo = sDU[a3], DVo = sDV[a3];
        if (PRED) {
          if (FIRST) {
            const double cff2 = (-1.0 / 12.0) * st.w2_p1;
            av_zt = 0.0; av_du1 = 0.0; av_dv1 = 0.0;
            av_du2 = cff2 * DUo;
            if (j >= 1) av_dv2 = cff2 * DVo;
          } else {
            const double cff1 = st.w1_m1;
            const double cff2 = (8.0 / 12.0) * st.w2_0 - (1.0 / 12.0) * st.w2_p1;
            av_zt = av_zt + cff1 * zr_o;
            av_du1 = av_du1 + cff1 * DUo;
            av_du2 = av_du2 + cff2 * DUo;
            if (j >= 1) {
              av_dv1 = av_dv1 + cff1 * DVo;
              av_dv2 = av_dv2 + cff2 * DVo;
            }
          }
        } else {
          const double cff2 = FIRST ? st.w2_0 : (5.0 / 12.0) * st.w2_0;
          av_du2 = av_du2 + cff2 * DUo;
          if (j >= 1) av_dv2 = av_dv2 + cff2 * DVo;
        }
        if (LAST) {
          st_w(f.Zt_avg1, j * P, i, av_zt, p);
          st_w(f.DU_avg1, j * P, i, av_du1, p);
          f.DU_avg2[o] = av_du2;
          if (j >= 1) { st_w(f.DV_avg1, j * P, i, av_dv1, p); f.DV_avg2[o] = av_dv2; }
        }
      }
      if (inner) {
        const double h0 = cH[a3], hW = cH[a3 - 1], hS = cH[a3S];
        const double rA0 = cRA[z0], rAW = cRA[zW], rAS = cRA[zS];
        const double pm0 = cPM[m3], pmW = cPM[m3 - 1], pmS = cPM[m3S], pn0 = cPN[m3], pnW = cPN[m3 - 1], pnS = cPN[m3S];
        const double onu = cONU[a3], omv = cOMV[a3];
        // ---- u-point (i,j) and v-point (i,j): arithmetic of both first (two independent chains), stores afterwards
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
        // coupling with the 3-D equations (:1884-2065)
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
        if (FIRST && PRED) { f.rufrc[o] = rf_u; rufrc_o = rf_u; f.ru[p.nstp][o] = rf_u; }
        st_u_closed(f.ubar[st.knew], j * P, i, j, xu, p);
        if (PRED) f.rubar[st.krhs][o] = rhs_u;
        if (xsend && (i >= xc.Iend - (XNW - 1) || i <= xc.Istr + (XNE - 1))) {
          push(1, i, j, xu);
          if (j == 1) push(1, i, 0, p.gamma2 * xu);
          if (j == Mm) push(1, i, Mm + 1, p.gamma2 * xu);
        }
        if (dov) {
          if (FIRST && PRED) { f.rvfrc[o] = rf_v; rvfrc_o = rf_v; f.rv[p.nstp][o] = rf_v; }
          st_v_closed(f.vbar[st.knew], j * P, i, j, xv, p);
          if (PRED) f.rvbar[st.krhs][o] = rhs_v;
          if (xsend && (i >= xc.Iend - (XNW - 1) || i <= xc.Istr + (XNE - 1))) {
            push(2, i, j, xv);
            if (j == 2) { push(2, i, 1, 0.0); push(2, i, 0, ldv(f.vbar[st.knew] + i)); }   // row 1 is the wall (v = 0); row 0 is never written
            if (j == Mm) push(2, i, Mm + 1, 0.0);
          }
        }
      }
    }
    // ---- publish: every store of this call is visible before the flag
    TRACE(5);
    __syncthreads();
    TRACE(6);
    if (tid == 0) st_release_u64(ctl.flags + (bx + nbx * by), base + (unsigned long long)c);   // release: cumulative over the barrier
    TRACE(7);
  }
  // ---- the CTA that finishes last advances the flag base and the exchange epoch for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctl.base[1], 1ULL) == (unsigned long long)nbx * nby - 1) {
      ctl.base[1] = 0ULL;
      ctl.base[0] = base + (unsigned long long)ctl.ncall + 1ULL;
      if (XCH && x.box != nullptr) { unsigned long long* hdr = (unsigned long long*)x.box; hdr[0] = xe0 + (unsigned long long)ctl.nsend; }
      __threadfence();
    }
  }
}

// Number of CTAs the loop kernel would need for this tile, or 0 when the tile is too large for every CTA to be resident at once
// (then the caller uses one k_step2d launch per call).
int step2d_loop_ctas(const Par& p) {
  static int cap[MAXDEV] = {0};
  int& c = cap[cur_dev()];
  if (c == 0) {
    const size_t smem = (size_t)SMEM_DOUBLES * sizeof(double);
    int dev = 0, nsm = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    const cudaError_t e1 = cudaFuncSetAttribute(k_step2d_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const cudaError_t e2 = cudaFuncSetAttribute(k_step2d_loop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 == cudaSuccess && e2 == cudaSuccess && coop &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_step2d_loop<true>, NTH, smem) == cudaSuccess && per_sm >= 1)
      c = nsm * per_sm;
    else { cudaGetLastError(); c = -1; }
  }
  const int nb = ((p.Iend - p.Istr + 1 + TX - 1) / TX) * ((p.Mm + TY - 1) / TY);
  return (c > 0 && nb <= c) ? nb : 0;
}

// Whole barotropic loop in one cooperative launch.  false: not launched (caller falls back to per-call launches).
bool launch_step2d_loop(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x, const LoopCtl& ctl) {
  if (!step2d_loop_ctas(p)) return false;
  dim3 g((p.Iend - p.Istr + 1 + TX - 1) / TX, (p.Mm + TY - 1) / TY);
  const size_t smem = (size_t)SMEM_DOUBLES * sizeof(double);
  Xchg xx;
  std::memset(&xx, 0, sizeof(xx));
  const bool xch = x && x->box;
  if (xch) xx = *x;
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = g; cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const cudaError_t e = xch ? cudaLaunchKernelEx(&cfg, k_step2d_loop<true>, p, f, xx, ctl) : cudaLaunchKernelEx(&cfg, k_step2d_loop<false>, p, f, xx, ctl);
  if (e != cudaSuccess) { std::fprintf(stderr, "roms_b200: cooperative launch of k_step2d_loop failed: %s\n", cudaGetErrorString(e)); return false; }
  return true;
}

}  // namespace rb

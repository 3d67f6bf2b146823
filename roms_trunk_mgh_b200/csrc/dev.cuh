// Device-side common definitions for the B200 (sm_100a) implementation of the ROMS nonlinear baroclinic step.
//
// HBM layout: every 2-D / 3-D field of one tile is stored i-fastest with a common row pitch P (doubles) and plane
// stride PL = P * nj; the pointer kept in Flds is pre-offset so that A[i + j*P + k*PL] addresses Fortran element
// A(i,j,k) directly.  P is a multiple of 16 doubles and the origin is shifted so that i = Istr sits on a 128-byte
// boundary: warps that walk the xi axis issue fully coalesced 256-byte requests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {

constexpr int MAXNT = 2;
constexpr int MAXN = 64;          // upper bound on vertical levels for thread-private column arrays

struct Par {
  // sizes / layout
  int Lm, Mm, N, NT;
  int P, PL;                      // row pitch, plane stride (doubles)
  int LBi, UBi, LBj, UBj;
  // tile ranges (get_bounds.F var_bounds; EW periodic, NS closed, NtileJ == 1)
  int Istr, Iend, Jstr, Jend;     // Jstr = 1, Jend = Mm
  int IstrU, JstrV;               // IstrU = Istr (periodic), JstrV = 2
  int JstrR, JendR;               // 0, Mm+1
  int Jstrm1, Jendp1, Jendp2, JstrVm1, JstrVm2;
  int ew_wrap;                    // 1: this tile owns the whole xi range -> periodic ghosts are filled by the producer
  // Split launches (multi-GPU overlap): a launch covers columns Istr..Iend minus a gap of gap_len columns that starts
  // gap_at columns after Istr.  Full launch: gap_len = 0.  "Edge" launch: the first and last EDGE_W columns of the tile.
  int gap_at, gap_len;
  // stepping
  int nstp, nnew, nrhs;           // 1..2
  int istart;                     // 0: iic == ntfirst, 1: iic == ntfirst+1, 2: later
  int iif, kstp, krhs, knew, ptsk, predictor, nfast;
  // options
  int nonlin_eos, curvgrid, uv_qdrag, salinity, hadv, vadv, itemp, isalt;
  int bv_frequency, eos_tderivative, solar_source, lmd_nonlocal;   // optional terms of rho_eos / pre_step3d (roms_b200_config)
  int bulk_fluxes, lmd_mixing;    // BULK_FLUXES / LMD_MIXING phases are part of the step (k_physics.cu)
  int fuse_tmix;                  // pre_step3d_t also applies t3dmix2_s (whole-step path only; 0 for single-phase calls)
  // scalars
  double dt, dtfast, g, rho0, R0, T0, S0, Tcoef, Scoef, gamma2, lambda, hc;
  double Akv_bak, Akt_bak[MAXNT];
  double w1_m1, w2_0, w2_p1;      // weight(1,iif-1), weight(2,iif), weight(2,iif+1) for the current step2d call
  double blk_ZQ, blk_ZT, blk_ZW;  // heights (m) of the atmospheric humidity / temperature / wind data (bulk_flux.F)
  int uv_adv;                     // rhs3d momentum advection: 0 default (U3 / C4), 1 UV_C4ADVECTION, 2 UV_SADVECTION, 3 UV_C2ADVECTION (also in step2d)
  int limit_bstress;              // LIMIT_BSTRESS in set_vbc
  int nospl_vvisc, nospl_vdiff;   // 1: SPLINES_VVISC / SPLINES_VDIFF not defined (centred implicit systems in step3d_uv / step3d_t)
  int qcorrection, limit_stflx_cooling, scorrection, pad2_;   // set_vbc surface-flux corrections (set_vbc.F:285-351)
  double Tnudg_salt;
  int bodyforce, levsfrc, levbfrc, vtransform;
  int atm_press, pad4_;           // ATM_PRESS: Pair in the pressure gradient   // BODYFORCE and its level ranges (mod_scalars.F levsfrc, levbfrc)
};

// Field table (all pointers pre-offset; [0] slots of time-indexed arrays are unused so Fortran indices apply)
struct Flds {
  // grid
  double *h, *f, *pm, *pn, *om_r, *on_r, *om_u, *on_u, *om_v, *on_v, *om_p, *on_p, *omn, *fomn, *pmon_r, *pnom_r, *pmon_u, *pnom_u,
      *pmon_v, *pnom_v, *pmon_p, *pnom_p, *dndx, *dmde, *rdrag, *rdrag2, *visc2_r, *visc2_p, *ZoBot;
  double* diff2[MAXNT];
  // 2-D state
  double *zeta[4], *ubar[4], *vbar[4], *rzeta[3], *rubar[3], *rvbar[3];
  double *Zt_avg1, *DU_avg1, *DU_avg2, *DV_avg1, *DV_avg2, *rufrc, *rvfrc, *rhoA, *rhoS, *sustr, *svstr, *bustr, *bvstr;
  double *stflx[MAXNT], *btflx[MAXNT], *stflux[MAXNT], *btflux[MAXNT];
  // 3-D state
  double *u[3], *v[3], *ru[3], *rv[3];
  double* t[4][MAXNT];
  double *rho, *pden, *Hz, *z_r, *Huon, *Hvom, *W, *wvel, *z_w, *Akv;
  double* Akt[MAXNT];
  // scratch
  double* P3;                      // prsgrd32 pressure (1:N)
  // optional: rho_eos outputs bvf (0:N), alpha, beta; pre_step3d inputs srflx, Jwtype, ghats (0:N)
  double *bvf, *alpha, *beta, *srflx, *Jwtype;
  double* ghats[MAXNT];
  // optional: atmosphere read by bulk_flux, its flux outputs, wind stress scratch at rho points; KPP boundary layer depth / index
  double *Uwind, *Vwind, *Tair, *Pair, *Hair, *rain, *cloud, *lrflx, *lhflx, *shflx, *Taux, *Tauy, *hsbl, *ksbl;
  // time-averaged fields (mod_average.F; allocated by roms_b200_set_avg)
  double *avgzeta, *avgu2d, *avgv2d, *avgu3d, *avgv3d, *avgrho, *avgw3d, *avgwvel;
  double* avgt[MAXNT];
  // 1-D (device)
  double *sc_r, *Cs_r, *sc_w, *Cs_w;
  // TS_DIF4: MIXING%diff4 = SQRT(ABS(tnu4)) (t3dmix4_s.h)
  double* diff4[MAXNT];
  // QCORRECTION / SCORRECTION / SRELAXATION data (mod_forces.F)
  double *sst, *dqdt, *sss;
};

constexpr int EDGE_W = 64;        // width of the tile edges computed ahead of the halo exchange (multiple of every CTA width)
// First column of the CTA whose columns start `off` columns into the launch (off is a multiple of the CTA width).
__device__ __forceinline__ int xcol0(const Par& p, int off) { return p.Istr + off + ((off >= p.gap_at) ? p.gap_len : 0); }
// Number of columns a launch covers (host side: grid sizing).
__host__ __device__ inline int xspan(const Par& p) { return p.Iend - p.Istr + 1 - p.gap_len; }

// ---- stores that also fill the periodic (xi) ghost images when this tile wraps onto itself -----------------------
// exchange_2d.F / exchange_3d.F: A(Lm+1:Lm+2) = A(1:2), A(-2:0) = A(Lm-2:Lm)
__device__ __forceinline__ void st_w(double* __restrict__ A, int o, int i, double x, const Par& p) {
  A[o + i] = x;
  if (p.ew_wrap) {
    if (i <= 2) A[o + i + p.Lm] = x;
    if (i >= p.Lm - 2) A[o + i - p.Lm] = x;
  }
}
// rho-type with zero-gradient closed walls (zetabc.F:536-545/685-694, t3dbc_im.F:477-489/611-623, bc_r2d/bc_w3d):
// o = offset of row j (without i)
__device__ __forceinline__ void st_r_grad(double* __restrict__ A, int o, int i, int j, double x, const Par& p) {
  st_w(A, o, i, x, p);
  if (j == 1) st_w(A, o - p.P, i, x, p);
  if (j == p.Mm) st_w(A, o + p.P, i, x, p);
}
// u-type with closed walls south/north: A(i,0) = gamma2*A(i,1), A(i,Mm+1) = gamma2*A(i,Mm) (u2dbc_im.F:963-979, u3dbc_im.F:507-529)
__device__ __forceinline__ void st_u_closed(double* __restrict__ A, int o, int i, int j, double x, const Par& p) {
  st_w(A, o, i, x, p);
  if (j == 1) st_w(A, o - p.P, i, p.gamma2 * x, p);
  if (j == p.Mm) st_w(A, o + p.P, i, p.gamma2 * x, p);
}
// v-type with closed walls: A(i,1) = 0, A(i,Mm+1) = 0 (v2dbc_im.F:436-441/785-790, v3dbc_im.F:222-230/364-372);
// called for j in 2..Mm
__device__ __forceinline__ void st_v_closed(double* __restrict__ A, int o, int i, int j, double x, const Par& p) {
  st_w(A, o, i, x, p);
  if (j == 2) st_w(A, o - p.P, i, 0.0, p);
  if (j == p.Mm) st_w(A, o + p.P, i, 0.0, p);
}

// Fire-and-forget prefetch of the line holding *p into L2: needs no destination register, so a thread can have many
// in flight.  The latency-bound column / multi-stage kernels use it to start the DRAM fetch of later stages early.
__device__ __forceinline__ void pf_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int PFD = 4;            // prefetch distance (levels ahead) of the column-marching kernels
// L2 prefetch of the element D levels above a[o] (one lane per 128-byte line issues it): the column-marching kernels keep
// one or two levels of operands in registers, which does not put enough bytes in flight to cover the DRAM latency; the
// prefetch turns the later register loads into L2 hits without costing registers.
template <int D>
__device__ __forceinline__ void pf_up(const double* a, int o, int k, int N, int PL) {
  if (D > 0) { if ((threadIdx.x & 15) == 0 && k + D <= N) pf_l2(a + o + D * PL); }
}

// Asynchronous 8-byte global -> shared copy (LDGSTS): no destination register, completion via wait_group.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int D>
__device__ __forceinline__ void pf_dn(const double* a, int o, int k, int PL) {
  if (D > 0) { if ((threadIdx.x & 15) == 0 && k - D >= 1) pf_l2(a + o - D * PL); }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Software-pipelined upward sweep over the levels of a column: level k+1 is requested before level k is computed.
// body(cur, nxt, k): nxt is level k+1, or a copy of level N when k == N.
// PINGPONG: the two level buffers swap roles every iteration (unrolled by two), which removes the register-to-register copy
// of a whole level per iteration but keeps two bodies' worth of values live; measured per kernel (profiles/README.md), it
// pays where the register budget has room (t3dmix2_s, step3d_uv) and costs spills where it has not (step3d_t, uv3dmix2).
template <bool PINGPONG, class Load, class Body>
__device__ __forceinline__ void sweep_levels(int N, Load load, Body body) {
  if (!PINGPONG) {
    auto cur = load(1);
    for (int k = 1; k <= N; ++k) {
      auto nxt = cur;
      if (k < N) nxt = load(k + 1);
      body(cur, nxt, k);
      cur = nxt;
    }
  } else {
    auto A = load(1);
    int k = 1;
    for (; k + 1 <= N; k += 2) {
      auto B = load(k + 1);
      body(A, B, k);
      if (k + 2 <= N) A = load(k + 2); else A = B;
      body(B, A, k + 1);
    }
    if (k <= N) body(A, A, k);
  }
}

// The same sweep with D levels in flight (a ring of D level buffers, unrolled by D): for kernels whose levels are small, so
// that one level ahead does not put enough bytes in flight to cover the DRAM latency.
template <int D, class Load, class Body>
__device__ __forceinline__ void sweep_levels_deep(int N, Load load, Body body) {
  decltype(load(1)) buf[D];
#pragma unroll
  for (int q = 0; q < D; ++q) buf[q] = load((1 + q <= N) ? 1 + q : N);
  for (int k = 1; k <= N; k += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int kk = k + q;
      if (kk <= N) {
        body(buf[q], (kk < N) ? buf[(q + 1) % D] : buf[q], kk);
        if (kk + D <= N) buf[q] = load(kk + D);
      }
    }
  }
}

// ---- halo exchange fused into the barotropic sub-step kernel (multi-GPU ring over NVLink peer memory) ----------------
// One exchange = the edge columns of up to XF 2-D fields (3 columns eastward, 2 westward, every row).  A sub-step kernel
// PUSHES the edge values it produces straight into the neighbours' mailboxes as it stores them, and the NEXT sub-step
// kernel starts by PULLING what its neighbours pushed into the ghost columns of its own arrays: no separate pack /
// exchange / unpack kernels between the ~58 dependent sub-steps of a baroclinic step.  Every double travels as a 16-byte
// line {lo, tag, hi, tag} (tag = low 32 bits of the exchange epoch; the flag-in-data scheme of NCCL's LL protocol, which
// only relies on 8-byte store atomicity), so no fences or separate flags are needed.  XSLOTS epochs are kept apart.
constexpr int XNW = 3, XNE = 2, XF = 4, XSLOTS = 4, XHDR = 16;     // XHDR doubles of header: [0] epoch, [1] CTA counter, [2] error
struct Xchg {
  unsigned long long* err;         // sticky device error word of the tile (state.h d_err): set when a wait gives up
  long long timeout_ns;            // how long a pull waits for the neighbour (<= 0: for ever)
  int send, recv;                  // push this sub-step's edge columns / first pull the previous sub-step's
  int nsend, nrecv;                // number of fields (3: zeta, ubar, vbar; 4: + rzeta)
  int Istr, Iend, nj;              // tile bounds (a split launch may cover only part of them), rows per column
  double* recvf[XF];               // arrays whose ghost columns the pull fills
  double* boxE; double* boxW;      // east / west neighbour's mailbox (peer mapped)
  double* box;                     // my mailbox; its header holds the epoch counter
};
__host__ __device__ inline size_t xslot_doubles(int nj) { return (size_t)2 * XF * nj * (XNW + XNE); }
__host__ __device__ inline size_t xbox_doubles(int nj) { return XHDR + XSLOTS * xslot_doubles(nj); }
// line of (field, row, column c) in the half of a slot filled by the west (c < XNW) / east (c < XNE) neighbour
__device__ __forceinline__ size_t xline_w(int nj, int fld, int j, int c) { return 2 * (size_t)((fld * nj + j) * XNW + c); }
__device__ __forceinline__ size_t xline_e(int nj, int fld, int j, int c) { return 2 * (size_t)(XF * nj * XNW + (fld * nj + j) * XNE + c); }

__device__ __forceinline__ void ll_store(double* line, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ll_load(const double* line, unsigned tag, double& v) {
  unsigned lo, t0, hi, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(line) : "memory");
  v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  return t0 == tag && t1 == tag;
}

// Wall-clock nanoseconds (independent of the SM clock): bounds the halo waits.
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// Spin until the line carries `tag`.  A neighbour that never delivers (a rank that died, a host that stopped stepping) ends
// the wait after timeout_ns: the sticky error word is set -- every host synchronisation point of the library turns it into
// exit_flag 8 -- and the value is whatever the line held.  false: timed out.
__device__ __forceinline__ bool ll_wait(const double* line, unsigned tag, double& v, long long timeout_ns, unsigned long long* err) {
  if (ll_load(line, tag, v)) return true;
  const long long t0 = gtime_ns();
  for (;;) {
#pragma unroll 1
    for (int q = 0; q < 64; ++q) { if (ll_load(line, tag, v)) return true; __nanosleep(20); }
    if (timeout_ns > 0 && gtime_ns() - t0 > timeout_ns) { *err = 1ULL; return false; }
  }
}

// ---- the whole barotropic loop as one persistent kernel (k_step2d_loop.cu) -----------------------------------------
// One entry per step2d call of LOOP_2D (main3d.F:592-700): the 2-D time indices of the call, its filter weights, and what the
// fused halo exchange does in it (rk / rr: time levels of zeta/ubar/vbar and rzeta whose ghost columns the pull fills).
struct LoopStep {
  int iif, kstp, krhs, knew, ptsk, predictor;
  int send, recv, nrecv, rk, rr, pad_;
  double w1_m1, w2_0, w2_p1;
};
struct LoopCtl {
  const LoopStep* steps; int ncall, nsend;
  unsigned long long* flags;       // one completion counter per CTA
  unsigned long long* base;        // [0] flag value of "call 0" of this launch (advanced by the kernel), [1] CTAs finished
  unsigned long long* err; long long timeout_ns;
};

// Function attributes (opt-in dynamic shared memory) are per device: launch wrappers remember what they set per device, so
// a process that drives several tiles on several GPUs (one handle each) gets them on every one.
constexpr int MAXDEV = 64;
inline int cur_dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < MAXDEV) ? d : 0; }

__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }   // Fortran MAX (first arg on ties)
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }

}  // namespace rb

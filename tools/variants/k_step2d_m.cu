// step2d_tile (ROMS/Nonlinear/step2d_LF_AM3.h:137-2528), row-marching form: one barotropic LF-AM3 predictor or corrector
// sub-step as a single kernel in which a thread owns one xi column and marches along eta.
//
// The tile kernel (k_step2d.cu) is bound by the L1 / shared-memory data pipe: ~250 eight-byte accesses per point, most of
// them re-reads of values a neighbouring thread produced.  Here a CTA owns TXM columns and a strip of JL rows and sweeps
// the strip row by row with the three stages of the sub-step skewed in eta,
//     iteration s :  A  Drhs, ubar, vbar, DUon, DVom of row s                    (:548-574)
//                    B  zeta_new, rho-point fluxes of row s-2; psi-point fluxes of row s-1   (:770-851, :1104-1430)
//                    C  averages, pressure gradient, flux divergences, coupling, time stepping of row s-2 (:614-682, :944-2430)
// so that everything a column needs from its own eta neighbours (row s-3 of stage B, the psi row of the previous
// iteration, h / rhoA / zeta of the row below) is still in that thread's registers, every global operand is loaded once
// per point (no eta halo re-loads, the operands stage B and stage C share are loaded once), and shared memory only carries
// what xi neighbours exchange: a 4-row ring of the stage-A fields and one row of 11 stage-B fluxes.  ~130 accesses per
// point.  Per-point arithmetic is expression-for-expression that of k_step2d.cu (and of the oracle), so results are
// bit-identical in the strict build.
#include "dev.cuh"
#include "kernels.h"

#include <cstdlib>

namespace rb {

#ifndef S2M_MINB
#define S2M_MINB 2          // resident CTAs per SM the 128-column instance is compiled for (register budget)
#endif
#ifndef S2M_MINB64
#define S2M_MINB64 4        // same for the 64-column instance
#endif

template <int TXM>
__global__ void __launch_bounds__(TXM + 32, (TXM == 128) ? S2M_MINB : S2M_MINB64) k_step2d_m(Par p, Flds f, int JL) {
  constexpr int CW = TXM + 5;                     // staged columns i0-3 .. i0+TXM+1
  extern __shared__ double smem[];
  double* sD = smem;                              // [4][CW] rings, slot = row & 3
  double* sU = sD + 4 * CW; double* sV = sU + 4 * CW; double* sDU = sV + 4 * CW; double* sDV = sDU + 4 * CW;
  double* xG = sDV + 4 * CW;                      // [CW] each: rho-point values of the current stage-B row, read by column i+1
  double* xGSA = xG + CW; double* xZw = xGSA + CW; double* xG2 = xZw + CW; double* xDnew = xG2 + CW;
  double* xaUFx = xDnew + CW; double* xcUFx = xaUFx + CW; double* xkUFx = xcUFx + CW; double* xvUFx = xkUFx + CW;
  double* yaVFx = xvUFx + CW;                     // [2][CW]: psi-point values (parity of the row), read by column i-1 one iteration later
  double* yvVFx = yaVFx + 2 * CW;

  const int tid = threadIdx.x;
  const int i0 = xcol0(p, blockIdx.x * TXM);
  const int j0 = blockIdx.y * JL;
  const int P = p.P, Mm = p.Mm;
  const int j1 = min(j0 + JL - 1, Mm + 1);
  // column of this thread: main threads own i0 .. i0+TXM-1; five lanes of the extra warp carry the halo columns
  const bool isMain = tid < TXM;
  const int hl = tid - TXM;                       // halo lane
  const int a = isMain ? tid + 3 : (hl < 3 ? hl : TXM + hl);          // staged column index (0 .. CW-1)
  const bool act = isMain || hl < 5;
  const int i = i0 - 3 + a;
  const bool doRho = isMain || hl == 2;           // rho column i0-1
  const bool doPsi = isMain || hl == 3;           // psi column i0+TXM
  const bool colok = act && i >= p.LBi && i <= p.UBi;

  const bool PRED = p.predictor != 0;
  const bool FIRST = (p.iif == 1);
  const bool active = (p.iif <= p.nfast);         // :755 (the nfast+1-th call only averages)
  const double* __restrict__ h = f.h;
  const double* __restrict__ zr = f.zeta[p.krhs];
  const double* __restrict__ zs = f.zeta[p.kstp];
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;
  const double* __restrict__ ur = f.ubar[p.krhs];
  const double* __restrict__ vr = f.vbar[p.krhs];
  const int qsafe = j0 * P + i0;
  const double c6 = 1.0 / 6.0;

  // values carried from one iteration to the next (own column)
  double Dprev = 0.0;                                                                 // Drhs of row s-1
  double zS_G = 0.0, zS_GSA = 0.0, zS_Zw = 0.0, zS_G2 = 0.0, zS_Dnew = 0.0;           // rho-point values of row j-1
  double zS_aVFe = 0.0, zS_cVFe = 0.0, zS_kVFe = 0.0, zS_vVFe = 0.0;
  double p0_aUFe = 0.0, p0_aVFx = 0.0, p0_vUFe = 0.0, p0_vVFx = 0.0;                  // psi-point values of row j
  double hS = 0.0, rAS = 0.0, zsS = 0.0, pnS = 0.0;                                   // h, rhoA, zeta(kstp), pn of row j-1

  for (int s = j0 - 3; s <= j1 + 2; ++s) {
    // ================= stage A: row s of Drhs, ubar, vbar, DUon, DVom ================================================
    {
      const bool ok = colok && s >= 0 && s <= Mm + 1;
      const bool okW = ok && i > p.LBi;
      const int q = ok ? (s * P + i) : qsafe;
      const int qW = okW ? q - 1 : q;
      const double zv = zr[q], hv = h[q], uv = ur[q], vv = vr[q], onu = f.on_u[q], omv = f.om_v[q];
      const double zw = zr[qW], hw = h[qW];
      const double D = ok ? (zv + hv) : 0.0;
      const double U = ok ? uv : 0.0, V = ok ? vv : 0.0;
      double du = 0.0, dv = 0.0;
      if (okW) {
        const double Dw = zw + hw;
        const double c = 0.5 * onu;
        const double c1 = c * (D + Dw);
        du = U * c1;
      }
      if (ok && s >= 1 && s > j0 - 3) {
        const double c = 0.5 * omv;
        const double c1 = c * (D + Dprev);
        dv = V * c1;
      }
      Dprev = D;
      if (act) {
        const int w = ((s + 8) & 3) * CW + a;
        sD[w] = D; sU[w] = U; sV[w] = V; sDU[w] = du; sDV[w] = dv;
      }
    }
    __syncthreads();

    const int r = s - 2;                          // rho row of stage B == output row of stage C
    const int rp = s - 1;                         // psi row of stage B
    const int o = r * P + i;
    const bool rowB = active && r >= 1 && r <= Mm && r >= j0 - 1 && r <= j1;
    const bool rhoOn = doRho && rowB && i <= p.Iend;
    const bool psiOn = doPsi && active && rp >= 1 && rp <= Mm + 1 && rp >= j0 && rp <= j1 + 1 && i <= p.Iend + 1;
    const bool outOn = isMain && r >= j0 && r <= j1 && i <= p.Iend;          // r <= j1 <= Mm+1
    const bool inner = outOn && active && r >= 1 && r <= Mm;
    const bool dov = inner && (r >= p.JstrV);

    // ---- global operands of stages B and C for this iteration, issued back to back ------------------------------------
    // shared by B (rho) and C
    double zs_q = 0.0, zr_q = 0.0, pm_q = 0.0, pn_q = 0.0, h_q = 0.0, rS = 0.0, rA = 0.0, pnW = 0.0, pmS = 0.0;
    // B (rho) only
    double fomn_q = 0.0, visc_q = 0.0, pmon_q = 0.0, pnom_q = 0.0, pnE = 0.0, pmN = 0.0, onr = 0.0, omr = 0.0, dndx_q = 0.0, dmde_q = 0.0, rz_s = 0.0, rz_p = 0.0;
    if (rhoOn) {
      zs_q = zs[o]; zr_q = zr[o]; pm_q = pm[o]; pn_q = pn[o]; h_q = h[o]; rS = f.rhoS[o]; rA = f.rhoA[o];
      fomn_q = f.fomn[o]; visc_q = f.visc2_r[o]; pmon_q = f.pmon_r[o]; pnom_q = f.pnom_r[o];
      pnE = pn[o + 1]; pnW = pn[o - 1]; pmN = pm[o + P]; pmS = pm[o - P]; onr = f.on_r[o]; omr = f.om_r[o];
      if (p.curvgrid) { dndx_q = f.dndx[o]; dmde_q = f.dmde[o]; }
      if (!FIRST && !PRED) { rz_s = f.rzeta[p.kstp][o]; rz_p = f.rzeta[p.ptsk][o]; }
    } else if (outOn) {
      zr_q = zr[o];                               // rows 0 and Mm+1 (and the averaging-only call): Zt_avg1 needs zeta(krhs)
    }
    // B (psi)
    double pvisc = 0.0, ppmon = 0.0, ppnom = 0.0, omp = 0.0, onp = 0.0, qn0 = 0.0, qnS = 0.0, qnW = 0.0, qnSW = 0.0, qm0 = 0.0, qmS = 0.0, qmW = 0.0, qmSW = 0.0;
    if (psiOn) {
      const int q = rp * P + i;
      pvisc = f.visc2_p[q]; ppmon = f.pmon_p[q]; ppnom = f.pnom_p[q]; omp = f.om_p[q]; onp = f.on_p[q];
      qn0 = pn[q]; qnS = pn[q - P]; qnW = pn[q - 1]; qnSW = pn[q - P - 1];
      qm0 = pm[q]; qmS = pm[q - P]; qmW = pm[q - 1]; qmSW = pm[q - P - 1];
    }
    // C only
    double av_du2 = 0.0, av_dv2 = 0.0, av_zt = 0.0, av_du1 = 0.0, av_dv1 = 0.0;
    double hW = 0.0, rAW = 0.0, pmW = 0.0, onu = 0.0, omv = 0.0, zsW = 0.0, us = 0.0, vs = 0.0, rufrc_o = 0.0, rvfrc_o = 0.0;
    double rub_s = 0.0, rub_p = 0.0, rvb_s = 0.0, rvb_p = 0.0, ru_n = 0.0, ru_so = 0.0, rv_n = 0.0, rv_so = 0.0;
    if (outOn) {
      av_du2 = f.DU_avg2[o]; av_dv2 = f.DV_avg2[o];
      if (PRED && !FIRST) { av_zt = f.Zt_avg1[o]; av_du1 = f.DU_avg1[o]; av_dv1 = f.DV_avg1[o]; }
    }
    if (inner) {
      hW = h[o - 1]; rAW = f.rhoA[o - 1]; pmW = pm[o - 1]; onu = f.on_u[o]; omv = f.om_v[o]; zsW = zs[o - 1];
      us = f.ubar[p.kstp][o]; vs = f.vbar[p.kstp][o]; rufrc_o = f.rufrc[o]; rvfrc_o = f.rvfrc[o];
      if (!FIRST && !PRED) { rub_s = f.rubar[p.kstp][o]; rub_p = f.rubar[p.ptsk][o]; rvb_s = f.rvbar[p.kstp][o]; rvb_p = f.rvbar[p.ptsk][o]; }
      if (FIRST && PRED && p.istart >= 1) { ru_n = f.ru[p.nnew][o]; rv_n = f.rv[p.nnew][o]; ru_so = f.ru[p.nstp][o]; rv_so = f.rv[p.nstp][o]; }
    }

    // ring slots of the stage-A rows this iteration reads
    const int wm3 = ((s + 5) & 3) * CW + a;       // row s-3
    const int wm2 = ((s + 6) & 3) * CW + a;       // row s-2
    const int wm1 = ((s + 7) & 3) * CW + a;       // row s-1
    const int w0 = ((s + 8) & 3) * CW + a;        // row s

    // ================= stage B, rho point (i, r) ======================================================================
    double Dnew = 0.0, zwrk = 0.0, gz = 0.0, gz2 = 0.0, gsa = 0.0;
    double a_ufx = 0.0, a_vfe = 0.0, c_ufx = 0.0, c_vfe = 0.0, k_ufx = 0.0, k_vfe = 0.0, v_ufx = 0.0, v_vfe = 0.0;
    double DUo = 0.0, DVo = 0.0;
    if (rhoOn) {
      // row r = s-2: xi neighbours from the ring; eta neighbours r-1 .. r+2 = s-3 .. s
      const double Um = sU[wm2 - 1], U0 = sU[wm2], U1 = sU[wm2 + 1], U2 = sU[wm2 + 2];
      const double DUm = sDU[wm2 - 1], DU0 = sDU[wm2], DU1 = sDU[wm2 + 1], DU2 = sDU[wm2 + 2];
      const double Vm = sV[wm3], V0 = sV[wm2], V1 = sV[wm1], V2 = sV[w0];
      const double DVm = sDV[wm3], DV0 = sDV[wm2], DV1 = sDV[wm1], DV2 = sDV[w0];
      const double D0 = sD[wm2];
      DUo = DU0; DVo = DV0;
      // new free surface (:770-851)
      const double dd = (DU0 - DU1) + (DV0 - DV1);
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
      if (isMain && r >= j0) {                                          // own points of this CTA
        st_r_grad(f.zeta[p.knew], r * P, i, r, zeta_new, p);
        if (PRED) st_w(f.rzeta[p.krhs], r * P, i, dd, p);
      }
      // advective UFx at rho(i,r) (:1104-1112)
      {
        const double gxu0 = Um - 2.0 * U0 + U1, gxu1 = U0 - 2.0 * U1 + U2;
        const double gxd0 = DUm - 2.0 * DU0 + DU1, gxd1 = DU0 - 2.0 * DU1 + DU2;
        a_ufx = 0.25 * (U0 + U1 - c6 * (gxu0 + gxu1)) * (DU0 + DU1 - c6 * (gxd0 + gxd1));
      }
      // advective VFe at rho(i,r) (:1263-1272): grad/Dgrad rows 2..Mm with wall copies (1)=(2), (Mm+1)=(Mm)
      {
        const double gyv0 = Vm - 2.0 * V0 + V1, gyv1 = V0 - 2.0 * V1 + V2;
        const double gyd0 = DVm - 2.0 * DV0 + DV1, gyd1 = DV0 - 2.0 * DV1 + DV2;
        const double gva = (r < 2) ? gyv1 : gyv0, gvb = (r + 1 > Mm) ? gyv0 : gyv1;
        const double gda = (r < 2) ? gyd1 : gyd0, gdb = (r + 1 > Mm) ? gyd0 : gyd1;
        a_vfe = 0.25 * (V0 + V1 - c6 * (gva + gvb)) * (DV0 + DV1 - c6 * (gda + gdb));
      }
      // Coriolis (:1291-1300) and curvilinear (:1333-1347) at rho(i,r)
      const double vS = V0 + V1, uS = U0 + U1;
      {
        const double c = 0.5 * D0 * fomn_q;
        c_ufx = c * vS; c_vfe = c * uS;
      }
      if (p.curvgrid) {
        const double c1 = 0.5 * vS, c2 = 0.5 * uS;
        const double c = D0 * (c1 * dndx_q - c2 * dmde_q);
        k_ufx = c * c1; k_vfe = c * c2;
      }
      // viscous stress at rho(i,r) (:1400-1414)
      {
        const double cr = visc_q * D0 * 0.5 *
                          (pmon_q * ((pn_q + pnE) * U1 - (pnW + pn_q) * U0) -
                           pnom_q * ((pm_q + pmN) * V1 - (pmS + pm_q) * V0));
        v_ufx = onr * onr * cr; v_vfe = omr * omr * cr;
      }
    } else if (outOn) {
      DUo = sDU[wm2]; DVo = sDV[wm2];             // averaging-only rows
    }
    if (doRho) {
      xG[a] = gz; xGSA[a] = gsa; xZw[a] = zwrk; xG2[a] = gz2; xDnew[a] = Dnew;
      xaUFx[a] = a_ufx; xcUFx[a] = c_ufx; xkUFx[a] = k_ufx; xvUFx[a] = v_ufx;
    }

    // ================= stage B, psi point (i, rp) =====================================================================
    double a_ufe = 0.0, a_vfx = 0.0, v_ufe = 0.0, v_vfx = 0.0;
    if (psiOn) {
      // row rp = s-1: eta neighbours rp-2 .. rp+1 = s-3 .. s
      const double Ub2 = sU[wm3], Ub1 = sU[wm2], U0 = sU[wm1], Ua1 = sU[w0];
      const double DUb2 = sDU[wm3], DUb1 = sDU[wm2], DU0 = sDU[wm1], DUa1 = sDU[w0];
      const double Vm2 = sV[wm1 - 2], Vm1 = sV[wm1 - 1], V0 = sV[wm1], V1 = sV[wm1 + 1];
      const double DVm2 = sDV[wm1 - 2], DVm1 = sDV[wm1 - 1], DV0 = sDV[wm1], DV1 = sDV[wm1 + 1];
      const double D00 = sD[wm1], DW0 = sD[wm1 - 1], D0S = sD[wm2], DWS = sD[wm2 - 1];
      // advective UFe at psi(i,rp) (:1141-1150): grad = d2y(ubar), rows 1..Mm with wall copies (0)=(1), (Mm+1)=(Mm)
      {
        const double gyu0 = Ub1 - 2.0 * U0 + Ua1;                       // row rp
        const double gyum = Ub2 - 2.0 * Ub1 + U0;                       // row rp-1
        const double ga = (rp > Mm) ? gyum : gyu0, gb = (rp - 1 < 1) ? gyu0 : gyum;
        const double gxd0 = DVm1 - 2.0 * DV0 + DV1, gxdm = DVm2 - 2.0 * DVm1 + DV0;
        a_ufe = 0.25 * (U0 + Ub1 - c6 * (ga + gb)) * (DV0 + DVm1 - c6 * (gxd0 + gxdm));
      }
      // advective VFx at psi(i,rp), rp = 2..Mm (:1213-1222)
      if (rp >= 2 && rp <= Mm) {
        const double gxv0 = Vm1 - 2.0 * V0 + V1, gxvm = Vm2 - 2.0 * Vm1 + V0;
        const double gyd0 = DUb1 - 2.0 * DU0 + DUa1, gydm = DUb2 - 2.0 * DUb1 + DU0;
        a_vfx = 0.25 * (V0 + Vm1 - c6 * (gxv0 + gxvm)) * (DU0 + DUb1 - c6 * (gyd0 + gydm));
      }
      // viscous stress at psi(i,rp) (:1394-1430)
      {
        const double Dp = 0.25 * (D00 + DW0 + D0S + DWS);
        const double cp = pvisc * Dp * 0.5 *
                          (ppmon * ((qnS + qn0) * V0 - (qnSW + qnW) * Vm1) +
                           ppnom * ((qmW + qm0) * U0 - (qmSW + qmS) * Ub1));
        v_ufe = omp * omp * cp; v_vfx = onp * onp * cp;
      }
    }
    if (doPsi) {
      const int y = ((rp + 8) & 1) * CW + a;
      yaVFx[y] = a_vfx; yvVFx[y] = v_vfx;
    }
    __syncthreads();

    // ================= stage C: rho point (i, j = r) ==================================================================
    if (outOn) {
      const int j = r;
      // fast-time averages (:614-682); rows 0..Mm+1 for Zt/DU, rows 1..Mm+1 for DV
      if (PRED) {
        if (FIRST) {
          const double cff2 = (-1.0 / 12.0) * p.w2_p1;
          st_w(f.Zt_avg1, j * P, i, 0.0, p);
          st_w(f.DU_avg1, j * P, i, 0.0, p);
          f.DU_avg2[o] = cff2 * DUo;
          if (j >= 1) { st_w(f.DV_avg1, j * P, i, 0.0, p); f.DV_avg2[o] = cff2 * DVo; }
        } else {
          const double cff1 = p.w1_m1;
          const double cff2 = (8.0 / 12.0) * p.w2_0 - (1.0 / 12.0) * p.w2_p1;
          st_w(f.Zt_avg1, j * P, i, av_zt + cff1 * zr_q, p);
          st_w(f.DU_avg1, j * P, i, av_du1 + cff1 * DUo, p);
          f.DU_avg2[o] = av_du2 + cff2 * DUo;
          if (j >= 1) {
            st_w(f.DV_avg1, j * P, i, av_dv1 + cff1 * DVo, p);
            f.DV_avg2[o] = av_dv2 + cff2 * DVo;
          }
        }
      } else {
        const double cff2 = FIRST ? p.w2_0 : (5.0 / 12.0) * p.w2_0;
        f.DU_avg2[o] = av_du2 + cff2 * DUo;
        if (j >= 1) f.DV_avg2[o] = av_dv2 + cff2 * DVo;
      }
    }
    if (inner) {
      const int j = r;
      const double h0 = h_q, rA0 = rA, pm0 = pm_q, pn0 = pn_q, zs0 = zs_q;
      // rho-point values of (i-1, j) and psi-point values of (i+1, j) from the neighbouring columns
      const double zW_G = xG[a - 1], zW_GSA = xGSA[a - 1], zW_Zw = xZw[a - 1], zW_G2 = xG2[a - 1], zW_Dnew = xDnew[a - 1];
      const double zW_aUFx = xaUFx[a - 1], zW_cUFx = xcUFx[a - 1], zW_kUFx = xkUFx[a - 1], zW_vUFx = xvUFx[a - 1];
      // ---- u-point (i,j)
      {
        const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
        double rhs_u = cff1 * onu *
                       ((hW + h0) * (zW_G - gz) +
                        (hW - h0) * (zW_GSA + gsa + cff2 * (rAW - rA0) * (zW_Zw - zwrk)) +
                        (zW_G2 - gz2));
        {
          const double a1 = a_ufx - zW_aUFx;
          const double a2 = a_ufe - p0_aUFe;                            // psi(i,j+1) - psi(i,j)
          const double fc = a1 + a2;
          rhs_u = rhs_u - fc;
        }
        rhs_u = rhs_u + 0.5 * (c_ufx + zW_cUFx);
        if (p.curvgrid) rhs_u = rhs_u + 0.5 * (k_ufx + zW_kUFx);
        {
          const double a1 = 0.5 * (pnW + pn0) * (v_ufx - zW_vUFx);
          const double a2 = 0.5 * (pmW + pm0) * (v_ufe - p0_vUFe);
          const double fc = a1 + a2;
          rhs_u = rhs_u + fc;
        }
        // coupling with the 3-D equations (:1884-2065); level k = 0 planes of ru carry the AB3 history of the 2-D forcing
        if (FIRST && PRED) {
          const double rf = rufrc_o - rhs_u;
          if (p.istart == 0) rhs_u = rhs_u + rf;
          else if (p.istart == 1) rhs_u = rhs_u + 1.5 * rf - 0.5 * ru_n;
          else rhs_u = rhs_u + (23.0 / 12.0) * rf - (16.0 / 12.0) * ru_n + (5.0 / 12.0) * ru_so;
          f.rufrc[o] = rf;
          f.ru[p.nstp][o] = rf;
        } else {
          rhs_u = rhs_u + rufrc_o;
        }
        // time stepping (:2098-2255), rhs history (:2420-2430), BCs (:2451-2460), periodic images (:2509-2524)
        const double Dstp = (zs0 + h0) + (zsW + hW);
        const double cff = (pm0 + pmW) * (pn0 + pnW);
        const double fc = 1.0 / (Dnew + zW_Dnew);
        double x;
        if (FIRST || PRED) {
          const double c1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
          x = (us * Dstp + cff * c1 * rhs_u) * fc;
        } else {
          const double c1 = 0.5 * p.dtfast * 5.0 / 12.0, c2 = 0.5 * p.dtfast * 8.0 / 12.0, c3 = 0.5 * p.dtfast * 1.0 / 12.0;
          x = (us * Dstp + cff * (c1 * rhs_u + c2 * rub_s - c3 * rub_p)) * fc;
        }
        st_u_closed(f.ubar[p.knew], j * P, i, j, x, p);
        if (PRED) f.rubar[p.krhs][o] = rhs_u;
      }
      // ---- v-point (i,j)
      if (dov) {
        const int yE = ((j + 8) & 1) * CW + a + 1;
        const double pE_aVFx = yaVFx[yE], pE_vVFx = yvVFx[yE];
        const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
        double rhs_v = cff1 * omv *
                       ((hS + h0) * (zS_G - gz) +
                        (hS - h0) * (zS_GSA + gsa + cff2 * (rAS - rA0) * (zS_Zw - zwrk)) +
                        (zS_G2 - gz2));
        {
          const double a1 = pE_aVFx - p0_aVFx;
          const double a2 = a_vfe - zS_aVFe;
          const double fc = a1 + a2;
          rhs_v = rhs_v - fc;
        }
        rhs_v = rhs_v - 0.5 * (c_vfe + zS_cVFe);
        if (p.curvgrid) rhs_v = rhs_v - 0.5 * (k_vfe + zS_kVFe);
        {
          const double a1 = 0.5 * (pnS + pn0) * (pE_vVFx - p0_vVFx);
          const double a2 = 0.5 * (pmS + pm0) * (v_vfe - zS_vVFe);
          const double fc = a1 - a2;
          rhs_v = rhs_v + fc;
        }
        if (FIRST && PRED) {
          const double rf = rvfrc_o - rhs_v;
          if (p.istart == 0) rhs_v = rhs_v + rf;
          else if (p.istart == 1) rhs_v = rhs_v + 1.5 * rf - 0.5 * rv_n;
          else rhs_v = rhs_v + (23.0 / 12.0) * rf - (16.0 / 12.0) * rv_n + (5.0 / 12.0) * rv_so;
          f.rvfrc[o] = rf;
          f.rv[p.nstp][o] = rf;
        } else {
          rhs_v = rhs_v + rvfrc_o;
        }
        const double Dstp = (zs0 + h0) + (zsS + hS);
        const double cff = (pm0 + pmS) * (pn0 + pnS);
        const double fc = 1.0 / (Dnew + zS_Dnew);
        double x;
        if (FIRST || PRED) {
          const double c1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
          x = (vs * Dstp + cff * c1 * rhs_v) * fc;
        } else {
          const double c1 = 0.5 * p.dtfast * 5.0 / 12.0, c2 = 0.5 * p.dtfast * 8.0 / 12.0, c3 = 0.5 * p.dtfast * 1.0 / 12.0;
          x = (vs * Dstp + cff * (c1 * rhs_v + c2 * rvb_s - c3 * rvb_p)) * fc;
        }
        st_v_closed(f.vbar[p.knew], j * P, i, j, x, p);
        if (PRED) f.rvbar[p.krhs][o] = rhs_v;
      }
    }
    // ---- carry this row's own-column values to the next iteration
    zS_G = gz; zS_GSA = gsa; zS_Zw = zwrk; zS_G2 = gz2; zS_Dnew = Dnew;
    zS_aVFe = a_vfe; zS_cVFe = c_vfe; zS_kVFe = k_vfe; zS_vVFe = v_vfe;
    p0_aUFe = a_ufe; p0_aVFx = a_vfx; p0_vUFe = v_ufe; p0_vVFx = v_vfx;
    hS = h_q; rAS = rA; zsS = zs_q; pnS = pn_q;
  }
}

// ---- tile kernel (k_step2d.cu), kept for A/B measurements: ROMS_B200_STEP2D=tile ----------------------------------------
void launch_step2d_tile(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x);

template <int TXM>
static int occupancy_m() {
  constexpr int CW = TXM + 5;
  const size_t smem = (size_t)(5 * 4 + 9 + 4) * CW * sizeof(double);
  static int occs[MAXDEV] = {0};
  int& occ = occs[cur_dev()];
  if (!occ) {
    cudaFuncSetAttribute(k_step2d_m<TXM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step2d_m<TXM>, TXM + 32, smem) != cudaSuccess || occ < 1) occ = 1;
  }
  return occ;
}

template <int TXM>
static void launch_m(const Par& p, const Flds& f, cudaStream_t s, int JL) {
  constexpr int CW = TXM + 5;
  const size_t smem = (size_t)(5 * 4 + 9 + 4) * CW * sizeof(double);
  dim3 g((xspan(p) + TXM - 1) / TXM, (p.Mm + 2 + JL - 1) / JL);
  k_step2d_m<TXM><<<g, TXM + 32, smem, s>>>(p, f, JL);
}

void launch_step2d(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x) {
  // default: the tile kernel.  The marching kernel is exact and moves half the bytes through L1, but a 2048x256 field gives it
  // only ~2.5 warps per scheduler, too few to hide the FP64 and DRAM latency of a row iteration (profiles/README.md): 92 us
  // per launch against 58 us.  ROMS_B200_STEP2D=march selects it.
  static const int mode = [] { const char* e = std::getenv("ROMS_B200_STEP2D"); return (e && e[0] == 'm') ? 1 : 0; }();
  if (mode == 0 || x) { launch_step2d_tile(p, f, s, x); return; }      // the fused halo exchange lives in the tile kernel
  static const int jl_env = [] { const char* e = std::getenv("ROMS_B200_S2M_JL"); return e ? std::atoi(e) : 0; }();      // tuning aids
  static const int tx_env = [] { const char* e = std::getenv("ROMS_B200_S2M_TX"); return e ? std::atoi(e) : 0; }();
  static const int nsm = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  // split launches (multi-GPU edges) need CTA widths that divide EDGE_W; otherwise the wider CTA has less halo overhead
  const int ncol = xspan(p);
  const bool wide = tx_env != 64 && p.gap_len == 0 && ncol >= 128 && ncol % 128 == 0;
  const int occ = wide ? occupancy_m<128>() : occupancy_m<64>();
  // strip length: one resident wave of CTAs if that leaves at least 4 rows per strip
  int JL = jl_env;
  if (JL <= 0) {
    const int ncb = (ncol + (wide ? 128 : 64) - 1) / (wide ? 128 : 64);
    const int strips = (nsm * occ) / ncb > 0 ? (nsm * occ) / ncb : 1;
    JL = (p.Mm + 2 + strips - 1) / strips;
    if (JL < 4) JL = 4;
  }
  if (wide) launch_m<128>(p, f, s, JL); else launch_m<64>(p, f, s, JL);
}

}  // namespace rb

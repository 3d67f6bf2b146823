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
#include "dev.cuh"
#include "kernels.h"

namespace rb {

constexpr int TX = 32, TY = 8;           // output tile
constexpr int HL = 3, HH = 2;            // low / high halo of the staged inputs
constexpr int SW = TX + HL + HH;         // staged width  (37)
constexpr int SH = TY + HL + HH;         // staged height (21)
constexpr int ZW = TX + 1, ZH = TY + 1;  // flux / zeta regions: one extra column and row
constexpr int NS = SW * SH, NZ = ZW * ZH;
constexpr int SMEM_DOUBLES = 5 * NS + 17 * NZ;

__global__ void __launch_bounds__(TX * TY, 3) k_step2d(Par p, Flds f) {
  extern __shared__ double smem[];
  double* sD = smem; double* sU = sD + NS; double* sV = sU + NS; double* sDU = sV + NS; double* sDV = sDU + NS;
  // regions with origin (i0-1, j0-1): zeta-stage and rho-point fluxes
  double* sDnew = sDV + NS; double* sZw = sDnew + NZ; double* sG = sZw + NZ; double* sG2 = sG + NZ; double* sGSA = sG2 + NZ;
  double* aUFx = sGSA + NZ; double* aVFe = aUFx + NZ; double* cUFx = aVFe + NZ; double* cVFe = cUFx + NZ;
  double* kUFx = cVFe + NZ; double* kVFe = kUFx + NZ; double* vUFx = kVFe + NZ; double* vVFe = vUFx + NZ;
  // regions with origin (i0, j0): psi-point fluxes
  double* aUFe = vVFe + NZ; double* aVFx = aUFe + NZ; double* vUFe = aVFx + NZ; double* vVFx = vUFe + NZ;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = p.Istr + blockIdx.x * TX, j0 = blockIdx.y * TY;       // tile origin (rho point of thread 0,0)
  const int P = p.P, Mm = p.Mm;
  const bool PRED = p.predictor != 0;
  const bool FIRST = (p.iif == 1);
  const bool active = (p.iif <= p.nfast);                              // :755 (the nfast+1-th call only averages)
  const double* __restrict__ h = f.h;
  const double* __restrict__ zr = f.zeta[p.krhs];
  const double* __restrict__ zs = f.zeta[p.kstp];
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;

  // ---- prefetch (L2) what stages 2-3 will read at this thread's own point, so that their DRAM latency overlaps stages 0-1
  {
    const int ip = i0 + tx, jp = j0 + ty;
    if (ip <= p.Iend && jp <= Mm + 1) {
      const int q = jp * P + ip;
      pf_l2(zs + q); pf_l2(pm + q); pf_l2(pn + q); pf_l2(f.rhoS + q); pf_l2(f.rhoA + q); pf_l2(f.fomn + q);
      pf_l2(f.visc2_r + q); pf_l2(f.pmon_r + q); pf_l2(f.pnom_r + q); pf_l2(f.on_r + q); pf_l2(f.om_r + q);
      pf_l2(f.visc2_p + q); pf_l2(f.pmon_p + q); pf_l2(f.pnom_p + q); pf_l2(f.om_p + q); pf_l2(f.on_p + q);
      pf_l2(f.DU_avg2 + q); pf_l2(f.DV_avg2 + q); pf_l2(f.rufrc + q); pf_l2(f.rvfrc + q);
      pf_l2(f.ubar[p.kstp] + q); pf_l2(f.vbar[p.kstp] + q);
      if (p.curvgrid) { pf_l2(f.dndx + q); pf_l2(f.dmde + q); }
      if (PRED) { pf_l2(f.Zt_avg1 + q); pf_l2(f.DU_avg1 + q); pf_l2(f.DV_avg1 + q); }
      else { pf_l2(f.rzeta[p.kstp] + q); pf_l2(f.rzeta[p.ptsk] + q); pf_l2(f.rubar[p.kstp] + q); pf_l2(f.rubar[p.ptsk] + q);
             pf_l2(f.rvbar[p.kstp] + q); pf_l2(f.rvbar[p.ptsk] + q); }
    }
  }
  // ---- stage 0: Drhs, ubar, vbar on the staged region
  {
    const double* __restrict__ ur = f.ubar[p.krhs];
    const double* __restrict__ vr = f.vbar[p.krhs];
    for (int s = tid; s < NS; s += TX * TY) {
      const int a = s % SW, b = s / SW;
      const int i = i0 - HL + a, j = j0 - HL + b;
      double d = 0.0, u = 0.0, v = 0.0;
      if (i >= p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1) {
        const int q = j * P + i;
        d = zr[q] + h[q]; u = ur[q]; v = vr[q];
      }
      sD[s] = d; sU[s] = u; sV[s] = v;
    }
  }
  __syncthreads();
  // ---- stage 1: DUon (needs Drhs(i-1)), DVom (needs Drhs(j-1))
  for (int s = tid; s < NS; s += TX * TY) {
    const int a = s % SW, b = s / SW;
    const int i = i0 - HL + a, j = j0 - HL + b;
    double du = 0.0, dv = 0.0;
    if (i > p.LBi && i <= p.UBi && j >= 0 && j <= Mm + 1 && a >= 1) {
      const double c = 0.5 * f.on_u[j * P + i];
      const double c1 = c * (sD[s] + sD[s - 1]);
      du = sU[s] * c1;
    }
    if (i >= p.LBi && i <= p.UBi && j >= 1 && j <= Mm + 1 && b >= 1) {
      const double c = 0.5 * f.om_v[j * P + i];
      const double c1 = c * (sD[s] + sD[s - SW]);
      dv = sV[s] * c1;
    }
    sDU[s] = du; sDV[s] = dv;
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

  if (active) {
    const double c6 = 1.0 / 6.0;
    // ---- stage 2a: rho-point quantities on the region with origin (i0-1, j0-1)
    const double fac = 1000.0 / p.rho0;
    for (int s = tid; s < NZ; s += TX * TY) {
      const int za = s % ZW, zb = s / ZW;
      const int i = i0 - 1 + za, j = j0 - 1 + zb;
      const int c0 = (zb + HL - 1) * SW + (za + HL - 1);               // staged index of (i,j)
      double Dnew = 0.0, zwrk = 0.0, gz = 0.0, gz2 = 0.0, gsa = 0.0;
      double a_ufx = 0.0, a_vfe = 0.0, c_ufx = 0.0, c_vfe = 0.0, k_ufx = 0.0, k_vfe = 0.0, v_ufx = 0.0, v_vfe = 0.0;
      if (j >= 1 && j <= Mm && i <= p.Iend) {
        const int q = j * P + i;
        // new free surface (:770-851)
        const double dd = (DU_(0, 0) - DU_(1, 0)) + (DV_(0, 0) - DV_(0, 1));
        double zeta_new;
        const double pmn = pm[q] * pn[q];
        if (FIRST) {
          zeta_new = zs[q] + pmn * p.dtfast * dd;
          zwrk = 0.5 * (zs[q] + zeta_new);
        } else if (PRED) {
          const double cff1 = 2.0 * p.dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
          zeta_new = zs[q] + pmn * cff1 * dd;
          zwrk = cff5 * zr[q] + cff4 * (zs[q] + zeta_new);
        } else {
          const double cff1 = p.dtfast * 5.0 / 12.0, cff2 = p.dtfast * 8.0 / 12.0, cff3 = p.dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
          const double cff = cff1 * dd;
          zeta_new = zs[q] + pmn * (cff + cff2 * f.rzeta[p.kstp][q] - cff3 * f.rzeta[p.ptsk][q]);
          zwrk = cff5 * zeta_new + cff4 * zr[q];
        }
        Dnew = zeta_new + h[q];
        const double rS = f.rhoS[q];
        gz = (fac + rS) * zwrk;
        gz2 = gz * zwrk;
        gsa = zwrk * (rS - f.rhoA[q]);
        if (za >= 1 && zb >= 1) {                                      // own points of this tile
          st_r_grad(f.zeta[p.knew], j * P, i, j, zeta_new, p);
          if (PRED) st_w(f.rzeta[p.krhs], j * P, i, dd, p);
        }
        // advective UFx at rho(i,j) (:1104-1112)
        a_ufx = 0.25 * (U_(0, 0) + U_(1, 0) - c6 * (GXU(0, 0) + GXU(1, 0))) * (DU_(0, 0) + DU_(1, 0) - c6 * (GXDU(0, 0) + GXDU(1, 0)));
        // advective VFe at rho(i,j) (:1263-1272): grad/Dgrad rows 2..Mm with wall copies (1)=(2), (Mm+1)=(Mm)
        {
          const int da = (j < 2) ? 1 : 0, db = (j + 1 > Mm) ? 0 : 1;
          a_vfe = 0.25 * (V_(0, 0) + V_(0, 1) - c6 * (GYV(0, da) + GYV(0, db))) * (DV_(0, 0) + DV_(0, 1) - c6 * (GYDV(0, da) + GYDV(0, db)));
        }
        // Coriolis (:1291-1300) and curvilinear (:1333-1347) at rho(i,j)
        const double D0 = D_(0, 0);
        const double vS = V_(0, 0) + V_(0, 1), uS = U_(0, 0) + U_(1, 0);
        {
          const double c = 0.5 * D0 * f.fomn[q];
          c_ufx = c * vS; c_vfe = c * uS;
        }
        if (p.curvgrid) {
          const double c1 = 0.5 * vS, c2 = 0.5 * uS;
          const double c = D0 * (c1 * f.dndx[q] - c2 * f.dmde[q]);
          k_ufx = c * c1; k_vfe = c * c2;
        }
        // viscous stress at rho(i,j) (:1400-1414)
        {
          const double cr = f.visc2_r[q] * D0 * 0.5 *
                            (f.pmon_r[q] * ((pn[q] + pn[q + 1]) * U_(1, 0) - (pn[q - 1] + pn[q]) * U_(0, 0)) -
                             f.pnom_r[q] * ((pm[q] + pm[q + P]) * V_(0, 1) - (pm[q - P] + pm[q]) * V_(0, 0)));
          const double onr = f.on_r[q], omr = f.om_r[q];
          v_ufx = onr * onr * cr; v_vfe = omr * omr * cr;
        }
      }
      sDnew[s] = Dnew; sZw[s] = zwrk; sG[s] = gz; sG2[s] = gz2; sGSA[s] = gsa;
      aUFx[s] = a_ufx; aVFe[s] = a_vfe; cUFx[s] = c_ufx; cVFe[s] = c_vfe; kUFx[s] = k_ufx; kVFe[s] = k_vfe; vUFx[s] = v_ufx; vVFe[s] = v_vfe;
    }
    // ---- stage 2b: psi-point fluxes on the region with origin (i0, j0)
    for (int s = tid; s < NZ; s += TX * TY) {
      const int za = s % ZW, zb = s / ZW;
      const int i = i0 + za, j = j0 + zb;
      const int c0 = (zb + HL) * SW + (za + HL);
      double a_ufe = 0.0, a_vfx = 0.0, v_ufe = 0.0, v_vfx = 0.0;
      if (j >= 1 && j <= Mm + 1 && i <= p.Iend + 1) {
        const int q = j * P + i;
        // advective UFe at psi(i,j) (:1141-1150): grad = d2y(ubar), rows 1..Mm with wall copies (0)=(1), (Mm+1)=(Mm)
        {
          const int d0 = (j > Mm) ? -1 : 0, dm = (j - 1 < 1) ? 0 : -1;
          a_ufe = 0.25 * (U_(0, 0) + U_(0, -1) - c6 * (GYU(0, d0) + GYU(0, dm))) * (DV_(0, 0) + DV_(-1, 0) - c6 * (GXDV(0, 0) + GXDV(-1, 0)));
        }
        // advective VFx at psi(i,j), j = 2..Mm (:1213-1222)
        if (j >= 2 && j <= Mm)
          a_vfx = 0.25 * (V_(0, 0) + V_(-1, 0) - c6 * (GXV(0, 0) + GXV(-1, 0))) * (DU_(0, 0) + DU_(0, -1) - c6 * (GYDU(0, 0) + GYDU(0, -1)));
        // viscous stress at psi(i,j) (:1394-1430)
        {
          const double Dp = 0.25 * (D_(0, 0) + D_(-1, 0) + D_(0, -1) + D_(-1, -1));
          const double cp = f.visc2_p[q] * Dp * 0.5 *
                            (f.pmon_p[q] * ((pn[q - P] + pn[q]) * V_(0, 0) - (pn[q - P - 1] + pn[q - 1]) * V_(-1, 0)) +
                             f.pnom_p[q] * ((pm[q - 1] + pm[q]) * U_(0, 0) - (pm[q - P - 1] + pm[q - P]) * U_(0, -1)));
          const double omp = f.om_p[q], onp = f.on_p[q];
          v_ufe = omp * omp * cp; v_vfx = onp * onp * cp;
        }
      }
      aUFe[s] = a_ufe; aVFx[s] = a_vfx; vUFe[s] = v_ufe; vVFx[s] = v_vfx;
    }
  }
  __syncthreads();

  // ---- stage 3: one thread per rho point of the tile
  const int i = i0 + tx, j = j0 + ty;
  if (i > p.Iend || j > Mm + 1) return;
  const int o = j * P + i;
  // fast-time averages (:614-682); rows 0..Mm+1 for Zt/DU, rows 1..Mm+1 for DV
  {
    const int c0 = (ty + HL) * SW + (tx + HL);
    const double DUo = DU_(0, 0), DVo = DV_(0, 0);
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
        st_w(f.Zt_avg1, j * P, i, f.Zt_avg1[o] + cff1 * zr[o], p);
        st_w(f.DU_avg1, j * P, i, f.DU_avg1[o] + cff1 * DUo, p);
        f.DU_avg2[o] = f.DU_avg2[o] + cff2 * DUo;
        if (j >= 1) {
          st_w(f.DV_avg1, j * P, i, f.DV_avg1[o] + cff1 * DVo, p);
          f.DV_avg2[o] = f.DV_avg2[o] + cff2 * DVo;
        }
      }
    } else {
      const double cff2 = FIRST ? p.w2_0 : (5.0 / 12.0) * p.w2_0;
      f.DU_avg2[o] = f.DU_avg2[o] + cff2 * DUo;
      if (j >= 1) f.DV_avg2[o] = f.DV_avg2[o] + cff2 * DVo;
    }
  }
  if (!active) return;
  if (j < 1 || j > Mm) return;
  const bool dov = (j >= p.JstrV);
  const int z0 = (ty + 1) * ZW + (tx + 1), zW = z0 - 1, zS = z0 - ZW;   // rho-region indices of (i,j), (i-1,j), (i,j-1)
  const int p0 = ty * ZW + tx, pE = p0 + 1, pN = p0 + ZW;               // psi-region indices of (i,j), (i+1,j), (i,j+1)
  const double* __restrict__ rhoA = f.rhoA;
  const double h0 = h[o], hW = h[o - 1], rA0 = rhoA[o];
  const double pmU = pm[o] + pm[o - 1], pnU = pn[o] + pn[o - 1];

  // ---- u-point (i,j)
  {
    const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
    double rhs_u = cff1 * f.on_u[o] *
                   ((hW + h0) * (sG[zW] - sG[z0]) +
                    (hW - h0) * (sGSA[zW] + sGSA[z0] + cff2 * (rhoA[o - 1] - rA0) * (sZw[zW] - sZw[z0])) +
                    (sG2[zW] - sG2[z0]));
    {
      const double a1 = aUFx[z0] - aUFx[zW];
      const double a2 = aUFe[pN] - aUFe[p0];
      const double fc = a1 + a2;
      rhs_u = rhs_u - fc;
    }
    rhs_u = rhs_u + 0.5 * (cUFx[z0] + cUFx[zW]);
    if (p.curvgrid) rhs_u = rhs_u + 0.5 * (kUFx[z0] + kUFx[zW]);
    {
      const double a1 = 0.5 * (pn[o - 1] + pn[o]) * (vUFx[z0] - vUFx[zW]);
      const double a2 = 0.5 * (pm[o - 1] + pm[o]) * (vUFe[pN] - vUFe[p0]);
      const double fc = a1 + a2;
      rhs_u = rhs_u + fc;
    }
    // coupling with the 3-D equations (:1884-2065)
    if (FIRST && PRED) {
      double* __restrict__ ru_s = f.ru[p.nstp];      // level k = 0 planes carry the AB3 history of the 2-D forcing
      const double rf = f.rufrc[o] - rhs_u;
      if (p.istart == 0) rhs_u = rhs_u + rf;
      else if (p.istart == 1) rhs_u = rhs_u + 1.5 * rf - 0.5 * f.ru[p.nnew][o];
      else rhs_u = rhs_u + (23.0 / 12.0) * rf - (16.0 / 12.0) * f.ru[p.nnew][o] + (5.0 / 12.0) * ru_s[o];
      f.rufrc[o] = rf;
      ru_s[o] = rf;
    } else {
      rhs_u = rhs_u + f.rufrc[o];
    }
    // time stepping (:2098-2255), rhs history (:2420-2430), BCs (:2451-2460), periodic images (:2509-2524)
    const double Dstp = (zs[o] + h0) + (zs[o - 1] + hW);
    const double cff = pmU * pnU;
    const double fc = 1.0 / (sDnew[z0] + sDnew[zW]);
    const double us = f.ubar[p.kstp][o];
    double x;
    if (FIRST || PRED) {
      const double c1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
      x = (us * Dstp + cff * c1 * rhs_u) * fc;
    } else {
      const double c1 = 0.5 * p.dtfast * 5.0 / 12.0, c2 = 0.5 * p.dtfast * 8.0 / 12.0, c3 = 0.5 * p.dtfast * 1.0 / 12.0;
      x = (us * Dstp + cff * (c1 * rhs_u + c2 * f.rubar[p.kstp][o] - c3 * f.rubar[p.ptsk][o])) * fc;
    }
    st_u_closed(f.ubar[p.knew], j * P, i, j, x, p);
    if (PRED) f.rubar[p.krhs][o] = rhs_u;
  }
  // ---- v-point (i,j)
  if (dov) {
    const double hS = h[o - P];
    const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
    double rhs_v = cff1 * f.om_v[o] *
                   ((hS + h0) * (sG[zS] - sG[z0]) +
                    (hS - h0) * (sGSA[zS] + sGSA[z0] + cff2 * (rhoA[o - P] - rA0) * (sZw[zS] - sZw[z0])) +
                    (sG2[zS] - sG2[z0]));
    {
      const double a1 = aVFx[pE] - aVFx[p0];
      const double a2 = aVFe[z0] - aVFe[zS];
      const double fc = a1 + a2;
      rhs_v = rhs_v - fc;
    }
    rhs_v = rhs_v - 0.5 * (cVFe[z0] + cVFe[zS]);
    if (p.curvgrid) rhs_v = rhs_v - 0.5 * (kVFe[z0] + kVFe[zS]);
    {
      const double a1 = 0.5 * (pn[o - P] + pn[o]) * (vVFx[pE] - vVFx[p0]);
      const double a2 = 0.5 * (pm[o - P] + pm[o]) * (vVFe[z0] - vVFe[zS]);
      const double fc = a1 - a2;
      rhs_v = rhs_v + fc;
    }
    if (FIRST && PRED) {
      double* __restrict__ rv_s = f.rv[p.nstp];
      const double rf = f.rvfrc[o] - rhs_v;
      if (p.istart == 0) rhs_v = rhs_v + rf;
      else if (p.istart == 1) rhs_v = rhs_v + 1.5 * rf - 0.5 * f.rv[p.nnew][o];
      else rhs_v = rhs_v + (23.0 / 12.0) * rf - (16.0 / 12.0) * f.rv[p.nnew][o] + (5.0 / 12.0) * rv_s[o];
      f.rvfrc[o] = rf;
      rv_s[o] = rf;
    } else {
      rhs_v = rhs_v + f.rvfrc[o];
    }
    const double Dstp = (zs[o] + h0) + (zs[o - P] + hS);
    const double cff = (pm[o] + pm[o - P]) * (pn[o] + pn[o - P]);
    const double fc = 1.0 / (sDnew[z0] + sDnew[zS]);
    const double vs = f.vbar[p.kstp][o];
    double x;
    if (FIRST || PRED) {
      const double c1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
      x = (vs * Dstp + cff * c1 * rhs_v) * fc;
    } else {
      const double c1 = 0.5 * p.dtfast * 5.0 / 12.0, c2 = 0.5 * p.dtfast * 8.0 / 12.0, c3 = 0.5 * p.dtfast * 1.0 / 12.0;
      x = (vs * Dstp + cff * (c1 * rhs_v + c2 * f.rvbar[p.kstp][o] - c3 * f.rvbar[p.ptsk][o])) * fc;
    }
    st_v_closed(f.vbar[p.knew], j * P, i, j, x, p);
    if (PRED) f.rvbar[p.krhs][o] = rhs_v;
  }
}

void launch_step2d(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(TX, TY);
  dim3 g((p.Iend - p.Istr + 1 + TX - 1) / TX, (p.Mm + 2 + TY - 1) / TY);
  const size_t smem = (size_t)SMEM_DOUBLES * sizeof(double);
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_step2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); once = true; }
  k_step2d<<<g, b, smem, s>>>(p, f);
}

}  // namespace rb

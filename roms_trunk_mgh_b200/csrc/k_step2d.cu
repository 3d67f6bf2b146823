// step2d_tile (ROMS/Nonlinear/step2d_LF_AM3.h:137-2528): one barotropic LF-AM3 predictor or corrector sub-step as a
// single kernel.  Everything a point needs is recomputed from the read-only old time levels (krhs, kstp), so no grid-wide
// synchronisation is required inside a sub-step: depth/transport (:548-574), fast-time averages (:614-682), new free
// surface (:770-851), barotropic pressure gradient with VAR_RHO_2D terms (:944-1019), 4th-order centred advection
// (:1081-1283), Coriolis (:1291-1325), curvilinear terms (:1333-1382), harmonic viscosity (:1394-1471), 2-D/3-D coupling
// (:1884-2065), LF / AM3 stepping (:2098-2255), rhs history (:2420-2430), closed-wall BCs and periodic images.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

struct ZetaPt { double zeta_new, Dnew, zwrk, gzeta, gzeta2, gzetaSA, rhs_zeta; };

__global__ void __launch_bounds__(256) k_step2d(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;                 // 0..Mm+1
  if (i > p.Iend || j > p.Mm + 1) return;
  const int P = p.P, Mm = p.Mm, o = j * P + i;
  const bool PRED = p.predictor != 0;
  const bool FIRST = (p.iif == 1);
  const double* __restrict__ h = f.h;
  const double* __restrict__ zr = f.zeta[p.krhs];
  const double* __restrict__ zs = f.zeta[p.kstp];
  const double* __restrict__ ur = f.ubar[p.krhs];
  const double* __restrict__ vr = f.vbar[p.krhs];
  const double* __restrict__ on_u = f.on_u;
  const double* __restrict__ om_v = f.om_v;
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;
  auto D = [&](int q) -> double { return zr[q] + h[q]; };                                     // Drhs (:550)
  auto DU = [&](int q) -> double { const double c = 0.5 * on_u[q]; const double c1 = c * (D(q) + D(q - 1)); return ur[q] * c1; };
  auto DV = [&](int q) -> double { const double c = 0.5 * om_v[q]; const double c1 = c * (D(q) + D(q - P)); return vr[q] * c1; };

  // ---- fast-time averages (:614-682); rows 0..Mm+1 for Zt/DU, rows 1..Mm+1 for DV
  {
    const double DUo = DU(o);
    const double DVo = (j >= 1) ? DV(o) : 0.0;
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
  if (p.iif > p.nfast) return;                                         // :755 (auxiliary nfast+1-th call)
  if (j < 1 || j > Mm) return;

  // ---- new free surface and the pressure-gradient ingredients at one rho point (:770-851)
  const double fac = 1000.0 / p.rho0;
  const double* __restrict__ rhoA = f.rhoA;
  const double* __restrict__ rhoS = f.rhoS;
  const double* __restrict__ rz_s = f.rzeta[p.kstp];
  const double* __restrict__ rz_p = f.rzeta[p.ptsk];
  auto zeta_at = [&](int q) -> ZetaPt {
    ZetaPt z;
    const double dd = (DU(q) - DU(q + 1)) + (DV(q) - DV(q + P));
    if (FIRST) {
      z.rhs_zeta = dd;
      z.zeta_new = zs[q] + pm[q] * pn[q] * p.dtfast * z.rhs_zeta;
      z.zwrk = 0.5 * (zs[q] + z.zeta_new);
    } else if (PRED) {
      const double cff1 = 2.0 * p.dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
      z.rhs_zeta = dd;
      z.zeta_new = zs[q] + pm[q] * pn[q] * cff1 * z.rhs_zeta;
      z.zwrk = cff5 * zr[q] + cff4 * (zs[q] + z.zeta_new);
    } else {
      const double cff1 = p.dtfast * 5.0 / 12.0, cff2 = p.dtfast * 8.0 / 12.0, cff3 = p.dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
      const double cff = cff1 * dd;
      z.rhs_zeta = 0.0;
      z.zeta_new = zs[q] + pm[q] * pn[q] * (cff + cff2 * rz_s[q] - cff3 * rz_p[q]);
      z.zwrk = cff5 * z.zeta_new + cff4 * zr[q];
    }
    z.Dnew = z.zeta_new + h[q];
    z.gzeta = (fac + rhoS[q]) * z.zwrk;
    z.gzeta2 = z.gzeta * z.zwrk;
    z.gzetaSA = z.zwrk * (rhoS[q] - rhoA[q]);
    return z;
  };
  const bool dov = (j >= p.JstrV);
  const ZetaPt z0 = zeta_at(o), zW = zeta_at(o - 1);
  ZetaPt zS = z0;                                                      // rho(i,j-1) only exists for v-points (j >= JstrV)
  if (dov) zS = zeta_at(o - P);
  st_r_grad(f.zeta[p.knew], j * P, i, j, z0.zeta_new, p);              // :860-929 (zetabc + periodic images)
  if (PRED) st_w(f.rzeta[p.krhs], j * P, i, z0.rhs_zeta, p);

  // ---- pressure gradient (:944-1019)
  double rhs_u, rhs_v = 0.0;
  {
    const double cff1 = 0.5 * p.g, cff2 = 1.0 / 3.0;
    rhs_u = cff1 * on_u[o] *
            ((h[o - 1] + h[o]) * (zW.gzeta - z0.gzeta) +
             (h[o - 1] - h[o]) * (zW.gzetaSA + z0.gzetaSA + cff2 * (rhoA[o - 1] - rhoA[o]) * (zW.zwrk - z0.zwrk)) +
             (zW.gzeta2 - z0.gzeta2));
    if (dov)
      rhs_v = cff1 * om_v[o] *
              ((h[o - P] + h[o]) * (zS.gzeta - z0.gzeta) +
               (h[o - P] - h[o]) * (zS.gzetaSA + z0.gzetaSA + cff2 * (rhoA[o - P] - rhoA[o]) * (zS.zwrk - z0.zwrk)) +
               (zS.gzeta2 - z0.gzeta2));
  }

  // ---- advection, 4th-order centred (:1081-1283)
  {
    const double c6 = 1.0 / 6.0;
    auto gxu = [&](int q) { return ur[q - 1] - 2.0 * ur[q] + ur[q + 1]; };          // grad  for UFx (u, xi)
    auto gxD = [&](int q) { return DU(q - 1) - 2.0 * DU(q) + DU(q + 1); };          // Dgrad for UFx
    // UFx at rho(i,j) and rho(i-1,j)
    const double DUw = DU(o - 1), DU0 = DU(o), DUe = DU(o + 1);
    const double gW = gxu(o - 1), g0 = gxu(o), gE = gxu(o + 1);
    const double DgW = gxD(o - 1), Dg0 = DU(o - 1) - 2.0 * DU0 + DUe, DgE = gxD(o + 1);
    const double UFx0 = 0.25 * (ur[o] + ur[o + 1] - c6 * (g0 + gE)) * (DU0 + DUe - c6 * (Dg0 + DgE));
    const double UFxW = 0.25 * (ur[o - 1] + ur[o] - c6 * (gW + g0)) * (DUw + DU0 - c6 * (DgW + Dg0));
    // UFe at psi(i,j), psi(i,j+1): grad = d2y(ubar) with wall copies grad(0)=grad(1), grad(Mm+1)=grad(Mm)
    auto gyu = [&](int jj) { const int q = jj * P + i; return ur[q - P] - 2.0 * ur[q] + ur[q + P]; };
    const int jm = (j - 1 < 1) ? 1 : j - 1, jp = (j + 1 > Mm) ? Mm : j + 1;
    const double gy_m = gyu(jm), gy_0 = gyu(j), gy_p = gyu(jp);
    auto gxDV = [&](int q) { return DV(q - 1) - 2.0 * DV(q) + DV(q + 1); };         // Dgrad for UFe
    const double UFe0 = 0.25 * (ur[o] + ur[o - P] - c6 * (gy_0 + gy_m)) * (DV(o) + DV(o - 1) - c6 * (gxDV(o) + gxDV(o - 1)));
    const double UFeN = 0.25 * (ur[o + P] + ur[o] - c6 * (gy_p + gy_0)) * (DV(o + P) + DV(o + P - 1) - c6 * (gxDV(o + P) + gxDV(o + P - 1)));
    {
      const double cff1 = UFx0 - UFxW;
      const double cff2 = UFeN - UFe0;
      const double fc = cff1 + cff2;
      rhs_u = rhs_u - fc;
    }
    if (dov) {
      auto gxv = [&](int q) { return vr[q - 1] - 2.0 * vr[q] + vr[q + 1]; };
      auto gyDU = [&](int q) { return DU(q - P) - 2.0 * DU(q) + DU(q + P); };
      // VFx at psi(i,j), psi(i+1,j)
      const double VFx0 = 0.25 * (vr[o] + vr[o - 1] - c6 * (gxv(o) + gxv(o - 1))) * (DU(o) + DU(o - P) - c6 * (gyDU(o) + gyDU(o - P)));
      const double VFxE = 0.25 * (vr[o + 1] + vr[o] - c6 * (gxv(o + 1) + gxv(o))) * (DU(o + 1) + DU(o + 1 - P) - c6 * (gyDU(o + 1) + gyDU(o + 1 - P)));
      // VFe at rho(i,j), rho(i,j-1): grad/Dgrad rows 2..Mm with copies (1)=(2), (Mm+1)=(Mm)
      auto gyv = [&](int jj) { const int q = jj * P + i; return vr[q - P] - 2.0 * vr[q] + vr[q + P]; };
      auto gyDV = [&](int jj) { const int q = jj * P + i; return DV(q - P) - 2.0 * DV(q) + DV(q + P); };
      const int ja = j, jb = (j + 1 > Mm) ? Mm : j + 1, jc = (j - 1 < 2) ? 2 : j - 1;
      const double gv_a = gyv(ja), gv_b = gyv(jb), gv_c = gyv(jc);
      const double Dv_a = gyDV(ja), Dv_b = gyDV(jb), Dv_c = gyDV(jc);
      const double VFe0 = 0.25 * (vr[o] + vr[o + P] - c6 * (gv_a + gv_b)) * (DV(o) + DV(o + P) - c6 * (Dv_a + Dv_b));
      const double VFeS = 0.25 * (vr[o - P] + vr[o] - c6 * (gv_c + gv_a)) * (DV(o - P) + DV(o) - c6 * (Dv_c + Dv_a));
      const double cff1 = VFxE - VFx0;
      const double cff2 = VFe0 - VFeS;
      const double fc = cff1 + cff2;
      rhs_v = rhs_v - fc;
    }
  }
  // ---- Coriolis (:1291-1325) and curvilinear terms (:1333-1382)
  {
    const double D0 = D(o), DW = D(o - 1), DS = D(o - P);
    const double vS0 = vr[o] + vr[o + P], uS0 = ur[o] + ur[o + 1];          // sums at rho(i,j)
    const double vSW = vr[o - 1] + vr[o + P - 1], uSW = ur[o - 1] + ur[o];  // rho(i-1,j)
    const double vSS = vr[o - P] + vr[o], uSS = ur[o - P] + ur[o - P + 1];  // rho(i,j-1)
    {
      const double c0 = 0.5 * D0 * f.fomn[o], cW = 0.5 * DW * f.fomn[o - 1];
      const double UFx0 = c0 * vS0, UFxW = cW * vSW;
      const double fac1 = 0.5 * (UFx0 + UFxW);
      rhs_u = rhs_u + fac1;
      if (dov) {
        const double cS = 0.5 * DS * f.fomn[o - P];
        const double VFe0 = c0 * uS0, VFeS = cS * uSS;
        const double fac1v = 0.5 * (VFe0 + VFeS);
        rhs_v = rhs_v - fac1v;
      }
    }
    if (p.curvgrid) {
      double c1 = 0.5 * vS0, c2 = 0.5 * uS0;
      double c = D0 * (c1 * f.dndx[o] - c2 * f.dmde[o]);
      const double UFx0 = c * c1, VFe0 = c * c2;
      c1 = 0.5 * vSW; c2 = 0.5 * uSW;
      c = DW * (c1 * f.dndx[o - 1] - c2 * f.dmde[o - 1]);
      const double UFxW = c * c1;
      const double fac1 = 0.5 * (UFx0 + UFxW);
      rhs_u = rhs_u + fac1;
      if (dov) {
        c1 = 0.5 * vSS; c2 = 0.5 * uSS;
        c = DS * (c1 * f.dndx[o - P] - c2 * f.dmde[o - P]);
        const double VFeS = c * c2;
        const double fac1v = 0.5 * (VFe0 + VFeS);
        rhs_v = rhs_v - fac1v;
      }
    }
  }
  // ---- harmonic viscosity (:1394-1471)
  {
    auto Dp = [&](int q) { return 0.25 * (D(q) + D(q - 1) + D(q - P) + D(q - P - 1)); };                  // Drhs_p at psi(q)
    auto rcf = [&](int q) {                                                                              // rho-point stress
      return f.visc2_r[q] * D(q) * 0.5 *
             (f.pmon_r[q] * ((pn[q] + pn[q + 1]) * ur[q + 1] - (pn[q - 1] + pn[q]) * ur[q]) -
              f.pnom_r[q] * ((pm[q] + pm[q + P]) * vr[q + P] - (pm[q - P] + pm[q]) * vr[q]));
    };
    auto pcf = [&](int q) {                                                                              // psi-point stress
      return f.visc2_p[q] * Dp(q) * 0.5 *
             (f.pmon_p[q] * ((pn[q - P] + pn[q]) * vr[q] - (pn[q - P - 1] + pn[q - 1]) * vr[q - 1]) +
              f.pnom_p[q] * ((pm[q - 1] + pm[q]) * ur[q] - (pm[q - P - 1] + pm[q - P]) * ur[q - P]));
    };
    const double cr0 = rcf(o), crW = rcf(o - 1);
    const double cp0 = pcf(o), cpN = pcf(o + P);
    {
      const double UFx0 = f.on_r[o] * f.on_r[o] * cr0, UFxW = f.on_r[o - 1] * f.on_r[o - 1] * crW;
      const double UFe0 = f.om_p[o] * f.om_p[o] * cp0, UFeN = f.om_p[o + P] * f.om_p[o + P] * cpN;
      const double cff1 = 0.5 * (pn[o - 1] + pn[o]) * (UFx0 - UFxW);
      const double cff2 = 0.5 * (pm[o - 1] + pm[o]) * (UFeN - UFe0);
      const double fc = cff1 + cff2;
      rhs_u = rhs_u + fc;
    }
    if (dov) {
      const double crS = rcf(o - P), cpE = pcf(o + 1);
      const double VFx0 = f.on_p[o] * f.on_p[o] * cp0, VFxE = f.on_p[o + 1] * f.on_p[o + 1] * cpE;
      const double VFe0 = f.om_r[o] * f.om_r[o] * cr0, VFeS = f.om_r[o - P] * f.om_r[o - P] * crS;
      const double cff1 = 0.5 * (pn[o - P] + pn[o]) * (VFxE - VFx0);
      const double cff2 = 0.5 * (pm[o - P] + pm[o]) * (VFe0 - VFeS);
      const double fc = cff1 - cff2;
      rhs_v = rhs_v + fc;
    }
  }
  // ---- coupling with the 3-D equations (:1884-2065)
  if (FIRST && PRED) {
    double* __restrict__ ru_s = f.ru[p.nstp];        // level k = 0 planes carry the AB3 history of the 2-D forcing
    double* __restrict__ rv_s = f.rv[p.nstp];
    const double* __restrict__ ru_n = f.ru[p.nnew];
    const double* __restrict__ rv_n = f.rv[p.nnew];
    {
      const double rf = f.rufrc[o] - rhs_u;
      if (p.istart == 0) rhs_u = rhs_u + rf;
      else if (p.istart == 1) rhs_u = rhs_u + 1.5 * rf - 0.5 * ru_n[o];
      else rhs_u = rhs_u + (23.0 / 12.0) * rf - (16.0 / 12.0) * ru_n[o] + (5.0 / 12.0) * ru_s[o];
      f.rufrc[o] = rf;
      ru_s[o] = rf;
    }
    if (dov) {
      const double rf = f.rvfrc[o] - rhs_v;
      if (p.istart == 0) rhs_v = rhs_v + rf;
      else if (p.istart == 1) rhs_v = rhs_v + 1.5 * rf - 0.5 * rv_n[o];
      else rhs_v = rhs_v + (23.0 / 12.0) * rf - (16.0 / 12.0) * rv_n[o] + (5.0 / 12.0) * rv_s[o];
      f.rvfrc[o] = rf;
      rv_s[o] = rf;
    }
  } else {
    rhs_u = rhs_u + f.rufrc[o];
    if (dov) rhs_v = rhs_v + f.rvfrc[o];
  }
  // ---- time stepping (:2098-2255), rhs history (:2420-2430), BCs (:2451-2460), periodic images (:2509-2524)
  {
    const double Dstp0 = zs[o] + h[o];
    const double* __restrict__ us = f.ubar[p.kstp];
    const double* __restrict__ vs = f.vbar[p.kstp];
    {
      const double DstpW = zs[o - 1] + h[o - 1];
      const double cff = (pm[o] + pm[o - 1]) * (pn[o] + pn[o - 1]);
      const double fc = 1.0 / (z0.Dnew + zW.Dnew);
      double x;
      if (FIRST || PRED) {
        const double cff1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
        x = (us[o] * (Dstp0 + DstpW) + cff * cff1 * rhs_u) * fc;
      } else {
        const double cff1 = 0.5 * p.dtfast * 5.0 / 12.0, cff2 = 0.5 * p.dtfast * 8.0 / 12.0, cff3 = 0.5 * p.dtfast * 1.0 / 12.0;
        x = (us[o] * (Dstp0 + DstpW) + cff * (cff1 * rhs_u + cff2 * f.rubar[p.kstp][o] - cff3 * f.rubar[p.ptsk][o])) * fc;
      }
      st_u_closed(f.ubar[p.knew], j * P, i, j, x, p);
      if (PRED) f.rubar[p.krhs][o] = rhs_u;
    }
    if (dov) {
      const double DstpS = zs[o - P] + h[o - P];
      const double cff = (pm[o] + pm[o - P]) * (pn[o] + pn[o - P]);
      const double fc = 1.0 / (z0.Dnew + zS.Dnew);
      double x;
      if (FIRST || PRED) {
        const double cff1 = FIRST ? 0.5 * p.dtfast : p.dtfast;
        x = (vs[o] * (Dstp0 + DstpS) + cff * cff1 * rhs_v) * fc;
      } else {
        const double cff1 = 0.5 * p.dtfast * 5.0 / 12.0, cff2 = 0.5 * p.dtfast * 8.0 / 12.0, cff3 = 0.5 * p.dtfast * 1.0 / 12.0;
        x = (vs[o] * (Dstp0 + DstpS) + cff * (cff1 * rhs_v + cff2 * f.rvbar[p.kstp][o] - cff3 * f.rvbar[p.ptsk][o])) * fc;
      }
      st_v_closed(f.vbar[p.knew], j * P, i, j, x, p);
      if (PRED) f.rvbar[p.krhs][o] = rhs_v;
    }
  }
}

void launch_step2d(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(64, 4);
  dim3 g((p.Iend - p.Istr + 1 + b.x - 1) / b.x, (p.Mm + 2 + b.y - 1) / b.y);
  k_step2d<<<g, b, 0, s>>>(p, f);
}

}  // namespace rb

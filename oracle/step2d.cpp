// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/step2d_LF_AM3.h:137-2528 (step2d_tile; included by step2d.F): barotropic LF-AM3 predictor /
// corrector with fast-time averaging.  Live options: SOLVE3D, VAR_RHO_2D (globaldefs.h:491-495), UV_ADV (4th-order
// centred, the default :1081-1283), UV_COR, CURVGRID (run-time flag here), UV_VIS2.  Shared-memory (non-DISTRIBUTE)
// ranges for Drhs/DUon/DVom (:548-574).
#include "roms_oracle.hpp"

namespace orc {

void step2d(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int krhs = m.krhs, kstp = m.kstp, knew = m.knew, iif = m.iif, nfast = m.nfast;
  const int nstp = m.nstp, nnew = m.nnew;
  const bool PREDICTOR = m.PREDICTOR_2D_STEP;
  const bool CORRECTOR = !PREDICTOR;
  const int ptsk = 3 - kstp;                                          // :502
  const double dtfast = m.dtfast, g = c.g, rho0 = c.rho0;
  const std::vector<double>& w1 = m.weight1; const std::vector<double>& w2 = m.weight2;
  F2 &h = m.h, &pm = m.pm, &pn = m.pn, &on_u = m.on_u, &om_v = m.om_v, &rhoA = m.rhoA, &rhoS = m.rhoS;
  F2 zeta_r = m.zeta[krhs], zeta_s = m.zeta[kstp], zeta_n = m.zeta[knew];
  F2 ubar_r = m.ubar[krhs], ubar_s = m.ubar[kstp], ubar_n = m.ubar[knew];
  F2 vbar_r = m.vbar[krhs], vbar_s = m.vbar[kstp], vbar_n = m.vbar[knew];
#define ORC_S2(name) S2 name(IminS, ImaxS, JminS, JmaxS)
  ORC_S2(Dgrad); ORC_S2(Dnew); ORC_S2(Drhs); ORC_S2(Drhs_p); ORC_S2(Dstp); ORC_S2(DUon); ORC_S2(DVom);
  ORC_S2(UFe); ORC_S2(UFx); ORC_S2(VFe); ORC_S2(VFx); ORC_S2(grad); ORC_S2(gzeta); ORC_S2(gzeta2); ORC_S2(gzetaSA);
  ORC_S2(rhs_ubar); ORC_S2(rhs_vbar); ORC_S2(rhs_zeta); ORC_S2(zeta_new); ORC_S2(zwrk);
#undef ORC_S2

  // ---- :548-574  total depth and transports at krhs
  for (int j = JstrVm2 - 1; j <= Jendp2; ++j)
    for (int i = IstrUm2 - 1; i <= Iendp2; ++i) Drhs(i, j) = zeta_r(i, j) + h(i, j);
  for (int j = JstrVm2 - 1; j <= Jendp2; ++j)
    for (int i = IstrUm2; i <= Iendp2; ++i) {
      double cff = 0.5 * on_u(i, j);
      double cff1 = cff * (Drhs(i, j) + Drhs(i - 1, j));
      DUon(i, j) = ubar_r(i, j) * cff1;
    }
  for (int j = JstrVm2; j <= Jendp2; ++j)
    for (int i = IstrUm2 - 1; i <= Iendp2; ++i) {
      double cff = 0.5 * om_v(i, j);
      double cff1 = cff * (Drhs(i, j) + Drhs(i, j - 1));
      DVom(i, j) = vbar_r(i, j) * cff1;
    }

  // ---- :614-682  fast-time averaging
  if (PREDICTOR) {
    if (iif == 1) {
      double cff2 = (-1.0 / 12.0) * w2[iif + 1];
      for (int j = JstrR; j <= JendR; ++j) {
        for (int i = IstrR; i <= IendR; ++i) m.Zt_avg1(i, j) = 0.0;
        for (int i = Istr; i <= IendR; ++i) { m.DU_avg1(i, j) = 0.0; m.DU_avg2(i, j) = cff2 * DUon(i, j); }
      }
      for (int j = Jstr; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) { m.DV_avg1(i, j) = 0.0; m.DV_avg2(i, j) = cff2 * DVom(i, j); }
    } else {
      double cff1 = w1[iif - 1];
      double cff2 = (8.0 / 12.0) * w2[iif] - (1.0 / 12.0) * w2[iif + 1];
      for (int j = JstrR; j <= JendR; ++j) {
        for (int i = IstrR; i <= IendR; ++i) m.Zt_avg1(i, j) = m.Zt_avg1(i, j) + cff1 * zeta_r(i, j);
        for (int i = Istr; i <= IendR; ++i) {
          m.DU_avg1(i, j) = m.DU_avg1(i, j) + cff1 * DUon(i, j);
          m.DU_avg2(i, j) = m.DU_avg2(i, j) + cff2 * DUon(i, j);
        }
      }
      for (int j = Jstr; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.DV_avg1(i, j) = m.DV_avg1(i, j) + cff1 * DVom(i, j);
          m.DV_avg2(i, j) = m.DV_avg2(i, j) + cff2 * DVom(i, j);
        }
    }
  } else {
    double cff2 = (iif == 1) ? w2[iif] : (5.0 / 12.0) * w2[iif];
    for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.DU_avg2(i, j) = m.DU_avg2(i, j) + cff2 * DUon(i, j);
    for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.DV_avg2(i, j) = m.DV_avg2(i, j) + cff2 * DVom(i, j);
  }
  // ---- :693-727
  if (iif == nfast + 1 && PREDICTOR) {
    exchange_r2d(m, b, m.Zt_avg1); exchange_u2d(m, b, m.DU_avg1); exchange_v2d(m, b, m.DV_avg1);
  }
  // ---- :755
  if (iif > nfast) return;

  // ---- :768-851  free surface at the new time level
  double fac = 1000.0 / rho0;
  if (iif == 1) {
    double cff1 = dtfast;
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        rhs_zeta(i, j) = (DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1));
        zeta_new(i, j) = zeta_s(i, j) + pm(i, j) * pn(i, j) * cff1 * rhs_zeta(i, j);
        Dnew(i, j) = zeta_new(i, j) + h(i, j);
        zwrk(i, j) = 0.5 * (zeta_s(i, j) + zeta_new(i, j));
        gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
        gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
        gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
      }
  } else if (PREDICTOR) {
    double cff1 = 2.0 * dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        rhs_zeta(i, j) = (DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1));
        zeta_new(i, j) = zeta_s(i, j) + pm(i, j) * pn(i, j) * cff1 * rhs_zeta(i, j);
        Dnew(i, j) = zeta_new(i, j) + h(i, j);
        zwrk(i, j) = cff5 * zeta_r(i, j) + cff4 * (zeta_s(i, j) + zeta_new(i, j));
        gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
        gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
        gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
      }
  } else if (CORRECTOR) {
    double cff1 = dtfast * 5.0 / 12.0, cff2 = dtfast * 8.0 / 12.0, cff3 = dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
    F2 rz_s = m.rzeta[kstp], rz_p = m.rzeta[ptsk];
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff = cff1 * ((DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1)));
        zeta_new(i, j) = zeta_s(i, j) + pm(i, j) * pn(i, j) * (cff + cff2 * rz_s(i, j) - cff3 * rz_p(i, j));
        Dnew(i, j) = zeta_new(i, j) + h(i, j);
        zwrk(i, j) = cff5 * zeta_new(i, j) + cff4 * zeta_r(i, j);
        gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
        gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
        gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
      }
  }
  // ---- :860-929
  for (int j = Jstr; j <= Jend; ++j) for (int i = Istr; i <= Iend; ++i) zeta_n(i, j) = zeta_new(i, j);
  if (PREDICTOR) {
    F2 rz = m.rzeta[krhs];
    for (int j = Jstr; j <= Jend; ++j) for (int i = Istr; i <= Iend; ++i) rz(i, j) = rhs_zeta(i, j);
    exchange_r2d(m, b, rz);
  }
  zetabc(m, b, knew);
  exchange_r2d(m, b, zeta_n);

  // ---- :939-1019  pressure gradient
  {
    double cff1 = 0.5 * g, cff2 = 1.0 / 3.0;
    for (int j = Jstr; j <= Jend; ++j) {
      for (int i = IstrU; i <= Iend; ++i)
        rhs_ubar(i, j) = cff1 * on_u(i, j) *
                         ((h(i - 1, j) + h(i, j)) * (gzeta(i - 1, j) - gzeta(i, j)) +
                          (h(i - 1, j) - h(i, j)) * (gzetaSA(i - 1, j) + gzetaSA(i, j) + cff2 * (rhoA(i - 1, j) - rhoA(i, j)) * (zwrk(i - 1, j) - zwrk(i, j))) +
                          (gzeta2(i - 1, j) - gzeta2(i, j)));
      if (j >= JstrV)
        for (int i = Istr; i <= Iend; ++i)
          rhs_vbar(i, j) = cff1 * om_v(i, j) *
                           ((h(i, j - 1) + h(i, j)) * (gzeta(i, j - 1) - gzeta(i, j)) +
                            (h(i, j - 1) - h(i, j)) * (gzetaSA(i, j - 1) + gzetaSA(i, j) + cff2 * (rhoA(i, j - 1) - rhoA(i, j)) * (zwrk(i, j - 1) - zwrk(i, j))) +
                            (gzeta2(i, j - 1) - gzeta2(i, j)));
    }
  }

  if (c.uv_adv == 3) {                                                  // UV_C2ADVECTION :1026-1080
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) UFx(i, j) = 0.25 * (DUon(i, j) + DUon(i + 1, j)) * (ubar_r(i, j) + ubar_r(i + 1, j));
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = IstrU; i <= Iend; ++i) UFe(i, j) = 0.25 * (DVom(i, j) + DVom(i - 1, j)) * (ubar_r(i, j) + ubar_r(i, j - 1));
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend + 1; ++i) VFx(i, j) = 0.25 * (DUon(i, j) + DUon(i, j - 1)) * (vbar_r(i, j) + vbar_r(i - 1, j));
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) VFe(i, j) = 0.25 * (DVom(i, j) + DVom(i, j + 1)) * (vbar_r(i, j) + vbar_r(i, j + 1));
  } else {
  // ---- :1081-1283  UV_ADV, fourth-order centred
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = IstrUm1; i <= Iendp1; ++i) {
      grad(i, j) = ubar_r(i - 1, j) - 2.0 * ubar_r(i, j) + ubar_r(i + 1, j);
      Dgrad(i, j) = DUon(i - 1, j) - 2.0 * DUon(i, j) + DUon(i + 1, j);
    }
  {
    double cff = 1.0 / 6.0;
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i)
        UFx(i, j) = 0.25 * (ubar_r(i, j) + ubar_r(i + 1, j) - cff * (grad(i, j) + grad(i + 1, j))) *
                    (DUon(i, j) + DUon(i + 1, j) - cff * (Dgrad(i, j) + Dgrad(i + 1, j)));
  }
  for (int j = Jstrm1; j <= Jendp1; ++j)
    for (int i = IstrU; i <= Iend; ++i) grad(i, j) = ubar_r(i, j - 1) - 2.0 * ubar_r(i, j) + ubar_r(i, j + 1);
  if (b.Southern_Edge) for (int i = IstrU; i <= Iend; ++i) grad(i, Jstr - 1) = grad(i, Jstr);
  if (b.Northern_Edge) for (int i = IstrU; i <= Iend; ++i) grad(i, Jend + 1) = grad(i, Jend);
  for (int j = Jstr; j <= Jend + 1; ++j)
    for (int i = IstrU - 1; i <= Iend; ++i) Dgrad(i, j) = DVom(i - 1, j) - 2.0 * DVom(i, j) + DVom(i + 1, j);
  {
    double cff = 1.0 / 6.0;
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = IstrU; i <= Iend; ++i)
        UFe(i, j) = 0.25 * (ubar_r(i, j) + ubar_r(i, j - 1) - cff * (grad(i, j) + grad(i, j - 1))) *
                    (DVom(i, j) + DVom(i - 1, j) - cff * (Dgrad(i, j) + Dgrad(i - 1, j)));
  }
  for (int j = JstrV; j <= Jend; ++j)
    for (int i = Istrm1; i <= Iendp1; ++i) grad(i, j) = vbar_r(i - 1, j) - 2.0 * vbar_r(i, j) + vbar_r(i + 1, j);
  for (int j = JstrV - 1; j <= Jend; ++j)
    for (int i = Istr; i <= Iend + 1; ++i) Dgrad(i, j) = DUon(i, j - 1) - 2.0 * DUon(i, j) + DUon(i, j + 1);
  {
    double cff = 1.0 / 6.0;
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend + 1; ++i)
        VFx(i, j) = 0.25 * (vbar_r(i, j) + vbar_r(i - 1, j) - cff * (grad(i, j) + grad(i - 1, j))) *
                    (DUon(i, j) + DUon(i, j - 1) - cff * (Dgrad(i, j) + Dgrad(i, j - 1)));
  }
  for (int j = JstrVm1; j <= Jendp1; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      grad(i, j) = vbar_r(i, j - 1) - 2.0 * vbar_r(i, j) + vbar_r(i, j + 1);
      Dgrad(i, j) = DVom(i, j - 1) - 2.0 * DVom(i, j) + DVom(i, j + 1);
    }
  if (b.Southern_Edge) for (int i = Istr; i <= Iend; ++i) { grad(i, Jstr) = grad(i, Jstr + 1); Dgrad(i, Jstr) = Dgrad(i, Jstr + 1); }
  if (b.Northern_Edge) for (int i = Istr; i <= Iend; ++i) { grad(i, Jend + 1) = grad(i, Jend); Dgrad(i, Jend + 1) = Dgrad(i, Jend); }
  {
    double cff = 1.0 / 6.0;
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i)
        VFe(i, j) = 0.25 * (vbar_r(i, j) + vbar_r(i, j + 1) - cff * (grad(i, j) + grad(i, j + 1))) *
                    (DVom(i, j) + DVom(i, j + 1) - cff * (Dgrad(i, j) + Dgrad(i, j + 1)));
  }
  }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = IstrU; i <= Iend; ++i) {
      double cff1 = UFx(i, j) - UFx(i - 1, j);
      double cff2 = UFe(i, j + 1) - UFe(i, j);
      double fc = cff1 + cff2;
      rhs_ubar(i, j) = rhs_ubar(i, j) - fc;
    }
  for (int j = JstrV; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      double cff1 = VFx(i + 1, j) - VFx(i, j);
      double cff2 = VFe(i, j) - VFe(i, j - 1);
      double fc = cff1 + cff2;
      rhs_vbar(i, j) = rhs_vbar(i, j) - fc;
    }

  // ---- :1291-1325  UV_COR
  for (int j = JstrV - 1; j <= Jend; ++j)
    for (int i = IstrU - 1; i <= Iend; ++i) {
      double cff = 0.5 * Drhs(i, j) * m.fomn(i, j);
      UFx(i, j) = cff * (vbar_r(i, j) + vbar_r(i, j + 1));
      VFe(i, j) = cff * (ubar_r(i, j) + ubar_r(i + 1, j));
    }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = IstrU; i <= Iend; ++i) { double fac1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); rhs_ubar(i, j) = rhs_ubar(i, j) + fac1; }
  for (int j = JstrV; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) { double fac1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rhs_vbar(i, j) = rhs_vbar(i, j) - fac1; }

  // ---- :1333-1382  CURVGRID && UV_ADV
  if (c.curvgrid) {
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff1 = 0.5 * (vbar_r(i, j) + vbar_r(i, j + 1));
        double cff2 = 0.5 * (ubar_r(i, j) + ubar_r(i + 1, j));
        double cff3 = cff1 * m.dndx(i, j);
        double cff4 = cff2 * m.dmde(i, j);
        double cff = Drhs(i, j) * (cff3 - cff4);
        UFx(i, j) = cff * cff1;
        VFe(i, j) = cff * cff2;
      }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) { double fac1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); rhs_ubar(i, j) = rhs_ubar(i, j) + fac1; }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) { double fac1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rhs_vbar(i, j) = rhs_vbar(i, j) - fac1; }
  }

  // ---- :1394-1471  UV_VIS2
  for (int j = Jstr; j <= Jend + 1; ++j)
    for (int i = Istr; i <= Iend + 1; ++i) Drhs_p(i, j) = 0.25 * (Drhs(i, j) + Drhs(i - 1, j) + Drhs(i, j - 1) + Drhs(i - 1, j - 1));
  for (int j = JstrV - 1; j <= Jend; ++j)
    for (int i = IstrU - 1; i <= Iend; ++i) {
      double cff = m.visc2_r(i, j) * Drhs(i, j) * 0.5 *
                   (m.pmon_r(i, j) * ((pn(i, j) + pn(i + 1, j)) * ubar_r(i + 1, j) - (pn(i - 1, j) + pn(i, j)) * ubar_r(i, j)) -
                    m.pnom_r(i, j) * ((pm(i, j) + pm(i, j + 1)) * vbar_r(i, j + 1) - (pm(i, j - 1) + pm(i, j)) * vbar_r(i, j)));
      UFx(i, j) = m.on_r(i, j) * m.on_r(i, j) * cff;
      VFe(i, j) = m.om_r(i, j) * m.om_r(i, j) * cff;
    }
  for (int j = Jstr; j <= Jend + 1; ++j)
    for (int i = Istr; i <= Iend + 1; ++i) {
      double cff = m.visc2_p(i, j) * Drhs_p(i, j) * 0.5 *
                   (m.pmon_p(i, j) * ((pn(i, j - 1) + pn(i, j)) * vbar_r(i, j) - (pn(i - 1, j - 1) + pn(i - 1, j)) * vbar_r(i - 1, j)) +
                    m.pnom_p(i, j) * ((pm(i - 1, j) + pm(i, j)) * ubar_r(i, j) - (pm(i - 1, j - 1) + pm(i, j - 1)) * ubar_r(i, j - 1)));
      UFe(i, j) = m.om_p(i, j) * m.om_p(i, j) * cff;
      VFx(i, j) = m.on_p(i, j) * m.on_p(i, j) * cff;
    }
  for (int j = Jstr; j <= Jend; ++j)
    for (int i = IstrU; i <= Iend; ++i) {
      double cff1 = 0.5 * (pn(i - 1, j) + pn(i, j)) * (UFx(i, j) - UFx(i - 1, j));
      double cff2 = 0.5 * (pm(i - 1, j) + pm(i, j)) * (UFe(i, j + 1) - UFe(i, j));
      double fc = cff1 + cff2;
      rhs_ubar(i, j) = rhs_ubar(i, j) + fc;
    }
  for (int j = JstrV; j <= Jend; ++j)
    for (int i = Istr; i <= Iend; ++i) {
      double cff1 = 0.5 * (pn(i, j - 1) + pn(i, j)) * (VFx(i + 1, j) - VFx(i, j));
      double cff2 = 0.5 * (pm(i, j - 1) + pm(i, j)) * (VFe(i, j) - VFe(i, j - 1));
      double fc = cff1 - cff2;
      rhs_vbar(i, j) = rhs_vbar(i, j) + fc;
    }

  // ---- :1884-2065  coupling with the 3-D momentum equations
  if (iif == 1 && PREDICTOR) {
    F3 ru_s = m.ru[nstp], ru_n = m.ru[nnew], rv_s = m.rv[nstp], rv_n = m.rv[nnew];
    if (m.iic == m.ntfirst) {
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU; i <= Iend; ++i) {
          m.rufrc(i, j) = m.rufrc(i, j) - rhs_ubar(i, j);
          rhs_ubar(i, j) = rhs_ubar(i, j) + m.rufrc(i, j);
          ru_s(i, j, 0) = m.rufrc(i, j);
        }
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          m.rvfrc(i, j) = m.rvfrc(i, j) - rhs_vbar(i, j);
          rhs_vbar(i, j) = rhs_vbar(i, j) + m.rvfrc(i, j);
          rv_s(i, j, 0) = m.rvfrc(i, j);
        }
    } else if (m.iic == m.ntfirst + 1) {
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU; i <= Iend; ++i) {
          m.rufrc(i, j) = m.rufrc(i, j) - rhs_ubar(i, j);
          rhs_ubar(i, j) = rhs_ubar(i, j) + 1.5 * m.rufrc(i, j) - 0.5 * ru_n(i, j, 0);
          ru_s(i, j, 0) = m.rufrc(i, j);
        }
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          m.rvfrc(i, j) = m.rvfrc(i, j) - rhs_vbar(i, j);
          rhs_vbar(i, j) = rhs_vbar(i, j) + 1.5 * m.rvfrc(i, j) - 0.5 * rv_n(i, j, 0);
          rv_s(i, j, 0) = m.rvfrc(i, j);
        }
    } else {
      const double cff1 = 23.0 / 12.0, cff2 = 16.0 / 12.0, cff3 = 5.0 / 12.0;
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU; i <= Iend; ++i) {
          m.rufrc(i, j) = m.rufrc(i, j) - rhs_ubar(i, j);
          rhs_ubar(i, j) = rhs_ubar(i, j) + cff1 * m.rufrc(i, j) - cff2 * ru_n(i, j, 0) + cff3 * ru_s(i, j, 0);
          ru_s(i, j, 0) = m.rufrc(i, j);
        }
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          m.rvfrc(i, j) = m.rvfrc(i, j) - rhs_vbar(i, j);
          rhs_vbar(i, j) = rhs_vbar(i, j) + cff1 * m.rvfrc(i, j) - cff2 * rv_n(i, j, 0) + cff3 * rv_s(i, j, 0);
          rv_s(i, j, 0) = m.rvfrc(i, j);
        }
    }
  } else {
    for (int j = Jstr; j <= Jend; ++j) for (int i = IstrU; i <= Iend; ++i) rhs_ubar(i, j) = rhs_ubar(i, j) + m.rufrc(i, j);
    for (int j = JstrV; j <= Jend; ++j) for (int i = Istr; i <= Iend; ++i) rhs_vbar(i, j) = rhs_vbar(i, j) + m.rvfrc(i, j);
  }

  // ---- :2098-2255  time step the 2-D momentum equations
  for (int j = JstrV - 1; j <= Jend; ++j) for (int i = IstrU - 1; i <= Iend; ++i) Dstp(i, j) = zeta_s(i, j) + h(i, j);
  if (iif == 1 || PREDICTOR) {
    double cff1 = (iif == 1) ? 0.5 * dtfast : dtfast;
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff = (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
        double fc = 1.0 / (Dnew(i, j) + Dnew(i - 1, j));
        ubar_n(i, j) = (ubar_s(i, j) * (Dstp(i, j) + Dstp(i - 1, j)) + cff * cff1 * rhs_ubar(i, j)) * fc;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
        double fc = 1.0 / (Dnew(i, j) + Dnew(i, j - 1));
        vbar_n(i, j) = (vbar_s(i, j) * (Dstp(i, j) + Dstp(i, j - 1)) + cff * cff1 * rhs_vbar(i, j)) * fc;
      }
  } else {
    double cff1 = 0.5 * dtfast * 5.0 / 12.0, cff2 = 0.5 * dtfast * 8.0 / 12.0, cff3 = 0.5 * dtfast * 1.0 / 12.0;
    F2 rub_s = m.rubar[kstp], rub_p = m.rubar[ptsk], rvb_s = m.rvbar[kstp], rvb_p = m.rvbar[ptsk];
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff = (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
        double fc = 1.0 / (Dnew(i, j) + Dnew(i - 1, j));
        ubar_n(i, j) = (ubar_s(i, j) * (Dstp(i, j) + Dstp(i - 1, j)) + cff * (cff1 * rhs_ubar(i, j) + cff2 * rub_s(i, j) - cff3 * rub_p(i, j))) * fc;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
        double fc = 1.0 / (Dnew(i, j) + Dnew(i, j - 1));
        vbar_n(i, j) = (vbar_s(i, j) * (Dstp(i, j) + Dstp(i, j - 1)) + cff * (cff1 * rhs_vbar(i, j) + cff2 * rvb_s(i, j) - cff3 * rvb_p(i, j))) * fc;
      }
  }
  // ---- :2420-2430
  if (PREDICTOR) {
    F2 rub = m.rubar[krhs], rvb = m.rvbar[krhs];
    for (int j = Jstr; j <= Jend; ++j) for (int i = IstrU; i <= Iend; ++i) rub(i, j) = rhs_ubar(i, j);
    for (int j = JstrV; j <= Jend; ++j) for (int i = Istr; i <= Iend; ++i) rvb(i, j) = rhs_vbar(i, j);
  }
  // ---- :2451-2524
  u2dbc(m, b, knew); v2dbc(m, b, knew);
  exchange_u2d(m, b, ubar_n); exchange_v2d(m, b, vbar_n);
}

}  // namespace orc

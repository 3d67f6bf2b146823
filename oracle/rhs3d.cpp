// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/rhs3d.F:174-1671 (rhs3d_tile): Coriolis :473-507, curvilinear :515-564, third-order upstream
// horizontal advection (default branch) :658-983, fourth-order centred vertical advection :1177-1265 / :1434-1522,
// column sums -> rufrc,rvfrc :1534-1667.  uv_adv = 1 (UV_C4ADVECTION): fourth-order centred horizontal fluxes :685-705,
// :761-781, :829-849, :902-921 and the 9/32, 1/32 vertical flux :1108-1175 / :1362-1429.
#include "roms_oracle.hpp"

namespace orc {

void rhs3d(Model& m, const Bnd& b) {
  const Cfg& c = m.c; ORC_UNPACK_BOUNDS(b);
  const int N = c.N;
  const double Gadv = -0.25;                                         // rhs3d.F:299
  const bool c4 = (c.uv_adv == 1);                                   // UV_C4ADVECTION
  const bool c2 = (c.uv_adv == 3);                                   // UV_C2ADVECTION
  const bool sadv = (c.uv_adv == 2);                                 // UV_SADVECTION (horizontal: the default branch)
  SK CF(IminS, ImaxS, 0, N), DC(IminS, ImaxS, 0, N);
  F3 &Hz = m.Hz, &Huon = m.Huon, &Hvom = m.Hvom, &W = m.W;
  F3 u = m.u[m.nrhs], v = m.v[m.nrhs], ru = m.ru[m.nrhs], rv = m.rv[m.nrhs];
  SK FC(IminS, ImaxS, 0, N);
  S2 Huee(IminS, ImaxS, JminS, JmaxS), Huxx(IminS, ImaxS, JminS, JmaxS), Hvee(IminS, ImaxS, JminS, JmaxS), Hvxx(IminS, ImaxS, JminS, JmaxS),
      UFx(IminS, ImaxS, JminS, JmaxS), UFe(IminS, ImaxS, JminS, JmaxS), VFx(IminS, ImaxS, JminS, JmaxS), VFe(IminS, ImaxS, JminS, JmaxS),
      uee(IminS, ImaxS, JminS, JmaxS), uxx(IminS, ImaxS, JminS, JmaxS), vee(IminS, ImaxS, JminS, JmaxS), vxx(IminS, ImaxS, JminS, JmaxS);

  if (c.bodyforce) {                                                 // BODYFORCE :326-466
    S2 wrk(IminS, ImaxS, JminS, JmaxS), Uwrk(IminS, ImaxS, JminS, JmaxS), Vwrk(IminS, ImaxS, JminS, JmaxS);
    F2 &pm = m.pm, &pn = m.pn;
    for (int pass = 0; pass < 2; ++pass) {                           // 0: surface stress over levsfrc:N (added), 1: bottom stress over 1:levbfrc (removed)
      const int k0 = pass ? 1 : c.levsfrc, k1 = pass ? c.levbfrc : N;
      for (int j = JstrV - 1; j <= Jend; ++j) for (int i = IstrU - 1; i <= Iend; ++i) wrk(i, j) = 0.0;
      if (pass == 0) { for (int k = N; k >= c.levsfrc; --k) for (int j = JstrV - 1; j <= Jend; ++j) for (int i = IstrU - 1; i <= Iend; ++i) wrk(i, j) = wrk(i, j) + Hz(i, j, k); }
      else { for (int k = 1; k <= c.levbfrc; ++k) for (int j = JstrV - 1; j <= Jend; ++j) for (int i = IstrU - 1; i <= Iend; ++i) wrk(i, j) = wrk(i, j) + Hz(i, j, k); }
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU; i <= Iend; ++i) {
          const double cff = 0.25 * (pm(i - 1, j) + pm(i, j)) * (pn(i - 1, j) + pn(i, j));
          const double cff1 = 1.0 / (cff * (wrk(i - 1, j) + wrk(i, j)));
          Uwrk(i, j) = (pass ? m.bustr(i, j) : m.sustr(i, j)) * cff1;
        }
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) {
          const double cff = 0.25 * (pm(i, j - 1) + pm(i, j)) * (pn(i, j - 1) + pn(i, j));
          const double cff1 = 1.0 / (cff * (wrk(i, j - 1) + wrk(i, j)));
          Vwrk(i, j) = (pass ? m.bvstr(i, j) : m.svstr(i, j)) * cff1;
        }
      for (int k = k0; k <= k1; ++k) {
        for (int j = Jstr; j <= Jend; ++j)
          for (int i = IstrU; i <= Iend; ++i) {
            const double cff = Uwrk(i, j) * (Hz(i, j, k) + Hz(i - 1, j, k));
            ru(i, j, k) = pass ? ru(i, j, k) - cff : ru(i, j, k) + cff;
          }
        for (int j = JstrV; j <= Jend; ++j)
          for (int i = Istr; i <= Iend; ++i) {
            const double cff = Vwrk(i, j) * (Hz(i, j, k) + Hz(i, j - 1, k));
            rv(i, j, k) = pass ? rv(i, j, k) - cff : rv(i, j, k) + cff;
          }
      }
    }
  }
  for (int k = 1; k <= N; ++k) {
    // ---- UV_COR :473-507
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff = 0.5 * Hz(i, j, k) * m.fomn(i, j);
        UFx(i, j) = cff * (v(i, j, k) + v(i, j + 1, k));
        VFe(i, j) = cff * (u(i, j, k) + u(i + 1, j, k));
      }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) { double cff1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); ru(i, j, k) = ru(i, j, k) + cff1; }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) { double cff1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rv(i, j, k) = rv(i, j, k) - cff1; }
    // ---- CURVGRID && UV_ADV :515-564
    if (c.curvgrid) {
      for (int j = JstrV - 1; j <= Jend; ++j)
        for (int i = IstrU - 1; i <= Iend; ++i) {
          double cff1 = 0.5 * (v(i, j, k) + v(i, j + 1, k));
          double cff2 = 0.5 * (u(i, j, k) + u(i + 1, j, k));
          double cff3 = cff1 * m.dndx(i, j);
          double cff4 = cff2 * m.dmde(i, j);
          double cff = Hz(i, j, k) * (cff3 - cff4);
          UFx(i, j) = cff * cff1;
          VFe(i, j) = cff * cff2;
        }
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU; i <= Iend; ++i) { double cff1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); ru(i, j, k) = ru(i, j, k) + cff1; }
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) { double cff1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rv(i, j, k) = rv(i, j, k) - cff1; }
    }
    if (c2) {                                                                                         // :605-657
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU - 1; i <= Iend; ++i) UFx(i, j) = 0.25 * (u(i, j, k) + u(i + 1, j, k)) * (Huon(i, j, k) + Huon(i + 1, j, k));
      for (int j = Jstr; j <= Jend + 1; ++j)
        for (int i = IstrU; i <= Iend; ++i) UFe(i, j) = 0.25 * (u(i, j - 1, k) + u(i, j, k)) * (Hvom(i - 1, j, k) + Hvom(i, j, k));
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend + 1; ++i) VFx(i, j) = 0.25 * (v(i - 1, j, k) + v(i, j, k)) * (Huon(i, j - 1, k) + Huon(i, j, k));
      for (int j = JstrV - 1; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i) VFe(i, j) = 0.25 * (v(i, j, k) + v(i, j + 1, k)) * (Hvom(i, j, k) + Hvom(i, j + 1, k));
    } else {
    // ---- UV_ADV, third-order upstream :658-983
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrUm1; i <= Iendp1; ++i) {
        uxx(i, j) = u(i - 1, j, k) - 2.0 * u(i, j, k) + u(i + 1, j, k);
        Huxx(i, j) = Huon(i - 1, j, k) - 2.0 * Huon(i, j, k) + Huon(i + 1, j, k);
      }
    // (:669-684 closed E/W wall copies: not live, EW periodic)
    if (c4) {                                                                                         // :685-705
      const double cff = 1.0 / 6.0;
      for (int j = Jstr; j <= Jend; ++j)
        for (int i = IstrU - 1; i <= Iend; ++i)
          UFx(i, j) = 0.25 * (u(i, j, k) + u(i + 1, j, k) - cff * (uxx(i, j) + uxx(i + 1, j))) *
                      (Huon(i, j, k) + Huon(i + 1, j, k) - cff * (Huxx(i, j) + Huxx(i + 1, j)));
    } else
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) {
        double cff1 = u(i, j, k) + u(i + 1, j, k);
        double cff = (cff1 > 0.0) ? uxx(i, j) : uxx(i + 1, j);
        UFx(i, j) = 0.25 * (cff1 + Gadv * cff) * (Huon(i, j, k) + Huon(i + 1, j, k) + Gadv * 0.5 * (Huxx(i, j) + Huxx(i + 1, j)));
      }
    for (int j = Jstrm1; j <= Jendp1; ++j)
      for (int i = IstrU; i <= Iend; ++i) uee(i, j) = u(i, j - 1, k) - 2.0 * u(i, j, k) + u(i, j + 1, k);
    if (b.Southern_Edge) for (int i = IstrU; i <= Iend; ++i) uee(i, Jstr - 1) = uee(i, Jstr);     // :742-748
    if (b.Northern_Edge) for (int i = IstrU; i <= Iend; ++i) uee(i, Jend + 1) = uee(i, Jend);     // :749-755
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = IstrU - 1; i <= Iend; ++i) Hvxx(i, j) = Hvom(i - 1, j, k) - 2.0 * Hvom(i, j, k) + Hvom(i + 1, j, k);
    if (c4) {                                                                                         // :761-781
      const double cff = 1.0 / 6.0;
      for (int j = Jstr; j <= Jend + 1; ++j)
        for (int i = IstrU; i <= Iend; ++i)
          UFe(i, j) = 0.25 * (u(i, j, k) + u(i, j - 1, k) - cff * (uee(i, j) + uee(i, j - 1))) *
                      (Hvom(i, j, k) + Hvom(i - 1, j, k) - cff * (Hvxx(i, j) + Hvxx(i - 1, j)));
    } else
    for (int j = Jstr; j <= Jend + 1; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff1 = u(i, j, k) + u(i, j - 1, k);
        double cff2 = Hvom(i, j, k) + Hvom(i - 1, j, k);
        double cff = (cff2 > 0.0) ? uee(i, j - 1) : uee(i, j);
        UFe(i, j) = 0.25 * (cff1 + Gadv * cff) * (cff2 + Gadv * 0.5 * (Hvxx(i, j) + Hvxx(i - 1, j)));
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istrm1; i <= Iendp1; ++i) vxx(i, j) = v(i - 1, j, k) - 2.0 * v(i, j, k) + v(i + 1, j, k);
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = Istr; i <= Iend + 1; ++i) Huee(i, j) = Huon(i, j - 1, k) - 2.0 * Huon(i, j, k) + Huon(i, j + 1, k);
    if (c4) {                                                                                         // :829-849
      const double cff = 1.0 / 6.0;
      for (int j = JstrV; j <= Jend; ++j)
        for (int i = Istr; i <= Iend + 1; ++i)
          VFx(i, j) = 0.25 * (v(i, j, k) + v(i - 1, j, k) - cff * (vxx(i, j) + vxx(i - 1, j))) *
                      (Huon(i, j, k) + Huon(i, j - 1, k) - cff * (Huee(i, j) + Huee(i, j - 1)));
    } else
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend + 1; ++i) {
        double cff1 = v(i, j, k) + v(i - 1, j, k);
        double cff2 = Huon(i, j, k) + Huon(i, j - 1, k);
        double cff = (cff2 > 0.0) ? vxx(i - 1, j) : vxx(i, j);
        VFx(i, j) = 0.25 * (cff1 + Gadv * cff) * (cff2 + Gadv * 0.5 * (Huee(i, j) + Huee(i, j - 1)));
      }
    for (int j = JstrVm1; j <= Jendp1; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        vee(i, j) = v(i, j - 1, k) - 2.0 * v(i, j, k) + v(i, j + 1, k);
        Hvee(i, j) = Hvom(i, j - 1, k) - 2.0 * Hvom(i, j, k) + Hvom(i, j + 1, k);
      }
    if (b.Southern_Edge) for (int i = Istr; i <= Iend; ++i) { vee(i, Jstr) = vee(i, Jstr + 1); Hvee(i, Jstr) = Hvee(i, Jstr + 1); }   // :886-893
    if (b.Northern_Edge) for (int i = Istr; i <= Iend; ++i) { vee(i, Jend + 1) = vee(i, Jend); Hvee(i, Jend + 1) = Hvee(i, Jend); }   // :894-901
    if (c4) {                                                                                         // :902-921
      const double cff = 1.0 / 6.0;
      for (int j = JstrV - 1; j <= Jend; ++j)
        for (int i = Istr; i <= Iend; ++i)
          VFe(i, j) = 0.25 * (v(i, j, k) + v(i, j + 1, k) - cff * (vee(i, j) + vee(i, j + 1))) *
                      (Hvom(i, j, k) + Hvom(i, j + 1, k) - cff * (Hvee(i, j) + Hvee(i, j + 1)));
    } else
    for (int j = JstrV - 1; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff1 = v(i, j, k) + v(i, j + 1, k);
        double cff = (cff1 > 0.0) ? vee(i, j) : vee(i, j + 1);
        VFe(i, j) = 0.25 * (cff1 + Gadv * cff) * (Hvom(i, j, k) + Hvom(i, j + 1, k) + Gadv * 0.5 * (Hvee(i, j) + Hvee(i, j + 1)));
      }
    }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = IstrU; i <= Iend; ++i) {
        double cff1 = UFx(i, j) - UFx(i - 1, j);
        double cff2 = UFe(i, j + 1) - UFe(i, j);
        double cff = cff1 + cff2;
        ru(i, j, k) = ru(i, j, k) - cff;
      }
    for (int j = JstrV; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        double cff1 = VFx(i + 1, j) - VFx(i, j);
        double cff2 = VFe(i, j) - VFe(i, j - 1);
        double cff = cff1 + cff2;
        rv(i, j, k) = rv(i, j, k) - cff;
      }
  }

  // ---- vertical advection + column sums
  for (int j = Jstr; j <= Jend; ++j) {
    if (c2) {                                                                                         // :1079-1107
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) FC(i, k) = 0.25 * (u(i, j, k) + u(i, j, k + 1)) * (W(i, j, k) + W(i - 1, j, k));
      for (int i = IstrU; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); ru(i, j, k) = ru(i, j, k) - cff; }
    } else if (sadv) {                                                                                // :1016-1078: parabolic splines
      const double cff1 = 9.0 / 16.0, cff2 = 1.0 / 16.0;
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i)
          DC(i, k) = cff1 * (Hz(i, j, k) + Hz(i - 1, j, k)) - cff2 * (Hz(i + 1, j, k) + Hz(i - 2, j, k));
      for (int i = IstrU; i <= Iend; ++i) { FC(i, 0) = 0.0; CF(i, 0) = 0.0; }
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i) {
          double cff = 1.0 / (2.0 * DC(i, k + 1) + DC(i, k) * (2.0 - FC(i, k - 1)));
          FC(i, k) = cff * DC(i, k + 1);
          CF(i, k) = cff * (6.0 * (u(i, j, k + 1) - u(i, j, k)) - DC(i, k) * CF(i, k - 1));
        }
      for (int i = IstrU; i <= Iend; ++i) CF(i, N) = 0.0;
      for (int k = N - 1; k >= 1; --k)
        for (int i = IstrU; i <= Iend; ++i) CF(i, k) = CF(i, k) - FC(i, k) * CF(i, k + 1);
      const double cff3 = 1.0 / 3.0, cff4 = 1.0 / 6.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = IstrU; i <= Iend; ++i)
          FC(i, k) = (cff1 * (W(i, j, k) + W(i - 1, j, k)) - cff2 * (W(i + 1, j, k) + W(i - 2, j, k))) *
                     (u(i, j, k) + DC(i, k) * (cff3 * CF(i, k) + cff4 * CF(i, k - 1)));
      for (int i = IstrU; i <= Iend; ++i) { FC(i, N) = 0.0; FC(i, 0) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); ru(i, j, k) = ru(i, j, k) - cff; }
    } else if (c4) {                                                                                  // :1108-1175
      const double cff1 = 9.0 / 32.0, cff2 = 1.0 / 32.0;
      for (int k = 2; k <= N - 2; ++k)
        for (int i = IstrU; i <= Iend; ++i)
          FC(i, k) = (cff1 * (u(i, j, k) + u(i, j, k + 1)) - cff2 * (u(i, j, k - 1) + u(i, j, k + 2))) * (W(i, j, k) + W(i - 1, j, k));
      for (int i = IstrU; i <= Iend; ++i) {
        FC(i, N) = 0.0;
        FC(i, N - 1) = (cff1 * (u(i, j, N - 1) + u(i, j, N)) - cff2 * (u(i, j, N - 2) + u(i, j, N))) * (W(i, j, N - 1) + W(i - 1, j, N - 1));
        FC(i, 1) = (cff1 * (u(i, j, 1) + u(i, j, 2)) - cff2 * (u(i, j, 1) + u(i, j, 3))) * (W(i, j, 1) + W(i - 1, j, 1));
        FC(i, 0) = 0.0;
      }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); ru(i, j, k) = ru(i, j, k) - cff; }
    } else {
      const double cff1 = 9.0 / 16.0, cff2 = 1.0 / 16.0;
      for (int k = 2; k <= N - 2; ++k)
        for (int i = IstrU; i <= Iend; ++i)
          FC(i, k) = (cff1 * (u(i, j, k) + u(i, j, k + 1)) - cff2 * (u(i, j, k - 1) + u(i, j, k + 2))) *
                     (cff1 * (W(i, j, k) + W(i - 1, j, k)) - cff2 * (W(i + 1, j, k) + W(i - 2, j, k)));
      for (int i = IstrU; i <= Iend; ++i) {
        FC(i, N) = 0.0;
        FC(i, N - 1) = (cff1 * (u(i, j, N - 1) + u(i, j, N)) - cff2 * (u(i, j, N - 2) + u(i, j, N))) *
                       (cff1 * (W(i, j, N - 1) + W(i - 1, j, N - 1)) - cff2 * (W(i + 1, j, N - 1) + W(i - 2, j, N - 1)));
        FC(i, 1) = (cff1 * (u(i, j, 1) + u(i, j, 2)) - cff2 * (u(i, j, 1) + u(i, j, 3))) *
                   (cff1 * (W(i, j, 1) + W(i - 1, j, 1)) - cff2 * (W(i + 1, j, 1) + W(i - 2, j, 1)));
        FC(i, 0) = 0.0;
      }
      for (int k = 1; k <= N; ++k)
        for (int i = IstrU; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); ru(i, j, k) = ru(i, j, k) - cff; }
    }
    if (j >= JstrV && c2) {                                                                           // :1330-1361
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) FC(i, k) = 0.25 * (v(i, j, k) + v(i, j, k + 1)) * (W(i, j, k) + W(i, j - 1, k));
      for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; FC(i, N) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); rv(i, j, k) = rv(i, j, k) - cff; }
    } else if (j >= JstrV && sadv) {                                                                  // :1267-1329
      const double cff1 = 9.0 / 16.0, cff2 = 1.0 / 16.0;
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i)
          DC(i, k) = (cff1 * (Hz(i, j, k) + Hz(i, j - 1, k)) - cff2 * (Hz(i, j + 1, k) + Hz(i, j - 2, k)));
      for (int i = Istr; i <= Iend; ++i) { FC(i, 0) = 0.0; CF(i, 0) = 0.0; }
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i) {
          double cff = 1.0 / (2.0 * DC(i, k + 1) + DC(i, k) * (2.0 - FC(i, k - 1)));
          FC(i, k) = cff * DC(i, k + 1);
          CF(i, k) = cff * (6.0 * (v(i, j, k + 1) - v(i, j, k)) - DC(i, k) * CF(i, k - 1));
        }
      for (int i = Istr; i <= Iend; ++i) CF(i, N) = 0.0;
      for (int k = N - 1; k >= 1; --k)
        for (int i = Istr; i <= Iend; ++i) CF(i, k) = CF(i, k) - FC(i, k) * CF(i, k + 1);
      const double cff3 = 1.0 / 3.0, cff4 = 1.0 / 6.0;
      for (int k = 1; k <= N - 1; ++k)
        for (int i = Istr; i <= Iend; ++i)
          FC(i, k) = (cff1 * (W(i, j, k) + W(i, j - 1, k)) - cff2 * (W(i, j + 1, k) + W(i, j - 2, k))) *
                     (v(i, j, k) + DC(i, k) * (cff3 * CF(i, k) + cff4 * CF(i, k - 1)));
      for (int i = Istr; i <= Iend; ++i) { FC(i, N) = 0.0; FC(i, 0) = 0.0; }
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); rv(i, j, k) = rv(i, j, k) - cff; }
    } else if (j >= JstrV && c4) {                                                                    // :1362-1429
      const double cff1 = 9.0 / 32.0, cff2 = 1.0 / 32.0;
      for (int k = 2; k <= N - 2; ++k)
        for (int i = Istr; i <= Iend; ++i)
          FC(i, k) = (cff1 * (v(i, j, k) + v(i, j, k + 1)) - cff2 * (v(i, j, k - 1) + v(i, j, k + 2))) * (W(i, j, k) + W(i, j - 1, k));
      for (int i = Istr; i <= Iend; ++i) {
        FC(i, N) = 0.0;
        FC(i, N - 1) = (cff1 * (v(i, j, N - 1) + v(i, j, N)) - cff2 * (v(i, j, N - 2) + v(i, j, N))) * (W(i, j, N - 1) + W(i, j - 1, N - 1));
        FC(i, 1) = (cff1 * (v(i, j, 1) + v(i, j, 2)) - cff2 * (v(i, j, 1) + v(i, j, 3))) * (W(i, j, 1) + W(i, j - 1, 1));
        FC(i, 0) = 0.0;
      }
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); rv(i, j, k) = rv(i, j, k) - cff; }
    } else if (j >= JstrV) {
      const double cff1 = 9.0 / 16.0, cff2 = 1.0 / 16.0;
      for (int k = 2; k <= N - 2; ++k)
        for (int i = Istr; i <= Iend; ++i)
          FC(i, k) = (cff1 * (v(i, j, k) + v(i, j, k + 1)) - cff2 * (v(i, j, k - 1) + v(i, j, k + 2))) *
                     (cff1 * (W(i, j, k) + W(i, j - 1, k)) - cff2 * (W(i, j + 1, k) + W(i, j - 2, k)));
      for (int i = Istr; i <= Iend; ++i) {
        FC(i, N) = 0.0;
        FC(i, N - 1) = (cff1 * (v(i, j, N - 1) + v(i, j, N)) - cff2 * (v(i, j, N - 2) + v(i, j, N))) *
                       (cff1 * (W(i, j, N - 1) + W(i, j - 1, N - 1)) - cff2 * (W(i, j + 1, N - 1) + W(i, j - 2, N - 1)));
        FC(i, 1) = (cff1 * (v(i, j, 1) + v(i, j, 2)) - cff2 * (v(i, j, 1) + v(i, j, 3))) *
                   (cff1 * (W(i, j, 1) + W(i, j - 1, 1)) - cff2 * (W(i, j + 1, 1) + W(i, j - 2, 1)));
        FC(i, 0) = 0.0;
      }
      for (int k = 1; k <= N; ++k)
        for (int i = Istr; i <= Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); rv(i, j, k) = rv(i, j, k) - cff; }
    }
    // ---- :1534-1598
    for (int i = IstrU; i <= Iend; ++i) m.rufrc(i, j) = ru(i, j, 1);
    for (int k = 2; k <= N; ++k) for (int i = IstrU; i <= Iend; ++i) m.rufrc(i, j) = m.rufrc(i, j) + ru(i, j, k);
    if (!c.bodyforce)                                                 // # ifndef BODYFORCE :1588-1599
    for (int i = IstrU; i <= Iend; ++i) {
      double cff = m.om_u(i, j) * m.on_u(i, j);
      double cff1 = m.sustr(i, j) * cff;
      double cff2 = -m.bustr(i, j) * cff;
      m.rufrc(i, j) = m.rufrc(i, j) + cff1 + cff2;
    }
    if (j >= JstrV) {   // :1600-1667
      for (int i = Istr; i <= Iend; ++i) m.rvfrc(i, j) = rv(i, j, 1);
      for (int k = 2; k <= N; ++k) for (int i = Istr; i <= Iend; ++i) m.rvfrc(i, j) = m.rvfrc(i, j) + rv(i, j, k);
      if (!c.bodyforce)
      for (int i = Istr; i <= Iend; ++i) {
        double cff = m.om_v(i, j) * m.on_v(i, j);
        double cff1 = m.svstr(i, j) * cff;
        double cff2 = -m.bvstr(i, j) * cff;
        m.rvfrc(i, j) = m.rvfrc(i, j) + cff1 + cff2;
      }
    }
  }
}

}  // namespace orc

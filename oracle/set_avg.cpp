// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/set_avg.F (AVERAGES): time-averaged output fields, for the state variables of the nonlinear chain --
// avgzeta (:280-290, :1308-1318, :2340-2362), avgu2d / avgv2d (:292-312, :1320-1345, :2364-2412), avgu3d / avgv3d
// (:330-354, :1358-1382, :2430-2474), avgw3d = W*pm*pn (:371-385, :1399-1413, :2505-2527), avgwvel (:386-398, :1414-1426,
// :2529-2551), avgrho (:400-412, :1428-1440, :2553-2575), avgt (:414-426, :1442-1454, :2576-2598).  Aout(...) = T for these.
// KOUT = kstp, NOUT = nrhs (globaldefs.h:504-508); called from main3d.F:494, right after set_zeta.
#include "roms_oracle.hpp"

namespace orc {

void set_avg(Model& m, const Bnd& b) {
  const Cfg& c = m.c;
  const int nAVG = c.nAVG, ntsAVG = c.ntsAVG, N = c.N;
  if (nAVG == 0) return;                                                             // :189
  ORC_UNPACK_BOUNDS(b);
  const int Kout = m.kstp, Nout = m.nrhs;
  const int iic = m.iic;
  const int nrrec = 0;
  F2 zeta = m.zeta[Kout], ubar = m.ubar[Kout], vbar = m.vbar[Kout];
  F3 u = m.u[Nout], v = m.v[Nout];
  if (((iic > ntsAVG) && ((iic - 1) % nAVG == 1)) || ((iic >= ntsAVG) && (nAVG == 1)) || ((nrrec > 0) && (iic == m.ntstart))) {   // :237-240
    // initialise
    for (int j = JstrR; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgzeta(i, j) = zeta(i, j);
    for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu2d(i, j) = ubar(i, j);
    for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv2d(i, j) = vbar(i, j);
    for (int k = 1; k <= N; ++k) {
      for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu3d(i, j, k) = u(i, j, k);
      for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv3d(i, j, k) = v(i, j, k);
    }
    for (int k = 0; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.avgw3d(i, j, k) = m.W(i, j, k) * m.pm(i, j) * m.pn(i, j);
          m.avgwvel(i, j, k) = m.wvel(i, j, k);
        }
    for (int k = 1; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.avgrho(i, j, k) = m.rho(i, j, k);
          for (int it = 0; it < c.NT; ++it) m.avgt[it](i, j, k) = m.t[Nout][it](i, j, k);
        }
  } else if (iic > ntsAVG) {                                                         // :1264
    // accumulate
    for (int j = JstrR; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgzeta(i, j) = m.avgzeta(i, j) + zeta(i, j);
    for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu2d(i, j) = m.avgu2d(i, j) + ubar(i, j);
    for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv2d(i, j) = m.avgv2d(i, j) + vbar(i, j);
    for (int k = 1; k <= N; ++k) {
      for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu3d(i, j, k) = m.avgu3d(i, j, k) + u(i, j, k);
      for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv3d(i, j, k) = m.avgv3d(i, j, k) + v(i, j, k);
    }
    for (int k = 0; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.avgw3d(i, j, k) = m.avgw3d(i, j, k) + m.W(i, j, k) * m.pm(i, j) * m.pn(i, j);
          m.avgwvel(i, j, k) = m.avgwvel(i, j, k) + m.wvel(i, j, k);
        }
    for (int k = 1; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.avgrho(i, j, k) = m.avgrho(i, j, k) + m.rho(i, j, k);
          for (int it = 0; it < c.NT; ++it) m.avgt[it](i, j, k) = m.avgt[it](i, j, k) + m.t[Nout][it](i, j, k);
        }
  }
  // convert the sums into averages when the window closes (:2298-2301), then the periodic copies
  if (((iic > ntsAVG) && ((iic - 1) % nAVG == 0) && ((iic != m.ntstart) || (nrrec == 0))) || ((iic >= ntsAVG) && (nAVG == 1))) {
    const double fac = 1.0 / (double)nAVG;                                           // :2327
    for (int j = JstrR; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgzeta(i, j) = fac * m.avgzeta(i, j);
    exchange_r2d(m, b, m.avgzeta);
    for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu2d(i, j) = fac * m.avgu2d(i, j);
    exchange_u2d(m, b, m.avgu2d);
    for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv2d(i, j) = fac * m.avgv2d(i, j);
    exchange_v2d(m, b, m.avgv2d);
    for (int k = 1; k <= N; ++k) {
      for (int j = JstrR; j <= JendR; ++j) for (int i = Istr; i <= IendR; ++i) m.avgu3d(i, j, k) = fac * m.avgu3d(i, j, k);
      for (int j = Jstr; j <= JendR; ++j) for (int i = IstrR; i <= IendR; ++i) m.avgv3d(i, j, k) = fac * m.avgv3d(i, j, k);
    }
    exchange_u3d(m, b, m.avgu3d); exchange_v3d(m, b, m.avgv3d);
    for (int k = 0; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) { m.avgw3d(i, j, k) = fac * m.avgw3d(i, j, k); m.avgwvel(i, j, k) = fac * m.avgwvel(i, j, k); }
    exchange_w3d(m, b, m.avgw3d); exchange_w3d(m, b, m.avgwvel);
    for (int k = 1; k <= N; ++k)
      for (int j = JstrR; j <= JendR; ++j)
        for (int i = IstrR; i <= IendR; ++i) {
          m.avgrho(i, j, k) = fac * m.avgrho(i, j, k);
          for (int it = 0; it < c.NT; ++it) m.avgt[it](i, j, k) = fac * m.avgt[it](i, j, k);
        }
    exchange_r3d(m, b, m.avgrho);
    for (int it = 0; it < c.NT; ++it) exchange_r3d(m, b, m.avgt[it]);
  }
}

}  // namespace orc

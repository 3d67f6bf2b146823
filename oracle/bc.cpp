// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).  Periodic ghost copies and the closed-wall
// branches of the physical boundary conditions (all five configurations are EW-periodic / NS-closed:
// ROMS/External/roms_{upwelling,seamount,benchmark1}.in LBC block).
#include "roms_oracle.hpp"

namespace orc {

// ROMS/Nonlinear/exchange_2d.F: r2d :229-412 (J range :286-293, copies :295-314); u2d :416-599; v2d :603-786;
// p2d :42-225.  NghostPoints = 2.  Only the EWperiodic && !NSperiodic path is live.
static void ew_copy2(const Model& m, const Bnd& b, F2 A, int Jmin, int Jmax) {
  const int Lm = m.c.Lm;
  if (b.Western_Edge)
    for (int j = Jmin; j <= Jmax; ++j) { A(Lm + 1, j) = A(1, j); A(Lm + 2, j) = A(2, j); }
  if (b.Eastern_Edge)
    for (int j = Jmin; j <= Jmax; ++j) { A(-2, j) = A(Lm - 2, j); A(-1, j) = A(Lm - 1, j); A(0, j) = A(Lm, j); }
}
void exchange_r2d(const Model& m, const Bnd& b, F2 A) { ew_copy2(m, b, A, b.JstrR, b.JendR); }
void exchange_u2d(const Model& m, const Bnd& b, F2 A) { ew_copy2(m, b, A, b.JstrR, b.JendR); }
void exchange_v2d(const Model& m, const Bnd& b, F2 A) { ew_copy2(m, b, A, b.Jstr, b.JendR); }
void exchange_p2d(const Model& m, const Bnd& b, F2 A) { ew_copy2(m, b, A, b.Jstr, b.JendR); }
// ROMS/Nonlinear/exchange_3d.F: r3d :259-468, u3d :471-679, v3d :683-892, w3d :896-1105 (same copies per k)
static void ew_copy3(const Model& m, const Bnd& b, F3 A, int Jmin, int Jmax) {
  for (int k = A.LBk; k <= A.UBk; ++k) ew_copy2(m, b, A.plane(k), Jmin, Jmax);
}
void exchange_r3d(const Model& m, const Bnd& b, F3 A) { ew_copy3(m, b, A, b.JstrR, b.JendR); }
void exchange_u3d(const Model& m, const Bnd& b, F3 A) { ew_copy3(m, b, A, b.JstrR, b.JendR); }
void exchange_v3d(const Model& m, const Bnd& b, F3 A) { ew_copy3(m, b, A, b.Jstr, b.JendR); }
void exchange_w3d(const Model& m, const Bnd& b, F3 A) { ew_copy3(m, b, A, b.JstrR, b.JendR); }

// ROMS/Nonlinear/zetabc.F:536-545 (south closed), :685-694 (north closed)
void zetabc(const Model& m, const Bnd& b, int kout) {
  F2 z = m.zeta[kout];
  if (b.Southern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) z(i, b.Jstr - 1) = z(i, b.Jstr);
  if (b.Northern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) z(i, b.Jend + 1) = z(i, b.Jend);
}
// ROMS/Nonlinear/u2dbc_im.F:963-979 (south closed), :1121-1137 (north closed)
void u2dbc(const Model& m, const Bnd& b, int kout) {
  F2 A = m.ubar[kout]; const double g2 = m.c.gamma2;
  if (b.Southern_Edge) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jstr - 1) = g2 * A(i, b.Jstr);
  if (b.Northern_Edge) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jend + 1) = g2 * A(i, b.Jend);
}
// ROMS/Nonlinear/v2dbc_im.F:436-441 (south closed), :785-790 (north closed)
void v2dbc(const Model& m, const Bnd& b, int kout) {
  F2 A = m.vbar[kout];
  if (b.Southern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr) = 0.0;
  if (b.Northern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1) = 0.0;
}
// ROMS/Nonlinear/u3dbc_im.F:507-529, :665-687
void u3dbc(const Model& m, const Bnd& b, int nout) {
  F3 A = m.u[nout]; const double g2 = m.c.gamma2;
  if (b.Southern_Edge)
    for (int k = 1; k <= m.c.N; ++k) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jstr - 1, k) = g2 * A(i, b.Jstr, k);
  if (b.Northern_Edge)
    for (int k = 1; k <= m.c.N; ++k) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jend + 1, k) = g2 * A(i, b.Jend, k);
}
// ROMS/Nonlinear/v3dbc_im.F:222-230, :364-372
void v3dbc(const Model& m, const Bnd& b, int nout) {
  F3 A = m.v[nout];
  if (b.Southern_Edge) for (int k = 1; k <= m.c.N; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr, k) = 0.0;
  if (b.Northern_Edge) for (int k = 1; k <= m.c.N; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1, k) = 0.0;
}
// ROMS/Nonlinear/t3dbc_im.F:477-489, :611-623
void t3dbc(const Model& m, const Bnd& b, int nout, int itrc) {
  F3 A = m.t[nout][itrc];
  if (b.Southern_Edge) for (int k = 1; k <= m.c.N; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr - 1, k) = A(i, b.Jstr, k);
  if (b.Northern_Edge) for (int k = 1; k <= m.c.N; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1, k) = A(i, b.Jend, k);
}
// ROMS/Nonlinear/bc_2d.F: bc_r2d :41-160, bc_u2d :164-338 (closed: gamma2), bc_v2d :342-516 (closed: 0); each ends
// with the periodic exchange
void bc_r2d(const Model& m, const Bnd& b, F2 A) {
  if (b.Northern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1) = A(i, b.Jend);
  if (b.Southern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr - 1) = A(i, b.Jstr);
  exchange_r2d(m, b, A);
}
void bc_u2d(const Model& m, const Bnd& b, F2 A) {
  const double g2 = m.c.gamma2;
  if (b.Northern_Edge) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jend + 1) = g2 * A(i, b.Jend);
  if (b.Southern_Edge) for (int i = b.IstrU; i <= b.Iend; ++i) A(i, b.Jstr - 1) = g2 * A(i, b.Jstr);
  exchange_u2d(m, b, A);
}
void bc_v2d(const Model& m, const Bnd& b, F2 A) {
  if (b.Northern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1) = 0.0;
  if (b.Southern_Edge) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr) = 0.0;
  exchange_v2d(m, b, A);
}
// ROMS/Nonlinear/bc_3d.F: bc_r3d :45-180, bc_w3d :588-723
void bc_r3d(const Model& m, const Bnd& b, F3 A) {
  if (b.Northern_Edge) for (int k = A.LBk; k <= A.UBk; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1, k) = A(i, b.Jend, k);
  if (b.Southern_Edge) for (int k = A.LBk; k <= A.UBk; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr - 1, k) = A(i, b.Jstr, k);
  exchange_r3d(m, b, A);
}
void bc_w3d(const Model& m, const Bnd& b, F3 A) {
  if (b.Northern_Edge) for (int k = A.LBk; k <= A.UBk; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jend + 1, k) = A(i, b.Jend, k);
  if (b.Southern_Edge) for (int k = A.LBk; k <= A.UBk; ++k) for (int i = b.Istr; i <= b.Iend; ++i) A(i, b.Jstr - 1, k) = A(i, b.Jstr, k);
  exchange_w3d(m, b, A);
}

}  // namespace orc

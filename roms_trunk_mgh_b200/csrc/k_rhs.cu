// rhs3d_tile (fused: Coriolis, curvilinear, 3rd-order upstream horizontal advection, 4th-order centred vertical
// advection, column sums) and uv3dmix2_s_tile.  Thread per column; the k loop carries the vertical flux FC(k-1) and
// the k-ordered rufrc/rvfrc sums in registers, so ru/rv are read and written exactly once.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

__device__ __forceinline__ double d2x(const double* __restrict__ A, int o) { return A[o - 1] - 2.0 * A[o] + A[o + 1]; }
__device__ __forceinline__ double d2y(const double* __restrict__ A, int o, int P) { return A[o - P] - 2.0 * A[o] + A[o + P]; }

// ROMS/Nonlinear/rhs3d.F:174-1671.  (A variant that gave the xi- and eta-momentum equations of a column to two threads of the
// same CTA -- half the live operands, 122 registers, twice the resident warps -- was exact but slower, 0.65 vs 0.56 ms: the
// kernel is bound by L1 / issue throughput, and the split adds ~10% loads.  A shared-memory tiled form that evaluates every face
// flux once -- tools/variants/k_rhs3d_tiled.cuh -- is exact too and also slower, 0.546 vs 0.517 ms.  profiles/README.md.)
#ifndef RHS_MINB
#define RHS_MINB 2
#endif
#ifndef RHS_BX
#define RHS_BX 64
#endif
#ifndef RHS_PF
#define RHS_PF 4          // L2 prefetch distance (levels) in k_rhs3d
#endif
#ifndef UVM_BX
#define UVM_BX 64
#endif
#ifndef UVM_PF
#define UVM_PF 4          // same for k_uv3dmix2
#endif
// UADV: 0 the default branch (third-order upstream horizontal :706-730 ..., fourth-order centred vertical :1177-1255); 1
// UV_C4ADVECTION (fourth-order centred horizontal :685-705, :761-781, :829-849, :902-921; vertical 9/32, 1/32 :1108-1175, :1362-1429);
// 2 UV_SADVECTION (the default horizontal branch; conservative parabolic splines in the vertical :1016-1078, :1267-1329);
// 3 UV_C2ADVECTION (second-order centred: horizontal :605-657, vertical :1079-1107, :1330-1361).
// BF: BODYFORCE -- the surface / bottom stress enters ru, rv as a body force over levels levsfrc:N / 1:levbfrc (:326-466) instead of rufrc (:1588-1599).
template <int UADV, bool BF = false>
__global__ void __launch_bounds__(128, RHS_MINB) k_rhs3d(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, Mm = p.Mm, o2 = j * P;
  const double Gadv = -0.25, C6 = 1.0 / 6.0;
  const double* __restrict__ u = f.u[p.nrhs];
  const double* __restrict__ v = f.v[p.nrhs];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  double* __restrict__ ru = f.ru[p.nrhs];
  double* __restrict__ rv = f.rv[p.nrhs];
  const bool dov = (j >= p.JstrV);
  // 2-D factors
  const double fomn0 = f.fomn[o2 + i], fomnW = f.fomn[o2 + i - 1], fomnS = f.fomn[o2 - P + i];
  double dndx0 = 0, dndxW = 0, dndxS = 0, dmde0 = 0, dmdeW = 0, dmdeS = 0;
  if (p.curvgrid) {
    dndx0 = f.dndx[o2 + i]; dndxW = f.dndx[o2 + i - 1]; dndxS = f.dndx[o2 - P + i];
    dmde0 = f.dmde[o2 + i]; dmdeW = f.dmde[o2 + i - 1]; dmdeS = f.dmde[o2 - P + i];
  }
  // Row offsets clamped at the closed walls: they serve the wall copies of the second differences (rhs3d.F:742-755,
  // :886-901) and keep every load in bounds so that all operands of a level can be fetched unconditionally.
  const int S2 = (j >= 2) ? -2 * P : -P, N2 = (j + 2 <= Mm + 1) ? 2 * P : P;
  double FCu_m = 0.0, FCv_m = 0.0, rufrc = 0.0, rvfrc = 0.0;
  // UV_SADVECTION (rhs3d.F:1016-1051, :1267-1302): the spline derivatives CF(0:N) of the u and v columns, by the reference's own
  // forward elimination / back substitution, before the level sweep (thread-local columns)
  double CFu[UADV == 2 ? MAXN + 1 : 1], CFv[UADV == 2 ? MAXN + 1 : 1];
  if (UADV == 2) {
    const double s1 = 9.0 / 16.0, s2 = 1.0 / 16.0;
    double FCs[MAXN + 1];
    const int ob = o2 + i;
    {
      FCs[0] = 0.0; CFu[0] = 0.0;
      double dck = s1 * (Hz[ob + p.PL] + Hz[ob + p.PL - 1]) - s2 * (Hz[ob + p.PL + 1] + Hz[ob + p.PL - 2]);
      for (int k = 1; k <= N - 1; ++k) {
        const int o1 = ob + (k + 1) * p.PL;
        const double dck1 = s1 * (Hz[o1] + Hz[o1 - 1]) - s2 * (Hz[o1 + 1] + Hz[o1 - 2]);
        const double cff = 1.0 / (2.0 * dck1 + dck * (2.0 - FCs[k - 1]));
        FCs[k] = cff * dck1;
        CFu[k] = cff * (6.0 * (u[o1] - u[o1 - p.PL]) - dck * CFu[k - 1]);
        dck = dck1;
      }
      CFu[N] = 0.0;
      for (int k = N - 1; k >= 1; --k) CFu[k] = CFu[k] - FCs[k] * CFu[k + 1];
    }
    if (dov) {
      FCs[0] = 0.0; CFv[0] = 0.0;
      double dck = (s1 * (Hz[ob + p.PL] + Hz[ob + p.PL - P]) - s2 * (Hz[ob + p.PL + P] + Hz[ob + p.PL - 2 * P]));
      for (int k = 1; k <= N - 1; ++k) {
        const int o1 = ob + (k + 1) * p.PL;
        const double dck1 = (s1 * (Hz[o1] + Hz[o1 - P]) - s2 * (Hz[o1 + P] + Hz[o1 - 2 * P]));
        const double cff = 1.0 / (2.0 * dck1 + dck * (2.0 - FCs[k - 1]));
        FCs[k] = cff * dck1;
        CFv[k] = cff * (6.0 * (v[o1] - v[o1 - p.PL]) - dck * CFv[k - 1]);
        dck = dck1;
      }
      CFv[N] = 0.0;
      for (int k = N - 1; k >= 1; --k) CFv[k] = CFv[k] - FCs[k] * CFv[k + 1];
    }
  }
  double bfUs = 0.0, bfUb = 0.0, bfVs = 0.0, bfVb = 0.0;
  if (BF) {                                                            // rhs3d.F:355-382, :411-438
    const int ob = o2 + i;
    const double cu = 0.25 * (f.pm[ob - 1] + f.pm[ob]) * (f.pn[ob - 1] + f.pn[ob]);
    const double cv = 0.25 * (f.pm[ob - P] + f.pm[ob]) * (f.pn[ob - P] + f.pn[ob]);
    double w0 = 0.0, wW = 0.0, wS = 0.0;
    for (int k = N; k >= p.levsfrc; --k) { const int o = ob + k * p.PL; w0 = w0 + Hz[o]; wW = wW + Hz[o - 1]; wS = wS + Hz[o - P]; }
    bfUs = f.sustr[ob] * (1.0 / (cu * (wW + w0)));
    if (dov) bfVs = f.svstr[ob] * (1.0 / (cv * (wS + w0)));
    w0 = 0.0; wW = 0.0; wS = 0.0;
    for (int k = 1; k <= p.levbfrc; ++k) { const int o = ob + k * p.PL; w0 = w0 + Hz[o]; wW = wW + Hz[o - 1]; wS = wS + Hz[o - P]; }
    bfUb = f.bustr[ob] * (1.0 / (cu * (wW + w0)));
    if (dov) bfVb = f.bvstr[ob] * (1.0 / (cv * (wS + w0)));
  }
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL + i;                     // includes i
    const int oU = (k < N) ? o + p.PL : o, oUU = (k + 2 <= N) ? o + 2 * p.PL : oU, oD = (k > 1) ? o - p.PL : o;
    // ---- all global operands of this level, issued back to back (one memory wait per level)
    pf_up<RHS_PF>(Hz, o, k, N, p.PL); pf_up<RHS_PF>(u, o, k, N, p.PL); pf_up<RHS_PF>(v, o, k, N, p.PL); pf_up<RHS_PF>(Huon, o, k, N, p.PL);
    pf_up<RHS_PF>(Hvom, o, k, N, p.PL); pf_up<RHS_PF>(W, o, k, N, p.PL); pf_up<RHS_PF>(ru, o, k, N, p.PL); pf_up<RHS_PF>(rv, o, k, N, p.PL);
    const double hz0 = Hz[o], hzW = Hz[o - 1], hzS = Hz[o - P];
    const double uW2 = u[o - 2], uW = u[o - 1], u0 = u[o], uE = u[o + 1], uE2 = u[o + 2];
    const double uS2 = u[o + S2], uS = u[o - P], uN = u[o + P], uN2 = u[o + N2], uSE = u[o - P + 1];
    const double vW2 = v[o - 2], vW = v[o - 1], v0 = v[o], vE = v[o + 1], vE2 = v[o + 2];
    const double vS2 = v[o + S2], vS = v[o - P], vN = v[o + P], vN2 = v[o + N2], vNW = v[o + P - 1];
    const double HuW2 = Huon[o - 2], HuW = Huon[o - 1], Hu0 = Huon[o], HuE = Huon[o + 1], HuE2 = Huon[o + 2];
    const double HuS2 = Huon[o + S2], HuS = Huon[o - P], HuN = Huon[o + P];
    const double HuES2 = Huon[o + 1 + S2], HuSE = Huon[o + 1 - P], HuNE = Huon[o + 1 + P];
    const double HvW2 = Hvom[o - 2], HvW = Hvom[o - 1], Hv0 = Hvom[o], HvE = Hvom[o + 1];
    const double HvNW2 = Hvom[o + P - 2], HvNW = Hvom[o + P - 1], HvN = Hvom[o + P], HvNE = Hvom[o + P + 1];
    const double HvS2 = Hvom[o + S2], HvS = Hvom[o - P], HvN2 = Hvom[o + N2];
    const double W0 = W[o], WW = W[o - 1], WE = W[o + 1], WW2 = W[o - 2], WS = W[o - P], WN = W[o + P], WS2 = W[o + S2];
    const double uUp = u[oU], uUp2 = u[oUU], uDn = u[oD], vUp = v[oU], vUp2 = v[oUU], vDn = v[oD];
    double rux = ru[o];
    double rvx = rv[o];
    if (BF) {                                                          // :383-410, :439-466 (before every other term, as the reference)
      if (k >= p.levsfrc) { rux = rux + bfUs * (hz0 + hzW); if (dov) rvx = rvx + bfVs * (hz0 + hzS); }
      if (k <= p.levbfrc) { rux = rux - bfUb * (hz0 + hzW); if (dov) rvx = rvx - bfVb * (hz0 + hzS); }
    }
    // ---- Coriolis (rhs3d.F:473-507): UFx at rho(i,j), rho(i-1,j); VFe at rho(i,j), rho(i,j-1)
    {
      const double c0 = 0.5 * hz0 * fomn0;
      const double UFx0 = c0 * (v0 + vN), VFe0 = c0 * (u0 + uE);
      const double cW = 0.5 * hzW * fomnW;
      const double UFxW = cW * (vW + vNW);
      rux = rux + 0.5 * (UFx0 + UFxW);
      if (dov) {
        const double cS = 0.5 * hzS * fomnS;
        const double VFeS = cS * (uS + uSE);
        rvx = rvx - 0.5 * (VFe0 + VFeS);
      }
    }
    // ---- curvilinear terms (rhs3d.F:515-564)
    if (p.curvgrid) {
      double c1 = 0.5 * (v0 + vN), c2 = 0.5 * (u0 + uE);
      double c = hz0 * (c1 * dndx0 - c2 * dmde0);
      const double UFx0 = c * c1, VFe0 = c * c2;
      c1 = 0.5 * (vW + vNW); c2 = 0.5 * (uW + u0);
      c = hzW * (c1 * dndxW - c2 * dmdeW);
      const double UFxW = c * c1;
      rux = rux + 0.5 * (UFx0 + UFxW);
      if (dov) {
        c1 = 0.5 * (vS + v0); c2 = 0.5 * (uS + uSE);
        c = hzS * (c1 * dndxS - c2 * dmdeS);
        const double VFeS = c * c2;
        rvx = rvx - 0.5 * (VFe0 + VFeS);
      }
    }
    // ---- horizontal advection of u (rhs3d.F:658-798, :946-963)
    {
      const double uxxW = uW2 - 2.0 * uW + u0, uxx0 = uW - 2.0 * u0 + uE, uxxE = u0 - 2.0 * uE + uE2;
      const double HxxW = HuW2 - 2.0 * HuW + Hu0, Hxx0 = HuW - 2.0 * Hu0 + HuE, HxxE = Hu0 - 2.0 * HuE + HuE2;
      // UFx(i,j) at rho(i,j) and UFx(i-1,j)
      double c1 = u0 + uE;
      double c = (c1 > 0.0) ? uxx0 : uxxE;
      double UFx0, UFxW;
      if (UADV == 3) {
        UFx0 = 0.25 * (u0 + uE) * (Hu0 + HuE);
        UFxW = 0.25 * (uW + u0) * (HuW + Hu0);
      } else if (UADV == 1) {
        UFx0 = 0.25 * (u0 + uE - C6 * (uxx0 + uxxE)) * (Hu0 + HuE - C6 * (Hxx0 + HxxE));
        UFxW = 0.25 * (uW + u0 - C6 * (uxxW + uxx0)) * (HuW + Hu0 - C6 * (HxxW + Hxx0));
      } else {
        UFx0 = 0.25 * (c1 + Gadv * c) * (Hu0 + HuE + Gadv * 0.5 * (Hxx0 + HxxE));
        c1 = uW + u0;
        c = (c1 > 0.0) ? uxxW : uxx0;
        UFxW = 0.25 * (c1 + Gadv * c) * (HuW + Hu0 + Gadv * 0.5 * (HxxW + Hxx0));
      }
      // UFe(i,j) and UFe(i,j+1) at psi points; uee(i,0) = uee(i,1), uee(i,Mm+1) = uee(i,Mm)
      const double uee_j = uS - 2.0 * u0 + uN;
      const double uee_jm1 = (j > 1) ? (uS2 - 2.0 * uS + u0) : uee_j;
      const double uee_jp1 = (j < Mm) ? (u0 - 2.0 * uN + uN2) : uee_j;
      const double Hvxx0 = HvW - 2.0 * Hv0 + HvE, HvxxW = HvW2 - 2.0 * HvW + Hv0;
      const double HvxxN = HvNW - 2.0 * HvN + HvNE, HvxxNW = HvNW2 - 2.0 * HvNW + HvN;
      double UFe0, UFeN;
      if (UADV == 3) {
        UFe0 = 0.25 * (uS + u0) * (HvW + Hv0);
        UFeN = 0.25 * (u0 + uN) * (HvNW + HvN);
      } else if (UADV == 1) {
        UFe0 = 0.25 * (u0 + uS - C6 * (uee_j + uee_jm1)) * (Hv0 + HvW - C6 * (Hvxx0 + HvxxW));
        UFeN = 0.25 * (uN + u0 - C6 * (uee_jp1 + uee_j)) * (HvN + HvNW - C6 * (HvxxN + HvxxNW));
      } else {
        c1 = u0 + uS;
        double c2 = Hv0 + HvW;
        c = (c2 > 0.0) ? uee_jm1 : uee_j;
        UFe0 = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Hvxx0 + HvxxW));
        c1 = uN + u0;
        c2 = HvN + HvNW;
        c = (c2 > 0.0) ? uee_j : uee_jp1;
        UFeN = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (HvxxN + HvxxNW));
      }
      const double a1 = UFx0 - UFxW;
      const double a2 = UFeN - UFe0;
      rux = rux - (a1 + a2);
    }
    // ---- horizontal advection of v (rhs3d.F:800-940, :965-982)
    if (dov) {
      const double vxxW = vW2 - 2.0 * vW + v0, vxx0 = vW - 2.0 * v0 + vE, vxxE = v0 - 2.0 * vE + vE2;
      const double Huee0 = HuS - 2.0 * Hu0 + HuN, HueeS = HuS2 - 2.0 * HuS + Hu0;
      const double HueeE = HuSE - 2.0 * HuE + HuNE, HueeSE = HuES2 - 2.0 * HuSE + HuE;
      // VFx(i,j), VFx(i+1,j) at psi points
      double c1, c2, c, VFx0, VFxE;
      if (UADV == 3) {
        VFx0 = 0.25 * (vW + v0) * (HuS + Hu0);
        VFxE = 0.25 * (v0 + vE) * (HuSE + HuE);
      } else if (UADV == 1) {
        VFx0 = 0.25 * (v0 + vW - C6 * (vxx0 + vxxW)) * (Hu0 + HuS - C6 * (Huee0 + HueeS));
        VFxE = 0.25 * (vE + v0 - C6 * (vxxE + vxx0)) * (HuE + HuSE - C6 * (HueeE + HueeSE));
      } else {
        c1 = v0 + vW;
        c2 = Hu0 + HuS;
        c = (c2 > 0.0) ? vxxW : vxx0;
        VFx0 = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Huee0 + HueeS));
        c1 = vE + v0;
        c2 = HuE + HuSE;
        c = (c2 > 0.0) ? vxx0 : vxxE;
        VFxE = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (HueeE + HueeSE));
      }
      // VFe(i,j), VFe(i,j-1) at rho points; vee/Hvee rows 2..Mm with copies (1)=(2), (Mm+1)=(Mm)
      const double vee_j = vS - 2.0 * v0 + vN, Hvee_j = HvS - 2.0 * Hv0 + HvN;
      const double vee_jp = (j < Mm) ? (v0 - 2.0 * vN + vN2) : vee_j, Hvee_jp = (j < Mm) ? (Hv0 - 2.0 * HvN + HvN2) : Hvee_j;
      const double vee_jm = (j > 2) ? (vS2 - 2.0 * vS + v0) : vee_j, Hvee_jm = (j > 2) ? (HvS2 - 2.0 * HvS + Hv0) : Hvee_j;
      double VFe0, VFeS;
      if (UADV == 3) {
        VFe0 = 0.25 * (v0 + vN) * (Hv0 + HvN);
        VFeS = 0.25 * (vS + v0) * (HvS + Hv0);
      } else if (UADV == 1) {
        VFe0 = 0.25 * (v0 + vN - C6 * (vee_j + vee_jp)) * (Hv0 + HvN - C6 * (Hvee_j + Hvee_jp));
        VFeS = 0.25 * (vS + v0 - C6 * (vee_jm + vee_j)) * (HvS + Hv0 - C6 * (Hvee_jm + Hvee_j));
      } else {
        c1 = v0 + vN;
        c = (c1 > 0.0) ? vee_j : vee_jp;
        VFe0 = 0.25 * (c1 + Gadv * c) * (Hv0 + HvN + Gadv * 0.5 * (Hvee_j + Hvee_jp));
        c1 = vS + v0;
        c = (c1 > 0.0) ? vee_jm : vee_j;
        VFeS = 0.25 * (c1 + Gadv * c) * (HvS + Hv0 + Gadv * 0.5 * (Hvee_jm + Hvee_j));
      }
      const double a1 = VFxE - VFx0;
      const double a2 = VFe0 - VFeS;
      rvx = rvx - (a1 + a2);
    }
    // ---- vertical advection (rhs3d.F:1177-1265, :1434-1522): FC(k) through the top of level k
    {
      const double c1 = (UADV == 1) ? 9.0 / 32.0 : 9.0 / 16.0, c2 = (UADV == 1) ? 1.0 / 32.0 : 1.0 / 16.0;
      double FCu = 0.0, FCv = 0.0;
      if (k < N) {
        const double ukm = (k > 1) ? uDn : u0;
        const double ukpp = (k + 2 <= N) ? uUp2 : uUp;
        if (UADV == 3) FCu = 0.25 * (u0 + uUp) * (W0 + WW);                  // rhs3d.F:1080-1088
        else if (UADV == 1) FCu = (c1 * (u0 + uUp) - c2 * (ukm + ukpp)) * (W0 + WW);
        else if (UADV == 2) {                                               // rhs3d.F:1052-1069
          const double DCk = c1 * (hz0 + hzW) - c2 * (Hz[o + 1] + Hz[o - 2]);
          FCu = (c1 * (W0 + WW) - c2 * (WE + WW2)) * (u0 + DCk * ((1.0 / 3.0) * CFu[k] + (1.0 / 6.0) * CFu[k - 1]));
        }
        else FCu = (c1 * (u0 + uUp) - c2 * (ukm + ukpp)) * (c1 * (W0 + WW) - c2 * (WE + WW2));
        if (dov) {
          const double vkm = (k > 1) ? vDn : v0;
          const double vkpp = (k + 2 <= N) ? vUp2 : vUp;
          if (UADV == 3) FCv = 0.25 * (v0 + vUp) * (W0 + WS);                // rhs3d.F:1331-1339
          else if (UADV == 1) FCv = (c1 * (v0 + vUp) - c2 * (vkm + vkpp)) * (W0 + WS);
          else if (UADV == 2) {                                             // rhs3d.F:1303-1320
            const double DCk = (c1 * (hz0 + hzS) - c2 * (Hz[o + P] + Hz[o - 2 * P]));
            FCv = (c1 * (W0 + WS) - c2 * (WN + WS2)) * (v0 + DCk * ((1.0 / 3.0) * CFv[k] + (1.0 / 6.0) * CFv[k - 1]));
          }
          else FCv = (c1 * (v0 + vUp) - c2 * (vkm + vkpp)) * (c1 * (W0 + WS) - c2 * (WN + WS2));
        }
      }
      rux = rux - (FCu - FCu_m);
      FCu_m = FCu;
      if (dov) { rvx = rvx - (FCv - FCv_m); FCv_m = FCv; }
    }
    ru[o] = rux;
    rufrc = (k == 1) ? rux : rufrc + rux;
    if (dov) { rv[o] = rvx; rvfrc = (k == 1) ? rvx : rvfrc + rvx; }
  }
  // ---- column sums + surface/bottom stresses (rhs3d.F:1534-1667)
  {
    const double c = f.om_u[o2 + i] * f.on_u[o2 + i];
    const double c1 = f.sustr[o2 + i] * c;
    const double c2 = -f.bustr[o2 + i] * c;
    f.rufrc[o2 + i] = BF ? rufrc : rufrc + c1 + c2;
  }
  if (dov) {
    const double c = f.om_v[o2 + i] * f.on_v[o2 + i];
    const double c1 = f.svstr[o2 + i] * c;
    const double c2 = -f.bvstr[o2 + i] * c;
    f.rvfrc[o2 + i] = BF ? rvfrc : rvfrc + c1 + c2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// uv3dmix2_s_tile (ROMS/Nonlinear/uv3dmix2_s.h:239-330): harmonic viscosity along s-surfaces, stress-tensor form.
// rufrc/rvfrc accumulate level by level (k = 1..N) exactly as the reference does.
#ifndef UVM_MINB
#define UVM_MINB 3
#endif
__global__ void __launch_bounds__(128, UVM_MINB) k_uv3dmix2(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, o2 = j * P + i;
  const double* __restrict__ u = f.u[p.nrhs];
  const double* __restrict__ v = f.v[p.nrhs];
  double* __restrict__ un = f.u[p.nnew];
  double* __restrict__ vn = f.v[p.nnew];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ pm = f.pm;
  const double* __restrict__ pn = f.pn;
  const bool dov = (j >= p.JstrV);
  // level-independent factors of the rho-point stress at a cell c and of the psi-point stress at the SW corner of c
  struct R2 { double pmon, pnom, pnE, pnW, pmN, pmS; };
  struct P2 { double pmon, pnom, pnv0, pnvW, pmu0, pmuS; };
  auto r2 = [&](int c) -> R2 {
    return R2{f.pmon_r[c], f.pnom_r[c], pn[c] + pn[c + 1], pn[c - 1] + pn[c], pm[c] + pm[c + P], pm[c - P] + pm[c]};
  };
  auto p2 = [&](int c) -> P2 {
    return P2{f.pmon_p[c], f.pnom_p[c], pn[c - P] + pn[c], pn[c - P - 1] + pn[c - 1], pm[c - 1] + pm[c], pm[c - P - 1] + pm[c - P]};
  };
  const R2 R0 = r2(o2), RW = r2(o2 - 1), RS = r2(dov ? o2 - P : o2);      // (the southern cell's factors read row j-2: only where v exists)
  const P2 Q0 = p2(o2), QN = p2(o2 + P), QE = p2(o2 + 1);
  auto rho_cff = [](const R2& m, double hz, double uE, double u0, double vN, double v0) -> double {
    return hz * 0.5 * (m.pmon * (m.pnE * uE - m.pnW * u0) - m.pnom * (m.pmN * vN - m.pmS * v0));
  };
  auto psi_cff = [](const P2& m, double hW, double h0, double hSW, double hS, double v0, double vW, double u0, double uS) -> double {
    return 0.125 * (hW + h0 + hSW + hS) * (m.pmon * (m.pnv0 * v0 - m.pnvW * vW) + m.pnom * (m.pmu0 * u0 - m.pmuS * uS));
  };
  const double onr2_0 = f.on_r[o2] * f.on_r[o2] * f.visc2_r[o2], onr2_W = f.on_r[o2 - 1] * f.on_r[o2 - 1] * f.visc2_r[o2 - 1];
  const double omr2_0 = f.om_r[o2] * f.om_r[o2] * f.visc2_r[o2], omr2_S = f.om_r[o2 - P] * f.om_r[o2 - P] * f.visc2_r[o2 - P];
  const double omp2_0 = f.om_p[o2] * f.om_p[o2] * f.visc2_p[o2], omp2_N = f.om_p[o2 + P] * f.om_p[o2 + P] * f.visc2_p[o2 + P];
  const double onp2_0 = f.on_p[o2] * f.on_p[o2] * f.visc2_p[o2], onp2_E = f.on_p[o2 + 1] * f.on_p[o2 + 1] * f.visc2_p[o2 + 1];
  const double pmU = pm[o2 - 1] + pm[o2], pnU = pn[o2 - 1] + pn[o2];
  const double pmV = pm[o2] + pm[o2 - P], pnV = pn[o2] + pn[o2 - P];
  const double pmVb = pm[o2 - P] + pm[o2], pnVb = pn[o2 - P] + pn[o2];
  double rufrc = f.rufrc[o2], rvfrc = dov ? f.rvfrc[o2] : 0.0;
  // operands of one level; level k+1 is requested before level k is computed
  struct Lvl { double h0, hW, hE, hS, hSW, hSE, hN, hNW, uW, u0, uE, uS, uSE, uN, vW, v0, vE, vS, vN, vNW, un0, vn0; };
  auto load_level = [&](int k) -> Lvl {
    const int o = o2 + k * p.PL;
    Lvl L;
    pf_up<UVM_PF>(Hz, o, k, N, p.PL); pf_up<UVM_PF>(u, o, k, N, p.PL); pf_up<UVM_PF>(v, o, k, N, p.PL); pf_up<UVM_PF>(un, o, k, N, p.PL); pf_up<UVM_PF>(vn, o, k, N, p.PL);
    L.h0 = Hz[o]; L.hW = Hz[o - 1]; L.hE = Hz[o + 1]; L.hS = Hz[o - P]; L.hSW = Hz[o - P - 1]; L.hSE = Hz[o - P + 1];
    L.hN = Hz[o + P]; L.hNW = Hz[o + P - 1];
    L.uW = u[o - 1]; L.u0 = u[o]; L.uE = u[o + 1]; L.uS = u[o - P]; L.uSE = u[o - P + 1]; L.uN = u[o + P];
    L.vW = v[o - 1]; L.v0 = v[o]; L.vE = v[o + 1]; L.vS = v[o - P]; L.vN = v[o + P]; L.vNW = v[o + P - 1];
    L.un0 = un[o];
    L.vn0 = dov ? vn[o] : 0.0;
    return L;
  };
  auto level = [&](const Lvl& cur, const Lvl&, int k) {
    const int o = o2 + k * p.PL;
    const double h0 = cur.h0, hW = cur.hW, hE = cur.hE, hS = cur.hS, hSW = cur.hSW, hSE = cur.hSE, hN = cur.hN, hNW = cur.hNW;
    const double uW = cur.uW, u0 = cur.u0, uE = cur.uE, uS = cur.uS, uSE = cur.uSE, uN = cur.uN;
    const double vW = cur.vW, v0 = cur.v0, vE = cur.vE, vS = cur.vS, vN = cur.vN, vNW = cur.vNW;
    const double un0 = cur.un0, vn0 = cur.vn0;
    const double cr0 = rho_cff(R0, h0, uE, u0, vN, v0), crW = rho_cff(RW, hW, u0, uW, vNW, vW);
    const double cp0 = psi_cff(Q0, hW, h0, hSW, hS, v0, vW, u0, uS), cpN = psi_cff(QN, hNW, hN, hW, h0, vN, vNW, uN, u0);
    {
      const double UFx0 = onr2_0 * cr0, UFxW = onr2_W * crW;
      const double UFe0 = omp2_0 * cp0, UFeN = omp2_N * cpN;
      const double cff = p.dt * 0.25 * pmU * pnU;
      const double cff1 = 0.5 * pnU * (UFx0 - UFxW);
      const double cff2 = 0.5 * pmU * (UFeN - UFe0);
      const double cff3 = cff * (cff1 + cff2);
      rufrc = rufrc + cff1 + cff2;
      un[o] = un0 + cff3;
    }
    if (dov) {
      const double crS = rho_cff(RS, hS, uSE, uS, v0, vS), cpE = psi_cff(QE, h0, hE, hS, hSE, vE, v0, uE, uSE);
      const double VFx0 = onp2_0 * cp0, VFxE = onp2_E * cpE;
      const double VFe0 = omr2_0 * cr0, VFeS = omr2_S * crS;
      const double cff = p.dt * 0.25 * pmV * pnV;
      const double cff1 = 0.5 * pnVb * (VFxE - VFx0);
      const double cff2 = 0.5 * pmVb * (VFe0 - VFeS);
      const double cff3 = cff * (cff1 - cff2);
      rvfrc = rvfrc + cff1 - cff2;
      vn[o] = vn0 + cff3;
    }
  };
  {
    sweep_levels<false>(N, load_level, level);
  }
  f.rufrc[o2] = rufrc;
  if (dov) f.rvfrc[o2] = rvfrc;
}

static inline dim3 g2(dim3 b, int ni, int nj, int nz = 1) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz); }
void launch_rhs3d(const Par& p, const Flds& f, cudaStream_t s) {
  dim3 b(RHS_BX, 128 / RHS_BX);
  if (p.bodyforce) {                                                   // BODYFORCE
    const dim3 g = g2(b, p.Iend - p.Istr + 1, p.Mm);
    if (p.uv_adv == 1) k_rhs3d<1, true><<<g, b, 0, s>>>(p, f);
    else if (p.uv_adv == 2) k_rhs3d<2, true><<<g, b, 0, s>>>(p, f);
    else if (p.uv_adv == 3) k_rhs3d<3, true><<<g, b, 0, s>>>(p, f);
    else k_rhs3d<0, true><<<g, b, 0, s>>>(p, f);
    return;
  }
  if (p.uv_adv == 1) k_rhs3d<1><<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  else if (p.uv_adv == 2) k_rhs3d<2><<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  else if (p.uv_adv == 3) k_rhs3d<3><<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
  else k_rhs3d<0><<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f);
}
void launch_uv3dmix2(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(UVM_BX, 128 / UVM_BX); k_uv3dmix2<<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f); }

}  // namespace rb

// rhs3d_tile (fused: Coriolis, curvilinear, 3rd-order upstream horizontal advection, 4th-order centred vertical
// advection, column sums) and uv3dmix2_s_tile.  Thread per column; the k loop carries the vertical flux FC(k-1) and
// the k-ordered rufrc/rvfrc sums in registers, so ru/rv are read and written exactly once.
#include "dev.cuh"
#include "kernels.h"

namespace rb {

__device__ __forceinline__ double d2x(const double* __restrict__ A, int o) { return A[o - 1] - 2.0 * A[o] + A[o + 1]; }
__device__ __forceinline__ double d2y(const double* __restrict__ A, int o, int P) { return A[o - P] - 2.0 * A[o] + A[o + P]; }

// ROMS/Nonlinear/rhs3d.F:174-1671
__global__ void __launch_bounds__(128) k_rhs3d(Par p, Flds f) {
  const int i = p.Istr + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > p.Iend || j > p.Mm) return;
  const int N = p.N, P = p.P, Mm = p.Mm, o2 = j * P;
  const double Gadv = -0.25;
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
  // clamped rows for the closed-wall copies of the second differences (rhs3d.F:742-755, :886-901)
  const int jm1c = (j - 1 < 1) ? 1 : j - 1;            // uee row for UFe(i,j):   uee(i,j-1), row 0 -> 1
  const int jp1c = (j + 1 > Mm) ? Mm : j + 1;          // uee row for UFe(i,j+1): uee(i,j+1), row Mm+1 -> Mm
  double FCu_m = 0.0, FCv_m = 0.0, rufrc = 0.0, rvfrc = 0.0;
  for (int k = 1; k <= N; ++k) {
    const int o = o2 + k * p.PL + i;                     // includes i
    if (k + PFD <= N) {                                  // start the DRAM fetch of level k+PFD
      const int q = o + PFD * p.PL;
      pf_l2(u + q); pf_l2(v + q); pf_l2(Hz + q); pf_l2(Huon + q); pf_l2(Hvom + q); pf_l2(W + q); pf_l2(ru + q); pf_l2(rv + q);
    }
    const double hz0 = Hz[o], hzW = Hz[o - 1], hzS = Hz[o - P];
    const double u0 = u[o], uE = u[o + 1], uW = u[o - 1], uN = u[o + P], uS = u[o - P];
    const double v0 = v[o], vN = v[o + P], vW = v[o - 1], vE = v[o + 1], vS = v[o - P];
    const double uSE = u[o - P + 1], vNW = v[o + P - 1];
    double rux = ru[o];
    double rvx = dov ? rv[o] : 0.0;
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
      const double uxxW = d2x(u, o - 1), uxx0 = d2x(u, o), uxxE = d2x(u, o + 1);
      const double HxxW = d2x(Huon, o - 1), Hxx0 = d2x(Huon, o), HxxE = d2x(Huon, o + 1);
      const double HuW = Huon[o - 1], Hu0 = Huon[o], HuE = Huon[o + 1];
      // UFx(i,j) at rho(i,j) and UFx(i-1,j)
      double c1 = u0 + uE;
      double c = (c1 > 0.0) ? uxx0 : uxxE;
      const double UFx0 = 0.25 * (c1 + Gadv * c) * (Hu0 + HuE + Gadv * 0.5 * (Hxx0 + HxxE));
      c1 = uW + u0;
      c = (c1 > 0.0) ? uxxW : uxx0;
      const double UFxW = 0.25 * (c1 + Gadv * c) * (HuW + Hu0 + Gadv * 0.5 * (HxxW + Hxx0));
      // UFe(i,j) and UFe(i,j+1) at psi points; uee rows clamped at the walls
      const int ob = o - j * P;                          // row 0 offset of this (i,k)
      const double uee_jm1 = d2y(u, ob + jm1c * P, P), uee_j = d2y(u, o, P), uee_jp1 = d2y(u, ob + jp1c * P, P);
      const double Hv0 = Hvom[o], HvW = Hvom[o - 1], HvN = Hvom[o + P], HvNW = Hvom[o + P - 1];
      const double Hvxx0 = d2x(Hvom, o), HvxxW = d2x(Hvom, o - 1), HvxxN = d2x(Hvom, o + P), HvxxNW = d2x(Hvom, o + P - 1);
      c1 = u0 + uS;
      double c2 = Hv0 + HvW;
      c = (c2 > 0.0) ? uee_jm1 : uee_j;
      const double UFe0 = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Hvxx0 + HvxxW));
      c1 = uN + u0;
      c2 = HvN + HvNW;
      c = (c2 > 0.0) ? uee_j : uee_jp1;
      const double UFeN = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (HvxxN + HvxxNW));
      const double a1 = UFx0 - UFxW;
      const double a2 = UFeN - UFe0;
      rux = rux - (a1 + a2);
    }
    // ---- horizontal advection of v (rhs3d.F:800-940, :965-982)
    if (dov) {
      const double vxxW = d2x(v, o - 1), vxx0 = d2x(v, o), vxxE = d2x(v, o + 1);
      const double Hu0 = Huon[o], HuS = Huon[o - P], HuE = Huon[o + 1], HuSE = Huon[o - P + 1];
      const double Huee0 = d2y(Huon, o, P), HueeS = d2y(Huon, o - P, P), HueeE = d2y(Huon, o + 1, P), HueeSE = d2y(Huon, o - P + 1, P);
      // VFx(i,j), VFx(i+1,j) at psi points
      double c1 = v0 + vW;
      double c2 = Hu0 + HuS;
      double c = (c2 > 0.0) ? vxxW : vxx0;
      const double VFx0 = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Huee0 + HueeS));
      c1 = vE + v0;
      c2 = HuE + HuSE;
      c = (c2 > 0.0) ? vxx0 : vxxE;
      const double VFxE = 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (HueeE + HueeSE));
      // VFe(i,j), VFe(i,j-1) at rho points; vee/Hvee defined for rows 2..Mm, copies vee(1)=vee(2), vee(Mm+1)=vee(Mm)
      const int ob = o - j * P;
      const int ja = (j < 2) ? 2 : j, jb = (j + 1 > Mm) ? Mm : j + 1, jc = (j - 1 < 2) ? 2 : j - 1;
      const double vee_j = d2y(v, ob + ja * P, P), vee_jp = d2y(v, ob + jb * P, P), vee_jm = d2y(v, ob + jc * P, P);
      const double Hvee_j = d2y(Hvom, ob + ja * P, P), Hvee_jp = d2y(Hvom, ob + jb * P, P), Hvee_jm = d2y(Hvom, ob + jc * P, P);
      const double Hv0 = Hvom[o], HvN = Hvom[o + P], HvS = Hvom[o - P];
      c1 = v0 + vN;
      c = (c1 > 0.0) ? vee_j : vee_jp;
      const double VFe0 = 0.25 * (c1 + Gadv * c) * (Hv0 + HvN + Gadv * 0.5 * (Hvee_j + Hvee_jp));
      c1 = vS + v0;
      c = (c1 > 0.0) ? vee_jm : vee_j;
      const double VFeS = 0.25 * (c1 + Gadv * c) * (HvS + Hv0 + Gadv * 0.5 * (Hvee_jm + Hvee_j));
      const double a1 = VFxE - VFx0;
      const double a2 = VFe0 - VFeS;
      rvx = rvx - (a1 + a2);
    }
    // ---- vertical advection (rhs3d.F:1177-1265, :1434-1522): FC(k) through the top of level k
    {
      const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
      double FCu = 0.0, FCv = 0.0;
      if (k < N) {
        const int ou = o + p.PL;
        const double ukm = (k > 1) ? u[o - p.PL] : u0;
        const double ukpp = (k + 2 <= N) ? u[ou + p.PL] : u[ou];
        FCu = (c1 * (u0 + u[ou]) - c2 * (ukm + ukpp)) * (c1 * (W[o] + W[o - 1]) - c2 * (W[o + 1] + W[o - 2]));
        if (dov) {
          const double vkm = (k > 1) ? v[o - p.PL] : v0;
          const double vkpp = (k + 2 <= N) ? v[ou + p.PL] : v[ou];
          FCv = (c1 * (v0 + v[ou]) - c2 * (vkm + vkpp)) * (c1 * (W[o] + W[o - P]) - c2 * (W[o + P] + W[o - 2 * P]));
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
    f.rufrc[o2 + i] = rufrc + c1 + c2;
  }
  if (dov) {
    const double c = f.om_v[o2 + i] * f.on_v[o2 + i];
    const double c1 = f.svstr[o2 + i] * c;
    const double c2 = -f.bvstr[o2 + i] * c;
    f.rvfrc[o2 + i] = rvfrc + c1 + c2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// uv3dmix2_s_tile (ROMS/Nonlinear/uv3dmix2_s.h:239-330): harmonic viscosity along s-surfaces, stress-tensor form.
// rufrc/rvfrc accumulate level by level (k = 1..N) exactly as the reference does.
__global__ void __launch_bounds__(128) k_uv3dmix2(Par p, Flds f) {
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
  // rho-point stress at cell c (offset oc): needs pn(c-1..c+1 in i), pm(c-1..c+1 in j)
  auto rho_cff = [&](int oc2, int oc3) -> double {
    return Hz[oc3] * 0.5 *
           (f.pmon_r[oc2] * ((pn[oc2] + pn[oc2 + 1]) * u[oc3 + 1] - (pn[oc2 - 1] + pn[oc2]) * u[oc3]) -
            f.pnom_r[oc2] * ((pm[oc2] + pm[oc2 + P]) * v[oc3 + P] - (pm[oc2 - P] + pm[oc2]) * v[oc3]));
  };
  // psi-point stress at corner c (SW corner of cell c)
  auto psi_cff = [&](int oc2, int oc3) -> double {
    return 0.125 * (Hz[oc3 - 1] + Hz[oc3] + Hz[oc3 - P - 1] + Hz[oc3 - P]) *
           (f.pmon_p[oc2] * ((pn[oc2 - P] + pn[oc2]) * v[oc3] - (pn[oc2 - P - 1] + pn[oc2 - 1]) * v[oc3 - 1]) +
            f.pnom_p[oc2] * ((pm[oc2 - 1] + pm[oc2]) * u[oc3] - (pm[oc2 - P - 1] + pm[oc2 - P]) * u[oc3 - P]));
  };
  const double onr2_0 = f.on_r[o2] * f.on_r[o2] * f.visc2_r[o2], onr2_W = f.on_r[o2 - 1] * f.on_r[o2 - 1] * f.visc2_r[o2 - 1];
  const double omr2_0 = f.om_r[o2] * f.om_r[o2] * f.visc2_r[o2], omr2_S = f.om_r[o2 - P] * f.om_r[o2 - P] * f.visc2_r[o2 - P];
  const double omp2_0 = f.om_p[o2] * f.om_p[o2] * f.visc2_p[o2], omp2_N = f.om_p[o2 + P] * f.om_p[o2 + P] * f.visc2_p[o2 + P];
  const double onp2_0 = f.on_p[o2] * f.on_p[o2] * f.visc2_p[o2], onp2_E = f.on_p[o2 + 1] * f.on_p[o2 + 1] * f.visc2_p[o2 + 1];
  const double pmU = pm[o2 - 1] + pm[o2], pnU = pn[o2 - 1] + pn[o2];
  const double pmV = pm[o2] + pm[o2 - P], pnV = pn[o2] + pn[o2 - P];
  const double pmVb = pm[o2 - P] + pm[o2], pnVb = pn[o2 - P] + pn[o2];
  double rufrc = f.rufrc[o2], rvfrc = dov ? f.rvfrc[o2] : 0.0;
  for (int k = 1; k <= N; ++k) {
    const int o3 = o2 + k * p.PL;
    if (k + PFD <= N) { const int q = o3 + PFD * p.PL; pf_l2(u + q); pf_l2(v + q); pf_l2(Hz + q); pf_l2(un + q); pf_l2(vn + q); }
    const double cr0 = rho_cff(o2, o3), crW = rho_cff(o2 - 1, o3 - 1);
    const double cp0 = psi_cff(o2, o3), cpN = psi_cff(o2 + P, o3 + P);
    {
      const double UFx0 = onr2_0 * cr0, UFxW = onr2_W * crW;
      const double UFe0 = omp2_0 * cp0, UFeN = omp2_N * cpN;
      const double cff = p.dt * 0.25 * pmU * pnU;
      const double cff1 = 0.5 * pnU * (UFx0 - UFxW);
      const double cff2 = 0.5 * pmU * (UFeN - UFe0);
      const double cff3 = cff * (cff1 + cff2);
      rufrc = rufrc + cff1 + cff2;
      un[o3] = un[o3] + cff3;
    }
    if (dov) {
      const double crS = rho_cff(o2 - P, o3 - P), cpE = psi_cff(o2 + 1, o3 + 1);
      const double VFx0 = onp2_0 * cp0, VFxE = onp2_E * cpE;
      const double VFe0 = omr2_0 * cr0, VFeS = omr2_S * crS;
      const double cff = p.dt * 0.25 * pmV * pnV;
      const double cff1 = 0.5 * pnVb * (VFxE - VFx0);
      const double cff2 = 0.5 * pmVb * (VFe0 - VFeS);
      const double cff3 = cff * (cff1 - cff2);
      rvfrc = rvfrc + cff1 - cff2;
      vn[o3] = vn[o3] + cff3;
    }
  }
  f.rufrc[o2] = rufrc;
  if (dov) f.rvfrc[o2] = rvfrc;
}

static inline dim3 g2(dim3 b, int ni, int nj, int nz = 1) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nz); }
void launch_rhs3d(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 2); k_rhs3d<<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f); }
void launch_uv3dmix2(const Par& p, const Flds& f, cudaStream_t s) { dim3 b(64, 2); k_uv3dmix2<<<g2(b, p.Iend - p.Istr + 1, p.Mm), b, 0, s>>>(p, f); }

}  // namespace rb

// REJECTED VARIANT (record only; not part of the library).  rhs3d_tile as a shared-memory tiled kernel: bit-identical to k_rhs3d on
// every parity test, but 0.546 ms against 0.517 ms on BENCHMARK3 (2 CTAs of 256 threads per SM, 128 registers; 1 CTA: 0.755, 3 CTAs:
// 1.15).  Every face flux is evaluated once instead of twice and 12 instead of 62 global loads are issued per point and level, but
// the four flux functions read ~75 values per thread and level from shared memory behind three barriers, which is more LSU work
// than the 62 L1-served loads of the column kernel; profiles/README.md.  To build: paste above launch_rhs3d in csrc/k_rhs.cu and
// launch k_rhs3d_tiled with dim3(RT_X, RT_Y) threads and sizeof(RhsTile) bytes of dynamic shared memory.
// ---------------------------------------------------------------------------------------------------------------
// rhs3d_tile with the horizontal operands and the advective fluxes shared through shared memory.  A CTA owns RT_X x RT_Y columns
// and marches k.  Per level: (1) every thread stages u, v, Huon, Hvom, W, Hz of ITS column, and 176 of the 256 threads one cell
// of the two-cell rim, in a (RT_X+4) x (RT_Y+4) tile; (2) every thread evaluates the four face fluxes of ITS cell once -- UFx at
// its rho point, UFe and VFx at its psi point, VFe at its rho point -- and 80 threads one flux of the one-face rim; (3) the
// divergences take the neighbours' fluxes from the tile.  Against the column kernel: every flux once instead of twice, every
// second difference from shared memory, 12 instead of 62 global loads per point and level.  The expressions and their order are
// those of k_rhs3d (bit-identical results); the loads of level k+1 are issued ahead of the barriers of level k.
#ifndef RHS_TILED
#define RHS_TILED 1
#endif
#ifndef RHS_TMINB
#define RHS_TMINB 2
#endif
constexpr int RT_X = 32, RT_Y = 8, RT_W = RT_X + 4, RT_H = RT_Y + 4, RT_NT = RT_X * RT_Y, RT_RIM = RT_W * RT_H - RT_NT;
static_assert(RT_RIM <= RT_NT && 2 * (RT_X + RT_Y) <= RT_NT, "one rim cell and one rim flux per thread");

struct RhsTile {
  double U[RT_H][RT_W], V[RT_H][RT_W], HU[RT_H][RT_W], HV[RT_H][RT_W], W[RT_H][RT_W], HZ[RT_H][RT_W];
  double FX[RT_Y][RT_X + 1];      // UFx at rho points i0-1 .. i0+RT_X-1            [ty][tx+1], rim column 0
  double FE[RT_Y + 1][RT_X];      // UFe at psi rows  j0 .. j0+RT_Y                  [ty][tx], rim row RT_Y
  double GX[RT_Y][RT_X + 1];      // VFx at psi columns i0 .. i0+RT_X                [ty][tx], rim column RT_X
  double GE[RT_Y + 1][RT_X];      // VFe at rho rows  j0-1 .. j0+RT_Y-1              [ty+1][tx], rim row 0
};
// Face fluxes at tile coordinates (x, y) = cell (i0 - 2 + x, j0 - 2 + y); jg = global row of the cell.  rhs3d.F:658-941.
__device__ __forceinline__ double rt_ufx(const RhsTile& t, int x, int y) {              // rho point between u(x), u(x+1)
  const double Gadv = -0.25;
  const double uxx0 = t.U[y][x - 1] - 2.0 * t.U[y][x] + t.U[y][x + 1], uxx1 = t.U[y][x] - 2.0 * t.U[y][x + 1] + t.U[y][x + 2];
  const double Hxx0 = t.HU[y][x - 1] - 2.0 * t.HU[y][x] + t.HU[y][x + 1], Hxx1 = t.HU[y][x] - 2.0 * t.HU[y][x + 1] + t.HU[y][x + 2];
  const double c1 = t.U[y][x] + t.U[y][x + 1];
  const double c = (c1 > 0.0) ? uxx0 : uxx1;
  return 0.25 * (c1 + Gadv * c) * (t.HU[y][x] + t.HU[y][x + 1] + Gadv * 0.5 * (Hxx0 + Hxx1));
}
__device__ __forceinline__ double rt_uee(const RhsTile& t, int x, int y, int jg, int j0g, int Mm) {   // uee(i, jg); (0) = (1), (Mm+1) = (Mm)
  const int r = (jg < 1) ? 1 : (jg > Mm ? Mm : jg);
  const int yy = y + (r - jg);
  (void)j0g;
  return t.U[yy - 1][x] - 2.0 * t.U[yy][x] + t.U[yy + 1][x];
}
__device__ __forceinline__ double rt_ufe(const RhsTile& t, int x, int y, int jg, int Mm) {          // psi point between u(y-1), u(y)
  const double Gadv = -0.25;
  const double Hvxx0 = t.HV[y][x - 1] - 2.0 * t.HV[y][x] + t.HV[y][x + 1], HvxxW = t.HV[y][x - 2] - 2.0 * t.HV[y][x - 1] + t.HV[y][x];
  const double c1 = t.U[y][x] + t.U[y - 1][x];
  const double c2 = t.HV[y][x] + t.HV[y][x - 1];
  const double c = (c2 > 0.0) ? rt_uee(t, x, y - 1, jg - 1, 0, Mm) : rt_uee(t, x, y, jg, 0, Mm);
  return 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Hvxx0 + HvxxW));
}
__device__ __forceinline__ double rt_vfx(const RhsTile& t, int x, int y) {              // psi point between v(x-1), v(x)
  const double Gadv = -0.25;
  const double vxxW = t.V[y][x - 2] - 2.0 * t.V[y][x - 1] + t.V[y][x], vxx0 = t.V[y][x - 1] - 2.0 * t.V[y][x] + t.V[y][x + 1];
  const double Huee0 = t.HU[y - 1][x] - 2.0 * t.HU[y][x] + t.HU[y + 1][x], HueeS = t.HU[y - 2][x] - 2.0 * t.HU[y - 1][x] + t.HU[y][x];
  const double c1 = t.V[y][x] + t.V[y][x - 1];
  const double c2 = t.HU[y][x] + t.HU[y - 1][x];
  const double c = (c2 > 0.0) ? vxxW : vxx0;
  return 0.25 * (c1 + Gadv * c) * (c2 + Gadv * 0.5 * (Huee0 + HueeS));
}
__device__ __forceinline__ void rt_vee(const RhsTile& t, int x, int y, int jg, int Mm, double& vee, double& Hvee) {   // rows 2..Mm; (1) = (2), (Mm+1) = (Mm)
  const int r = (jg < 2) ? 2 : (jg > Mm ? Mm : jg);
  const int yy = y + (r - jg);
  vee = t.V[yy - 1][x] - 2.0 * t.V[yy][x] + t.V[yy + 1][x];
  Hvee = t.HV[yy - 1][x] - 2.0 * t.HV[yy][x] + t.HV[yy + 1][x];
}
__device__ __forceinline__ double rt_vfe(const RhsTile& t, int x, int y, int jg, int Mm) {          // rho point between v(y), v(y+1)
  const double Gadv = -0.25;
  double vee0, Hvee0, vee1, Hvee1;
  rt_vee(t, x, y, jg, Mm, vee0, Hvee0);
  rt_vee(t, x, y + 1, jg + 1, Mm, vee1, Hvee1);
  const double c1 = t.V[y][x] + t.V[y + 1][x];
  const double c = (c1 > 0.0) ? vee0 : vee1;
  return 0.25 * (c1 + Gadv * c) * (t.HV[y][x] + t.HV[y + 1][x] + Gadv * 0.5 * (Hvee0 + Hvee1));
}

__global__ void __launch_bounds__(RT_NT, RHS_TMINB) k_rhs3d_tiled(Par p, Flds f) {
  extern __shared__ double rt_smem[];
  RhsTile& t = *reinterpret_cast<RhsTile*>(rt_smem);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * RT_X + tx;
  const int i0 = p.Istr + blockIdx.x * RT_X, j0 = 1 + blockIdx.y * RT_Y;
  const int i = i0 + tx, j = j0 + ty;
  const bool own = (i <= p.Iend && j <= p.Mm);
  const int N = p.N, P = p.P, PL = p.PL, Mm = p.Mm;
  auto clampi = [&](int ii) { return ii > p.Iend + 2 ? p.Iend + 2 : ii; };
  auto clampj = [&](int jj) { return jj < 0 ? 0 : (jj > Mm + 1 ? Mm + 1 : jj); };
  const double* __restrict__ u = f.u[p.nrhs];
  const double* __restrict__ v = f.v[p.nrhs];
  const double* __restrict__ Hz = f.Hz;
  const double* __restrict__ Huon = f.Huon;
  const double* __restrict__ Hvom = f.Hvom;
  const double* __restrict__ W = f.W;
  double* __restrict__ ru = f.ru[p.nrhs];
  double* __restrict__ rv = f.rv[p.nrhs];
  const bool dov = (j >= p.JstrV);
  // own column (clamped for the threads beyond the range: they stage valid data and store nothing)
  const int oo = clampj(j) * P + clampi(i);
  // rim cell of this thread: rows 0,1 and RT_H-2,RT_H-1 (RT_W cells each), then columns 0,1 and RT_W-2,RT_W-1 of the middle rows
  int rx = 0, ry = 0;
  const bool rim = tid < RT_RIM;
  if (tid < 2 * RT_W) { ry = tid / RT_W; rx = tid % RT_W; }
  else if (tid < 4 * RT_W) { ry = RT_H - 2 + (tid - 2 * RT_W) / RT_W; rx = (tid - 2 * RT_W) % RT_W; }
  else if (rim) { const int q = tid - 4 * RT_W; ry = 2 + q / 4; const int cc = q % 4; rx = (cc < 2) ? cc : RT_W - 4 + cc; }
  const int orim = clampj(j0 - 2 + ry) * P + clampi(i0 - 2 + rx);
  // rim flux of this thread (tid < 2*(RT_X+RT_Y)): UFx column i0-1, UFe row j0+RT_Y, VFx column i0+RT_X, VFe row j0-1
  const int fk = (tid < RT_Y) ? 0 : (tid < RT_Y + RT_X) ? 1 : (tid < 2 * RT_Y + RT_X) ? 2 : (tid < 2 * (RT_X + RT_Y)) ? 3 : -1;
  const int fq = (fk == 0) ? tid : (fk == 1) ? tid - RT_Y : (fk == 2) ? tid - RT_Y - RT_X : tid - 2 * RT_Y - RT_X;
  // 2-D factors
  double fomn0 = 0, fomnW = 0, fomnS = 0, dndx0 = 0, dndxW = 0, dndxS = 0, dmde0 = 0, dmdeW = 0, dmdeS = 0;
  if (own) {
    const int o2 = j * P + i;
    fomn0 = f.fomn[o2]; fomnW = f.fomn[o2 - 1]; fomnS = f.fomn[o2 - P];
    if (p.curvgrid) {
      dndx0 = f.dndx[o2]; dndxW = f.dndx[o2 - 1]; dndxS = f.dndx[o2 - P];
      dmde0 = f.dmde[o2]; dmdeW = f.dmde[o2 - 1]; dmdeS = f.dmde[o2 - P];
    }
  }
  double FCu_m = 0.0, FCv_m = 0.0, rufrc = 0.0, rvfrc = 0.0;
  // software pipeline: the operands of level k+1 (own column and rim cell) and u, v of level k+3 (vertical advection) are
  // requested before level k is staged
  struct Cell { double hu, hv, w, hz; };
  auto ld = [&](int o) -> Cell { return Cell{Huon[o], Hvom[o], W[o], Hz[o]}; };
  Cell nc = ld(oo + PL);
  Cell nr = nc;
  double nru_ = 0.0, nrv_ = 0.0;                              // u, v of the rim cell
  if (rim) { nr = ld(orim + PL); nru_ = u[orim + PL]; nrv_ = v[orim + PL]; }
  double ua = 0.0, ub = u[oo + PL], uc = u[oo + 2 * PL], ud = u[oo + 3 * PL];      // u(k-1), u(k), u(k+1), u(k+2)
  double va = 0.0, vb = v[oo + PL], vc = v[oo + 2 * PL], vd = v[oo + 3 * PL];
  double nrux = ru[oo + PL], nrvx = rv[oo + PL];
  const int X = tx + 2, Y = ty + 2;
  for (int k = 1; k <= N; ++k) {
    const Cell c = nc, r = nr;
    const double r_u = nru_, r_v = nrv_;
    const double rux0 = nrux, rvx0 = nrvx;
    double un = ud, vn = vd;                                  // becomes u(k+3), v(k+3)
    if (k < N) {
      const int o = oo + (k + 1) * PL;
      pf_up<RHS_PF>(Hz, o, k + 1, N, PL); pf_up<RHS_PF>(u, o, k + 1, N, PL); pf_up<RHS_PF>(v, o, k + 1, N, PL); pf_up<RHS_PF>(Huon, o, k + 1, N, PL);
      pf_up<RHS_PF>(Hvom, o, k + 1, N, PL); pf_up<RHS_PF>(W, o, k + 1, N, PL); pf_up<RHS_PF>(ru, o, k + 1, N, PL); pf_up<RHS_PF>(rv, o, k + 1, N, PL);
      nc = ld(o);
      if (rim) { const int q = orim + (k + 1) * PL; nr = ld(q); nru_ = u[q]; nrv_ = v[q]; }
      nrux = ru[o]; nrvx = rv[o];
      if (k + 3 <= N) { un = u[oo + (k + 3) * PL]; vn = v[oo + (k + 3) * PL]; }
    }
    // (1) stage level k
    t.U[Y][X] = ub; t.V[Y][X] = vb; t.HU[Y][X] = c.hu; t.HV[Y][X] = c.hv; t.W[Y][X] = c.w; t.HZ[Y][X] = c.hz;
    if (rim) { t.U[ry][rx] = r_u; t.V[ry][rx] = r_v; t.HU[ry][rx] = r.hu; t.HV[ry][rx] = r.hv; t.W[ry][rx] = r.w; t.HZ[ry][rx] = r.hz; }
    __syncthreads();
    // (2) face fluxes, once each
    t.FX[ty][tx + 1] = rt_ufx(t, X, Y);
    t.FE[ty][tx] = rt_ufe(t, X, Y, j, Mm);
    t.GX[ty][tx] = rt_vfx(t, X, Y);
    t.GE[ty + 1][tx] = rt_vfe(t, X, Y, j, Mm);
    if (fk == 0) t.FX[fq][0] = rt_ufx(t, 1, fq + 2);
    else if (fk == 1) t.FE[RT_Y][fq] = rt_ufe(t, fq + 2, RT_Y + 2, j0 + RT_Y, Mm);
    else if (fk == 2) t.GX[fq][RT_X] = rt_vfx(t, RT_X + 2, fq + 2);
    else if (fk == 3) t.GE[0][fq] = rt_vfe(t, fq + 2, 1, j0 - 1, Mm);
    __syncthreads();
    // (3) the right-hand sides of this level
    if (own) {
      const int o = oo + k * PL;
      double rux = rux0, rvx = rvx0;
      const double hz0 = c.hz, hzW = t.HZ[Y][X - 1], hzS = t.HZ[Y - 1][X];
      const double u0 = ub, v0 = vb;
      const double uW = t.U[Y][X - 1], uE = t.U[Y][X + 1], uS = t.U[Y - 1][X], uSE = t.U[Y - 1][X + 1];
      const double vW = t.V[Y][X - 1], vN = t.V[Y + 1][X], vS = t.V[Y - 1][X], vNW = t.V[Y + 1][X - 1];
      {                                                     // Coriolis (rhs3d.F:473-507)
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
      if (p.curvgrid) {                                     // curvilinear terms (rhs3d.F:515-564)
        double c1 = 0.5 * (v0 + vN), c2 = 0.5 * (u0 + uE);
        double cc = hz0 * (c1 * dndx0 - c2 * dmde0);
        const double UFx0 = cc * c1, VFe0 = cc * c2;
        c1 = 0.5 * (vW + vNW); c2 = 0.5 * (uW + u0);
        cc = hzW * (c1 * dndxW - c2 * dmdeW);
        const double UFxW = cc * c1;
        rux = rux + 0.5 * (UFx0 + UFxW);
        if (dov) {
          c1 = 0.5 * (vS + v0); c2 = 0.5 * (uS + uSE);
          cc = hzS * (c1 * dndxS - c2 * dmdeS);
          const double VFeS = cc * c2;
          rvx = rvx - 0.5 * (VFe0 + VFeS);
        }
      }
      {                                                     // horizontal advection of u (rhs3d.F:658-798, :946-963)
        const double a1 = t.FX[ty][tx + 1] - t.FX[ty][tx];
        const double a2 = t.FE[ty + 1][tx] - t.FE[ty][tx];
        rux = rux - (a1 + a2);
      }
      if (dov) {                                            // horizontal advection of v (rhs3d.F:800-940, :965-982)
        const double a1 = t.GX[ty][tx + 1] - t.GX[ty][tx];
        const double a2 = t.GE[ty + 1][tx] - t.GE[ty][tx];
        rvx = rvx - (a1 + a2);
      }
      {                                                     // vertical advection (rhs3d.F:1177-1265, :1434-1522)
        const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
        double FCu = 0.0, FCv = 0.0;
        if (k < N) {
          const double W0 = c.w, WW = t.W[Y][X - 1], WE = t.W[Y][X + 1], WW2 = t.W[Y][X - 2], WS = t.W[Y - 1][X], WN = t.W[Y + 1][X], WS2 = t.W[Y - 2][X];
          const double ukm = (k > 1) ? ua : u0;
          const double ukpp = (k + 2 <= N) ? ud : uc;
          FCu = (c1 * (u0 + uc) - c2 * (ukm + ukpp)) * (c1 * (W0 + WW) - c2 * (WE + WW2));
          if (dov) {
            const double vkm = (k > 1) ? va : v0;
            const double vkpp = (k + 2 <= N) ? vd : vc;
            FCv = (c1 * (v0 + vc) - c2 * (vkm + vkpp)) * (c1 * (W0 + WS) - c2 * (WN + WS2));
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
    __syncthreads();                                        // the tile is overwritten by the next level
    ua = ub; ub = uc; uc = ud; ud = un;
    va = vb; vb = vc; vc = vd; vd = vn;
  }
  // ---- column sums + surface/bottom stresses (rhs3d.F:1534-1667)
  if (own) {
    const int o2 = j * P;
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
}


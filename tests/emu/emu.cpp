// TEST INFRASTRUCTURE (tests/emu): a host harness around the GENERATED host copies of the barrier-free CUDA kernels
// (gen.py).  It owns a field table laid out like the device one (common pitch, pre-offset pointers, zeroed guard rows), fills
// `Par` the way csrc/api.cu does for a single EW-periodic tile, and runs one phase = the same launch wrappers the library
// calls, every thread executed in turn on the CPU.  Built -O2 -ffp-contract=off like the oracle's parity build, the results
// must equal the oracle's bit for bit: the CPU-side check of a kernel's indices, branches and operation order.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "cuda_runtime.h"
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
EmuBlockBarrier emu_block_barrier;
EmuWarp emu_warps[64];
void __syncthreads() { emu_block_barrier.wait(); }
double* emu_dynsmem() { static std::vector<double> buf(64 * 1024); return buf.data(); }      // 512 KB >= any CTA's dynamic shared memory
unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { static std::mutex m; std::lock_guard<std::mutex> g(m); const unsigned long long o = *a; *a = o + v; return o; }
#include "kernels.h"

using namespace rb;

namespace {
struct Slot { double** slot; int LBk, nk; std::vector<double> store; };
struct Emu {
  Par p; Flds f;
  int ni, nj, LBi, LBj, ioff, dj_gradps, mix_geo_ts, ana_vmix, ts_dif4;
  std::map<std::string, Slot> reg;
  std::vector<double> diag_partial, diag_out;
  void add(const std::string& name, double** slot, int LBk, int nk) {
    Slot& s = reg[name];
    s.slot = slot; s.LBk = LBk; s.nk = nk;
    s.store.assign((size_t)p.PL * nk + 64, 0.0);
    *slot = s.store.data() + 32 + ioff - LBi + (2 - LBj) * p.P - (long)LBk * p.PL;
  }
};
}  // namespace

extern "C" {

void* emu_create(const int* io, const double* dv) {
  Emu* e = new Emu();
  Par& p = e->p; Flds& f = e->f;
  std::memset(&p, 0, sizeof(p)); std::memset(&f, 0, sizeof(f));
  p.Lm = io[0]; p.Mm = io[1]; p.N = io[2]; p.NT = io[3];
  const int Lm = p.Lm, Mm = p.Mm, N = p.N;
  // xi-tile `tile` of NtileI (io[25], io[26]; 1, 0 = the whole periodic domain): get_bounds.F tile_bounds_2d as csrc/api.cu make_bounds
  const int NtileI = io[25] > 0 ? io[25] : 1, tile = io[26];
  int is = 1, ie = Lm;
  if (NtileI > 1) {
    const int chunk = (Lm + NtileI - 1) / NtileI, margin = (NtileI * chunk - Lm) / 2;
    is = 1 + tile * chunk - margin; ie = is + chunk - 1;
    if (is < 1) is = 1;
    if (ie > Lm) ie = Lm;
  }
  // device extents of a tile: Istr-3 .. Iend+2 (csrc/api.cu: LBi_dev = Istr - 3, UBi = Iend + Nghost); the single tile -2 .. Lm+2
  e->LBi = (NtileI > 1) ? is - 3 : -2; e->LBj = 0; e->ni = ((NtileI > 1) ? ie + 2 : Lm + 2) - e->LBi + 1; e->nj = Mm + 2;
  e->ioff = (16 - ((is - e->LBi) % 16)) % 16;                      // i = Istr on a 128-byte boundary, as on the device
  p.P = ((e->ioff + e->ni + 15) / 16) * 16; p.PL = p.P * (e->nj + 4);
  p.LBi = e->LBi; p.UBi = e->LBi + e->ni - 1; p.LBj = 0; p.UBj = Mm + 1;
  // get_bounds.F var_bounds of an EW-periodic, NS-closed grid: no clipping in xi on any tile, the eta clips of the edge rows
  p.Istr = is; p.Iend = ie; p.Jstr = 1; p.Jend = Mm; p.IstrU = is; p.JstrV = 2; p.JstrR = 0; p.JendR = Mm + 1;
  p.Jstrm1 = 1; p.Jendp1 = Mm; p.Jendp2 = Mm + 1; p.JstrVm1 = 2; p.JstrVm2 = 1;
  p.ew_wrap = (NtileI == 1) ? 1 : 0; p.gap_at = 0x7fffffff; p.gap_len = 0;
  p.nonlin_eos = io[4]; p.curvgrid = io[5]; p.uv_qdrag = io[6]; p.salinity = io[7]; p.hadv = io[8]; p.vadv = io[9]; p.itemp = io[10]; p.isalt = io[11];
  p.bv_frequency = io[12]; p.eos_tderivative = io[13]; p.solar_source = io[14]; p.lmd_nonlocal = io[15]; p.bulk_fluxes = io[16]; p.lmd_mixing = io[17];
  p.uv_adv = io[18]; e->ts_dif4 = io[19]; e->dj_gradps = io[20]; e->mix_geo_ts = io[21]; e->ana_vmix = io[22];
  p.dt = dv[0]; p.g = dv[1]; p.rho0 = dv[2]; p.R0 = dv[3]; p.T0 = dv[4]; p.S0 = dv[5]; p.Tcoef = dv[6]; p.Scoef = dv[7];
  p.gamma2 = dv[8]; p.lambda = dv[9]; p.hc = dv[10]; p.Akv_bak = dv[11]; p.Akt_bak[0] = dv[12]; p.Akt_bak[1] = dv[13];
  p.blk_ZQ = dv[14]; p.blk_ZT = dv[15]; p.blk_ZW = dv[16];
  p.dtfast = p.dt / (double)io[23]; p.limit_bstress = io[24]; p.nospl_vvisc = io[27]; p.nospl_vdiff = io[28];
  p.qcorrection = io[29]; p.limit_stflx_cooling = io[30]; p.scorrection = io[31]; p.Tnudg_salt = dv[17];
  p.bodyforce = io[32]; p.levsfrc = io[33]; p.levbfrc = io[34]; p.vtransform = (io[35] == 1) ? 1 : 2; p.atm_press = io[36];
#define A2(name) e->add(#name, &f.name, 0, 1)
#define A3(name, k0, nk) e->add(#name, &f.name, k0, nk)
  A2(h); A2(f); A2(pm); A2(pn); A2(om_r); A2(on_r); A2(om_u); A2(on_u); A2(om_v); A2(on_v); A2(om_p); A2(on_p); A2(omn); A2(fomn);
  A2(pmon_r); A2(pnom_r); A2(pmon_u); A2(pnom_u); A2(pmon_v); A2(pnom_v); A2(pmon_p); A2(pnom_p); A2(dndx); A2(dmde); A2(rdrag); A2(rdrag2);
  A2(visc2_r); A2(visc2_p); A2(Zt_avg1); A2(DU_avg1); A2(DU_avg2); A2(DV_avg1); A2(DV_avg2); A2(rufrc); A2(rvfrc); A2(rhoA); A2(rhoS);
  A2(sustr); A2(svstr); A2(bustr); A2(bvstr); A2(ZoBot); A2(alpha); A2(beta); A2(srflx); A2(Jwtype);
  A2(Uwind); A2(Vwind); A2(Tair); A2(Pair); A2(Hair); A2(rain); A2(cloud); A2(lrflx); A2(lhflx); A2(shflx); A2(Taux); A2(Tauy); A2(hsbl); A2(ksbl); A2(sst); A2(dqdt); A2(sss);
  A3(rho, 1, N); A3(pden, 1, N); A3(Hz, 1, N); A3(z_r, 1, N); A3(Huon, 1, N); A3(Hvom, 1, N); A3(W, 0, N + 1); A3(wvel, 0, N + 1);
  A3(z_w, 0, N + 1); A3(Akv, 0, N + 1); A3(P3, 1, N); A3(bvf, 0, N + 1);
#undef A2
#undef A3
  for (int k = 1; k <= 3; ++k) {
    const std::string s = std::to_string(k);
    e->add("zeta" + s, &f.zeta[k], 0, 1); e->add("ubar" + s, &f.ubar[k], 0, 1); e->add("vbar" + s, &f.vbar[k], 0, 1);
  }
  for (int k = 1; k <= 2; ++k) {
    const std::string s = std::to_string(k);
    e->add("rzeta" + s, &f.rzeta[k], 0, 1); e->add("rubar" + s, &f.rubar[k], 0, 1); e->add("rvbar" + s, &f.rvbar[k], 0, 1);
    e->add("u" + s, &f.u[k], 1, N); e->add("v" + s, &f.v[k], 1, N); e->add("ru" + s, &f.ru[k], 0, N + 1); e->add("rv" + s, &f.rv[k], 0, N + 1);
  }
  for (int it = 0; it < p.NT; ++it) {
    const std::string s = std::to_string(it);
    for (int k = 1; k <= 3; ++k) e->add("t" + std::to_string(k) + "_" + s, &f.t[k][it], 1, N);
    e->add("Akt_" + s, &f.Akt[it], 0, N + 1); e->add("diff2_" + s, &f.diff2[it], 0, 1); e->add("diff4_" + s, &f.diff4[it], 0, 1);
    e->add("stflx_" + s, &f.stflx[it], 0, 1); e->add("btflx_" + s, &f.btflx[it], 0, 1);
    e->add("stflux_" + s, &f.stflux[it], 0, 1); e->add("btflux_" + s, &f.btflux[it], 0, 1); e->add("ghats_" + s, &f.ghats[it], 0, N + 1);
  }
  static thread_local std::vector<double> sc;
  sc.assign(4 * (MAXN + 1), 0.0);
  f.sc_r = sc.data(); f.Cs_r = sc.data() + (MAXN + 1); f.sc_w = sc.data() + 2 * (MAXN + 1); f.Cs_w = sc.data() + 3 * (MAXN + 1);
  return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }

// host layout: (nk, nj, ni) = Fortran A(LBi:UBi, LBj:UBj, LBk:)
int emu_xfer(void* h, const char* name, double* host, int up) {
  Emu* e = (Emu*)h;
  auto it = e->reg.find(name);
  if (it == e->reg.end()) return 2;
  Slot& s = it->second;
  double* A = *s.slot;
  for (int k = 0; k < s.nk; ++k)
    for (int j = 0; j < e->nj; ++j) {
      double* d = A + e->LBi + (long)(j + e->LBj) * e->p.P + (long)(k + s.LBk) * e->p.PL;
      double* hp = host + ((size_t)k * e->nj + j) * e->ni;
      if (up) std::memcpy(d, hp, sizeof(double) * e->ni); else std::memcpy(hp, d, sizeof(double) * e->ni);
    }
  return 0;
}
// the 12 scalars of roms_b200_diag from the device-side results of phase 18 (csrc/api.cu finish_diag)
void emu_diag(void* h, double* out12) {
  const double* d = ((Emu*)h)->diag_out.data();
  const double vol = d[2];
  out12[0] = d[0] / vol; out12[1] = d[1] / vol; out12[2] = out12[0] + out12[1]; out12[3] = vol;
  out12[4] = d[7]; out12[5] = d[4]; out12[6] = d[5]; out12[7] = d[6];
  out12[8] = d[11]; out12[9] = d[12]; out12[10] = d[9]; out12[11] = d[10];
}
// first array column (Fortran index) and number of columns of this tile's arrays
void emu_extent(void* h, int* out2) { Emu* e = (Emu*)h; out2[0] = e->LBi; out2[1] = e->ni; }
int emu_levels(void* h, const char* name) { Emu* e = (Emu*)h; auto it = e->reg.find(name); return it == e->reg.end() ? -1 : it->second.nk; }
void emu_scoord(void* h, int which, const double* v, int n) {
  Emu* e = (Emu*)h;
  double* d = which == 0 ? e->f.sc_r : which == 1 ? e->f.Cs_r : which == 2 ? e->f.sc_w : e->f.Cs_w;
  std::memcpy(d, v, sizeof(double) * n);
}
// the 2-D indices of one step2d call (main3d.F:599-661) and its filter weights weight(1,iif-1), weight(2,iif), weight(2,iif+1)
void emu_indices2d(void* h, int iif, int kstp, int krhs, int knew, int predictor, int nfast, double w1_m1, double w2_0, double w2_p1) {
  Par& p = ((Emu*)h)->p;
  p.iif = iif; p.kstp = kstp; p.krhs = krhs; p.knew = knew; p.ptsk = 3 - kstp; p.predictor = predictor; p.nfast = nfast;
  p.w1_m1 = w1_m1; p.w2_0 = w2_0; p.w2_p1 = w2_p1;
}
void emu_indices(void* h, int nstp, int nnew, int nrhs, int istart) { Par& p = ((Emu*)h)->p; p.nstp = nstp; p.nnew = nnew; p.nrhs = nrhs; p.istart = istart; }

// one phase, as csrc/api.cu run_phase_async issues it for a single tile (phase numbers of include/roms_b200.h)
int emu_run(void* h, int phase) {
  Emu* e = (Emu*)h;
  const Par& p = e->p; const Flds& f = e->f;
  cudaStream_t s = nullptr;
  switch (phase) {
    case 1: launch_set_massflux(p, f, s); break;
    case 2: launch_rho_eos(p, f, s); break;
    case 3: launch_set_vbc(p, f, s); break;
    case 4: if (e->ana_vmix) launch_ana_vmix(p, f, s); break;
    case 5: case 16: launch_omega(p, f, s); break;
    case 6: launch_wvelocity(p, f, p.nstp, s); break;
    case 7: launch_set_zeta(p, f, s); break;
    case 8: launch_pre_step3d_t(p, f, s); launch_pre_step3d_uv(p, f, s); break;
    case 9: launch_prsgrd(p, f, e->dj_gradps, s); break;
    case 10:
      if (e->mix_geo_ts) launch_t3dmix2_geo(p, f, s); else launch_t3dmix2_s(p, f, s);
      break;
    case 26: launch_t3dmix4_s(p, f, s); break;
    case 11: launch_rhs3d(p, f, s); break;
    case 12: launch_uv3dmix2(p, f, s); break;
    case 13: launch_step2d(p, f, s, nullptr); break;
    case 14: launch_set_depth(p, f, s); break;
    case 15: launch_step3d_uv(p, f, s); break;
    case 17: launch_step3d_t(p, f, s); break;
    case 18:                                                       // diag.F: the three kernels, then csrc/api.cu finish_diag
      e->diag_partial.assign((size_t)diag_partial_doubles(p) + 16, 0.0); e->diag_out.assign(16, 0.0);
      launch_diag(p, f, e->diag_partial.data(), e->diag_out.data(), p.knew, s);
      break;
    case 23: launch_bulk_flux(p, f, s); break;
    case 24: launch_lmd_vmix(p, f, s); break;
    case 25: launch_bvf_mix(p, f, s); break;
    default: return 5;
  }
  return 0;
}
}  // extern "C"

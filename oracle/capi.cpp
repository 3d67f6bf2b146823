// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).  extern "C" surface for tests/ (ctypes) and for the
// bench.py cpu_baseline / --impl reference legs.
#include "roms_oracle.hpp"
#include <cstring>
#include <map>
#include <chrono>

using namespace orc;

namespace {
struct Handle { Model m; bool ready = false; };

bool lookup(Model& m, const std::string& name, double** p, int dims[6]) {
  auto set2 = [&](F2 a) { *p = a.p; dims[0] = a.LBi; dims[1] = a.UBi - a.LBi + 1; dims[2] = a.LBj; dims[3] = a.UBj - a.LBj + 1; dims[4] = 0; dims[5] = 1; return true; };
  auto set3 = [&](F3 a) { *p = a.p; dims[0] = a.LBi; dims[1] = a.UBi - a.LBi + 1; dims[2] = a.LBj; dims[3] = a.UBj - a.LBj + 1; dims[4] = a.LBk; dims[5] = a.UBk - a.LBk + 1; return true; };
  static const char* n2[] = {"h", "f", "pm", "pn", "om_r", "on_r", "om_u", "on_u", "om_v", "on_v", "om_p", "on_p", "omn", "fomn", "pmon_r", "pnom_r",
                             "pmon_u", "pnom_u", "pmon_v", "pnom_v", "pmon_p", "pnom_p", "dndx", "dmde", "rdrag", "rdrag2", "ZoBot", "visc2_r", "visc2_p",
                             "Zt_avg1", "DU_avg1", "DU_avg2", "DV_avg1", "DV_avg2", "rufrc", "rvfrc", "rhoA", "rhoS", "sustr", "svstr", "bustr", "bvstr",
                             "avgzeta", "avgu2d", "avgv2d", "alpha", "beta", "srflx", "Jwtype",
                             "Uwind", "Vwind", "Tair", "Pair", "Hair", "rain", "cloud", "lrflx", "lhflx", "shflx", "hsbl", "ksbl", "latr", "lonr", "sst", "dqdt", "sss"};
  F2* a2[] = {&m.h, &m.f, &m.pm, &m.pn, &m.om_r, &m.on_r, &m.om_u, &m.on_u, &m.om_v, &m.on_v, &m.om_p, &m.on_p, &m.omn, &m.fomn, &m.pmon_r, &m.pnom_r,
              &m.pmon_u, &m.pnom_u, &m.pmon_v, &m.pnom_v, &m.pmon_p, &m.pnom_p, &m.dndx, &m.dmde, &m.rdrag, &m.rdrag2, &m.ZoBot, &m.visc2_r, &m.visc2_p,
              &m.Zt_avg1, &m.DU_avg1, &m.DU_avg2, &m.DV_avg1, &m.DV_avg2, &m.rufrc, &m.rvfrc, &m.rhoA, &m.rhoS, &m.sustr, &m.svstr, &m.bustr, &m.bvstr,
              &m.avgzeta, &m.avgu2d, &m.avgv2d, &m.alpha, &m.beta, &m.srflx, &m.Jwtype,
              &m.Uwind, &m.Vwind, &m.Tair, &m.Pair, &m.Hair, &m.rain, &m.cloud, &m.lrflx, &m.lhflx, &m.shflx, &m.hsbl, &m.ksbl, &m.latr, &m.lonr, &m.sst, &m.dqdt, &m.sss};
  for (size_t i = 0; i < sizeof(n2) / sizeof(n2[0]); ++i) if (name == n2[i]) return set2(*a2[i]);
  static const char* n3[] = {"rho", "pden", "Hz", "z_r", "Huon", "Hvom", "W", "wvel", "z_w", "Akv", "avgu3d", "avgv3d", "avgrho", "avgw3d", "avgwvel", "bvf"};
  F3* a3[] = {&m.rho, &m.pden, &m.Hz, &m.z_r, &m.Huon, &m.Hvom, &m.W, &m.wvel, &m.z_w, &m.Akv, &m.avgu3d, &m.avgv3d, &m.avgrho, &m.avgw3d, &m.avgwvel, &m.bvf};
  for (size_t i = 0; i < sizeof(n3) / sizeof(n3[0]); ++i) if (name == n3[i]) return set3(*a3[i]);
  // indexed names: zeta1..3, ubar1..3, vbar1..3, rzeta1..2, rubar1..2, rvbar1..2, u1..2, v1..2, ru1..2, rv1..2,
  // t<tl>_<itrc>, Akt_<itrc>, diff2_<itrc>, stflx_<itrc>, btflx_<itrc>, stflux_<itrc>, btflux_<itrc>
  auto idx = [&](const char* base, int maxi) -> int {
    size_t L = std::strlen(base);
    if (name.size() == L + 1 && name.compare(0, L, base) == 0) { int k = name[L] - '0'; if (k >= 1 && k <= maxi) return k; }
    return 0;
  };
  int k;
  if ((k = idx("zeta", 3))) return set2(m.zeta[k]);
  if ((k = idx("ubar", 3))) return set2(m.ubar[k]);
  if ((k = idx("vbar", 3))) return set2(m.vbar[k]);
  if ((k = idx("rzeta", 2))) return set2(m.rzeta[k]);
  if ((k = idx("rubar", 2))) return set2(m.rubar[k]);
  if ((k = idx("rvbar", 2))) return set2(m.rvbar[k]);
  if ((k = idx("ru", 2))) return set3(m.ru[k]);
  if ((k = idx("rv", 2))) return set3(m.rv[k]);
  if ((k = idx("u", 2))) return set3(m.u[k]);
  if ((k = idx("v", 2))) return set3(m.v[k]);
  if (name.size() == 4 && name[0] == 't' && name[2] == '_') {
    int tl = name[1] - '0', it = name[3] - '0';
    if (tl >= 1 && tl <= 3 && it >= 0 && it < m.c.NT) return set3(m.t[tl][it]);
  }
  auto trc = [&](const char* base) -> int {
    size_t L = std::strlen(base);
    if (name.size() == L + 1 && name.compare(0, L, base) == 0) { int it = name[L] - '0'; if (it >= 0 && it < m.c.NT) return it; }
    return -1;
  };
  int it;
  if ((it = trc("Akt_")) >= 0) return set3(m.Akt[it]);
  if ((it = trc("avgt_")) >= 0) return set3(m.avgt[it]);
  if ((it = trc("ghats_")) >= 0) return set3(m.ghats[it]);
  if ((it = trc("diff2_")) >= 0) return set2(m.diff2[it]);
  if ((it = trc("diff4_")) >= 0) return set2(m.diff4[it]);
  if ((it = trc("stflx_")) >= 0) return set2(m.stflx[it]);
  if ((it = trc("btflx_")) >= 0) return set2(m.btflx[it]);
  if ((it = trc("stflux_")) >= 0) return set2(m.stflux[it]);
  if ((it = trc("btflux_")) >= 0) return set2(m.btflux[it]);
  return false;
}
}  // namespace

extern "C" {

void* orc_create(int app, int Lm, int Mm, int N, int NtileI, int NtileJ) {
  Handle* h = new Handle();
  h->m.c = make_cfg(app, Lm, Mm, N);
  h->m.c.NtileI = NtileI; h->m.c.NtileJ = NtileJ;
  return h;
}
void orc_destroy(void* hp) { delete (Handle*)hp; }

// options must be set before orc_init
int orc_set_option(void* hp, const char* key, double val) {
  Cfg& c = ((Handle*)hp)->m.c; std::string k(key);
  if (k == "nonlin_eos") c.nonlin_eos = (int)val; else if (k == "dj_gradps") c.dj_gradps = (int)val;
  else if (k == "curvgrid") c.curvgrid = (int)val; else if (k == "mix_geo_ts") c.mix_geo_ts = (int)val;
  else if (k == "uv_qdrag") c.uv_qdrag = (int)val; else if (k == "hadv") c.hadv = (int)val; else if (k == "vadv") c.vadv = (int)val;
  else if (k == "ana_vmix") c.ana_vmix = (int)val; else if (k == "wvelocity_every_step") c.wvelocity_every_step = (int)val;
  else if (k == "dt") c.dt = val; else if (k == "ndtfast") c.ndtfast = (int)val; else if (k == "visc2") c.visc2 = val;
  else if (k == "tnu2") { c.tnu2[0] = c.tnu2[1] = val; } else if (k == "gamma2") c.gamma2 = val;
  else if (k == "Akv_bak") c.Akv_bak = val; else if (k == "Akt_bak") { c.Akt_bak[0] = c.Akt_bak[1] = val; }
  else if (k == "rdrg") c.rdrg = val; else if (k == "rdrg2") c.rdrg2 = val; else if (k == "Zob") c.Zob = val;
  else if (k == "bv_frequency") c.bv_frequency = (int)val; else if (k == "eos_tderivative") c.eos_tderivative = (int)val;
  else if (k == "solar_source") c.solar_source = (int)val; else if (k == "lmd_nonlocal") c.lmd_nonlocal = (int)val;
  else if (k == "nAVG") c.nAVG = (int)val; else if (k == "ntsAVG") c.ntsAVG = (int)val;
  else if (k == "bvf_mixing") c.bvf_mixing = (int)val;
  else if (k == "uv_adv") c.uv_adv = (int)val; else if (k == "ts_dif4") c.ts_dif4 = (int)val;
  else if (k == "tnu4") { c.tnu4[0] = c.tnu4[1] = val; } else if (k == "limit_bstress") c.limit_bstress = (int)val;
  else if (k == "qcorrection") c.qcorrection = (int)val; else if (k == "limit_stflx_cooling") c.limit_stflx_cooling = (int)val;
  else if (k == "scorrection") c.scorrection = (int)val; else if (k == "Tnudg_salt") c.Tnudg_salt = val;
  else if (k == "Vtransform") c.Vtransform = (int)val; else if (k == "atm_press") c.atm_press = (int)val;
  else if (k == "bodyforce") c.bodyforce = (int)val; else if (k == "levsfrc") c.levsfrc = (int)val; else if (k == "levbfrc") c.levbfrc = (int)val;
  else if (k == "nospl_vvisc") c.nospl_vvisc = (int)val; else if (k == "nospl_vdiff") c.nospl_vdiff = (int)val;
  else if (k == "rdrg2") c.rdrg2 = val; else if (k == "rdrg") c.rdrg = val;
  else if (k == "bulk_fluxes") c.bulk_fluxes = (int)val; else if (k == "lmd_mixing") c.lmd_mixing = (int)val;
  else return 1;
  return 0;
}
double orc_get_option(void* hp, const char* key) {
  Model& m = ((Handle*)hp)->m; Cfg& c = m.c; std::string k(key);
  if (k == "nonlin_eos") return c.nonlin_eos; if (k == "dj_gradps") return c.dj_gradps; if (k == "curvgrid") return c.curvgrid;
  if (k == "mix_geo_ts") return c.mix_geo_ts; if (k == "uv_qdrag") return c.uv_qdrag; if (k == "hadv") return c.hadv; if (k == "vadv") return c.vadv;
  if (k == "ana_vmix") return c.ana_vmix; if (k == "dt") return c.dt; if (k == "ndtfast") return c.ndtfast; if (k == "visc2") return c.visc2;
  if (k == "tnu2") return c.tnu2[0]; if (k == "gamma2") return c.gamma2; if (k == "Akv_bak") return c.Akv_bak; if (k == "Akt_bak") return c.Akt_bak[0];
  if (k == "rdrg") return c.rdrg; if (k == "rdrg2") return c.rdrg2; if (k == "salinity") return c.salinity; if (k == "NT") return c.NT;
  if (k == "Lm") return c.Lm; if (k == "Mm") return c.Mm; if (k == "N") return c.N; if (k == "rho0") return c.rho0; if (k == "g") return c.g;
  if (k == "R0") return c.R0; if (k == "T0") return c.T0; if (k == "S0") return c.S0; if (k == "Tcoef") return c.Tcoef; if (k == "Scoef") return c.Scoef;
  if (k == "theta_s") return c.theta_s; if (k == "theta_b") return c.theta_b; if (k == "Tcline") return c.Tcline; if (k == "lambda") return c.lambda;
  if (k == "nfast") return m.nfast; if (k == "dtfast") return m.dtfast; if (k == "hc") return m.hc; if (k == "wvelocity_every_step") return c.wvelocity_every_step;
  if (k == "bv_frequency") return c.bv_frequency; if (k == "eos_tderivative") return c.eos_tderivative;
  if (k == "solar_source") return c.solar_source; if (k == "lmd_nonlocal") return c.lmd_nonlocal;
  if (k == "bvf_mixing") return c.bvf_mixing; if (k == "itemp") return c.itemp; if (k == "isalt") return c.isalt;
  if (k == "qcorrection") return c.qcorrection; if (k == "limit_stflx_cooling") return c.limit_stflx_cooling;
  if (k == "scorrection") return c.scorrection; if (k == "Tnudg_salt") return c.Tnudg_salt;
  if (k == "Vtransform") return c.Vtransform; if (k == "atm_press") return c.atm_press;
  if (k == "bodyforce") return c.bodyforce; if (k == "levsfrc") return c.levsfrc; if (k == "levbfrc") return c.levbfrc;
  if (k == "nospl_vvisc") return c.nospl_vvisc; if (k == "nospl_vdiff") return c.nospl_vdiff;
  if (k == "limit_bstress") return c.limit_bstress; if (k == "rdrg2") return c.rdrg2; if (k == "rdrg") return c.rdrg;
  if (k == "uv_adv") return c.uv_adv; if (k == "ts_dif4") return c.ts_dif4; if (k == "tnu4") return c.tnu4[0];
  if (k == "bulk_fluxes") return c.bulk_fluxes; if (k == "lmd_mixing") return c.lmd_mixing;
  if (k == "blk_ZQ") return c.blk_ZQ; if (k == "blk_ZT") return c.blk_ZT; if (k == "blk_ZW") return c.blk_ZW;
  if (k == "app") return c.app; if (k == "nAVG") return c.nAVG; if (k == "ntsAVG") return c.ntsAVG;
  return -1.0e300;
}

void orc_init(void* hp) { Handle* h = (Handle*)hp; h->m.allocate(); initialize(h->m); h->ready = true; }
void orc_step(void* hp, int nsteps, int nthreads) { Model& m = ((Handle*)hp)->m; for (int s = 0; s < nsteps; ++s) main3d_step(m, nthreads); }
void orc_run_phase(void* hp, int phase, int nthreads) { run_phase(((Handle*)hp)->m, phase, nthreads); }

// time-index state machine: idx = {iic, ntstart, ntfirst, nstp, nnew, nrhs, iif, indx1, kstp, krhs, knew, PREDICTOR, exit_flag}
void orc_get_indices(void* hp, int* idx, double* tm) {
  Model& m = ((Handle*)hp)->m;
  int v[13] = {m.iic, m.ntstart, m.ntfirst, m.nstp, m.nnew, m.nrhs, m.iif, m.indx1, m.kstp, m.krhs, m.knew, m.PREDICTOR_2D_STEP ? 1 : 0, m.exit_flag};
  std::memcpy(idx, v, sizeof(v)); tm[0] = m.time; tm[1] = m.tdays;
}
void orc_set_indices(void* hp, const int* v, const double* tm) {
  Model& m = ((Handle*)hp)->m;
  m.iic = v[0]; m.ntstart = v[1]; m.ntfirst = v[2]; m.nstp = v[3]; m.nnew = v[4]; m.nrhs = v[5]; m.iif = v[6]; m.indx1 = v[7];
  m.kstp = v[8]; m.krhs = v[9]; m.knew = v[10]; m.PREDICTOR_2D_STEP = v[11] != 0; m.exit_flag = v[12]; m.time = tm[0]; m.tdays = tm[1];
}

// zero-copy access: returns 0 on success; dims = {LBi, ni, LBj, nj, LBk, nk}
int orc_field(void* hp, const char* name, double** ptr, int* dims) { return lookup(((Handle*)hp)->m, name, ptr, dims) ? 0 : 1; }
// 1-D: which = 0 sc_r, 1 Cs_r, 2 sc_w, 3 Cs_w (N+1 values each, index k), 4 weight(1,:), 5 weight(2,:) (2*ndtfast+2 values)
int orc_vector(void* hp, int which, double* out, int n) {
  Model& m = ((Handle*)hp)->m; const std::vector<double>* v = nullptr;
  switch (which) { case 0: v = &m.sc_r; break; case 1: v = &m.Cs_r; break; case 2: v = &m.sc_w; break; case 3: v = &m.Cs_w; break;
                   case 4: v = &m.weight1; break; case 5: v = &m.weight2; break; default: return 1; }
  if ((int)v->size() < n) n = (int)v->size();
  std::memcpy(out, v->data(), n * sizeof(double)); return 0;
}

// out = {avgke, avgpe, avgkp, volume, max_speed, maxCu, maxCv, maxCw, ubarmax, vbarmax, umax, vmax}
void orc_diag(void* hp, double* out) {
  Model& m = ((Handle*)hp)->m; diag(m);
  double v[12] = {m.avgke, m.avgpe, m.avgkp, m.volume, m.max_speed, m.maxCu, m.maxCv, m.maxCw, m.ubarmax, m.vbarmax, m.umax, m.vmax};
  std::memcpy(out, v, sizeof(v));
}

// get_bounds.F restatement: out[0..] = tile, Itile, Jtile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS, JmaxS, Istr, IstrB, IstrP, IstrR, IstrT,
// IstrM, IstrU, Iend, IendB, IendP, IendR, IendT, Jstr, JstrB, JstrP, JstrR, JstrT, JstrM, JstrV, Jend, JendB, JendP, JendR, JendT, Istrm3,
// Istrm2, Istrm1, IstrUm2, IstrUm1, Iendp1, Iendp2, Iendp2i, Iendp3, Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1, Jendp1, Jendp2, Jendp2i,
// Jendp3, W,E,S,N edge flags  (57 ints)
void orc_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, int distribute, int* out) {
  Cfg c; c.Lm = Lm; c.Mm = Mm; c.NtileI = NtileI; c.NtileJ = NtileJ; c.EWperiodic = true; c.NSperiodic = false;
  Bnd b; compute_bounds(c, tile, distribute != 0, b);
  int v[57] = {b.tile, b.Itile, b.Jtile, b.LBi, b.UBi, b.LBj, b.UBj, b.IminS, b.ImaxS, b.JminS, b.JmaxS, b.Istr, b.IstrB, b.IstrP, b.IstrR, b.IstrT,
               b.IstrM, b.IstrU, b.Iend, b.IendB, b.IendP, b.IendR, b.IendT, b.Jstr, b.JstrB, b.JstrP, b.JstrR, b.JstrT, b.JstrM, b.JstrV, b.Jend,
               b.JendB, b.JendP, b.JendR, b.JendT, b.Istrm3, b.Istrm2, b.Istrm1, b.IstrUm2, b.IstrUm1, b.Iendp1, b.Iendp2, b.Iendp2i, b.Iendp3,
               b.Jstrm3, b.Jstrm2, b.Jstrm1, b.JstrVm2, b.JstrVm1, b.Jendp1, b.Jendp2, b.Jendp2i, b.Jendp3,
               b.Western_Edge, b.Eastern_Edge, b.Southern_Edge, b.Northern_Edge};
  std::memcpy(out, v, sizeof(v));
}

void orc_physics_point(int which, const double* in, double* out) { physics_point(which, in, out); }
void orc_eos_point(double Tt, double Ts, double Tp, double* out3) { eos_point(Tt, Ts, Tp, &out3[0], &out3[1], &out3[2]); }

// set_weights stand-alone: returns nfast; chk[5] = FORMAT 40 integrals; w1,w2 sized 2*ndtfast+2
int orc_set_weights(int ndtfast, double dt, double* chk, double* w1, double* w2) {
  Model m; m.c.ndtfast = ndtfast; m.c.dt = dt; set_weights(m, chk);
  if (w1) std::memcpy(w1, m.weight1.data(), m.weight1.size() * sizeof(double));
  if (w2) std::memcpy(w2, m.weight2.data(), m.weight2.size() * sizeof(double));
  return m.nfast;
}

// timed run for the CPU baseline: returns wall seconds for nsteps baroclinic steps
double orc_timed_steps(void* hp, int nsteps, int nthreads) {
  Model& m = ((Handle*)hp)->m;
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < nsteps; ++s) main3d_step(m, nthreads);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

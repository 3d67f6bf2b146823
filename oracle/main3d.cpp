// ORACLE -- TEST INFRASTRUCTURE ONLY (see roms_oracle.hpp).
// ROMS/Nonlinear/main3d.F:189-917 (one baroclinic step), ROMS/Nonlinear/set_data.F (analytical forcing refresh),
// ROMS/Nonlinear/diag.F:207-560 and ROMS/Functionals/ana_diag.h:108-156.  Tiles are spread over host threads the way
// the reference's shared-memory mode does (one parallel loop per phase == "!$OMP BARRIER" after every phase).
#include "roms_oracle.hpp"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <functional>
#include <memory>

namespace orc {

// Persistent worker pool (this image's g++ has no libgomp, so std::thread replaces "!$OMP PARALLEL").
namespace {
class Pool {
 public:
  explicit Pool(int n) : n_(n) { for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); }); }
  ~Pool() { { std::unique_lock<std::mutex> l(mu_); stop_ = true; ++gen_; } cv_.notify_all(); for (auto& t : th_) t.join(); }
  int size() const { return n_; }
  void run(int ntask, const std::function<void(int)>& fn) {
    { std::unique_lock<std::mutex> l(mu_); fn_ = &fn; ntask_ = ntask; next_.store(0); pending_ = n_; ++gen_; }
    cv_.notify_all();
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }
 private:
  void loop(int) {
    unsigned long seen = 0;
    for (;;) {
      { std::unique_lock<std::mutex> l(mu_); cv_.wait(l, [&] { return gen_ != seen; }); seen = gen_; if (stop_) return; }
      for (;;) { int t = next_.fetch_add(1); if (t >= ntask_) break; (*fn_)(t); }
      { std::unique_lock<std::mutex> l(mu_); if (--pending_ == 0) done_.notify_all(); }
    }
  }
  int n_; std::vector<std::thread> th_; std::mutex mu_; std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr; int ntask_ = 0; std::atomic<int> next_{0}; int pending_ = 0; unsigned long gen_ = 0; bool stop_ = false;
};
std::unique_ptr<Pool> g_pool;
}  // namespace

template <class Fn>
static void for_tiles(Model& m, int nthreads, Fn fn) {
  const int nt = (int)m.tiles.size();
  if (nthreads <= 1 || nt == 1) { for (int t_ = 0; t_ < nt; ++t_) fn(m.tiles[t_]); return; }
  if (!g_pool || g_pool->size() != nthreads) g_pool.reset(new Pool(nthreads));
  std::function<void(int)> f = [&](int t_) { fn(m.tiles[t_]); };
  g_pool->run(nt, f);
}

// diag.F:207-437 (all tiles, then the global reduction) + blow-up flags :506-538 (limits mod_scalars.F:548-549)
void diag(Model& m) {
  const Cfg& c = m.c; const int N = c.N; const int idia = m.nstp;
  F3 u = m.u[idia], v = m.v[idia];
  double volume = 0, avgke = 0, avgpe = 0, maxspeed = 0, maxrho = -1.0e37, max_C = 0, max_Cu = 0, max_Cv = 0, max_Cw = 0;
  for (const Bnd& b : m.tiles) {
    ORC_UNPACK_BOUNDS(b);
    S2 ke2d(IminS, ImaxS, JminS, JmaxS), pe2d(IminS, ImaxS, JminS, JmaxS);
    double my_max_C = 0, my_max_Cu = 0, my_max_Cv = 0, my_max_Cw = 0, my_maxspeed = 0, my_maxrho = -1.0e37;
    for (int j = Jstr; j <= Jend; ++j) {
      for (int i = Istr; i <= Iend; ++i) { ke2d(i, j) = 0.0; pe2d(i, j) = 0.5 * c.g * m.z_w(i, j, N) * m.z_w(i, j, N); }
      double cff = c.g / c.rho0;
      for (int k = N; k >= 1; --k)
        for (int i = Istr; i <= Iend; ++i) {
          double u2v2 = u(i, j, k) * u(i, j, k) + u(i + 1, j, k) * u(i + 1, j, k) + v(i, j, k) * v(i, j, k) + v(i, j + 1, k) * v(i, j + 1, k);
          ke2d(i, j) = ke2d(i, j) + m.Hz(i, j, k) * 0.25 * u2v2;
          pe2d(i, j) = pe2d(i, j) + cff * m.Hz(i, j, k) * (m.rho(i, j, k) + 1000.0) * (m.z_r(i, j, k) - m.z_w(i, j, 0));
          double my_Cu = 0.5 * std::fabs(u(i, j, k) + u(i + 1, j, k)) * c.dt * m.pm(i, j);
          double my_Cv = 0.5 * std::fabs(v(i, j, k) + v(i, j + 1, k)) * c.dt * m.pn(i, j);
          double my_Cw = 0.5 * std::fabs(m.wvel(i, j, k - 1) + m.wvel(i, j, k)) * c.dt / m.Hz(i, j, k);
          double my_C = my_Cu + my_Cv + my_Cw;
          if (my_C > my_max_C) { my_max_C = my_C; my_max_Cu = my_Cu; my_max_Cv = my_Cv; my_max_Cw = my_Cw; }
          my_maxspeed = std::max(my_maxspeed, std::sqrt(0.5 * u2v2));
          my_maxrho = std::max(my_maxrho, m.rho(i, j, k));
        }
    }
    for (int i = Istr; i <= Iend; ++i) { pe2d(i, Jend + 1) = 0.0; pe2d(i, Jstr - 1) = 0.0; ke2d(i, Jstr - 1) = 0.0; }
    for (int j = Jstr; j <= Jend; ++j)
      for (int i = Istr; i <= Iend; ++i) {
        pe2d(i, Jend + 1) = pe2d(i, Jend + 1) + m.omn(i, j) * (m.z_w(i, j, N) - m.z_w(i, j, 0));
        pe2d(i, Jstr - 1) = pe2d(i, Jstr - 1) + m.omn(i, j) * pe2d(i, j);
        ke2d(i, Jstr - 1) = ke2d(i, Jstr - 1) + m.omn(i, j) * ke2d(i, j);
      }
    double my_volume = 0, my_avgpe = 0, my_avgke = 0;
    for (int i = Istr; i <= Iend; ++i) { my_volume += pe2d(i, Jend + 1); my_avgpe += pe2d(i, Jstr - 1); my_avgke += ke2d(i, Jstr - 1); }
    volume += my_volume; avgke += my_avgke; avgpe += my_avgpe;
    maxspeed = std::max(maxspeed, my_maxspeed); maxrho = std::max(maxrho, my_maxrho);
    if (my_max_C > max_C) { max_C = my_max_C; max_Cu = my_max_Cu; max_Cv = my_max_Cv; max_Cw = my_max_Cw; }
  }
  m.volume = volume; m.avgke = avgke / volume; m.avgpe = avgpe / volume; m.avgkp = m.avgke + m.avgpe;
  m.max_speed = maxspeed; m.maxCu = max_Cu; m.maxCv = max_Cv; m.maxCw = max_Cw;
  const double max_speed_lim = 20.0, max_rho_lim = 200.0;           // mod_scalars.F:548-549
  if (!(m.avgke == m.avgke) || !(m.avgpe == m.avgpe) || std::isinf(m.avgke) || std::isinf(m.avgpe)) m.exit_flag = 1;
  if (maxspeed > max_speed_lim) m.exit_flag = 1;
  if (maxrho > max_rho_lim) m.exit_flag = 1;
  // ana_diag.h:116-142 (SEAMOUNT; harmless elsewhere): signed maxima over the global index range
  {
    const int Lm = c.Lm, Mm = c.Mm; F3 un = m.u[m.nnew], vn = m.v[m.nnew];
    double umax = 0, vmax = 0, ubarmax = 0, vbarmax = 0;
    for (int k = 1; k <= N; ++k) {
      for (int j = 0; j <= Mm + 1; ++j) for (int i = 1; i <= Lm + 1; ++i) umax = std::max(umax, un(i, j, k));
      for (int j = 1; j <= Mm + 1; ++j) for (int i = 0; i <= Lm + 1; ++i) vmax = std::max(vmax, vn(i, j, k));
    }
    for (int j = 0; j <= Mm + 1; ++j) for (int i = 1; i <= Lm + 1; ++i) ubarmax = std::max(ubarmax, m.ubar[m.knew](i, j));
    for (int j = 1; j <= Mm + 1; ++j) for (int i = 0; i <= Lm + 1; ++i) vbarmax = std::max(vbarmax, m.vbar[m.knew](i, j));
    m.umax = umax; m.vmax = vmax; m.ubarmax = ubarmax; m.vbarmax = vbarmax;
  }
}

// The barotropic sub-cycle: main3d.F:592-700
static void step2d_loop(Model& m, int nthreads) {
  for (int my_iif = 1; my_iif <= m.nfast + 1; ++my_iif) {
    int next_indx1 = 3 - m.indx1;
    if (!m.PREDICTOR_2D_STEP && my_iif <= m.nfast + 1) {
      m.PREDICTOR_2D_STEP = true;
      m.iif = my_iif;
      m.kstp = (m.iif == 1) ? m.indx1 : 3 - m.indx1;
      m.knew = 3;
      m.krhs = m.indx1;
    }
    if (my_iif <= m.nfast + 1) for_tiles(m, nthreads, [&](const Bnd& b) { step2d(m, b); });
    if (m.PREDICTOR_2D_STEP) {
      m.PREDICTOR_2D_STEP = false;
      m.knew = next_indx1;
      m.kstp = 3 - m.knew;
      m.krhs = 3;
      if (m.iif < m.nfast + 1) m.indx1 = next_indx1;
    }
    if (m.iif < m.nfast + 1) for_tiles(m, nthreads, [&](const Bnd& b) { step2d(m, b); });
  }
}

void run_phase(Model& m, int phase, int nthreads) {
  switch (phase) {
    case PH_SET_DATA:
      // set_data.F: with BULK_FLUXES the atmosphere replaces ana_smflux and the heat part of ana_stflux (:406-412, :552-564)
      for_tiles(m, nthreads, [&](const Bnd& b) {
        if (m.c.bulk_fluxes) ana_atmosphere(m, b); else ana_smflux(m, b);
        ana_stflux_btflux(m, b);
      });
      break;
    case PH_BULK_FLUX: if (m.c.bulk_fluxes) for_tiles(m, nthreads, [&](const Bnd& b) { bulk_flux(m, b); }); break;
    case PH_BVF_MIX: if (m.c.bvf_mixing) for_tiles(m, nthreads, [&](const Bnd& b) { bvf_mix(m, b); }); break;
    case PH_LMD_VMIX:
      if (m.c.lmd_mixing) {
        for_tiles(m, nthreads, [&](const Bnd& b) { lmd_vmix(m, b); });
        for_tiles(m, nthreads, [&](const Bnd& b) { lmd_vmix_bc(m, b); });
      }
      break;
    case PH_SET_MASSFLUX: for_tiles(m, nthreads, [&](const Bnd& b) { set_massflux(m, b); }); break;
    case PH_RHO_EOS: for_tiles(m, nthreads, [&](const Bnd& b) { rho_eos(m, b); }); break;
    case PH_DIAG: diag(m); break;
    case PH_SET_VBC: for_tiles(m, nthreads, [&](const Bnd& b) { set_vbc(m, b); }); break;
    case PH_ANA_VMIX: if (m.c.ana_vmix) for_tiles(m, nthreads, [&](const Bnd& b) { ana_vmix(m, b); }); break;
    case PH_OMEGA: case PH_OMEGA2: for_tiles(m, nthreads, [&](const Bnd& b) { omega(m, b); }); break;
    case PH_WVELOCITY: for_tiles(m, nthreads, [&](const Bnd& b) { wvelocity(m, b, m.nstp); }); break;
    case PH_SET_ZETA: for_tiles(m, nthreads, [&](const Bnd& b) { set_zeta(m, b); }); break;
    case PH_SET_AVG: for_tiles(m, nthreads, [&](const Bnd& b) { set_avg(m, b); }); break;
    case PH_PRE_STEP3D: for_tiles(m, nthreads, [&](const Bnd& b) { pre_step3d(m, b); }); break;
    case PH_PRSGRD: for_tiles(m, nthreads, [&](const Bnd& b) { prsgrd(m, b); }); break;
    case PH_T3DMIX: for_tiles(m, nthreads, [&](const Bnd& b) { t3dmix2(m, b); }); break;                              // rhs3d.F:81-88
    case PH_T3DMIX4: if (m.c.ts_dif4) for_tiles(m, nthreads, [&](const Bnd& b) { t3dmix4(m, b); }); break;            // rhs3d.F:89-97
    case PH_RHS3D: for_tiles(m, nthreads, [&](const Bnd& b) { rhs3d(m, b); }); break;
    case PH_UV3DMIX: for_tiles(m, nthreads, [&](const Bnd& b) { uv3dmix2(m, b); }); break;
    case PH_STEP2D: for_tiles(m, nthreads, [&](const Bnd& b) { step2d(m, b); }); break;
    case PH_STEP2D_LOOP: step2d_loop(m, nthreads); break;
    case PH_SET_DEPTH: for_tiles(m, nthreads, [&](const Bnd& b) { set_depth(m, b); }); break;
    case PH_STEP3D_UV: for_tiles(m, nthreads, [&](const Bnd& b) { step3d_uv(m, b); }); break;
    case PH_STEP3D_T: for_tiles(m, nthreads, [&](const Bnd& b) { step3d_t(m, b); }); break;
    case PH_INI:                                                       // main3d.F:189-191, :269-285 (first step only)
      m.nstp = 1 + ((m.iic - m.ntstart) % 2); m.nnew = 3 - m.nstp; m.nrhs = m.nstp;
      for_tiles(m, nthreads, [&](const Bnd& b) { ini_zeta(m, b); set_depth(m, b); });
      for_tiles(m, nthreads, [&](const Bnd& b) { ini_fields(m, b); });
      break;
    default: std::fprintf(stderr, "oracle: unknown phase %d\n", phase); std::abort();
  }
}

// main3d.F:189-917.  rhs3d (the driver, rhs3d.F:74-159) calls pre_step3d, prsgrd, t3dmix2, rhs3d_tile, uv3dmix2 per tile
// in that order; there is no barrier between them, but each only reads off-tile data produced in earlier phases.
void main3d_step(Model& m, int nthreads) {
  m.nstp = 1 + ((m.iic - m.ntstart) % 2); m.nnew = 3 - m.nstp; m.nrhs = m.nstp;      // :189-191
  m.tdays = m.time / 86400.0;                                                        // :196
  run_phase(m, PH_SET_DATA, nthreads);                                               // :222
  if (m.iic == m.ntstart) {                                                          // :269-285
    for_tiles(m, nthreads, [&](const Bnd& b) { ini_zeta(m, b); set_depth(m, b); });
    for_tiles(m, nthreads, [&](const Bnd& b) { ini_fields(m, b); });
  }
  for_tiles(m, nthreads, [&](const Bnd& b) { set_massflux(m, b); rho_eos(m, b); });   // :307-309
  run_phase(m, PH_DIAG, nthreads);                                                   // :314
  run_phase(m, PH_BULK_FLUX, nthreads);                                              // :384-390
  run_phase(m, PH_SET_VBC, nthreads);                                                // :394
  if (!m.c.ana_vmix) { if (m.c.lmd_mixing) run_phase(m, PH_LMD_VMIX, nthreads); else run_phase(m, PH_BVF_MIX, nthreads); }   // :467-469
  for_tiles(m, nthreads, [&](const Bnd& b) {                                         // :465-475
    if (m.c.ana_vmix) ana_vmix(m, b);
    omega(m, b);
    if (m.c.wvelocity_every_step) wvelocity(m, b, m.nstp);
  });
  for_tiles(m, nthreads, [&](const Bnd& b) { set_zeta(m, b); set_avg(m, b); });      // :489-495
  for_tiles(m, nthreads, [&](const Bnd& b) {                                         // :563 -> rhs3d.F:74-159
    pre_step3d(m, b); prsgrd(m, b); t3dmix2(m, b); if (m.c.ts_dif4) t3dmix4(m, b); rhs3d(m, b); uv3dmix2(m, b);
  });
  step2d_loop(m, nthreads);                                                          // :592-700
  run_phase(m, PH_SET_DEPTH, nthreads);                                              // :736
  run_phase(m, PH_STEP3D_UV, nthreads);                                              // :762
  run_phase(m, PH_OMEGA2, nthreads);                                                 // :789
  run_phase(m, PH_STEP3D_T, nthreads);                                               // :814
  m.iic += 1; m.time += m.c.dt;                                                      // :914-915
}

}  // namespace orc

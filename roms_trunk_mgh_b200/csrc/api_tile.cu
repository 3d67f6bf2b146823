// Per-routine host-pointer entry points (form 1 of the boundary): the argument lists mirror the reference's _tile
// routines; each call uploads the whole Fortran arrays it is given, runs the routine's kernels on the device and
// downloads the outputs.  No host pointer is retained.  Built on the resident API only.
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/roms_b200.h"

namespace {

struct Tmp {
  roms_b200_handle h = nullptr;
  size_t n2 = 0;
  int N = 0;
  int rc = 0;
  explicit Tmp(const roms_b200_tile_t* b) {
    if (!b) { rc = 2; return; }
    rc = roms_b200_create(&b->cfg, &h);
    if (rc) return;
    int ab[4];
    roms_b200_array_bounds(h, ab);
    n2 = (size_t)(ab[1] - ab[0] + 1) * (ab[3] - ab[2] + 1);
    N = b->cfg.N;
    int idx[13] = {b->iic, 1, b->ntfirst, b->nstp, b->nnew, b->nrhs, b->iif, 1, b->kstp, b->krhs, b->knew, b->predictor, 0};
    double tm[2] = {0.0, 0.0};
    rc = roms_b200_set_indices(h, idx, tm);
  }
  ~Tmp() { if (h) roms_b200_destroy(h); }
  void up(const std::string& name, const double* p, int nk) { if (!rc && p) rc = roms_b200_set_field(h, name.c_str(), p, n2 * nk); }
  void down(const std::string& name, double* p, int nk) { if (!rc && p) rc = roms_b200_get_field(h, name.c_str(), p, n2 * nk); }
  void run(int phase) { if (!rc) rc = roms_b200_run_phase(h, phase); }
};

}  // namespace

extern "C" {

// rho_eos_tile: ROMS/Nonlinear/rho_eos.F:111 (nonlinear) / :576 (linear)
int roms_b200_rho_eos_tile(const roms_b200_tile_t* b, const double* Hz, const double* z_r, const double* z_w, const double* t,
                           const double* s, double* rhoA, double* rhoS, double* pden, double* rho) {
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  const std::string tl = std::to_string(b->nrhs);
  T.up("Hz", Hz, N); T.up("z_r", z_r, N); T.up("z_w", z_w, N + 1);
  T.up("t" + tl + "_" + std::to_string(b->cfg.itemp - 1), t, N);
  if (s && b->cfg.NT >= 2) T.up("t" + tl + "_" + std::to_string(b->cfg.isalt - 1), s, N);
  T.run(ROMS_B200_RHO_EOS);
  T.down("rhoA", rhoA, 1); T.down("rhoS", rhoS, 1); T.down("pden", pden, N); T.down("rho", rho, N);
  return T.rc;
}

// prsgrd32_tile / prsgrd31_tile: ROMS/Nonlinear/prsgrd32.h:106, prsgrd31.h:97
int roms_b200_prsgrd_tile(const roms_b200_tile_t* b, const double* Hz, const double* om_v, const double* on_u, const double* z_r,
                          const double* z_w, const double* rho, double* ru, double* rv) {
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  const std::string tl = std::to_string(b->nrhs);
  T.up("Hz", Hz, N); T.up("om_v", om_v, 1); T.up("on_u", on_u, 1); T.up("z_r", z_r, N); T.up("z_w", z_w, N + 1); T.up("rho", rho, N);
  T.up("ru" + tl, ru, N + 1); T.up("rv" + tl, rv, N + 1);       // level 0 planes and untouched points are preserved
  T.run(ROMS_B200_PRSGRD);
  T.down("ru" + tl, ru, N + 1); T.down("rv" + tl, rv, N + 1);
  return T.rc;
}

// set_massflux_tile: ROMS/Nonlinear/set_massflux.F:73
int roms_b200_set_massflux_tile(const roms_b200_tile_t* b, const double* u, const double* v, const double* Hz, const double* om_v,
                                const double* on_u, double* Huon, double* Hvom) {
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  const std::string tl = std::to_string(b->nrhs);
  T.up("u" + tl, u, N); T.up("v" + tl, v, N); T.up("Hz", Hz, N); T.up("om_v", om_v, 1); T.up("on_u", on_u, 1);
  T.up("Huon", Huon, N); T.up("Hvom", Hvom, N);
  T.run(ROMS_B200_SET_MASSFLUX);
  T.down("Huon", Huon, N); T.down("Hvom", Hvom, N);
  return T.rc;
}

// omega_tile: ROMS/Nonlinear/omega.F:73
int roms_b200_omega_tile(const roms_b200_tile_t* b, const double* Huon, const double* Hvom, const double* z_w, double* W) {
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  T.up("Huon", Huon, N); T.up("Hvom", Hvom, N); T.up("z_w", z_w, N + 1); T.up("W", W, N + 1);
  T.run(ROMS_B200_OMEGA);
  T.down("W", W, N + 1);
  return T.rc;
}

// set_depth_tile: ROMS/Nonlinear/set_depth.F:82 (Vtransform = 2)
int roms_b200_set_depth_tile(const roms_b200_tile_t* b, const double* h, const double* Zt_avg1, const double* sc_r, const double* Cs_r,
                             const double* sc_w, const double* Cs_w, double* Hz, double* z_r, double* z_w) {
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  T.up("h", h, 1); T.up("Zt_avg1", Zt_avg1, 1);
  if (!T.rc) T.rc = roms_b200_set_scoord(T.h, 0, sc_r, N + 1);
  if (!T.rc) T.rc = roms_b200_set_scoord(T.h, 1, Cs_r, N + 1);
  if (!T.rc) T.rc = roms_b200_set_scoord(T.h, 2, sc_w, N + 1);
  if (!T.rc) T.rc = roms_b200_set_scoord(T.h, 3, Cs_w, N + 1);
  T.up("Hz", Hz, N); T.up("z_r", z_r, N); T.up("z_w", z_w, N + 1);
  T.run(ROMS_B200_SET_DEPTH);
  T.down("Hz", Hz, N); T.down("z_r", z_r, N); T.down("z_w", z_w, N + 1);
  return T.rc;
}

}  // extern "C"

// Per-routine host-pointer entry points (form 1 of the boundary): the argument lists mirror the reference's _tile
// routines; each call uploads the whole Fortran arrays it is given, runs the routine's kernels on the device and
// downloads the outputs.  No host pointer is retained.  Built on the resident API only.
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/roms_b200.h"

extern "C" int roms_b200_create_lazy_(const roms_b200_config* cfg, roms_b200_handle* out);   // api.cu

namespace {

struct Tmp {
  roms_b200_handle h = nullptr;
  size_t n2 = 0;
  int N = 0;
  int rc = 0;
  explicit Tmp(const roms_b200_tile_t* b) {
    if (!b) { rc = 2; return; }
    // a _tile call computes the whole xi-column of tiles; with NtileJ > 1 a host would call it once per Jtile and the
    // read-modify-write routines would be applied NtileJ times: the per-routine form needs NtileJ == 1
    if (b->cfg.NtileJ != 1) { rc = 5; return; }
    rc = roms_b200_create_lazy_(&b->cfg, &h);          // only the arrays handed over get device storage
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
  T.up("rhoA", rhoA, 1); T.up("rhoS", rhoS, 1); T.up("pden", pden, N); T.up("rho", rho, N);     // untouched elements are preserved
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

// ---- generic form: any routine of the chain -----------------------------------------------------------------------
// The dummy-argument lists of the remaining _tile routines run to 40-70 whole arrays (step2d_LF_AM3.h:137-215,
// rhs3d.F:174-252, step3d_uv.F:111-170, ...); they are passed by name instead of by position.  The names are the
// reference's own (OCEAN/GRID/COUPLING/MIXING/FORCES members), with the time level spelled out as roms_b200_set_field does.
namespace {
struct RoutineArgs { int phase; const char* in; const char* out; };
// 3-D time levels: both levels of u, v, ru, rv and all three of t are listed (which of them a call reads is decided by
// nstp/nnew/nrhs in roms_b200_tile_t); `*` = one entry per tracer; a leading `?` marks an argument that only the optional terms
// of the routine touch (BV_FREQUENCY / expansion coefficients in rho_eos, SOLAR_SOURCE / LMD_NONLOCAL in pre_step3d): pass it
// when the configuration has the array (roms_b200_config switches), leave it out otherwise.
const RoutineArgs kRoutineArgs[] = {
    {ROMS_B200_SET_MASSFLUX, "u1,u2,v1,v2,Hz,on_u,om_v", "Huon,Hvom"},
    {ROMS_B200_RHO_EOS, "t1_*,t2_*,z_r,z_w,Hz", "rho,pden,rhoA,rhoS,?bvf,?alpha,?beta"},
    {ROMS_B200_SET_VBC, "u1,u2,v1,v2,t1_*,t2_*,rdrag,rdrag2,stflux_*,btflux_*,?ZoBot,?z_r,?z_w,?Hz,?sst,?dqdt,?sss", "bustr,bvstr,stflx_*,btflx_*"},
    {ROMS_B200_ANA_VMIX, "z_w", "Akv,Akt_*"},
    {ROMS_B200_OMEGA, "Huon,Hvom,z_w", "W"},
    {ROMS_B200_WVELOCITY, "u1,u2,v1,v2,z_r,z_w,W,DU_avg1,DV_avg1,pm,pn", "wvel"},
    {ROMS_B200_SET_ZETA, "Zt_avg1", "zeta1,zeta2"},
    {ROMS_B200_PRE_STEP3D,
     "Hz,Huon,Hvom,W,z_r,Akv,Akt_*,pm,pn,stflx_*,btflx_*,sustr,svstr,bustr,bvstr,ru1,ru2,rv1,rv2,u1,u2,v1,v2,t1_*,t2_*,"
     "?z_w,?srflx,?Jwtype,?ghats_*",
     "t3_*,t1_*,t2_*,u1,u2,v1,v2"},
    {ROMS_B200_PRSGRD, "Hz,z_r,z_w,rho,on_u,om_v,?Pair", "ru1,ru2,rv1,rv2"},
    {ROMS_B200_T3DMIX, "Hz,z_r,pm,pn,on_u,om_v,pmon_u,pnom_v,diff2_*,t1_*,t2_*", "t1_*,t2_*"},
    {ROMS_B200_T3DMIX4, "Hz,pm,pn,pmon_u,pnom_v,diff4_*,t1_*,t2_*", "t1_*,t2_*"},
    {ROMS_B200_RHS3D, "Hz,Huon,Hvom,W,u1,u2,v1,v2,fomn,dndx,dmde,om_u,on_u,om_v,on_v,sustr,svstr,bustr,bvstr,ru1,ru2,rv1,rv2",
     "ru1,ru2,rv1,rv2,rufrc,rvfrc"},
    {ROMS_B200_UV3DMIX,
     "Hz,u1,u2,v1,v2,pm,pn,om_r,on_r,om_p,on_p,pmon_r,pnom_r,pmon_p,pnom_p,visc2_r,visc2_p,rufrc,rvfrc", "u1,u2,v1,v2,rufrc,rvfrc"},
    {ROMS_B200_STEP2D,
     "h,pm,pn,on_u,om_v,fomn,dndx,dmde,om_r,on_r,om_p,on_p,pmon_r,pnom_r,pmon_p,pnom_p,visc2_r,visc2_p,rhoA,rhoS,rufrc,rvfrc,"
     "zeta1,zeta2,zeta3,ubar1,ubar2,ubar3,vbar1,vbar2,vbar3,rzeta1,rzeta2,rubar1,rubar2,rvbar1,rvbar2,Zt_avg1,DU_avg1,DU_avg2,"
     "DV_avg1,DV_avg2,ru1,ru2,rv1,rv2",
     "zeta1,zeta2,zeta3,ubar1,ubar2,ubar3,vbar1,vbar2,vbar3,rzeta1,rzeta2,rubar1,rubar2,rvbar1,rvbar2,Zt_avg1,DU_avg1,DU_avg2,"
     "DV_avg1,DV_avg2,rufrc,rvfrc,ru1,ru2,rv1,rv2"},
    {ROMS_B200_SET_DEPTH, "h,Zt_avg1", "z_r,z_w,Hz"},
    {ROMS_B200_STEP3D_UV, "Akv,Hz,ru1,ru2,rv1,rv2,DU_avg1,DU_avg2,DV_avg1,DV_avg2,pm,pn,on_u,om_v,u1,u2,v1,v2,Huon,Hvom,?z_r",
     "u1,u2,v1,v2,Huon,Hvom,ubar1,ubar2,vbar1,vbar2"},
    {ROMS_B200_OMEGA2, "Huon,Hvom,z_w", "W"},
    {ROMS_B200_STEP3D_T, "Hz,Huon,Hvom,W,Akt_*,pm,pn,t1_*,t2_*,t3_*,?z_r", "t1_*,t2_*"},
    // cfg.bulk_fluxes / cfg.lmd_mixing (with the switches lmd_mixing needs) must be set in roms_b200_tile_t.cfg for these two
    {ROMS_B200_BULK_FLUX, "t1_*,t2_*,Uwind,Vwind,Tair,Pair,Hair,rain,cloud,srflx", "lrflx,lhflx,shflx,stflux_*,sustr,svstr"},
    {ROMS_B200_LMD_VMIX, "f,Hz,z_w,u1,u2,v1,v2,pden,bvf,alpha,beta,srflx,Jwtype,stflx_*,sustr,svstr,bustr,bvstr,hsbl,Akv,Akt_*",
     "Akv,Akt_*,ghats_*,hsbl,ksbl"},
    {ROMS_B200_BVF_MIX, "bvf", "Akv,Akt_*"},
};
}  // namespace

/* "in:<names>;out:<names>" of a routine (see kRoutineArgs), or NULL for an unknown phase. */
const char* roms_b200_routine_args(int phase) {
  // built once (thread-safe static initialisation): the returned pointers stay valid for the life of the library
  static const std::vector<std::string> table = [] {
    std::vector<std::string> t(32);
    for (const RoutineArgs& r : kRoutineArgs) t[r.phase & 31] = std::string("in:") + r.in + ";out:" + r.out;
    return t;
  }();
  if (phase < 0 || phase >= 32 || table[phase].empty()) return nullptr;
  return table[phase].c_str();
}

/* Run ONE routine on whole Fortran arrays in host memory.  mode[i]: 1 = input, 2 = output (its current content is uploaded
 * first, so elements the routine does not touch are preserved, as with an INTENT(inout) dummy), 3 = both.  scoord4 =
 * sc_r, Cs_r, sc_w, Cs_w with N+1 entries each (set_depth; NULL otherwise); weight1/2 = set_weights.F products with
 * nweight entries each (step2d; NULL otherwise). */
int roms_b200_routine_tile(const roms_b200_tile_t* b, int phase, int nargs, const char* const* names, double* const* arrays,
                           const int* mode, const double* scoord4, int nfast, const double* weight1, const double* weight2, int nweight) {
  if (!b || nargs < 0 || (nargs > 0 && (!names || !arrays || !mode))) return 2;
  if (!roms_b200_routine_args(phase)) return 5;
  Tmp T(b);
  if (T.rc) return T.rc;
  const int N = T.N;
  std::vector<int> nk(nargs, 0);
  for (int a = 0; a < nargs && !T.rc; ++a) {
    if (!names[a] || !arrays[a] || mode[a] < 1 || mode[a] > 3) return 2;
    T.rc = roms_b200_field_levels(T.h, names[a], nullptr, &nk[a]);
    if (!T.rc) T.up(names[a], arrays[a], nk[a]);
  }
  if (scoord4) for (int w = 0; w < 4 && !T.rc; ++w) T.rc = roms_b200_set_scoord(T.h, w, scoord4 + (size_t)w * (N + 1), N + 1);
  if (weight1 && weight2 && !T.rc) T.rc = roms_b200_set_weights(T.h, nfast, weight1, weight2, nweight);
  T.run(phase);
  for (int a = 0; a < nargs && !T.rc; ++a)
    if (mode[a] & 2) T.down(names[a], arrays[a], nk[a]);
  return T.rc;
}

}  // extern "C"

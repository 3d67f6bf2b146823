// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (roms_trunk_mgh_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// CPU restatement (C++) of the ROMS nonlinear baroclinic time step (ROMS/Nonlinear/main3d.F chain).
// Parity status: the reference is Fortran and cannot be compiled in this environment (no Fortran
// compiler, no NetCDF).  The restatement is pinned only by the reference's own known answers:
// EOS check values (ROMS/Nonlinear/rho_eos.F:21-29), set_weights integrals (ROMS/Utility/set_weights.F
// FORMAT 40), nfast, tiling invariance (ROMS/Bin/verify.sh:985-1045) and physical invariants.
// Beyond those: PARITY UNPINNED.
//
// Conventions: arrays keep Fortran index ranges (i fastest, then j, then k); A(i,j,k) accessors take
// Fortran indices directly.  Every routine cites the reference file:line it follows.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>

namespace orc {

#ifdef ORC_BOUNDS_CHECK
#define ORC_CHK(cond, what) do { if (!(cond)) { std::fprintf(stderr, "oracle bounds violation: %s (%s:%d)\n", what, __FILE__, __LINE__); std::abort(); } } while (0)
#else
#define ORC_CHK(cond, what) ((void)0)
#endif

// 2-D view A(LBi:UBi, LBj:UBj)
struct F2 {
  double* p = nullptr; int LBi = 0, UBi = -1, LBj = 0, UBj = -1; int ni = 0;
  F2() {}
  F2(double* p_, int LBi_, int UBi_, int LBj_, int UBj_) : p(p_), LBi(LBi_), UBi(UBi_), LBj(LBj_), UBj(UBj_), ni(UBi_ - LBi_ + 1) {}
  inline double& operator()(int i, int j) const {
    ORC_CHK(i >= LBi && i <= UBi && j >= LBj && j <= UBj, "F2");
    return p[(i - LBi) + (size_t)ni * (j - LBj)];
  }
  size_t size() const { return (size_t)ni * (UBj - LBj + 1); }
};
// 3-D view A(LBi:UBi, LBj:UBj, LBk:UBk)
struct F3 {
  double* p = nullptr; int LBi = 0, UBi = -1, LBj = 0, UBj = -1, LBk = 1, UBk = 0; int ni = 0; size_t nij = 0;
  F3() {}
  F3(double* p_, int LBi_, int UBi_, int LBj_, int UBj_, int LBk_, int UBk_)
      : p(p_), LBi(LBi_), UBi(UBi_), LBj(LBj_), UBj(UBj_), LBk(LBk_), UBk(UBk_), ni(UBi_ - LBi_ + 1),
        nij((size_t)(UBi_ - LBi_ + 1) * (UBj_ - LBj_ + 1)) {}
  inline double& operator()(int i, int j, int k) const {
    ORC_CHK(i >= LBi && i <= UBi && j >= LBj && j <= UBj && k >= LBk && k <= UBk, "F3");
    return p[(i - LBi) + (size_t)ni * (j - LBj) + nij * (k - LBk)];
  }
  F2 plane(int k) const { return F2(p + nij * (k - LBk), LBi, UBi, LBj, UBj); }
  size_t size() const { return nij * (UBk - LBk + 1); }
};

// private (per-tile) scratch: owns storage
struct S2 {  // (IminS:ImaxS, JminS:JmaxS)
  std::vector<double> d; int i0, i1, j0, j1, ni;
  S2(int i0_, int i1_, int j0_, int j1_) : d((size_t)(i1_ - i0_ + 1) * (j1_ - j0_ + 1), 0.0), i0(i0_), i1(i1_), j0(j0_), j1(j1_), ni(i1_ - i0_ + 1) {}
  inline double& operator()(int i, int j) {
    ORC_CHK(i >= i0 && i <= i1 && j >= j0 && j <= j1, "S2");
    return d[(i - i0) + (size_t)ni * (j - j0)];
  }
};
struct S3 {  // (IminS:ImaxS, JminS:JmaxS, k0:k1)
  std::vector<double> d; int i0, i1, j0, j1, k0, k1, ni; size_t nij;
  S3(int i0_, int i1_, int j0_, int j1_, int k0_, int k1_)
      : d((size_t)(i1_ - i0_ + 1) * (j1_ - j0_ + 1) * (k1_ - k0_ + 1), 0.0), i0(i0_), i1(i1_), j0(j0_), j1(j1_), k0(k0_), k1(k1_),
        ni(i1_ - i0_ + 1), nij((size_t)(i1_ - i0_ + 1) * (j1_ - j0_ + 1)) {}
  inline double& operator()(int i, int j, int k) {
    ORC_CHK(i >= i0 && i <= i1 && j >= j0 && j <= j1 && k >= k0 && k <= k1, "S3");
    return d[(i - i0) + (size_t)ni * (j - j0) + nij * (k - k0)];
  }
};
typedef S2 SK;  // (IminS:ImaxS, 0:N) slabs use S2 with j-range = k-range

// Tile index sets: ROMS/Modules/mod_param.F:162-230 (T_BOUNDS), ROMS/Include/set_bounds.h, tile.h
struct Bnd {
  int tile, Itile, Jtile;
  int LBi, UBi, LBj, UBj;
  int IminS, ImaxS, JminS, JmaxS;
  int Istr, IstrB, IstrP, IstrR, IstrT, IstrM, IstrU;
  int Iend, IendB, IendP, IendR, IendT;
  int Jstr, JstrB, JstrP, JstrR, JstrT, JstrM, JstrV;
  int Jend, JendB, JendP, JendR, JendT;
  int Istrm3, Istrm2, Istrm1, IstrUm2, IstrUm1;
  int Iendp1, Iendp2, Iendp2i, Iendp3;
  int Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1;
  int Jendp1, Jendp2, Jendp2i, Jendp3;
  bool Western_Edge, Eastern_Edge, Southern_Edge, Northern_Edge;
  bool SouthWest_Corner, SouthEast_Corner, NorthWest_Corner, NorthEast_Corner;
  bool SouthWest_Test, SouthEast_Test, NorthWest_Test, NorthEast_Test;
};

#define ORC_UNPACK_BOUNDS(b)                                                                          \
  const int Istr = (b).Istr, IstrB = (b).IstrB, IstrP = (b).IstrP, IstrR = (b).IstrR, IstrT = (b).IstrT,     \
            IstrM = (b).IstrM, IstrU = (b).IstrU, Iend = (b).Iend, IendB = (b).IendB, IendP = (b).IendP,     \
            IendR = (b).IendR, IendT = (b).IendT, Jstr = (b).Jstr, JstrB = (b).JstrB, JstrP = (b).JstrP,     \
            JstrR = (b).JstrR, JstrT = (b).JstrT, JstrM = (b).JstrM, JstrV = (b).JstrV, Jend = (b).Jend,     \
            JendB = (b).JendB, JendP = (b).JendP, JendR = (b).JendR, JendT = (b).JendT,                      \
            Istrm3 = (b).Istrm3, Istrm2 = (b).Istrm2, Istrm1 = (b).Istrm1, IstrUm2 = (b).IstrUm2,             \
            IstrUm1 = (b).IstrUm1, Iendp1 = (b).Iendp1, Iendp2 = (b).Iendp2, Iendp2i = (b).Iendp2i,           \
            Iendp3 = (b).Iendp3, Jstrm3 = (b).Jstrm3, Jstrm2 = (b).Jstrm2, Jstrm1 = (b).Jstrm1,               \
            JstrVm2 = (b).JstrVm2, JstrVm1 = (b).JstrVm1, Jendp1 = (b).Jendp1, Jendp2 = (b).Jendp2,           \
            Jendp2i = (b).Jendp2i, Jendp3 = (b).Jendp3, IminS = (b).IminS, ImaxS = (b).ImaxS,                 \
            JminS = (b).JminS, JmaxS = (b).JmaxS;                                                            \
  (void)Istr; (void)IstrB; (void)IstrP; (void)IstrR; (void)IstrT; (void)IstrM; (void)IstrU; (void)Iend;      \
  (void)IendB; (void)IendP; (void)IendR; (void)IendT; (void)Jstr; (void)JstrB; (void)JstrP; (void)JstrR;     \
  (void)JstrT; (void)JstrM; (void)JstrV; (void)Jend; (void)JendB; (void)JendP; (void)JendR; (void)JendT;     \
  (void)Istrm3; (void)Istrm2; (void)Istrm1; (void)IstrUm2; (void)IstrUm1; (void)Iendp1; (void)Iendp2;        \
  (void)Iendp2i; (void)Iendp3; (void)Jstrm3; (void)Jstrm2; (void)Jstrm1; (void)JstrVm2; (void)JstrVm1;       \
  (void)Jendp1; (void)Jendp2; (void)Jendp2i; (void)Jendp3; (void)IminS; (void)ImaxS; (void)JminS; (void)JmaxS

// Application identifiers (ROMS/Include/{upwelling,seamount,benchmark}.h)
enum App { APP_UPWELLING = 0, APP_SEAMOUNT = 1, APP_BENCHMARK = 2 };
enum HAdv { HADV_U3 = 0, HADV_A4 = 1, HADV_C4 = 2, HADV_C2 = 3 };
enum VAdv { VADV_C4 = 0, VADV_A4 = 1, VADV_C2 = 2, VADV_SPLINES = 3 };

// Run configuration = the live cpp switches + roms_*.in keywords of the three applications.
struct Cfg {
  int app = APP_UPWELLING;
  int Lm = 41, Mm = 80, N = 16, NT = 2;
  int NtileI = 1, NtileJ = 1;
  int Nghost = 2;                 // NghostPoints (ROMS/Utility/inp_par.F:264-281)
  bool EWperiodic = true, NSperiodic = false;
  double dt = 300.0; int ndtfast = 30;
  // cpp switches
  int nonlin_eos = 0;             // NONLIN_EOS
  int dj_gradps = 1;              // DJ_GRADPS -> prsgrd32, else prsgrd31
  int curvgrid = 0;               // CURVGRID
  int spherical = 0;              // SPHERICAL (set-up only)
  int mix_geo_ts = 0;             // MIX_GEO_TS (else MIX_S_TS)
  int uv_qdrag = 0;               // 0 UV_LDRAG, 1 UV_QDRAG, 2 UV_LOGDRAG (set_vbc.F:541-652)
  double Zob = 0.02;              // bottom roughness (m), roms_*.in Zob -> GRID%ZoBot (mod_grid.F:1256)
  int salinity = 1;               // SALINITY
  int ana_vmix = 0;               // ANA_VMIX (UPWELLING profile)
  int wvelocity_every_step = 1;   // main3d.F:475
  int hadv = HADV_U3, vadv = VADV_C4;
  int uv_adv = 0;                 // momentum advection: 0 default (U3 horizontal, C4 vertical), 1 UV_C4ADVECTION (rhs3d.F:685-921, :1108-1175), 2 UV_SADVECTION (:1016-1078, :1267-1329), 3 UV_C2ADVECTION (:605-657, :1079-1107; step2d_LF_AM3.h:1026-1080)
  int nospl_vvisc = 0, nospl_vdiff = 0;   // 1: SPLINES_VVISC / SPLINES_VDIFF UNdefined -> centred implicit vertical viscosity / diffusion
                                          // (step3d_uv.F:397-462, :730-795; step3d_t.F:1430-1499) instead of the parabolic splines
  int qcorrection = 0;            // QCORRECTION: stflx(itemp) += dqdt (SST - sst) (set_vbc.F:285-299)
  int limit_stflx_cooling = 0;    // LIMIT_STFLX_COOLING: no further cooling below -2 degC (:301-328)
  int scorrection = 0;            // 1 SCORRECTION, 2 SRELAXATION (:344-351), with Tnudg(isalt) = Tnudg_salt (1/s)
  double Tnudg_salt = 0.0;
  int bodyforce = 0, levsfrc = 0, levbfrc = 0;   // BODYFORCE: surface / bottom stress as a body force over levels levsfrc:N / 1:levbfrc
                                                 // (rhs3d.F:326-466, :1588-1599; pre_step3d.F:931-937, :1036-1042)
  int atm_press = 0;              // ATM_PRESS: the atmospheric pressure Pair (mb) in the pressure gradient (prsgrd31/32/40)
  int limit_bstress = 0;          // LIMIT_BSTRESS (set_vbc.F:533-540): |bottom stress| <= 0.75 |u| Hz / dt
  int ts_dif4 = 0;                // TS_DIF4 (+ MIX_S_TS): t3dmix4_s.h after t3dmix2 (rhs3d.F:81-97)
  // physical parameters
  double rho0 = 1025.0, g = 9.81;
  double R0 = 1027.0, T0 = 14.0, S0 = 35.0, Tcoef = 1.7e-4, Scoef = 0.0;
  double tnu2[2] = {0.0, 0.0}; double visc2 = 5.0;
  double tnu4[2] = {0.0, 0.0};    // TNU4 (m4/s); diff4 holds SQRT(ABS(tnu4)) (read_phypar.F:6905)
  double Akt_bak[2] = {1e-6, 1e-6}; double Akv_bak = 1e-5;
  double rdrg = 3e-4, rdrg2 = 3e-3;
  double gamma2 = 1.0;
  double theta_s = 3.0, theta_b = 0.0, Tcline = 25.0;
  int Vtransform = 2, Vstretching = 4;
  double lambda = 1.0;            // mod_scalars.F (implicit vertical diffusion weight)
  int itemp = 1, isalt = 2;       // tracer indices (1-based)
  // optional terms INSIDE the routines of the chain that the shipped BENCHMARK cpp set switches on (benchmark.h); their
  // inputs (srflx, ghats from bulk_flux / lmd_skpp) are supplied by the host
  int bv_frequency = 0;           // BV_FREQUENCY: rho_eos also returns bvf (rho_eos.F:402-418 / :751-758)
  int eos_tderivative = 0;        // LMD_SKPP || BULK_FLUXES: rho_eos also returns alpha, beta (:420-462 / :760-773)
  int solar_source = 0;           // SOLAR_SOURCE: shortwave penetration in pre_step3d (:312-333, :866-883), lmd_swfrac.F
  int lmd_nonlocal = 0;           // LMD_NONLOCAL: KPP nonlocal transport in pre_step3d (:850-865)
  // the forcing / mixing physics of the shipped BENCHMARK set (physics.cpp)
  int bulk_fluxes = 0;            // BULK_FLUXES: bulk_flux computes stflux(itemp), sustr, svstr from the atmosphere (main3d.F:384-390)
  int bvf_mixing = 0;             // BVF_MIXING: Akv, Akt from the Brunt-Vaisala frequency (bvf_mix.F; needs bv_frequency; main3d.F:468-469)
  int lmd_mixing = 0;             // LMD_MIXING (+LMD_RIMIX, LMD_CONVEC, LMD_SKPP, LMD_NONLOCAL, RI_SPLINES): lmd_vmix (main3d.F:467)
  double blk_ZQ = 10.0, blk_ZT = 10.0, blk_ZW = 10.0;   // roms_benchmark1.in BLK_ZQ, BLK_ZT, BLK_ZW
  int nAVG = 0, ntsAVG = 1;       // AVERAGES: window length in steps (0: off) and starting step (roms_*.in NAVG, NTSAVG)
};

Cfg make_cfg(int app, int Lm = 0, int Mm = 0, int N = 0);   // defaults from roms_<app>.in (Lm=0 -> shipped sizes)

void compute_bounds(const Cfg& c, int tile, bool distribute, Bnd& b);  // get_bounds.F

// Whole-model state (shared-memory / serial layout: one global array set, tiles are index ranges)
struct Model {
  Cfg c;
  int LBi, UBi, LBj, UBj;
  std::vector<Bnd> tiles;
  // ---- 1-D
  std::vector<double> sc_r, Cs_r, sc_w, Cs_w;   // sc_r/Cs_r index 1..N (slot 0 unused), sc_w/Cs_w 0..N
  double hc = 0;
  std::vector<double> weight1, weight2;         // weight(1,i), weight(2,i), i=1..2*ndtfast (slot 0 unused)
  int nfast = 0; double dtfast = 0;
  // ---- storage
  std::vector<std::vector<double>> pool;
  // ---- 2-D grid (mod_grid.F)
  F2 h, f, pm, pn, om_r, on_r, om_u, on_u, om_v, on_v, om_p, on_p, omn, fomn, pmon_r, pnom_r, pmon_u, pnom_u,
      pmon_v, pnom_v, pmon_p, pnom_p, dndx, dmde, xr, yr, latr, lonr, rdrag, rdrag2, ZoBot;
  F2 visc2_r, visc2_p; F2 diff2[2]; F2 diff4[2]; // mod_mixing.F
  // ---- 2-D state (mod_ocean.F, mod_coupling.F, mod_forces.F)
  F2 zeta[4], ubar[4], vbar[4];                 // [1..3]
  F2 rzeta[3], rubar[3], rvbar[3];              // [1..2]
  F2 Zt_avg1, DU_avg1, DU_avg2, DV_avg1, DV_avg2, rufrc, rvfrc, rhoA, rhoS;
  F2 sustr, svstr, bustr, bvstr; F2 stflx[2], btflx[2], stflux[2], btflux[2];
  // ---- 3-D
  F3 u[3], v[3];                                // [1..2] (k=1..N)
  F3 t[4][2];                                   // [time 1..3][itrc 0..NT-1]
  F3 ru[3], rv[3];                              // [1..2], k=0..N
  F3 rho, pden, Hz, z_r, Huon, Hvom;            // k=1..N
  F3 W, wvel, z_w, Akv; F3 Akt[2];              // k=0..N
  F3 bvf; F2 alpha, beta;                       // rho_eos optional outputs (bvf k=0..N)
  F2 srflx, Jwtype; F3 ghats[2];                // mod_forces.F srflx; mod_mixing.F Jwtype, ghats (k=0..N)
  F2 Uwind, Vwind, Tair, Pair, Hair, rain, cloud, lrflx, lhflx, shflx;   // mod_forces.F (BULK_FLUXES)
  F2 sst, dqdt, sss;                            // mod_forces.F: QCORRECTION / SCORRECTION / SRELAXATION data
  F2 hsbl, ksbl;                                // mod_mixing.F (LMD_SKPP); ksbl is INTEGER in the reference, kept as whole doubles
  // ---- time-averaged fields (mod_average.F; set_avg.cpp)
  F2 avgzeta, avgu2d, avgv2d; F3 avgu3d, avgv3d, avgrho, avgt[2];   // k=1..N
  F3 avgw3d, avgwvel;                                               // k=0..N
  // ---- time stepping (mod_stepping.F:64-72; initial.F:126-170)
  int iic = 0, ntstart = 1, ntfirst = 1, ntend = 0;
  int nstp = 1, nnew = 1, nrhs = 1;
  int iif = 1, indx1 = 1, kstp = 1, krhs = 1, knew = 1; bool PREDICTOR_2D_STEP = false;
  double time = 0.0, tdays = 0.0;
  int exit_flag = 0;
  // diag outputs (diag.F)
  double avgke = 0, avgpe = 0, avgkp = 0, volume = 0, max_speed = 0, maxCu = 0, maxCv = 0, maxCw = 0;
  // SEAMOUNT ana_diag series (ana_diag.h:108-156)
  double ubarmax = 0, vbarmax = 0, umax = 0, vmax = 0;

  F2 new2(); F3 new3(int k0, int k1);
  void allocate();
};

// ---- set-up (CPU, one-off)
void set_scoord(Model& m);                    // ROMS/Utility/set_scoord.F
void set_weights(Model& m, double out_chk[5]); // ROMS/Utility/set_weights.F
void ana_grid(Model& m, const Bnd& b);        // ROMS/Functionals/ana_grid.h
void metrics(Model& m, const Bnd& b);         // ROMS/Utility/metrics.F
void ini_hmixcoef(Model& m, const Bnd& b);    // ROMS/Utility/ini_hmixcoef.F
void ana_initial(Model& m, const Bnd& b);     // ROMS/Functionals/ana_initial.h
void ana_smflux(Model& m, const Bnd& b);      // ROMS/Functionals/ana_smflux.h
void ana_stflux_btflux(Model& m, const Bnd& b);
void ana_vmix(Model& m, const Bnd& b);        // ROMS/Functionals/ana_vmix.h
void ini_zeta(Model& m, const Bnd& b);        // ROMS/Nonlinear/ini_fields.F:836-1137
void ini_fields(Model& m, const Bnd& b);      // ROMS/Nonlinear/ini_fields.F:106-777
void initialize(Model& m);                    // initial.F call order
void set_avg(Model& m, const Bnd& b);         // ROMS/Nonlinear/set_avg.F
void ana_atmosphere(Model& m, const Bnd& b);  // set_data.F:197-394 -> ana_cloud/tair/humid/srflux/winds/rain/pair (BENCHMARK)
void bulk_flux(Model& m, const Bnd& b);       // ROMS/Nonlinear/bulk_flux.F
void bvf_mix(Model& m, const Bnd& b);         // ROMS/Nonlinear/bvf_mix.F
void lmd_vmix(Model& m, const Bnd& b);        // ROMS/Nonlinear/lmd_vmix.F, lmd_skpp.F, lmd_swfrac.F
void physics_point(int which, const double* in, double* out);   // bulk_psiu/psit, lmd_swfrac, KPP velocity scales at a point
void lmd_vmix_bc(Model& m, const Bnd& b);     // ... its closing bc_w3d / exchange (second stage, see physics.cpp)

// ---- periodic exchanges / boundary conditions
void exchange_r2d(const Model& m, const Bnd& b, F2 A);
void exchange_u2d(const Model& m, const Bnd& b, F2 A);
void exchange_v2d(const Model& m, const Bnd& b, F2 A);
void exchange_p2d(const Model& m, const Bnd& b, F2 A);
void exchange_r3d(const Model& m, const Bnd& b, F3 A);
void exchange_u3d(const Model& m, const Bnd& b, F3 A);
void exchange_v3d(const Model& m, const Bnd& b, F3 A);
void exchange_w3d(const Model& m, const Bnd& b, F3 A);
void zetabc(const Model& m, const Bnd& b, int kout);
void u2dbc(const Model& m, const Bnd& b, int kout);
void v2dbc(const Model& m, const Bnd& b, int kout);
void u3dbc(const Model& m, const Bnd& b, int nout);
void v3dbc(const Model& m, const Bnd& b, int nout);
void t3dbc(const Model& m, const Bnd& b, int nout, int itrc);
void bc_r2d(const Model& m, const Bnd& b, F2 A);
void bc_u2d(const Model& m, const Bnd& b, F2 A);
void bc_v2d(const Model& m, const Bnd& b, F2 A);
void bc_r3d(const Model& m, const Bnd& b, F3 A);
void bc_w3d(const Model& m, const Bnd& b, F3 A);

// ---- the chain (one call per tile)
void set_massflux(Model& m, const Bnd& b);
void rho_eos(Model& m, const Bnd& b);
void set_vbc(Model& m, const Bnd& b);
void omega(Model& m, const Bnd& b);
void wvelocity(Model& m, const Bnd& b, int Ninp);
void set_zeta(Model& m, const Bnd& b);
void pre_step3d(Model& m, const Bnd& b);
void prsgrd(Model& m, const Bnd& b);
void t3dmix2(Model& m, const Bnd& b);
void t3dmix4(Model& m, const Bnd& b);         // ROMS/Nonlinear/t3dmix4_s.h (TS_DIF4)
void rhs3d(Model& m, const Bnd& b);       // rhs3d_tile only
void uv3dmix2(Model& m, const Bnd& b);
void step2d(Model& m, const Bnd& b);
void set_depth(Model& m, const Bnd& b);
void step3d_uv(Model& m, const Bnd& b);
void step3d_t(Model& m, const Bnd& b);
void diag(Model& m);                      // all tiles + reduction

// EOS point function used for the rho_eos.F:21-29 check values
void eos_point(double Tt, double Ts, double Tp, double* den, double* den1, double* bulk);

// phase ids for run_phase (shared numbering with the product's C-ABI, include/roms_b200.h)
enum Phase {
  PH_SET_MASSFLUX = 1, PH_RHO_EOS = 2, PH_SET_VBC = 3, PH_ANA_VMIX = 4, PH_OMEGA = 5, PH_WVELOCITY = 6, PH_SET_ZETA = 7,
  PH_PRE_STEP3D = 8, PH_PRSGRD = 9, PH_T3DMIX = 10, PH_RHS3D = 11, PH_UV3DMIX = 12, PH_STEP2D = 13, PH_SET_DEPTH = 14,
  PH_STEP3D_UV = 15, PH_OMEGA2 = 16, PH_STEP3D_T = 17, PH_DIAG = 18, PH_SET_DATA = 19, PH_STEP2D_LOOP = 20, PH_INI = 21, PH_SET_AVG = 22,
  PH_BULK_FLUX = 23, PH_LMD_VMIX = 24, PH_BVF_MIX = 25, PH_T3DMIX4 = 26
};
void run_phase(Model& m, int phase, int nthreads);
void main3d_step(Model& m, int nthreads);    // one baroclinic step (main3d.F:189-917)

}  // namespace orc

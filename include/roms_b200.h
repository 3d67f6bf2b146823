/*
 * roms_b200.h -- C-ABI of the B200-native ROMS nonlinear baroclinic time step (main3d chain).
 *
 * This is the drop-in boundary for ONE path of ROMS: the phases that ROMS/Nonlinear/main3d.F:307-814 runs every
 * baroclinic step.  The reference has no FFI; its seam is the `xxx(ng,tile)` -> `xxx_tile(ng,tile,LBi,UBi,LBj,UBj,
 * IminS,ImaxS,JminS,JmaxS,<time idx>,<arrays...>)` convention (e.g. ROMS/Nonlinear/rhs3d.F:25 -> :174,
 * ROMS/Nonlinear/prsgrd32.h:40 -> :106).  A Fortran host binds these entry points with ISO_C_BINDING
 * (see INTEGRATION.md for the INTERFACE blocks and the one-line CALL replacements).
 *
 * Conventions
 *  - plain C: pointers, ints, doubles.  No C++/torch types.  All reals are IEEE binary64 (ROMS r8, mod_kinds.F).
 *  - every array argument is a WHOLE Fortran array in Fortran (column-major) order, i fastest:
 *      2-D  A(LBi:UBi, LBj:UBj)            n = ni*nj
 *      3-D  A(LBi:UBi, LBj:UBj, 1:N|0:N)   n = ni*nj*nk      (one time level / one tracer per call)
 *    host pointers are never retained after the call returns.
 *  - return value = ROMS exit_flag codes (ROMS/Modules/mod_scalars.F:523-532): 0 NoError, 1 blow-up, 2 input error,
 *    5 configuration error, 8 fatal algorithm error (CUDA / NCCL failure, or a halo exchange whose neighbour never
 *    delivered: sticky, every later synchronising call returns 8 too).  The Fortran shim assigns it to exit_flag.
 *  - the library is re-entrant per handle; one handle == one (ng,tile) == one GPU.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns 8.
 */
#ifndef ROMS_B200_H
#define ROMS_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct roms_b200_state* roms_b200_handle;

/* ROMS_APPLICATION (only selects defaults of the analytical per-step forcing kept on the device: ANA_VMIX). */
enum { ROMS_B200_APP_UPWELLING = 0, ROMS_B200_APP_SEAMOUNT = 1, ROMS_B200_APP_BENCHMARK = 2 };
/* Hadvection / Vadvection keywords of roms_*.in (ROMS/Modules/mod_param.F:382-394, T_ADV) */
enum { ROMS_B200_HADV_U3 = 0, ROMS_B200_HADV_A4 = 1, ROMS_B200_HADV_C4 = 2, ROMS_B200_HADV_C2 = 3 };
enum { ROMS_B200_VADV_C4 = 0, ROMS_B200_VADV_A4 = 1, ROMS_B200_VADV_C2 = 2, ROMS_B200_VADV_SPLINES = 3 };

/* Live cpp switches (ROMS/Include/{upwelling,seamount,benchmark}.h) + roms_*.in keywords for this path. */
typedef struct roms_b200_config {
  int Lm, Mm, N, NT;            /* mod_param.F: interior points, levels, tracers                                  */
  int NtileI, NtileJ, tile;     /* domain partition and this handle's tile (get_bounds.F:933-1007).  A handle owns the whole
                                   xi-column of tiles that contains `tile` (all Jtile): the eta partition only splits loop
                                   ranges, results do not depend on it.  The per-routine _tile forms need NtileJ == 1.   */
  int ndtfast;                  /* NDTFAST; nfast and weights are uploaded with roms_b200_set_weights              */
  double dt;                    /* DT (s)                                                                          */
  int nonlin_eos;               /* NONLIN_EOS (rho_eos.F:111) else linear EOS (rho_eos.F:576)                      */
  int dj_gradps;                /* pressure-gradient algorithm (prsgrd.F:16-26): 1 DJ_GRADPS -> prsgrd32.h, 0 prsgrd31.h, 2 PJ_GRADP ->
                                   prsgrd40.h (finite-volume Jacobian), 3 WJ_GRADP -> prsgrd31.h with the weighted Jacobian (:236-254)  */
  int curvgrid;                 /* CURVGRID terms (rhs3d.F:515-564, step2d_LF_AM3.h:1333-1382)                     */
  int mix_geo_ts;               /* MIX_GEO_TS -> t3dmix2_geo.h, else MIX_S_TS -> t3dmix2_s.h                       */
  int uv_qdrag;                 /* bottom stress: 0 UV_LDRAG (set_vbc.F:629-652), 1 UV_QDRAG (:591-624), 2 UV_LOGDRAG (:541-586,
                                   roughness length "ZoBot" = GRID%ZoBot uploaded by the host)                          */
  int salinity;                 /* SALINITY                                                                        */
  int ana_vmix;                 /* ANA_VMIX, UPWELLING profile (ana_vmix.h:200-208,327-337), refreshed every step  */
  int wvelocity_every_step;     /* main3d.F:475                                                                    */
  int hadv, vadv;               /* tracer advection schemes                                                        */
  double rho0, g;               /* mod_scalars.F                                                                   */
  double R0, T0, S0, Tcoef, Scoef;   /* linear EOS                                                                 */
  double Akt_bak[2], Akv_bak;   /* background mixing (ANA_VMIX)                                                    */
  double gamma2;                /* slipperiness (u2dbc_im.F:973)                                                   */
  double lambda;                /* implicit weight of vertical diffusion (mod_scalars.F), 1.0                      */
  double hc;                    /* s-coordinate critical depth (set_scoord.F:170-178)                              */
  int itemp, isalt;             /* 1-based tracer indices                                                          */
  int device;                   /* CUDA device ordinal                                                             */
  /* Optional terms INSIDE the routines of the chain that the shipped BENCHMARK cpp set (ROMS/Include/benchmark.h) switches
   * on.  Their inputs come from parameterisations that stay on the host (bulk_flux.F, lmd_vmix.F) and are uploaded like any
   * field: "srflx" (FORCES%srflx), "Jwtype" (MIXING%Jwtype), "ghats_<itrc>" (MIXING%ghats, 0:N).                            */
  int bv_frequency;             /* BV_FREQUENCY: rho_eos also returns "bvf" (0:N) (rho_eos.F:402-418 / :751-758)      */
  int eos_tderivative;          /* LMD_SKPP || BULK_FLUXES: rho_eos also returns "alpha", "beta" (:420-462 / :760-773) */
  int solar_source;             /* SOLAR_SOURCE: shortwave penetration in pre_step3d (:312-333, :866-883; lmd_swfrac.F) */
  int lmd_nonlocal;             /* LMD_NONLOCAL: KPP nonlocal transport in pre_step3d (:850-865)                       */
  /* The parameterisations themselves, on the device (main3d.F:384-390, :467).  bulk_fluxes: the host uploads the atmosphere
   * "Uwind","Vwind","Tair","Pair","Hair","rain","cloud","srflx" (FORCES, bulk_flux.F:100-125) and ROMS_B200_BULK_FLUX returns
   * "lrflx","lhflx","shflx","stflux_<itemp-1>","sustr","svstr".  lmd_mixing: ROMS_B200_LMD_VMIX (lmd_vmix.F with LMD_RIMIX,
   * LMD_CONVEC, LMD_SKPP, LMD_NONLOCAL, RI_SPLINES) returns "Akv","Akt_<itrc>","ghats_<itrc>","hsbl","ksbl" (ksbl as whole
   * doubles; the reference holds INTEGERs); it needs bv_frequency, eos_tderivative, solar_source and lmd_nonlocal set.        */
  int bulk_fluxes;              /* BULK_FLUXES + LONGWAVE (bulk_flux.F:381-948)                                        */
  int lmd_mixing;               /* LMD_MIXING (lmd_vmix.F:100-659, lmd_skpp.F:246-923, lmd_swfrac.F)                   */
  double blk_ZQ, blk_ZT, blk_ZW;/* BLK_ZQ, BLK_ZT, BLK_ZW (m): heights of the humidity / temperature / wind data       */
  int bvf_mixing;               /* BVF_MIXING (bvf_mix.F:92-127; main3d.F:468-469): Akv, Akt from "bvf"; needs bv_frequency   */
  int nospl_vvisc, nospl_vdiff; /* 1: SPLINES_VVISC / SPLINES_VDIFF NOT defined: centred implicit vertical viscosity / diffusion
                                   (step3d_uv.F:397-462, :730-795; step3d_t.F:1196-1198, :1430-1499) instead of the parabolic splines;
                                   0 (what upwelling.h, seamount.h and benchmark.h select): the splines                          */
  int vtransform;               /* Vtransform of roms_*.in in set_depth: 2 (or 0) set_depth.F:210-262, 1 the original transformation :160-208;
                                   hc is the host's SCALARS(ng)%hc either way (MIN(hmin, Tcline) for 1: set_scoord.F:157-163)            */
  int bodyforce, levsfrc, levbfrc; /* BODYFORCE: surface / bottom stress as a body force over levels levsfrc:N / 1:levbfrc (roms_*.in LEVSFRC,
                                   LEVBFRC) in rhs3d (rhs3d.F:326-466, :1588-1599) and no stress boundary flux in pre_step3d (:931-937)   */
  int atm_press;                /* ATM_PRESS: the atmospheric pressure "Pair" (mb, FORCES) enters the pressure gradient (prsgrd31.h:213-215, :294-296,
                                   prsgrd32.h:265-267, prsgrd40.h:194-196)                                                                      */
  int limit_bstress;            /* LIMIT_BSTRESS (set_vbc.F:533-540 and the three drag laws): |bustr| <= 0.75 |u(k=1)| Hz(k=1) / dt        */
  int uv_adv;                   /* momentum advection in rhs3d: 0 the default branch (third-order upstream horizontal, fourth-order
                                   centred vertical), 1 UV_C4ADVECTION (rhs3d.F:685-705, :761-781, :829-849, :902-921, :1108-1175,
                                   :1362-1429; step2d keeps its fourth-order centred default, step2d_LF_AM3.h:1065), 2 UV_SADVECTION
                                   (default horizontal branch, parabolic splines in the vertical: rhs3d.F:1016-1078, :1267-1329),
                                   3 UV_C2ADVECTION (second-order centred in rhs3d.F:605-657, :1079-1107, :1330-1361 AND in step2d,
                                   step2d_LF_AM3.h:1026-1080; LOOP_2D then runs as per-call kernels)                                */
  int qcorrection;              /* QCORRECTION (set_vbc.F:285-299): stflx(itemp) += dqdt (SST - sst); fields "dqdt", "sst" (FORCES)        */
  int limit_stflx_cooling;      /* LIMIT_STFLX_COOLING (:301-328): a cooling heat flux is suppressed where SST < -2 degC                       */
  int scorrection;              /* 1 SCORRECTION, 2 SRELAXATION (:344-351) with Tnudg_salt = Tnudg(isalt) (1/s); field "sss"; Hz is read      */
  int ts_dif4;                  /* TS_DIF4 + MIX_S_TS: biharmonic tracer mixing along s-surfaces (t3dmix4_s.h:215-476), phase
                                   ROMS_B200_T3DMIX4, after the harmonic operator (rhs3d.F:81-97); field "diff4_<itrc>" =
                                   MIXING%diff4 = SQRT(ABS(tnu4)) (read_phypar.F:6905); not with mix_geo_ts                   */
  double Tnudg_salt;            /* Tnudg(isalt,ng) of SCORRECTION / SRELAXATION, 1/s (roms_*.in TNUDG, converted by read_phypar.F)                 */
} roms_b200_config;

/* Fills *cfg with the shipped defaults of roms_<app>.in (Lm,Mm,N = 0 keeps the shipped grid size). */
int roms_b200_default_config(int app, int Lm, int Mm, int N, roms_b200_config* cfg);

/* Tile index sets.  Replaces get_bounds / var_bounds / tile.h (ROMS/Utility/get_bounds.F:2-258, :933-1853).
 * distribute != 0 gives the DISTRIBUTE array bounds (:165-185), else the shared-memory ones (:229-253).
 * out[57] in the order documented in roms_b200_bounds_names(). */
int roms_b200_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, int distribute, int* out57);
const char* roms_b200_bounds_names(void);   /* comma-separated names of the 57 integers */

/* ---- resident form -------------------------------------------------------------------------------------------- */
int roms_b200_create(const roms_b200_config* cfg, roms_b200_handle* out);
int roms_b200_destroy(roms_b200_handle h);
/* array bounds of this handle's tile: LBi, UBi, LBj, UBj */
int roms_b200_array_bounds(roms_b200_handle h, int* out4);

/* Field transfer by name.  Names follow the reference's module members: grid (mod_grid.F) "h","f","pm","pn","om_r",
 * "on_r","om_u","on_u","om_v","on_v","om_p","on_p","omn","fomn","pmon_r","pnom_r","pmon_u","pnom_u","pmon_v","pnom_v",
 * "pmon_p","pnom_p","dndx","dmde","rdrag","rdrag2" ("ZoBot" with UV_LOGDRAG); mixing (mod_mixing.F) "visc2_r","visc2_p","diff2_<itrc>","Akv",
 * "Akt_<itrc>"; ocean (mod_ocean.F) "zeta<1-3>","ubar<1-3>","vbar<1-3>","rzeta<1-2>","rubar<1-2>","rvbar<1-2>",
 * "u<1-2>","v<1-2>","t<1-3>_<itrc>","ru<1-2>","rv<1-2>","rho","pden","W","wvel"; depths "Hz","z_r","z_w","Huon","Hvom";
 * coupling (mod_coupling.F) "Zt_avg1","DU_avg1","DU_avg2","DV_avg1","DV_avg2","rufrc","rvfrc","rhoA","rhoS";
 * forces (mod_forces.F) "sustr","svstr","bustr","bvstr","stflx_<itrc>","btflx_<itrc>","stflux_<itrc>","btflux_<itrc>";
 * with the switches of roms_b200_config that create them: "bvf","alpha","beta","srflx","Jwtype","ghats_<itrc>","Uwind","Vwind",
 * "Tair","Pair","Hair","rain","cloud","lrflx","lhflx","shflx","hsbl","ksbl".
 * <itrc> is 0-based.  n = number of doubles in the whole Fortran array (checked).
 * With a ring attached (roms_b200_attach_nccl) roms_b200_set_field is COLLECTIVE: every upload is followed by the halo
 * exchange of that field (mp_exchange2d/3d), so all tiles must upload the same fields in the same order -- as the
 * reference's distributed build does when every rank runs the same set_data / get_data.  roms_b200_step_forced /
 * roms_b200_step_fields upload WITHOUT an exchange: the surface forcing arrays must carry valid ghost columns (LBi:UBi), as
 * they do in the reference after set_data; the step itself is collective. */
int roms_b200_set_field(roms_b200_handle h, const char* name, const double* host, size_t n);
int roms_b200_get_field(roms_b200_handle h, const char* name, double* host, size_t n);
/* s-coordinate vectors (mod_scalars.F SCALARS%): which = 0 sc_r, 1 Cs_r, 2 sc_w, 3 Cs_w; n = N+1 values indexed by k */
int roms_b200_set_scoord(roms_b200_handle h, int which, const double* v, int n);
/* barotropic filter (set_weights.F): weight(1,:), weight(2,:) with 2*ndtfast+2 values each (slot 0 unused), nfast */
int roms_b200_set_weights(roms_b200_handle h, int nfast, const double* w1, const double* w2, int n);

/* time-index state machine (mod_stepping.F:64-72, main3d.F:189-191,599-611,656-661):
 * idx[13] = iic, ntstart, ntfirst, nstp, nnew, nrhs, iif, indx1, kstp, krhs, knew, PREDICTOR_2D_STEP, exit_flag;
 * tm[2] = time (s), tdays */
int roms_b200_set_indices(roms_b200_handle h, const int* idx13, const double* tm2);
int roms_b200_get_indices(roms_b200_handle h, int* idx13, double* tm2);

/* One phase of main3d on this tile, enqueue + synchronise.  Phase ids: */
enum {
  ROMS_B200_SET_MASSFLUX = 1,  /* set_massflux.F:73      */  ROMS_B200_RHO_EOS = 2,      /* rho_eos.F:111/:576     */
  ROMS_B200_SET_VBC = 3,       /* set_vbc.F:104          */  ROMS_B200_ANA_VMIX = 4,     /* ana_vmix.h             */
  ROMS_B200_OMEGA = 5,         /* omega.F:73             */  ROMS_B200_WVELOCITY = 6,    /* wvelocity.F:61         */
  ROMS_B200_SET_ZETA = 7,      /* set_zeta.F:59          */  ROMS_B200_PRE_STEP3D = 8,   /* pre_step3d.F:123       */
  ROMS_B200_PRSGRD = 9,        /* prsgrd32.h:106/31:97   */  ROMS_B200_T3DMIX = 10,      /* t3dmix2_s.h:89/_geo:90 */
  ROMS_B200_RHS3D = 11,        /* rhs3d.F:174            */  ROMS_B200_UV3DMIX = 12,     /* uv3dmix2_s.h:114       */
  ROMS_B200_STEP2D = 13,       /* step2d_LF_AM3.h:137    */  ROMS_B200_SET_DEPTH = 14,   /* set_depth.F:82         */
  ROMS_B200_STEP3D_UV = 15,    /* step3d_uv.F:111        */  ROMS_B200_OMEGA2 = 16,      /* main3d.F:789           */
  ROMS_B200_STEP3D_T = 17,     /* step3d_t.F:108         */  ROMS_B200_DIAG = 18,        /* diag.F:80              */
  ROMS_B200_SET_DATA = 19,     /* (host forcing; no-op)  */  ROMS_B200_STEP2D_LOOP = 20, /* main3d.F:592-700       */
  ROMS_B200_SET_AVG = 22,      /* set_avg.F:128          */
  ROMS_B200_BULK_FLUX = 23,    /* bulk_flux.F:59         */
  ROMS_B200_LMD_VMIX = 24,     /* lmd_vmix.F:33 (lmd_vmix_tile, lmd_skpp, lmd_finish) */
  ROMS_B200_BVF_MIX = 25,      /* bvf_mix.F:27           */
  ROMS_B200_T3DMIX4 = 26       /* t3dmix4_s.h:41 (TS_DIF4 + MIX_S_TS; rhs3d.F:89-97) */
};
int roms_b200_run_phase(roms_b200_handle h, int phase);

/* nsteps baroclinic steps, device resident (main3d.F:189-917 without get_data/output).  Asynchronous: returns after
 * enqueueing; roms_b200_sync waits.  Forcing (sustr, svstr, stflux, btflux) is whatever was last uploaded. */
int roms_b200_main3d_step(roms_b200_handle h, int nsteps);
int roms_b200_sync(roms_b200_handle h);
/* AVERAGES (ROMS/Nonlinear/set_avg.F, called from main3d.F:494 right after set_zeta): time averages of the chain's state
 * variables over windows of nAVG steps starting after step ntsAVG (roms_*.in NAVG, NTSAVG; nAVG = 0 switches them off),
 * accumulated on the device by every step.  The averages are fields like any other (roms_b200_get_field): "avgzeta", "avgu2d",
 * "avgv2d" (2-D), "avgu3d", "avgv3d", "avgrho", "avgt_<itrc>" (1:N), "avgw3d" = W*pm*pn, "avgwvel" (0:N); a window's averages are
 * complete after the step with MOD(iic-1, nAVG) == 0, which is when the reference writes them (wrt_avg). */
int roms_b200_set_avg(roms_b200_handle h, int nAVG, int ntsAVG);
/* Run-time switches of a handle (set before stepping; the captured time-step graphs are dropped):
 *   "cuda_graphs"      1 (default) replay one CUDA graph per baroclinic step, 0 plain stream launches
 *   "step2d_exchange"  how the xi-halo of the barotropic sub-steps travels on the NVLink peer path (before attach only):
 *                      2 (default) inside the step2d kernels with edge-first split launches, 1 inside the kernels with one
 *                      launch per sub-step, 0 stand-alone exchange kernels
 *   "step2d_loop_kernel" 1 (default) run LOOP_2D (main3d.F:592-700) as one persistent kernel whenever all CTAs of the tile
 *                      can be resident at once (small tiles: BENCHMARK3 on 8 GPUs, BENCHMARK1 on one), 0 one launch per call
 *   "fuse_phases"      1 (default) roms_b200_main3d_step / step_forced fuse routines that share operands (t3dmix2_s into
 *                      pre_step3d's tracer pass), 0 one kernel group per
 *                      routine as roms_b200_run_phase always does; the strict build gives the same bits either way
 *   "ghost_compute"    1 (default) in a ring, bulk_flux and set_vbc compute their ghost columns (inputs are valid there) instead
 *                      of exchanging them; 0 exchange after every phase as the reference's mp_exchange calls do.  Same bits
 *                      either way on every column a routine reads; with 1 the outermost ghost column of bustr, bvstr, sustr
 *                      may be stale (nothing reads it)
 *   "overlap"          1 (default) edge-first two-stream overlap of halo exchanges with interior compute (before attach only)
 *   "halo_timeout_s"   seconds a kernel waits for a neighbour's halo before it gives up and raises exit_flag 8
 *                      (default 30; <= 0 waits for ever)
 * Unknown key: 2; a value that cannot be applied in the handle's state: 5. */
int roms_b200_set_option(roms_b200_handle h, const char* key, double value);
/* One step the way main3d sees it from the host: H2D of this step's surface forcing (the set_data products
 * sustr(LBi:UBi,LBj:UBj), svstr, and optionally stflux for each tracer; NULL keeps the resident value), the step,
 * then D2H of the diag scalars.  out12 = avgke, avgpe, avgkp, volume, max_speed, maxCu, maxCv, maxCw,
 * ubarmax, vbarmax, umax, vmax (diag.F:293-437, ana_diag.h:116-142).  Synchronous. */
int roms_b200_step_forced(roms_b200_handle h, const double* sustr, const double* svstr, const double* stflux_temp,
                          size_t n2d, double* out12);
/* The same with any set of 2-D forcing arrays, by name -- with cfg.bulk_fluxes the set_data products are the atmosphere
 * ("Uwind","Vwind","Tair","Pair","Hair","rain","cloud","srflx") instead of the stresses.  arrays[i] == NULL keeps the
 * resident value of names[i]; nfields <= 16. */
int roms_b200_step_fields(roms_b200_handle h, int nfields, const char* const* names, const double* const* arrays,
                          size_t n2d, double* out12);
int roms_b200_diag(roms_b200_handle h, double* out12);
/* Pin a caller-owned host range (cudaHostRegister) so that roms_b200_step_forced copies from it directly instead of
 * staging through the library's own pinned buffer.  Meant for the module arrays the Fortran host allocates once
 * (FORCES(ng)%sustr/svstr/stflux, mod_forces.F:185-463): register after ROMS_allocate_arrays, the range must stay
 * mapped until roms_b200_unregister_host / roms_b200_destroy. */
int roms_b200_register_host(roms_b200_handle h, void* p, size_t bytes);
int roms_b200_unregister_host(roms_b200_handle h, void* p);

/* timing helpers for bench.py: elapsed device ms between two internal CUDA events bracketing the last
 * roms_b200_main3d_step call; per-phase accumulated device ms since the last reset (wclock regions, timers.F). */
int roms_b200_last_step_ms(roms_b200_handle h, float* ms);
/* on = 1: every phase bracketed by events and synchronised (plain stream launches); on = 2: events between consecutive
 * phases recorded inside the captured CUDA graph of the step -- the split of the configuration that is actually timed
 * (the phase times add up to the step); 0: off.  Phase ids index ms_by_phase32. */
int roms_b200_profile_enable(roms_b200_handle h, int on);
int roms_b200_profile_get(roms_b200_handle h, double* ms_by_phase32, long long* launches);
long long roms_b200_launch_count(roms_b200_handle h);

/* multi-GPU: ring neighbours in xi (NtileI>1, NtileJ==1).  comm is an ncclComm_t created by the caller (one rank per
 * tile); the library then performs mp_exchange2d/3d/4d (ROMS/Utility/mp_exchange.F:1413-2128) with grouped
 * ncclSend/ncclRecv on its own stream. */
int roms_b200_attach_nccl(roms_b200_handle h, void* nccl_comm, int rank, int nranks);
int roms_b200_nccl_unique_id(char* out128);
int roms_b200_nccl_init_rank(const char* id128, int rank, int nranks, void** comm_out);
/* Optional NVLink peer path for the same exchanges (the NCCL path stays the fall-back): every tile owns a mailbox in HBM
 * that its two ring neighbours map through CUDA IPC and store into directly, so one exchange is a single kernel (remote
 * stores + flag, then flag wait + local copy) with no NCCL rendezvous -- what the ~118 small
 * mp_exchange2d calls of the barotropic loop (step2d_LF_AM3.h:586,884,924,2519) need.  Protocol, all ranks:
 * peer_export -> 64-byte IPC handle; exchange handles on the host; peer_attach(west's, east's); agree that every rank
 * succeeded; peer_enable(1).  peer_error returns 1 if a wait for a neighbour ever timed out (results invalid). */
int roms_b200_peer_export(roms_b200_handle h, char* out64);
int roms_b200_peer_attach(roms_b200_handle h, const char* west64, const char* east64);
int roms_b200_peer_enable(roms_b200_handle h, int on);
int roms_b200_peer_error(roms_b200_handle h);
/* Test hook: raises the sticky device error word exactly as a timed-out halo wait does (mp_exchange.F:1698-1705 sets
 * exit_flag and returns; here roms_b200_sync / run_phase / step_forced / diag / get_field return 8 from then on). */
int roms_b200_peer_error_inject(roms_b200_handle h);

/* ---- per-routine host-pointer form (mirrors the _tile argument lists; used by the parity tests) ---------------- */
/* roms_b200_tile_t carries what tile.h/set_bounds.h give a _tile routine. */
typedef struct roms_b200_tile_t {
  roms_b200_config cfg;        /* cfg.tile selects the tile                                                       */
  int iic, ntfirst;            /* start-up branches (pre_step3d.F:947, step3d_uv.F:305, step2d_LF_AM3.h:1885)      */
  int nstp, nnew, nrhs;        /* 3-D time indices                                                                */
  int iif, kstp, krhs, knew, predictor;   /* 2-D time indices, PREDICTOR_2D_STEP                                   */
} roms_b200_tile_t;

/* rho_eos_tile (rho_eos.F:111 / :576): t = t(:,:,:,nrhs,itemp), s = t(:,:,:,nrhs,isalt) or NULL */
int roms_b200_rho_eos_tile(const roms_b200_tile_t* b, const double* Hz, const double* z_r, const double* z_w,
                           const double* t, const double* s, double* rhoA, double* rhoS, double* pden, double* rho);
/* prsgrd32_tile / prsgrd31_tile (prsgrd32.h:106, prsgrd31.h:97): ru, rv = ru(:,:,0:N,nrhs), rv(:,:,0:N,nrhs) */
int roms_b200_prsgrd_tile(const roms_b200_tile_t* b, const double* Hz, const double* om_v, const double* on_u,
                          const double* z_r, const double* z_w, const double* rho, double* ru, double* rv);
/* set_massflux_tile (set_massflux.F:73) */
int roms_b200_set_massflux_tile(const roms_b200_tile_t* b, const double* u, const double* v, const double* Hz,
                                const double* om_v, const double* on_u, double* Huon, double* Hvom);
/* omega_tile (omega.F:73) */
int roms_b200_omega_tile(const roms_b200_tile_t* b, const double* Huon, const double* Hvom, const double* z_w, double* W);
/* set_depth_tile (set_depth.F:82), Vtransform = 2: sc/Cs vectors have N+1 entries indexed by k */
int roms_b200_set_depth_tile(const roms_b200_tile_t* b, const double* h, const double* Zt_avg1, const double* sc_r,
                             const double* Cs_r, const double* sc_w, const double* Cs_w, double* Hz, double* z_r, double* z_w);


/* ---- generic per-routine form: every routine of the chain ------------------------------------------------------ */
/* The remaining _tile routines (set_vbc.F:104, wvelocity.F:61, set_zeta.F:59, pre_step3d.F:123, t3dmix2_s.h:89 /
 * t3dmix2_geo.h:90, rhs3d.F:174, uv3dmix2_s.h:114, step2d_LF_AM3.h:137, step3d_uv.F:111, step3d_t.F:108, ana_vmix.h)
 * take 20-70 whole arrays each; this entry point passes them by NAME (the names of roms_b200_set_field: the members of
 * OCEAN/GRID/COUPLING/MIXING/FORCES, time level spelled out, tracer index as suffix _0, _1) instead of by position.
 * phase = a ROMS_B200_* routine id; mode[i]: 1 input, 2 output (current content uploaded first, so untouched elements
 * are preserved like an INTENT(inout) dummy), 3 both; every array is the whole Fortran array (LBi:UBi,LBj:UBj[,k]).
 * scoord4 = sc_r, Cs_r, sc_w, Cs_w (N+1 entries each, set_depth only, else NULL); weight1/weight2 = weight(1,:),
 * weight(2,:) of set_weights.F with nweight entries each and nfast (step2d only, else NULL).  Nothing is retained. */
int roms_b200_routine_tile(const roms_b200_tile_t* b, int phase, int nargs, const char* const* names, double* const* arrays,
                           const int* mode, const double* scoord4, int nfast, const double* weight1, const double* weight2,
                           int nweight);
/* "in:<comma-separated names>;out:<names>" a routine needs, NULL for an unknown phase.  `*` = one entry per tracer.  A leading
 * `?` marks an argument only the optional terms of the routine touch (bvf / alpha / beta of rho_eos; z_w, srflx, Jwtype, ghats of
 * pre_step3d with SOLAR_SOURCE / LMD_NONLOCAL): pass it (without the `?`) when the configuration has the array, leave it out
 * otherwise -- an optional input that is left out reads as zero, an optional output that is left out is discarded. */
const char* roms_b200_routine_args(int phase);
/* vertical extent of a named field: first level (0 or 1) and number of planes (1, N or N+1) */
int roms_b200_field_levels(roms_b200_handle h, const char* name, int* LBk, int* nk);

#ifdef __cplusplus
}
#endif
#endif /* ROMS_B200_H */

/*
 * c_driver.c -- a plain-C host that plays ROMS/Nonlinear/main3d.F:189-191,:307-814 against the C-ABI of
 * include/roms_b200.h, the way the Fortran shim (fortran/mod_b200.F) does.  TEST INFRASTRUCTURE: it exercises the drop-in
 * boundary from a compiled host with no Python in between.
 *
 *   c_driver <state_in> <state_out> routine|resident <nsteps>
 *
 * routine : every _tile routine of the chain goes through roms_b200_routine_tile with the whole host arrays passed BY NAME
 *           (list from roms_b200_routine_args), in main3d's call order; the time-index state machine of main3d / LOOP_2D
 *           (mod_stepping.F:64-72, main3d.F:189-191,592-700) runs HERE, on the host, exactly as in the reference.
 * resident: roms_b200_create / set_field / set_scoord / set_weights / set_indices / main3d_step / sync / get_field.
 * layout  : (`c_driver layout`) prints sizeof / offsetof of the two boundary structs for the struct-layout test.
 *
 * State file: "RB2S", int32 nrec, then nrec x { char name[24]; int32 nk, nj, ni; double data[nk*nj*ni] }.  Records whose
 * name starts with '@' are not fields: @config (raw roms_b200_config), @indices (13 ints + time + tdays as doubles),
 * @sc_r @Cs_r @sc_w @Cs_w (N+1 each), @w1 @w2 (weights), @nfast.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "roms_b200.h"

#define MAXREC 256
typedef struct { char name[24]; int nk, nj, ni; double* data; } rec_t;
static rec_t recs[MAXREC];
static int nrec = 0;

static rec_t* find(const char* name) {
  for (int i = 0; i < nrec; ++i) if (strcmp(recs[i].name, name) == 0) return &recs[i];
  return NULL;
}
static rec_t* need(const char* name) {
  rec_t* r = find(name);
  if (!r) { fprintf(stderr, "c_driver: record '%s' missing\n", name); exit(3); }
  return r;
}

static int read_state(const char* path) {
  FILE* f = fopen(path, "rb");
  char magic[4];
  int32_t n = 0;
  if (!f || fread(magic, 1, 4, f) != 4 || memcmp(magic, "RB2S", 4) != 0 || fread(&n, 4, 1, f) != 1 || n > MAXREC) return 1;
  for (int i = 0; i < n; ++i) {
    rec_t* r = &recs[i];
    int32_t d[3];
    if (fread(r->name, 1, 24, f) != 24 || fread(d, 4, 3, f) != 3) return 1;
    r->nk = d[0]; r->nj = d[1]; r->ni = d[2];
    const size_t cnt = (size_t)r->nk * r->nj * r->ni;
    r->data = (double*)malloc(cnt * sizeof(double));
    if (!r->data || fread(r->data, sizeof(double), cnt, f) != cnt) return 1;
  }
  nrec = n;
  fclose(f);
  return 0;
}

static int write_state(const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  int32_t n = nrec;
  fwrite("RB2S", 1, 4, f); fwrite(&n, 4, 1, f);
  for (int i = 0; i < nrec; ++i) {
    const rec_t* r = &recs[i];
    int32_t d[3] = {r->nk, r->nj, r->ni};
    fwrite(r->name, 1, 24, f); fwrite(d, 4, 3, f);
    fwrite(r->data, sizeof(double), (size_t)r->nk * r->nj * r->ni, f);
  }
  return fclose(f);
}

/* ---- mod_stepping.F state, owned by the host as in the reference ---- */
static roms_b200_config cfg;
static int iic, ntstart, ntfirst, nstp, nnew, nrhs, iif, indx1, kstp, krhs, knew, predictor, exit_flag;
static double time_s, tdays;
static int nfast, nweight;

#define CHECK(call) do { exit_flag = (call); if (exit_flag) { fprintf(stderr, "c_driver: %s -> exit_flag %d\n", #call, exit_flag); return exit_flag; } } while (0)

/* One routine through the by-name entry point: what `CALL b200_routine (ng, tile, phase)` does in fortran/mod_b200.F. */
static int routine(int phase) {
  const char* spec = roms_b200_routine_args(phase);
  if (!spec) return 5;
  static char names[128][24];
  const char* cn[128];
  double* arrs[128];
  int mode[128];
  int nargs = 0;
  const char* p = spec + 3;                         /* after "in:" */
  int bit = 1;
  while (*p) {
    if (*p == ';') { p += 5; bit = 2; continue; }   /* ";out:" */
    char tok[24];
    int l = 0;
    while (*p && *p != ',' && *p != ';') tok[l++] = *p++;
    tok[l] = 0;
    if (*p == ',') ++p;
    const int optional = (tok[0] == '?');           /* present only with the cpp switch that creates the array */
    if (optional) memmove(tok, tok + 1, strlen(tok));
    char* star = strchr(tok, '*');
    for (int it = 0; it < (star ? cfg.NT : 1); ++it) {
      char one[24];
      strcpy(one, tok);
      if (star) one[star - tok] = (char)('0' + it);
      if (optional && !find(one)) continue;
      int j = 0;
      while (j < nargs && strcmp(names[j], one) != 0) ++j;
      if (j == nargs) { strcpy(names[nargs], one); cn[nargs] = names[nargs]; arrs[nargs] = need(one)->data; mode[nargs] = 0; ++nargs; }
      mode[j] |= bit;
    }
  }
  roms_b200_tile_t b;
  memset(&b, 0, sizeof(b));
  b.cfg = cfg; b.iic = iic; b.ntfirst = ntfirst; b.nstp = nstp; b.nnew = nnew; b.nrhs = nrhs;
  b.iif = iif; b.kstp = kstp; b.krhs = krhs; b.knew = knew; b.predictor = predictor;
  double* sc4 = NULL;
  if (phase == ROMS_B200_SET_DEPTH) {
    const int n1 = cfg.N + 1;
    sc4 = (double*)malloc(4 * n1 * sizeof(double));
    memcpy(sc4, need("@sc_r")->data, n1 * sizeof(double)); memcpy(sc4 + n1, need("@Cs_r")->data, n1 * sizeof(double));
    memcpy(sc4 + 2 * n1, need("@sc_w")->data, n1 * sizeof(double)); memcpy(sc4 + 3 * n1, need("@Cs_w")->data, n1 * sizeof(double));
  }
  const int st2 = phase == ROMS_B200_STEP2D;
  const int rc = roms_b200_routine_tile(&b, phase, nargs, cn, arrs, mode, sc4, nfast, st2 ? need("@w1")->data : NULL,
                                        st2 ? need("@w2")->data : NULL, st2 ? nweight : 0);
  free(sc4);
  return rc;
}

/* main3d.F:189-191 + :307-814 for one baroclinic step, routine by routine */
static int main3d_by_routine(void) {
  nstp = 1 + ((iic - ntstart) % 2); nnew = 3 - nstp; nrhs = nstp;          /* :189-191 */
  tdays = time_s / 86400.0;
  CHECK(routine(ROMS_B200_SET_MASSFLUX));                                  /* :307 */
  CHECK(routine(ROMS_B200_RHO_EOS));
  if (cfg.bulk_fluxes) CHECK(routine(ROMS_B200_BULK_FLUX));                /* :384-390 */
  CHECK(routine(ROMS_B200_SET_VBC));                                       /* :429 */
  if (cfg.ana_vmix) CHECK(routine(ROMS_B200_ANA_VMIX));                    /* :464-470 */
  else if (cfg.lmd_mixing) CHECK(routine(ROMS_B200_LMD_VMIX));
  else if (cfg.bvf_mixing) CHECK(routine(ROMS_B200_BVF_MIX));
  CHECK(routine(ROMS_B200_OMEGA));                                         /* :474 */
  if (cfg.wvelocity_every_step) CHECK(routine(ROMS_B200_WVELOCITY));
  CHECK(routine(ROMS_B200_SET_ZETA));                                      /* :531 */
  CHECK(routine(ROMS_B200_PRE_STEP3D));                                    /* rhs3d.F:25-167 */
  CHECK(routine(ROMS_B200_PRSGRD));
  CHECK(routine(ROMS_B200_T3DMIX));
  if (cfg.ts_dif4) CHECK(routine(ROMS_B200_T3DMIX4));
  CHECK(routine(ROMS_B200_RHS3D));
  CHECK(routine(ROMS_B200_UV3DMIX));
  for (int my_iif = 1; my_iif <= nfast + 1; ++my_iif) {                    /* LOOP_2D, :592-700 */
    const int next_indx1 = 3 - indx1;
    if (!predictor && my_iif <= nfast + 1) {
      predictor = 1; iif = my_iif;
      kstp = (iif == 1) ? indx1 : 3 - indx1;
      knew = 3; krhs = indx1;
    }
    if (my_iif <= nfast + 1) CHECK(routine(ROMS_B200_STEP2D));
    if (predictor) {
      predictor = 0; knew = next_indx1; kstp = 3 - knew; krhs = 3;
      if (iif < nfast + 1) indx1 = next_indx1;
    }
    if (iif < nfast + 1) CHECK(routine(ROMS_B200_STEP2D));
  }
  CHECK(routine(ROMS_B200_SET_DEPTH));                                     /* :744 */
  CHECK(routine(ROMS_B200_STEP3D_UV));
  CHECK(routine(ROMS_B200_OMEGA2));                                        /* :789 */
  CHECK(routine(ROMS_B200_STEP3D_T));
  iic += 1; time_s += cfg.dt;
  return 0;
}

static int run_resident(int nsteps) {
  roms_b200_handle h = NULL;
  CHECK(roms_b200_create(&cfg, &h));
  for (int i = 0; i < nrec; ++i)
    if (recs[i].name[0] != '@') CHECK(roms_b200_set_field(h, recs[i].name, recs[i].data, (size_t)recs[i].nk * recs[i].nj * recs[i].ni));
  const char* sc[4] = {"@sc_r", "@Cs_r", "@sc_w", "@Cs_w"};
  for (int w = 0; w < 4; ++w) CHECK(roms_b200_set_scoord(h, w, need(sc[w])->data, cfg.N + 1));
  CHECK(roms_b200_set_weights(h, nfast, need("@w1")->data, need("@w2")->data, nweight));
  int idx[13] = {iic, ntstart, ntfirst, nstp, nnew, nrhs, iif, indx1, kstp, krhs, knew, predictor, exit_flag};
  double tm[2] = {time_s, tdays};
  CHECK(roms_b200_set_indices(h, idx, tm));
  CHECK(roms_b200_main3d_step(h, nsteps));
  CHECK(roms_b200_sync(h));
  for (int i = 0; i < nrec; ++i)
    if (recs[i].name[0] != '@') CHECK(roms_b200_get_field(h, recs[i].name, recs[i].data, (size_t)recs[i].nk * recs[i].nj * recs[i].ni));
  CHECK(roms_b200_get_indices(h, idx, tm));
  iic = idx[0]; nstp = idx[3]; nnew = idx[4]; nrhs = idx[5]; iif = idx[6]; indx1 = idx[7]; kstp = idx[8]; krhs = idx[9]; knew = idx[10];
  predictor = idx[11]; exit_flag = idx[12]; time_s = tm[0]; tdays = tm[1];
  return roms_b200_destroy(h);
}

static void print_layout(void) {
#define OFFC(m) printf("config.%s %zu\n", #m, offsetof(roms_b200_config, m))
#define OFFT(m) printf("tile.%s %zu\n", #m, offsetof(roms_b200_tile_t, m))
  printf("sizeof.config %zu\nsizeof.tile %zu\n", sizeof(roms_b200_config), sizeof(roms_b200_tile_t));
  OFFC(Lm); OFFC(Mm); OFFC(N); OFFC(NT); OFFC(NtileI); OFFC(NtileJ); OFFC(tile); OFFC(ndtfast); OFFC(dt); OFFC(nonlin_eos); OFFC(dj_gradps);
  OFFC(curvgrid); OFFC(mix_geo_ts); OFFC(uv_qdrag); OFFC(salinity); OFFC(ana_vmix); OFFC(wvelocity_every_step); OFFC(hadv); OFFC(vadv);
  OFFC(rho0); OFFC(g); OFFC(R0); OFFC(T0); OFFC(S0); OFFC(Tcoef); OFFC(Scoef); OFFC(Akt_bak); OFFC(Akv_bak); OFFC(gamma2); OFFC(lambda);
  OFFC(hc); OFFC(itemp); OFFC(isalt); OFFC(device);
  OFFC(bv_frequency); OFFC(eos_tderivative); OFFC(solar_source); OFFC(lmd_nonlocal);
  OFFC(bulk_fluxes); OFFC(lmd_mixing); OFFC(blk_ZQ); OFFC(blk_ZT); OFFC(blk_ZW); OFFC(bvf_mixing); OFFC(nospl_vvisc); OFFC(nospl_vdiff); OFFC(vtransform); OFFC(bodyforce); OFFC(levsfrc); OFFC(levbfrc); OFFC(atm_press); OFFC(limit_bstress); OFFC(uv_adv); OFFC(qcorrection); OFFC(limit_stflx_cooling); OFFC(scorrection); OFFC(ts_dif4); OFFC(Tnudg_salt);
  OFFT(cfg); OFFT(iic); OFFT(ntfirst); OFFT(nstp); OFFT(nnew); OFFT(nrhs); OFFT(iif); OFFT(kstp); OFFT(krhs); OFFT(knew); OFFT(predictor);
}

int main(int argc, char** argv) {
  if (argc == 2 && strcmp(argv[1], "layout") == 0) { print_layout(); return 0; }
  if (argc != 5) { fprintf(stderr, "usage: c_driver <state_in> <state_out> routine|resident <nsteps> | c_driver layout\n"); return 2; }
  if (read_state(argv[1])) { fprintf(stderr, "c_driver: cannot read %s\n", argv[1]); return 2; }
  memcpy(&cfg, need("@config")->data, sizeof(cfg));
  const double* ix = need("@indices")->data;
  iic = (int)ix[0]; ntstart = (int)ix[1]; ntfirst = (int)ix[2]; nstp = (int)ix[3]; nnew = (int)ix[4]; nrhs = (int)ix[5]; iif = (int)ix[6];
  indx1 = (int)ix[7]; kstp = (int)ix[8]; krhs = (int)ix[9]; knew = (int)ix[10]; predictor = (int)ix[11]; exit_flag = (int)ix[12];
  time_s = ix[13]; tdays = ix[14];
  nfast = (int)need("@nfast")->data[0];
  nweight = need("@w1")->ni;
  const int nsteps = atoi(argv[4]);
  int rc = 0;
  if (strcmp(argv[3], "routine") == 0) {
    for (int s = 0; s < nsteps && !rc; ++s) rc = main3d_by_routine();
  } else if (strcmp(argv[3], "resident") == 0) {
    rc = run_resident(nsteps);
  } else return 2;
  if (rc) return 10 + rc;
  double* o = need("@indices")->data;
  const double v[15] = {iic, ntstart, ntfirst, nstp, nnew, nrhs, iif, indx1, kstp, krhs, knew, predictor, exit_flag, time_s, tdays};
  memcpy(o, v, sizeof(v));
  return write_state(argv[2]) ? 4 : 0;
}

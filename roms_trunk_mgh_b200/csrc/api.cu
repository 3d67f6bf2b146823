// C-ABI (include/roms_b200.h) + device-resident state + the main3d orchestration for one tile on one B200.
// The product path has no CPU fallback: every compute entry point needs a CUDA device and returns exit_flag 8 otherwise.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include "state.h"
#include "kernels.h"

using namespace rb;
using namespace rbi;

namespace {

enum { NoError = 0, BlowUp = 1, InputError = 2, ConfigError = 5, FatalError = 8 };

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      std::fprintf(stderr, "roms_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return FatalError;                                                                                 \
    }                                                                                                    \
  } while (0)

// ---- tile index sets: ROMS/Utility/get_bounds.F (tile_bounds_2d :933-1007, var_bounds :1009-1853, get_bounds :60-258)
void tile_range(int n, int ntile, int t, int& s, int& e) {
  const int chunk = (n + ntile - 1) / ntile;
  const int margin = (ntile * chunk - n) / 2;
  s = 1 + t * chunk - margin;
  e = s + chunk - 1;
  s = std::max(s, 1);
  e = std::min(e, n);
}

// EW periodic, NS closed (the LBC set of all supported applications)
void make_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, bool distribute, Bounds& b) {
  const int Ng = 2;
  b.tile = tile; b.Jtile = tile / NtileI; b.Itile = tile - b.Jtile * NtileI;
  int is, ie, js, je;
  tile_range(Lm, NtileI, b.Itile, is, ie);
  tile_range(Mm, NtileJ, b.Jtile, js, je);
  b.west = b.Itile == 0; b.east = b.Itile == NtileI - 1; b.south = b.Jtile == 0; b.north = b.Jtile == NtileJ - 1;
  // xi: periodic -> no clipping on any tile
  b.Istr = b.IstrP = b.IstrR = b.IstrT = b.IstrU = b.IstrB = b.IstrM = is;
  b.Istrm3 = is - 3; b.Istrm2 = is - 2; b.Istrm1 = is - 1; b.IstrUm2 = is - 2; b.IstrUm1 = is - 1;
  b.Iend = b.IendR = b.IendP = b.IendT = b.IendB = ie;
  b.Iendp1 = ie + 1; b.Iendp2 = ie + 2; b.Iendp2i = ie + 2; b.Iendp3 = ie + 3;
  // eta: closed walls on the southern / northern edge tiles
  b.Jstr = b.JstrP = js;
  if (b.south) {
    b.JstrR = js - 1; b.JstrT = b.JstrR; b.JstrV = js + 1; b.JstrB = b.JstrT + 1; b.JstrM = b.JstrP + 1;
    b.Jstrm3 = std::max(0, js - 3); b.Jstrm2 = std::max(0, js - 2); b.JstrVm2 = std::max(1, b.JstrV - 2);
    b.Jstrm1 = std::max(1, js - 1); b.JstrVm1 = std::max(2, b.JstrV - 1);
  } else {
    b.JstrR = b.JstrT = b.JstrV = b.JstrB = b.JstrM = js;
    b.Jstrm3 = js - 3; b.Jstrm2 = js - 2; b.JstrVm2 = js - 2; b.Jstrm1 = js - 1; b.JstrVm1 = js - 1;
  }
  b.Jend = je;
  if (b.north) {
    b.JendR = je + 1; b.JendP = b.JendR; b.JendT = b.JendR; b.JendB = b.JendT - 1;
    b.Jendp1 = std::min(je + 1, Mm); b.Jendp2i = std::min(je + 2, Mm); b.Jendp2 = std::min(je + 2, Mm + 1); b.Jendp3 = std::min(je + 3, Mm + 1);
  } else {
    b.JendR = b.JendP = b.JendT = b.JendB = je;
    b.Jendp1 = je + 1; b.Jendp2i = je + 2; b.Jendp2 = je + 2; b.Jendp3 = je + 3;
  }
  b.IminS = is - 3; b.ImaxS = ie + 3; b.JminS = js - 3; b.JmaxS = je + 3;
  const int Imin = -Ng, Imax = Lm + Ng, Jmin = 0, Jmax = Mm + 1;
  if (distribute) {
    b.LBi = (b.Itile == 0) ? Imin : is - Ng;
    b.UBi = (b.Itile == NtileI - 1) ? Imax : ie + Ng;
    b.LBj = (b.Jtile == 0) ? Jmin : js - Ng;
    b.UBj = (b.Jtile == NtileJ - 1) ? Jmax : je + Ng;
  } else {
    b.LBi = Imin; b.UBi = Imax; b.LBj = Jmin; b.UBj = Jmax;
  }
}

}  // namespace

namespace {

// Device storage of one registered field (guard rows included); sets *fi.slot.
int materialize(roms_b200_state* h, FieldInfo& fi) {
  if (fi.base) return NoError;
  const size_t n = (size_t)h->par.PL * fi.nk;
  // Two zeroed guard rows on either side of every field.  Stencil kernels load the operands of clamped neighbours
  // unconditionally (the values are then discarded), so a row just outside rows LBj..UBj of the first / last plane must be
  // addressable whatever the surrounding address space looks like -- an address just before a cudaMalloc block is only
  // mapped by accident (a BENCHMARK3 run faulted in k_uv3dmix2 when the preceding block had been freed).
  const size_t guard = ((size_t)2 * h->par.P + 31) / 32 * 32;
  double* raw = nullptr;
  const size_t bytes = (((n + 2 * guard) * sizeof(double) + 255) / 256) * 256;
  CK(cudaMalloc(&raw, bytes));
  h->allocs.push_back(raw);
  CK(cudaMemsetAsync(raw, 0, (n + 2 * guard) * sizeof(double), h->stream));
  fi.base = raw + guard;
  // element (i,j,k) lives at base[(i-LBi+ioff) + (j-LBj)*P + (k-LBk)*PL]
  *fi.slot = fi.base + h->ioff - h->LBi_dev - (ptrdiff_t)h->b.LBj * h->par.P - (ptrdiff_t)fi.LBk * h->par.PL;
  return NoError;
}

// Register a field.  Normally its storage is allocated at once; a handle created for ONE per-routine call (api_tile.cu) is
// `lazy`: a field gets storage when the caller hands it over (roms_b200_set_field) -- a routine's published argument list is
// everything it touches -- so a _tile call on BENCHMARK3 allocates the routine's 5-50 arrays instead of the whole 5 GB state.
int alloc_field(roms_b200_state* h, const std::string& name, double** slot, int LBk, int nk) {
  *slot = nullptr;
  FieldInfo fi{slot, LBk, nk, nullptr};
  if (!h->lazy || name == "P3" || name == "Taux" || name == "Tauy") { const int rc = materialize(h, fi); if (rc) return rc; }
  h->reg[name] = fi;
  return NoError;
}

void fill_par(roms_b200_state* h) {
  Par& p = h->par;
  p.nstp = h->nstp; p.nnew = h->nnew; p.nrhs = h->nrhs;
  p.istart = (h->iic == h->ntfirst) ? 0 : (h->iic == h->ntfirst + 1 ? 1 : 2);
  p.iif = h->iif; p.kstp = h->kstp; p.krhs = h->krhs; p.knew = h->knew; p.ptsk = 3 - h->kstp; p.predictor = h->predictor; p.nfast = h->nfast;
  p.dtfast = h->dtfast;
  auto W1 = [&](int i) { return (i >= 0 && i < (int)h->w1.size()) ? h->w1[i] : 0.0; };
  auto W2 = [&](int i) { return (i >= 0 && i < (int)h->w2.size()) ? h->w2[i] : 0.0; };
  p.w1_m1 = W1(h->iif - 1); p.w2_0 = W2(h->iif); p.w2_p1 = W2(h->iif + 1);
}

// profile = 2: mark the start of `phase` (or, with phase < 0, the end of the step) on the main stream
void mark_phase(roms_b200_state* h, int phase) {
  if (h->profile != 2 || h->n_ev_ph >= 40) return;
  const int q = h->n_ev_ph++;
  if (!h->ev_ph[q]) cudaEventCreate(&h->ev_ph[q]);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(h->stream, &cs);
  cudaEventRecordWithFlags(h->ev_ph[q], h->stream, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
  h->ev_ph_phase[q] = phase;
}
// profile = 2: after the step has completed, charge the interval between consecutive marks to the phase that started it
void collect_marks(roms_b200_state* h) {
  for (int q = 0; q + 1 < h->n_ev_ph; ++q) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_ph[q], h->ev_ph[q + 1]) == cudaSuccess && h->ev_ph_phase[q] >= 0) h->phase_ms[h->ev_ph_phase[q] & 31] += ms;
  }
  cudaGetLastError();
}

// set_avg.F:237-240 / :1264 / :2298-2301 for the step about to run: bit 0 initialise, bit 1 accumulate, bit 2 normalise
int avg_mode(const roms_b200_state* h) {
  if (h->navg <= 0) return 0;
  const int iic = h->iic, n = h->navg, s0 = h->ntsavg;
  int m = 0;
  if ((iic > s0 && (iic - 1) % n == 1) || (iic >= s0 && n == 1)) m |= 1;
  else if (iic > s0) m |= 2;
  if ((iic > s0 && (iic - 1) % n == 0) || (iic >= s0 && n == 1)) m |= 4;
  return m;
}

struct PhaseTimer {
  roms_b200_state* h; int phase; cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(roms_b200_state* h_, int ph) : h(h_), phase(ph) {
    if (h->profile == 1) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, h->stream); }
    else if (h->profile == 2 && h->in_step) mark_phase(h, ph);
  }
  ~PhaseTimer() {
    if (h->profile == 1) {
      cudaEventRecord(b, h->stream); cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b); h->phase_ms[phase & 31] += ms;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
  }
};

// Fields whose xi-ghost columns must be refreshed after a phase (the mp_exchange call sites of SURVEY.md section 2.4).
std::vector<std::string> halo_fields(roms_b200_state* h, int phase) {
  std::vector<std::string> v;
  auto T = [&](int tl) { for (int it = 0; it < h->cfg.NT; ++it) v.push_back("t" + std::to_string(tl) + "_" + std::to_string(it)); };
  const std::string nn = std::to_string(h->nnew), kn = std::to_string(h->knew), kr = std::to_string(h->krhs);
  switch (phase) {
    case ROMS_B200_SET_MASSFLUX: v = {"Huon", "Hvom"}; break;                       // set_massflux.F:177
    case ROMS_B200_RHO_EOS:                                                         // rho_eos.F:489-526
      v = {"rho", "pden", "rhoA", "rhoS"};
      if (h->cfg.bv_frequency) v.push_back("bvf");
      if (h->cfg.eos_tderivative) { v.push_back("alpha"); v.push_back("beta"); }
      break;
    case ROMS_B200_SET_VBC: v = {"bustr", "bvstr"}; break;                          // set_vbc.F:664
    case ROMS_B200_BULK_FLUX:                                                       // bulk_flux.F:949-960
      if (h->cfg.bulk_fluxes) v = {"lrflx", "lhflx", "shflx", "stflux_" + std::to_string(h->cfg.itemp - 1), "sustr", "svstr"};
      break;
    case ROMS_B200_BVF_MIX:                                                         // bvf_mix.F:133-160
      if (h->cfg.bvf_mixing) { v = {"Akv"}; for (int it = 0; it < (h->cfg.salinity ? 2 : 1); ++it) v.push_back("Akt_" + std::to_string(it)); }
      break;
    case ROMS_B200_LMD_VMIX:                                                        // lmd_skpp.F:641-647, lmd_vmix.F:644-655
      if (h->cfg.lmd_mixing) { v = {"hsbl", "Akv"}; for (int it = 0; it < (h->cfg.salinity ? 2 : 1); ++it) v.push_back("Akt_" + std::to_string(it)); }
      break;
    case ROMS_B200_ANA_VMIX:
      if (h->cfg.ana_vmix) { v = {"Akv"}; for (int it = 0; it < h->cfg.NT; ++it) v.push_back("Akt_" + std::to_string(it)); }
      break;
    case ROMS_B200_OMEGA: case ROMS_B200_OMEGA2: v = {"W"}; break;                  // omega.F:220
    case ROMS_B200_WVELOCITY: v = {"wvel"}; break;                                  // wvelocity.F:258
    case ROMS_B200_SET_ZETA: v = {"zeta1", "zeta2"}; break;                         // set_zeta.F:118
    case ROMS_B200_PRE_STEP3D: T(3); break;                                         // pre_step3d.F:1145
    case ROMS_B200_STEP2D:                                                          // step2d_LF_AM3.h:586-2519
      if (h->iif > h->nfast) v = {"Zt_avg1", "DU_avg1", "DV_avg1"};                 // :714
      else if (h->predictor) v = {"zeta" + kn, "rzeta" + kr, "ubar" + kn, "vbar" + kn};
      else v = {"zeta" + kn, "ubar" + kn, "vbar" + kn};
      break;
    case ROMS_B200_SET_DEPTH: v = {"z_w", "z_r", "Hz"}; break;                      // set_depth.F:269-279
    case ROMS_B200_SET_AVG:                                                         // set_avg.F:2347-2598 (when the window closes)
      if (avg_mode(h) & 4) {
        v = {"avgzeta", "avgu2d", "avgv2d", "avgu3d", "avgv3d", "avgrho", "avgw3d", "avgwvel"};
        for (int it = 0; it < h->cfg.NT; ++it) v.push_back("avgt_" + std::to_string(it));
      }
      break;
    case ROMS_B200_STEP3D_UV: v = {"u" + nn, "v" + nn, "Huon", "Hvom", "ubar1", "ubar2", "vbar1", "vbar2"}; break;   // step3d_uv.F:1464-1471
    case ROMS_B200_STEP3D_T: T(h->nnew); break;                                     // step3d_t.F:1626
    default: break;
  }
  return v;
}

// ---- two-stream schedule of a tile that sits in a ring -------------------------------------------------------------
// main stream : full-tile kernels of the phases that need no exchange, and the INTERIOR part (all but the first / last
//               EDGE_W columns) of the phases whose output is exchanged
// comm stream : (high priority) the EDGE part of those phases, then their halo exchange
// An interior kernel never touches ghost columns, so it only waits for the previous edge kernel (ev_edge), not for the
// exchange that follows it: exchange n overlaps interior n AND interior n+1.  An edge kernel needs the ghosts of exchange
// n-1 (same stream) and the interior results so far (ev_main).  Full-tile kernels wait for the latest exchange (ev_halo).
// Every point is still computed exactly once from the same inputs, so results do not depend on the split.
void join_halo(roms_b200_state* h) {
  if (h->halo_pending) { cudaStreamWaitEvent(h->stream, h->ev_halo, 0); h->halo_pending = false; h->edge_pending = false; }
}

template <class F>
void launch_full(roms_b200_state* h, F fn) {
  join_halo(h);
  fn(h->par, h->stream);
}

// Launch `fn` over the tile and refresh the xi-ghost columns of the fields `phase` produces (mp_exchange2d/3d/4d).
// in_kernel_exchange: the kernels fn launches move the halo themselves (fused step2d exchange); only the split schedule is kept
template <class F>
int launch_with_halo(roms_b200_state* h, int phase, F fn, bool in_kernel_exchange = false) {
  const Par& p = h->par;
  if (!h->halo) { fn(p, h->stream); return NoError; }
  std::vector<std::string> names = halo_fields(h, phase);
  const int ni = p.Iend - p.Istr + 1;
  if (names.empty() || !h->overlap || ni < 4 * EDGE_W) {
    launch_full(h, fn);
    return in_kernel_exchange ? (int)NoError : halo_exchange(h, names, h->stream);
  }
  if (in_kernel_exchange) names.clear();
  if (h->edge_pending) { cudaStreamWaitEvent(h->stream, h->ev_edge, 0); h->edge_pending = false; }
  cudaEventRecord(h->ev_main, h->stream);
  cudaStreamWaitEvent(h->comm_stream, h->ev_main, 0);
  Par e = p; e.gap_at = EDGE_W; e.gap_len = ni - 2 * EDGE_W;
  fn(e, h->comm_stream);
  cudaEventRecord(h->ev_edge, h->comm_stream);
  const int rc = halo_exchange(h, names, h->comm_stream);
  cudaEventRecord(h->ev_halo, h->comm_stream);
  Par q = p; q.Istr = p.Istr + EDGE_W; q.Iend = p.Iend - EDGE_W;
  fn(q, h->stream);
  h->edge_pending = true; h->halo_pending = true;
  h->launches += 1;
  return rc;
}

// Ring tiles: a phase whose output at a ghost column only depends on inputs that are valid there can compute its ghost columns
// itself (gw columns west of Istr, ge east of Iend) instead of exchanging them -- the same expressions on bitwise identical
// inputs give the neighbour's bits.  Used for the two 2-D phases whose exchange is not hidden behind interior work: bulk_flux
// (0.059 -> 0.037 ms on a 256-column tile) and set_vbc (0.017 -> 0.010).  For the 3-D phases (set_massflux, rho_eos, omega,
// set_depth: measured) the edge-first overlap already hides the exchange, and a whole-tile launch that has to wait for the
// previous phase's exchange is slower.  Valid input ranges: the device keeps three ghost columns west (the eastward exchanges
// move three) and two east of the tile.
template <class F>
int launch_ghosts(roms_b200_state* h, int gw, int ge, F fn) {
  join_halo(h);
  if (h->edge_pending) { cudaStreamWaitEvent(h->stream, h->ev_edge, 0); h->edge_pending = false; }
  Par q = h->par;
  q.Istr -= gw; q.Iend += ge;
  fn(q, h->stream);
  return NoError;
}
bool ghost_compute(const roms_b200_state* h) { return h->halo && h->ghost_compute; }

bool fused_tmix(const roms_b200_state* h) { return h->in_step && h->fuse_phases && !h->cfg.mix_geo_ts; }

// The 2-D time-index machine of LOOP_2D (main3d.F:592-700) for one call sequence that starts from (indx1, predictor = 0):
// fn(call number) is invoked once per step2d call with h->iif/kstp/krhs/knew/predictor set for it.
template <class F>
int loop2d_machine(roms_b200_state* h, F fn) {
  for (int my_iif = 1; my_iif <= h->nfast + 1; ++my_iif) {
    const int next_indx1 = 3 - h->indx1;
    if (!h->predictor && my_iif <= h->nfast + 1) {
      h->predictor = 1; h->iif = my_iif;
      h->kstp = (h->iif == 1) ? h->indx1 : 3 - h->indx1;
      h->knew = 3; h->krhs = h->indx1;
    }
    if (my_iif <= h->nfast + 1) { if (fn()) return FatalError; }
    if (h->predictor) {
      h->predictor = 0; h->knew = next_indx1; h->kstp = 3 - h->knew; h->krhs = 3;
      if (h->iif < h->nfast + 1) h->indx1 = next_indx1;
    }
    if (h->iif < h->nfast + 1) { if (fn()) return FatalError; }
  }
  return NoError;
}

// Call tables of the persistent loop kernel for a loop that starts with indx1 = 1 and 2 (k_step2d_loop.cu).  Built outside
// any graph capture (roms_b200_set_weights); with_ring: the fused halo exchange pushes in calls 1..2*nfast and pulls in calls
// 2..2*nfast+1.
int build_loop_tables(roms_b200_state* h) {
  const int ncall = 2 * h->nfast + 1;
  const int s_indx1 = h->indx1, s_iif = h->iif, s_kstp = h->kstp, s_krhs = h->krhs, s_knew = h->knew, s_pred = h->predictor;
  for (int start = 1; start <= 2; ++start) {
    std::vector<LoopStep> tab;
    h->indx1 = start; h->predictor = 0;
    loop2d_machine(h, [&]() {
      fill_par(h);
      const Par& p = h->par;
      LoopStep st;
      std::memset(&st, 0, sizeof(st));
      st.iif = p.iif; st.kstp = p.kstp; st.krhs = p.krhs; st.knew = p.knew; st.ptsk = p.ptsk; st.predictor = p.predictor;
      st.w1_m1 = p.w1_m1; st.w2_0 = p.w2_0; st.w2_p1 = p.w2_p1;
      const int c = (int)tab.size() + 1;
      st.send = c <= ncall - 1; st.recv = c >= 2;
      if (c >= 2) { const LoopStep& pr = tab.back(); st.nrecv = pr.predictor ? 4 : 3; st.rk = pr.knew; st.rr = pr.krhs; }
      tab.push_back(st);
      return 0;
    });
    if ((int)tab.size() != ncall) return FatalError;
    if (h->d_loop_tab[start]) { cudaFree(h->d_loop_tab[start]); h->d_loop_tab[start] = nullptr; }
    CK(cudaMalloc(&h->d_loop_tab[start], ncall * sizeof(LoopStep)));
    CK(cudaMemcpy(h->d_loop_tab[start], tab.data(), ncall * sizeof(LoopStep), cudaMemcpyHostToDevice));
  }
  h->indx1 = s_indx1; h->iif = s_iif; h->kstp = s_kstp; h->krhs = s_krhs; h->knew = s_knew; h->predictor = s_pred;
  return NoError;
}

// LOOP_2D as one persistent kernel when every CTA of the tile can be resident at once.  Returns -1 when not applicable.
int try_step2d_loop_kernel(roms_b200_state* h) {
  if (h->cfg.uv_adv == 3) return -1;                         // UV_C2ADVECTION: per-call kernels (the loop kernel carries the default fluxes only)
  if (!h->loop_kernel || h->predictor != 0 || h->indx1 < 1 || h->indx1 > 2 || !h->d_loop_tab[h->indx1] || !h->d_loop_flags) return -1;
  fill_par(h);
  // every tile of a ring must take the same decision (the tail of the exchange protocol differs): decide on the widest tile
  Par widest = h->par;
  widest.Iend = widest.Istr + (h->cfg.Lm + h->cfg.NtileI - 1) / h->cfg.NtileI - 1;
  const int nctas = step2d_loop_ctas(widest);
  if (nctas <= 0 || nctas > 4096) return -1;
  Xchg x;
  std::memset(&x, 0, sizeof(x));
  if (h->halo && !fused_xchg_fill(h, x)) return -1;          // a ring without the NVLink peer path: per-call launches + NCCL
  join_halo(h);
  if (h->edge_pending) { cudaStreamWaitEvent(h->stream, h->ev_edge, 0); h->edge_pending = false; }
  LoopCtl ctl;
  ctl.steps = h->d_loop_tab[h->indx1]; ctl.ncall = 2 * h->nfast + 1; ctl.nsend = h->halo ? 2 * h->nfast : 0;
  ctl.flags = h->d_loop_flags; ctl.base = h->d_loop_base; ctl.err = h->d_err; ctl.timeout_ns = (long long)(h->halo_timeout_s * 1e9);
  if (!launch_step2d_loop(h->par, h->fl, h->stream, h->halo ? &x : nullptr, ctl)) { cudaGetLastError(); return -1; }
  h->launches += 1;
  loop2d_machine(h, []() { return 0; });                     // leave the host's copy of the index machine where the loop ends
  if (h->halo) { if (halo_exchange(h, {"Zt_avg1", "DU_avg1", "DV_avg1"}, h->stream)) return FatalError; }   // step2d_LF_AM3.h:714
  return NoError;
}

// Lazy handles (per-routine calls): the optional arrays of a routine that the caller did not hand over still need storage --
// rho_eos writes bvf / alpha / beta, pre_step3d reads srflx / Jwtype / ghats (zeros: no shortwave, no nonlocal transport).
int ensure_optional(roms_b200_state* h, int phase) {
  if (!h->lazy) return NoError;
  std::vector<std::string> names;
  if (phase == ROMS_B200_RHO_EOS) names = {"bvf", "alpha", "beta"};
  else if (phase == ROMS_B200_SET_VBC) { if (h->cfg.uv_qdrag == 2) names = {"ZoBot", "z_r", "z_w"}; if (h->cfg.limit_bstress || h->cfg.scorrection) names.push_back("Hz"); }
  else if (phase == ROMS_B200_BULK_FLUX) names = {"lrflx", "lhflx", "shflx", "sustr", "svstr", "stflux_" + std::to_string(h->cfg.itemp - 1)};
  else if (phase == ROMS_B200_LMD_VMIX) { names = {"hsbl", "ksbl", "Akv"}; for (int it = 0; it < h->cfg.NT; ++it) { names.push_back("ghats_" + std::to_string(it)); names.push_back("Akt_" + std::to_string(it)); } }
  else if (phase == ROMS_B200_PRE_STEP3D) { names = {"srflx", "Jwtype", "z_w"}; for (int it = 0; it < h->cfg.NT; ++it) names.push_back("ghats_" + std::to_string(it)); }
  for (const std::string& n : names) {
    auto it = h->reg.find(n);
    if (it != h->reg.end()) { const int rc = materialize(h, it->second); if (rc) return rc; }
  }
  return NoError;
}

int run_phase_async(roms_b200_state* h, int phase) {
  { const int rc = ensure_optional(h, phase); if (rc) return rc; }
  fill_par(h);
  const Par& p = h->par; const Flds& f = h->fl; cudaStream_t s = h->stream;
  PhaseTimer pt(h, phase);
  int rc = NoError;
  switch (phase) {
    case ROMS_B200_SET_DATA: break;
    case ROMS_B200_SET_MASSFLUX:                                                    // Huon(i) needs Hz(i-1)
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_set_massflux(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_RHO_EOS:                                                         // column-local
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_rho_eos(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_SET_VBC:                                                         // bustr(i) needs v(i-1), bvstr(i) needs u(i+1)
      if (ghost_compute(h)) rc = launch_ghosts(h, 2, 1, [&](const Par& q, cudaStream_t st) { launch_set_vbc(q, f, st); });
      else rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_set_vbc(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_ANA_VMIX:
      if (h->cfg.ana_vmix) { rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_ana_vmix(q, f, st); }); h->launches += 1; }
      break;
    case ROMS_B200_BULK_FLUX:
      if (!h->cfg.bulk_fluxes) return ConfigError;
      // the stresses at u / v points need the rho-point values of the neighbouring column: one launch over the whole tile
      // in a ring the fluxes of the ghost columns are computed locally: rho points Istr-2 .. Iend+2, stresses Istr-1 .. Iend+2
      if (ghost_compute(h)) rc = launch_ghosts(h, 1, 2, [&](const Par& q, cudaStream_t st) { launch_bulk_flux(q, f, st); });
      else {
        launch_full(h, [&](const Par& q, cudaStream_t st) { launch_bulk_flux(q, f, st); });
        if (h->halo) rc = halo_exchange(h, halo_fields(h, phase), h->stream);
      }
      h->launches += 2; break;
    case ROMS_B200_BVF_MIX:
      if (!h->cfg.bvf_mixing) return ConfigError;
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_bvf_mix(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_LMD_VMIX:
      if (!h->cfg.lmd_mixing) return ConfigError;
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_lmd_vmix(q, f, st); });
      h->launches += (p.Iend == p.Lm) ? 2 : 1; break;
    case ROMS_B200_OMEGA: case ROMS_B200_OMEGA2:                                    // W(i) needs Huon(i+1); read at i-2 .. i+1 (rhs3d)
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_omega(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_WVELOCITY:
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_wvelocity(q, f, h->nstp, st); }); h->launches += 1; break;
    case ROMS_B200_SET_ZETA:
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_set_zeta(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_PRE_STEP3D:
      h->par.fuse_tmix = fused_tmix(h) ? 1 : 0;
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_pre_step3d_t(q, f, st); });
      h->par.fuse_tmix = 0;
      launch_full(h, [&](const Par& q, cudaStream_t st) { launch_pre_step3d_uv(q, f, st); });
      h->launches += 2; break;
    case ROMS_B200_PRSGRD: launch_full(h, [&](const Par& q, cudaStream_t st) { launch_prsgrd(q, f, h->cfg.dj_gradps, st); }); h->launches += h->cfg.dj_gradps == 1 ? 2 : 1; break;
    case ROMS_B200_T3DMIX:
      if (fused_tmix(h)) break;                                                     // already applied by pre_step3d_t
      if (h->cfg.mix_geo_ts) {
        if (!h->all_diff2_zero) { launch_full(h, [&](const Par& q, cudaStream_t st) { launch_t3dmix2_geo(q, f, st); }); h->launches += 1; }   // diff2 == 0: exact no-op
      } else { launch_full(h, [&](const Par& q, cudaStream_t st) { launch_t3dmix2_s(q, f, st); }); h->launches += 1; }
      break;
    case ROMS_B200_T3DMIX4:                                                         // rhs3d.F:89-97
      if (!h->cfg.ts_dif4) return ConfigError;
      launch_full(h, [&](const Par& q, cudaStream_t st) { launch_t3dmix4_s(q, f, st); }); h->launches += 1;
      break;
    case ROMS_B200_RHS3D: launch_full(h, [&](const Par& q, cudaStream_t st) { launch_rhs3d(q, f, st); }); h->launches += 1; break;
    case ROMS_B200_UV3DMIX: launch_full(h, [&](const Par& q, cudaStream_t st) { launch_uv3dmix2(q, f, st); }); h->launches += 1; break;
    case ROMS_B200_STEP2D: rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_step2d(q, f, st); }); h->launches += 1; break;
    case ROMS_B200_SET_DEPTH:                                                       // column-local
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_set_depth(q, f, st); });
      h->launches += 1; break;
    case ROMS_B200_SET_AVG: {
      const int m = avg_mode(h);
      if (!m) break;
      // KOUT = kstp, NOUT = nrhs (globaldefs.h:504-508)
      rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) {
        launch_set_avg(q, f, (m & 1) ? 0 : 1, (m & 4) ? 1 : 0, 1.0 / (double)h->navg, h->kstp, h->nrhs, st); });
      h->launches += 1; break;
    }
    case ROMS_B200_STEP3D_UV: rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_step3d_uv(q, f, st); }); h->launches += 2; break;
    case ROMS_B200_STEP3D_T: rc = launch_with_halo(h, phase, [&](const Par& q, cudaStream_t st) { launch_step3d_t(q, f, st); }); h->launches += 1; break;
    case ROMS_B200_DIAG: launch_full(h, [&](const Par& q, cudaStream_t st) { launch_diag(q, f, h->d_diag_partial, h->d_diag_out, h->knew, st); }); h->launches += 3; break;
    case ROMS_B200_STEP2D_LOOP: {
      // main3d.F:592-700
      { const int lk = try_step2d_loop_kernel(h); if (lk >= 0) { rc = lk; break; } }
      // Sub-step calls are numbered c = 1 .. 2*nfast+1.  On the NVLink peer path the xi-halo of calls 1 .. 2*nfast-1 travels
      // inside the kernels themselves (push at the end of call c, pull at the start of call c+1: dev.cuh Xchg); the last
      // corrector and the averaging-only call keep the stand-alone exchange so that the 3-D kernels find complete ghosts.
      Xchg xbase;
      const bool fused = fused_xchg_fill(h, xbase) && (p.Iend - p.Istr + 1) >= 8;
      int call = 0;
      double* prev_send[XF] = {nullptr, nullptr, nullptr, nullptr};
      int prev_n = 0;
      auto sub_step = [&]() {
        fill_par(h);
        h->launches += 1;
        ++call;
        if (!fused) return launch_with_halo(h, ROMS_B200_STEP2D, [&](const Par& q, cudaStream_t st) { launch_step2d(q, f, st); });
        Xchg x = xbase;
        x.recv = (call >= 2 && call <= 2 * h->nfast) ? 1 : 0;
        x.send = (call <= 2 * h->nfast - 1) ? 1 : 0;
        x.nrecv = prev_n;
        for (int q = 0; q < XF; ++q) x.recvf[q] = prev_send[q];
        x.nsend = h->predictor ? 4 : 3;
        prev_send[0] = f.zeta[h->knew]; prev_send[1] = f.ubar[h->knew]; prev_send[2] = f.vbar[h->knew]; prev_send[3] = f.rzeta[h->krhs];
        prev_n = x.nsend;
        if (x.send) {
          // mode 1: one full-tile launch, the exchange is inside the kernel; mode 2: keep the edge-first two-stream split (the
          // edge launch pulls and pushes, the interior launch of the next sub-step only waits for this edge launch)
          if (h->fused_mode == 2) {
            return launch_with_halo(h, ROMS_B200_STEP2D, [&](const Par& q, cudaStream_t st) { launch_step2d(q, f, st, q.gap_len > 0 || q.Istr == h->par.Istr ? &x : nullptr); }, true);
          }
          launch_full(h, [&](const Par& q, cudaStream_t st) { launch_step2d(q, f, st, &x); });
          return (int)NoError;
        }
        return launch_with_halo(h, ROMS_B200_STEP2D, [&](const Par& q, cudaStream_t st) { launch_step2d(q, f, st, x.recv ? &x : nullptr); });
      };
      if (loop2d_machine(h, sub_step)) return FatalError;
      break;
    }
    default: return ConfigError;
  }
  if (rc) return FatalError;
  if (phase == ROMS_B200_DIAG && h->halo) { if (halo_reduce_diag(h)) return FatalError; }
  if (h->profile == 1) join_halo(h); // per-phase timing: charge the exposed part of the exchange to its phase
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { std::fprintf(stderr, "roms_b200: launch error in phase %d: %s\n", phase, cudaGetErrorString(e)); return FatalError; }
  return NoError;
}

// main3d.F:189-917 for one step (without the first-step ini_zeta/ini_fields block and without get_data/output)
int step_phases_body(roms_b200_state* h, bool with_diag);
int step_phases(roms_b200_state* h, bool with_diag) {
  h->in_step = true;
  const int rc = step_phases_body(h, with_diag);
  h->in_step = false;
  return rc;
}
int step_phases_body(roms_b200_state* h, bool with_diag) {
  static const int seq1[] = {ROMS_B200_SET_MASSFLUX, ROMS_B200_RHO_EOS};
  for (int ph : seq1) { int rc = run_phase_async(h, ph); if (rc) return rc; }
  if (with_diag) { int rc = run_phase_async(h, ROMS_B200_DIAG); if (rc) return rc; }
  if (h->ev_forcing) {
    // forcing uploaded by roms_b200_step_forced on the copy stream: needed from set_vbc on.  Inside a graph capture this
    // becomes an event-wait node on the (external) event, evaluated at every replay.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(h->stream, &cs);
    cudaStreamWaitEvent(h->stream, h->ev_forcing, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : 0);
  }
  if (h->cfg.bulk_fluxes) { int rc = run_phase_async(h, ROMS_B200_BULK_FLUX); if (rc) return rc; }      // main3d.F:384-390
  { int rc = run_phase_async(h, ROMS_B200_SET_VBC); if (rc) return rc; }
  {                                                                                  // :464-470: ANA_VMIX, else LMD_MIXING, else BVF_MIXING
    const int ph = h->cfg.ana_vmix ? ROMS_B200_ANA_VMIX : h->cfg.lmd_mixing ? ROMS_B200_LMD_VMIX : h->cfg.bvf_mixing ? ROMS_B200_BVF_MIX : ROMS_B200_ANA_VMIX;
    int rc = run_phase_async(h, ph); if (rc) return rc;
  }
  { int rc = run_phase_async(h, ROMS_B200_OMEGA); if (rc) return rc; }
  if (h->cfg.wvelocity_every_step) { int rc = run_phase_async(h, ROMS_B200_WVELOCITY); if (rc) return rc; }
  static const int seq3[] = {ROMS_B200_SET_ZETA,  ROMS_B200_SET_AVG, ROMS_B200_PRE_STEP3D, ROMS_B200_PRSGRD,    ROMS_B200_T3DMIX,   ROMS_B200_RHS3D, ROMS_B200_UV3DMIX,
                             ROMS_B200_STEP2D_LOOP, ROMS_B200_SET_DEPTH, ROMS_B200_STEP3D_UV, ROMS_B200_OMEGA2, ROMS_B200_STEP3D_T};
  for (int ph : seq3) {
    int rc = run_phase_async(h, ph); if (rc) return rc;
    if (ph == ROMS_B200_T3DMIX && h->cfg.ts_dif4) { rc = run_phase_async(h, ROMS_B200_T3DMIX4); if (rc) return rc; }   // rhs3d.F:89-97
  }
  join_halo(h);                      // the step ends with both streams joined (also required to end a graph capture)
  mark_phase(h, -1);
  return NoError;
}

// One baroclinic step.  After the two start-up steps (AB3 / LF-AM3 start-up branches) the launch sequence of a step
// depends only on (nstp, indx1, with_diag): it is captured once per such state into a CUDA graph -- ~80 kernels on one
// GPU, ~400 kernel / NCCL / event nodes on two streams with a ring attached -- and replayed afterwards, which removes
// the host launch cost from the critical path (the barotropic sub-steps are ~10 us kernels on an 8-GPU tiling).
int one_step(roms_b200_state* h, bool with_diag) {
  h->nstp = 1 + ((h->iic - h->ntstart) % 2); h->nnew = 3 - h->nstp; h->nrhs = h->nstp;
  h->tdays = h->time / 86400.0;
  const bool steady = h->iic >= h->ntfirst + 2 && h->predictor == 0;
  if (h->use_graphs && steady && h->profile != 1) {
    const int key = h->nstp | (h->indx1 << 2) | ((with_diag ? 1 : 0) << 4) | ((h->profile == 2 ? 1 : 0) << 5) | (avg_mode(h) << 6);
    auto it = h->graphs.find(key);
    StepGraph* g = (it == h->graphs.end()) ? nullptr : (StepGraph*)it->second;
    if (!g) {
      const int s_indx1 = h->indx1, s_iif = h->iif, s_kstp = h->kstp, s_krhs = h->krhs, s_knew = h->knew, s_pred = h->predictor;
      const long long l0 = h->launches;
      cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
      bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
      h->n_ev_ph = 0;
      int rc = ok ? step_phases(h, with_diag) : FatalError;
      if (ok) ok = cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph != nullptr;
      if (ok && rc == NoError) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (!ok || rc != NoError) {
        // capture is not possible here (e.g. an NCCL build without graph support): go back to plain launches for good
        cudaGetLastError();
        std::fprintf(stderr, "roms_b200: CUDA graph capture of the time step failed; using stream launches\n");
        h->use_graphs = 0;
        h->indx1 = s_indx1; h->iif = s_iif; h->kstp = s_kstp; h->krhs = s_krhs; h->knew = s_knew; h->predictor = s_pred; h->launches = l0;
        const int rc2 = step_phases(h, with_diag);
        if (rc2) return rc2;
        h->iic += 1; h->time += h->cfg.dt;
        return NoError;
      }
      g = new StepGraph{exec, h->indx1, h->iif, h->kstp, h->krhs, h->knew, h->predictor, h->launches - l0, h->n_ev_ph};
      h->graphs[key] = g;
      h->launches = l0;
    }
    CK(cudaGraphLaunch(g->exec, h->stream));
    h->indx1 = g->indx1; h->iif = g->iif; h->kstp = g->kstp; h->krhs = g->krhs; h->knew = g->knew; h->predictor = g->predictor;
    h->launches += g->launches;
    h->iic += 1; h->time += h->cfg.dt;
    if (h->profile == 2) { h->n_ev_ph = g->n_marks; CK(cudaStreamSynchronize(h->stream)); collect_marks(h); }
    return NoError;
  }
  h->n_ev_ph = 0;
  const int rc = step_phases(h, with_diag);
  if (h->profile == 2 && !rc) { CK(cudaStreamSynchronize(h->stream)); collect_marks(h); }
  if (rc) return rc;
  h->iic += 1; h->time += h->cfg.dt;
  return NoError;
}

// true if [p, p+bytes) lies inside a host range the caller pinned with roms_b200_register_host
bool is_registered(roms_b200_state* h, const void* p, size_t bytes) {
  const char* a = (const char*)p;
  for (auto& r : h->host_pinned) if (a >= (const char*)r.first && a + bytes <= (const char*)r.first + r.second) return true;
  return false;
}

// Wait for the stream and read the sticky device error word back with it.  A halo wait that gave up (dev.cuh ll_wait) left
// garbage in ghost columns: exit_flag becomes 8 and stays 8 (mod_scalars.F:523-532 "fatal algorithm error"), which is what a
// Fortran host polling FoundError(exit_flag, NoError, ...) after each call needs in order to stop.
int sync_and_check(roms_b200_state* h) {
  CK(cudaMemcpyAsync(h->h_err, h->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return rbi::check_device_error(h);
}

int finish_diag(roms_b200_state* h, double* out12) {
  CK(cudaMemcpyAsync(h->h_diag_out, h->d_diag_out, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  { const int rc = sync_and_check(h); if (rc) return rc; }
  const double* d = h->h_diag_out;
  const double vol = d[2];
  out12[0] = d[0] / vol; out12[1] = d[1] / vol; out12[2] = out12[0] + out12[1]; out12[3] = vol;
  out12[4] = d[7]; out12[5] = d[4]; out12[6] = d[5]; out12[7] = d[6];
  out12[8] = d[11]; out12[9] = d[12]; out12[10] = d[9]; out12[11] = d[10];
  // blow-up detection: diag.F:506-538, limits mod_scalars.F:548-549
  const bool bad = !(out12[0] == out12[0]) || !(out12[1] == out12[1]) || std::abs(out12[0]) > 1e300 || std::abs(out12[1]) > 1e300;
  if (bad || d[7] > 20.0 || d[8] > 200.0) h->exit_flag = BlowUp;
  return h->exit_flag;
}

}  // namespace

namespace rbi {
int check_device_error(roms_b200_state* h) {
  if (h->h_err && *h->h_err != 0ULL) {
    if (h->exit_flag != FatalError) std::fprintf(stderr, "roms_b200: tile %d: a halo exchange timed out waiting for a neighbour; the state is invalid (exit_flag 8)\n", h->cfg.tile);
    h->exit_flag = FatalError;
  }
  return h->exit_flag == FatalError ? (int)FatalError : (int)NoError;
}
}  // namespace rbi

extern "C" {

int roms_b200_default_config(int app, int Lm, int Mm, int N, roms_b200_config* c) {
  if (!c) return InputError;
  std::memset(c, 0, sizeof(*c));
  c->NtileI = 1; c->NtileJ = 1; c->tile = 0; c->rho0 = 1025.0; c->g = 9.81; c->lambda = 1.0; c->itemp = 1; c->isalt = 2;
  c->dj_gradps = 1; c->wvelocity_every_step = 1; c->R0 = 1027.0; c->Tcoef = 1.7e-4; c->device = 0;
  if (app == ROMS_B200_APP_UPWELLING) {          // ROMS/Include/upwelling.h, ROMS/External/roms_upwelling.in
    c->Lm = 41; c->Mm = 80; c->N = 16; c->NT = 2; c->dt = 300.0; c->ndtfast = 30; c->salinity = 1; c->ana_vmix = 1;
    c->hadv = ROMS_B200_HADV_U3; c->vadv = ROMS_B200_VADV_C4; c->T0 = 14.0; c->S0 = 35.0; c->Scoef = 0.0;
    c->Akt_bak[0] = c->Akt_bak[1] = 1e-6; c->Akv_bak = 1e-5; c->gamma2 = 1.0; c->hc = 25.0;
  } else if (app == ROMS_B200_APP_SEAMOUNT) {    // seamount.h, roms_seamount.in
    c->Lm = 49; c->Mm = 48; c->N = 13; c->NT = 1; c->dt = 60.0; c->ndtfast = 20; c->mix_geo_ts = 1; c->uv_qdrag = 1;
    c->hadv = ROMS_B200_HADV_A4; c->vadv = ROMS_B200_VADV_A4; c->T0 = 10.0; c->S0 = 32.0; c->Scoef = 7.6e-4;
    c->Akt_bak[0] = c->Akt_bak[1] = 1e-6; c->Akv_bak = 1e-5; c->gamma2 = -1.0; c->hc = 100.0;
  } else if (app == ROMS_B200_APP_BENCHMARK) {   // benchmark.h grid/IC with the reduced physics set, roms_benchmark1.in
    c->Lm = 512; c->Mm = 64; c->N = 30; c->NT = 2; c->dt = 150.0; c->ndtfast = 20; c->nonlin_eos = 1; c->curvgrid = 1;
    c->uv_qdrag = 1; c->salinity = 1; c->hadv = ROMS_B200_HADV_U3; c->vadv = ROMS_B200_VADV_C4; c->T0 = 10.0; c->S0 = 35.0;
    c->Scoef = 7.6e-4; c->Akt_bak[0] = c->Akt_bak[1] = 1e-5; c->Akv_bak = 1e-4; c->gamma2 = 1.0; c->hc = 400.0;
  } else return ConfigError;
  if (Lm > 0) c->Lm = Lm;
  if (Mm > 0) c->Mm = Mm;
  if (N > 0) c->N = N;
  return NoError;
}

static const char* kBoundNames =
    "tile,Itile,Jtile,LBi,UBi,LBj,UBj,IminS,ImaxS,JminS,JmaxS,Istr,IstrB,IstrP,IstrR,IstrT,IstrM,IstrU,Iend,IendB,IendP,IendR,IendT,"
    "Jstr,JstrB,JstrP,JstrR,JstrT,JstrM,JstrV,Jend,JendB,JendP,JendR,JendT,Istrm3,Istrm2,Istrm1,IstrUm2,IstrUm1,Iendp1,Iendp2,Iendp2i,"
    "Iendp3,Jstrm3,Jstrm2,Jstrm1,JstrVm2,JstrVm1,Jendp1,Jendp2,Jendp2i,Jendp3,Western_Edge,Eastern_Edge,Southern_Edge,Northern_Edge";
const char* roms_b200_bounds_names(void) { return kBoundNames; }

int roms_b200_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, int distribute, int* o) {
  if (!o || Lm < 1 || Mm < 1 || NtileI < 1 || NtileJ < 1 || tile < 0 || tile >= NtileI * NtileJ) return InputError;
  Bounds b; make_bounds(Lm, Mm, NtileI, NtileJ, tile, distribute != 0, b);
  const int v[57] = {b.tile, b.Itile, b.Jtile, b.LBi, b.UBi, b.LBj, b.UBj, b.IminS, b.ImaxS, b.JminS, b.JmaxS, b.Istr, b.IstrB, b.IstrP, b.IstrR,
                     b.IstrT, b.IstrM, b.IstrU, b.Iend, b.IendB, b.IendP, b.IendR, b.IendT, b.Jstr, b.JstrB, b.JstrP, b.JstrR, b.JstrT, b.JstrM,
                     b.JstrV, b.Jend, b.JendB, b.JendP, b.JendR, b.JendT, b.Istrm3, b.Istrm2, b.Istrm1, b.IstrUm2, b.IstrUm1, b.Iendp1, b.Iendp2,
                     b.Iendp2i, b.Iendp3, b.Jstrm3, b.Jstrm2, b.Jstrm1, b.JstrVm2, b.JstrVm1, b.Jendp1, b.Jendp2, b.Jendp2i, b.Jendp3,
                     b.west, b.east, b.south, b.north};
  std::memcpy(o, v, sizeof(v));
  return NoError;
}

static int create_impl(const roms_b200_config* cfg, roms_b200_handle* out, bool lazy);
int roms_b200_create(const roms_b200_config* cfg, roms_b200_handle* out) { return create_impl(cfg, out, false); }
// internal (api_tile.cu): a handle whose fields get device storage when they are first uploaded
int roms_b200_create_lazy_(const roms_b200_config* cfg, roms_b200_handle* out) { return create_impl(cfg, out, true); }
static int create_impl(const roms_b200_config* cfg, roms_b200_handle* out, bool lazy) {
  if (!cfg || !out) return InputError;
  *out = nullptr;
  if (cfg->N < 4 || cfg->N > MAXN || cfg->NT < 1 || cfg->NT > MAXNT || cfg->Lm < 8 || cfg->Mm < 4) return ConfigError;
  if (cfg->NtileI < 1 || cfg->NtileJ < 1 || cfg->tile < 0 || cfg->tile >= cfg->NtileI * cfg->NtileJ) return ConfigError;
  if (cfg->uv_qdrag < 0 || cfg->uv_qdrag > 2) return ConfigError;
  // lmd_skpp reads bvf, alpha / beta, srflx and writes ghats: the switches that provide those arrays must be on with it
  if (cfg->lmd_mixing && !(cfg->bv_frequency && cfg->eos_tderivative && cfg->solar_source && cfg->lmd_nonlocal)) return ConfigError;
  if (cfg->bvf_mixing && !cfg->bv_frequency) return ConfigError;
  if (cfg->uv_adv < 0 || cfg->uv_adv > 3) return ConfigError;
  if (cfg->dj_gradps < 0 || cfg->dj_gradps > 3 || cfg->vtransform < 0 || cfg->vtransform > 2) return ConfigError;
  if (cfg->bodyforce && (cfg->levsfrc < 1 || cfg->levsfrc > cfg->N || cfg->levbfrc < 1 || cfg->levbfrc > cfg->N)) return ConfigError;
  if (cfg->scorrection < 0 || cfg->scorrection > 2 || (cfg->scorrection && !(cfg->salinity && cfg->NT >= 2))) return ConfigError;
  if (cfg->ts_dif4 && cfg->mix_geo_ts) return ConfigError;                          // t3dmix4_geo.h is not built
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    std::fprintf(stderr, "roms_b200: no CUDA device; this library has no CPU fallback\n");
    return FatalError;
  }
  CK(cudaSetDevice(cfg->device));
  roms_b200_state* h = new roms_b200_state();
  h->lazy = lazy;
#define CKD(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      std::fprintf(stderr, "roms_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      roms_b200_destroy(h);                                                                              \
      return FatalError;                                                                                 \
    }                                                                                                    \
  } while (0)
  h->cfg = *cfg;
  // NtileI x NtileJ partitions (get_bounds.F:985-1004; the shipped roms_benchmark3.in is 2 x 2): a handle owns a whole xi-COLUMN
  // of tiles, i.e. tiles Itile + Jtile*NtileI for every Jtile.  The partition only distributes loop ranges -- results do not
  // depend on it (ROMS/Bin/verify.sh:985-1045) -- so the device treats the column as one tile; any tile number of the column
  // selects it.  (A distributed-memory host that splits eta between ranks is not supported: its arrays hold only a J sub-range.)
  h->itile = cfg->tile % cfg->NtileI;
  make_bounds(cfg->Lm, cfg->Mm, cfg->NtileI, 1, h->itile, cfg->NtileI > 1, h->b);
  const Bounds& b = h->b;
  h->ni = b.UBi - b.LBi + 1; h->nj = b.UBj - b.LBj + 1;
  h->LBi_dev = b.Istr - 3;                       // == b.LBi on the western tile; one extra ghost column elsewhere
  h->ni_dev = b.UBi - h->LBi_dev + 1;
  // origin shift: put i = Istr on a 128-byte boundary
  h->ioff = (16 - ((b.Istr - h->LBi_dev) % 16)) % 16;
  Par& p = h->par;
  std::memset(&p, 0, sizeof(p));
  p.Lm = cfg->Lm; p.Mm = cfg->Mm; p.N = cfg->N; p.NT = cfg->NT;
  p.P = ((h->ioff + h->ni_dev + 15) / 16) * 16; p.PL = p.P * h->nj;
  if ((long long)p.PL * (cfg->N + 1) >= (1LL << 31)) { delete h; return ConfigError; }
  p.LBi = h->LBi_dev; p.UBi = b.UBi; p.LBj = b.LBj; p.UBj = b.UBj;
  p.Istr = b.Istr; p.Iend = b.Iend; p.Jstr = b.Jstr; p.Jend = b.Jend; p.IstrU = b.IstrU; p.JstrV = b.JstrV; p.JstrR = b.JstrR; p.JendR = b.JendR;
  p.Jstrm1 = b.Jstrm1; p.Jendp1 = b.Jendp1; p.Jendp2 = b.Jendp2; p.JstrVm1 = b.JstrVm1; p.JstrVm2 = b.JstrVm2;
  p.ew_wrap = (cfg->NtileI == 1) ? 1 : 0;
  p.gap_at = 0x7fffffff; p.gap_len = 0;
  p.nonlin_eos = cfg->nonlin_eos; p.curvgrid = cfg->curvgrid; p.uv_qdrag = cfg->uv_qdrag; p.salinity = cfg->salinity;
  p.hadv = cfg->hadv; p.vadv = cfg->vadv; p.itemp = cfg->itemp; p.isalt = cfg->isalt;
  p.bv_frequency = cfg->bv_frequency; p.eos_tderivative = cfg->eos_tderivative; p.solar_source = cfg->solar_source; p.lmd_nonlocal = cfg->lmd_nonlocal;
  p.bulk_fluxes = cfg->bulk_fluxes; p.lmd_mixing = cfg->lmd_mixing; p.uv_adv = cfg->uv_adv; p.limit_bstress = cfg->limit_bstress;
  p.nospl_vvisc = cfg->nospl_vvisc ? 1 : 0; p.nospl_vdiff = cfg->nospl_vdiff ? 1 : 0;
  p.qcorrection = cfg->qcorrection ? 1 : 0; p.limit_stflx_cooling = cfg->limit_stflx_cooling ? 1 : 0; p.scorrection = cfg->scorrection;
  p.pad2_ = 0; p.Tnudg_salt = cfg->Tnudg_salt;
  p.bodyforce = cfg->bodyforce ? 1 : 0; p.levsfrc = cfg->levsfrc; p.levbfrc = cfg->levbfrc; p.vtransform = (cfg->vtransform == 1) ? 1 : 2;
  p.atm_press = cfg->atm_press ? 1 : 0; p.pad4_ = 0;
  p.blk_ZQ = cfg->blk_ZQ > 0.0 ? cfg->blk_ZQ : 10.0; p.blk_ZT = cfg->blk_ZT > 0.0 ? cfg->blk_ZT : 10.0; p.blk_ZW = cfg->blk_ZW > 0.0 ? cfg->blk_ZW : 10.0;
  p.dt = cfg->dt; p.g = cfg->g; p.rho0 = cfg->rho0; p.R0 = cfg->R0; p.T0 = cfg->T0; p.S0 = cfg->S0; p.Tcoef = cfg->Tcoef; p.Scoef = cfg->Scoef;
  p.gamma2 = cfg->gamma2; p.lambda = cfg->lambda; p.hc = cfg->hc; p.Akv_bak = cfg->Akv_bak;
  for (int it = 0; it < MAXNT; ++it) p.Akt_bak[it] = cfg->Akt_bak[it];
  h->dtfast = cfg->dt / (double)cfg->ndtfast;
  std::memset(h->phase_ms, 0, sizeof(h->phase_ms));
  CKD(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKD(cudaEventCreate(&h->ev0)); CKD(cudaEventCreate(&h->ev1));
  Flds& f = h->fl;
  std::memset(&f, 0, sizeof(f));
  const int N = cfg->N;
  int rc = 0;
#define A2(name) rc |= alloc_field(h, #name, &f.name, 0, 1)
#define A3(name, k0, nk) rc |= alloc_field(h, #name, &f.name, k0, nk)
  A2(h); A2(f); A2(pm); A2(pn); A2(om_r); A2(on_r); A2(om_u); A2(on_u); A2(om_v); A2(on_v); A2(om_p); A2(on_p); A2(omn); A2(fomn);
  A2(pmon_r); A2(pnom_r); A2(pmon_u); A2(pnom_u); A2(pmon_v); A2(pnom_v); A2(pmon_p); A2(pnom_p); A2(dndx); A2(dmde); A2(rdrag); A2(rdrag2);
  A2(visc2_r); A2(visc2_p); A2(Zt_avg1); A2(DU_avg1); A2(DU_avg2); A2(DV_avg1); A2(DV_avg2); A2(rufrc); A2(rvfrc); A2(rhoA); A2(rhoS);
  A2(sustr); A2(svstr); A2(bustr); A2(bvstr);
  A3(rho, 1, N); A3(pden, 1, N); A3(Hz, 1, N); A3(z_r, 1, N); A3(Huon, 1, N); A3(Hvom, 1, N); A3(W, 0, N + 1); A3(wvel, 0, N + 1);
  A3(z_w, 0, N + 1); A3(Akv, 0, N + 1); A3(P3, 1, N);
#undef A2
#undef A3
  for (int k = 1; k <= 3; ++k) {
    rc |= alloc_field(h, "zeta" + std::to_string(k), &f.zeta[k], 0, 1);
    rc |= alloc_field(h, "ubar" + std::to_string(k), &f.ubar[k], 0, 1);
    rc |= alloc_field(h, "vbar" + std::to_string(k), &f.vbar[k], 0, 1);
  }
  for (int k = 1; k <= 2; ++k) {
    rc |= alloc_field(h, "rzeta" + std::to_string(k), &f.rzeta[k], 0, 1);
    rc |= alloc_field(h, "rubar" + std::to_string(k), &f.rubar[k], 0, 1);
    rc |= alloc_field(h, "rvbar" + std::to_string(k), &f.rvbar[k], 0, 1);
    rc |= alloc_field(h, "u" + std::to_string(k), &f.u[k], 1, N);
    rc |= alloc_field(h, "v" + std::to_string(k), &f.v[k], 1, N);
    rc |= alloc_field(h, "ru" + std::to_string(k), &f.ru[k], 0, N + 1);
    rc |= alloc_field(h, "rv" + std::to_string(k), &f.rv[k], 0, N + 1);
  }
  for (int it = 0; it < cfg->NT; ++it) {
    const std::string s = std::to_string(it);
    for (int k = 1; k <= 3; ++k) rc |= alloc_field(h, "t" + std::to_string(k) + "_" + s, &f.t[k][it], 1, N);
    rc |= alloc_field(h, "Akt_" + s, &f.Akt[it], 0, N + 1);
    rc |= alloc_field(h, "diff2_" + s, &f.diff2[it], 0, 1);
    if (cfg->ts_dif4) rc |= alloc_field(h, "diff4_" + s, &f.diff4[it], 0, 1);
    rc |= alloc_field(h, "stflx_" + s, &f.stflx[it], 0, 1);
    rc |= alloc_field(h, "btflx_" + s, &f.btflx[it], 0, 1);
    rc |= alloc_field(h, "stflux_" + s, &f.stflux[it], 0, 1);
    rc |= alloc_field(h, "btflux_" + s, &f.btflux[it], 0, 1);
  }
  if (cfg->uv_qdrag == 2) rc |= alloc_field(h, "ZoBot", &f.ZoBot, 0, 1);
  if (cfg->qcorrection) { rc |= alloc_field(h, "sst", &f.sst, 0, 1); rc |= alloc_field(h, "dqdt", &f.dqdt, 0, 1); }
  if (cfg->scorrection) rc |= alloc_field(h, "sss", &f.sss, 0, 1);
  if (cfg->bv_frequency) rc |= alloc_field(h, "bvf", &f.bvf, 0, N + 1);
  if (cfg->eos_tderivative) { rc |= alloc_field(h, "alpha", &f.alpha, 0, 1); rc |= alloc_field(h, "beta", &f.beta, 0, 1); }
  if (cfg->solar_source || cfg->bulk_fluxes) rc |= alloc_field(h, "srflx", &f.srflx, 0, 1);
  if (cfg->solar_source) rc |= alloc_field(h, "Jwtype", &f.Jwtype, 0, 1);
  if (cfg->atm_press && !cfg->bulk_fluxes) rc |= alloc_field(h, "Pair", &f.Pair, 0, 1);
  if (cfg->bulk_fluxes) {
#define A2(name) rc |= alloc_field(h, #name, &f.name, 0, 1)
    A2(Uwind); A2(Vwind); A2(Tair); A2(Pair); A2(Hair); A2(rain); A2(cloud); A2(lrflx); A2(lhflx); A2(shflx); A2(Taux); A2(Tauy);
#undef A2
  }
  if (cfg->lmd_mixing) { rc |= alloc_field(h, "hsbl", &f.hsbl, 0, 1); rc |= alloc_field(h, "ksbl", &f.ksbl, 0, 1); }
  if (cfg->lmd_nonlocal) for (int it = 0; it < cfg->NT; ++it) rc |= alloc_field(h, "ghats_" + std::to_string(it), &f.ghats[it], 0, N + 1);
  if (rc) { roms_b200_destroy(h); return FatalError; }
  double* sc = nullptr;
  CKD(cudaMalloc(&sc, 4 * (MAXN + 1) * sizeof(double)));
  CKD(cudaMemsetAsync(sc, 0, 4 * (MAXN + 1) * sizeof(double), h->stream));
  h->allocs.push_back(sc);
  f.sc_r = sc; f.Cs_r = sc + (MAXN + 1); f.sc_w = sc + 2 * (MAXN + 1); f.Cs_w = sc + 3 * (MAXN + 1);
  CKD(cudaMalloc(&h->d_diag_partial, (size_t)diag_partial_doubles(p) * sizeof(double)));
  h->allocs.push_back(h->d_diag_partial);
  CKD(cudaMalloc(&h->d_diag_out, 16 * sizeof(double)));
  h->allocs.push_back(h->d_diag_out);
  CKD(cudaMallocHost(&h->h_diag_out, 16 * sizeof(double)));
  CKD(cudaMalloc(&h->d_err, 64));
  h->allocs.push_back(h->d_err);
  CKD(cudaMemsetAsync(h->d_err, 0, 64, h->stream));
  CKD(cudaMalloc(&h->d_loop_flags, 4096 * sizeof(unsigned long long)));
  h->allocs.push_back(h->d_loop_flags);
  CKD(cudaMemsetAsync(h->d_loop_flags, 0, 4096 * sizeof(unsigned long long), h->stream));
  CKD(cudaMalloc(&h->d_loop_base, 64));
  h->allocs.push_back(h->d_loop_base);
  CKD(cudaMemsetAsync(h->d_loop_base, 0, 64, h->stream));
  CKD(cudaMallocHost(&h->h_err, 64));
  std::memset(h->h_err, 0, 64);
  CKD(cudaStreamSynchronize(h->stream));
#undef CKD
  *out = h;
  return NoError;
}

int roms_b200_destroy(roms_b200_handle h) {
  if (!h) return NoError;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  drop_graphs(h);
  halo_destroy(h);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); cudaEventDestroy(h->ev_forcing); cudaEventDestroy(h->ev_step_in); }
  for (void* p : h->allocs) cudaFree(p);
  for (LoopStep* t : h->d_loop_tab) if (t) cudaFree(t);
  if (h->h_diag_out) cudaFreeHost(h->h_diag_out);
  if (h->h_err) cudaFreeHost(h->h_err);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  for (auto& r : h->host_pinned) cudaHostUnregister(r.first);
  h->host_pinned.clear();
  for (cudaEvent_t e : h->ev_ph) if (e) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return NoError;
}

int roms_b200_array_bounds(roms_b200_handle h, int* o) {
  if (!h || !o) return InputError;
  o[0] = h->b.LBi; o[1] = h->b.UBi; o[2] = h->b.LBj; o[3] = h->b.UBj;
  return NoError;
}

static int xfer(roms_b200_handle h, const char* name, double* host, size_t n, bool up) {
  if (!h || !name || !host) return InputError;
  auto it = h->reg.find(name);
  if (it == h->reg.end()) { std::fprintf(stderr, "roms_b200: unknown field '%s'\n", name); return InputError; }
  FieldInfo& fi = it->second;
  const size_t want = (size_t)h->ni * h->nj * fi.nk;
  if (n != want) { std::fprintf(stderr, "roms_b200: field '%s' expects %zu doubles, got %zu\n", name, want, n); return InputError; }
  CK(cudaSetDevice(h->cfg.device));
  if (!fi.base) { if (!up) return InputError; const int rc = materialize(h, fi); if (rc) return rc; }   // lazy handle: first hand-over
  double* dev = fi.base + h->ioff + (h->b.LBi - h->LBi_dev);
  const size_t dp = (size_t)h->par.P * sizeof(double), sp = (size_t)h->ni * sizeof(double);
  if (up) CK(cudaMemcpy2DAsync(dev, dp, host, sp, sp, (size_t)h->nj * fi.nk, cudaMemcpyHostToDevice, h->stream));
  else CK(cudaMemcpy2DAsync(host, sp, dev, dp, sp, (size_t)h->nj * fi.nk, cudaMemcpyDeviceToHost, h->stream));
  if (up && h->halo) { if (halo_exchange(h, {std::string(name)}, h->stream)) return FatalError; }
  { const int rc = sync_and_check(h); if (rc) return rc; }
  if (up && std::strncmp(name, "diff2_", 6) == 0) {
    bool z = true;
    for (size_t q = 0; q < n; ++q) if (host[q] != 0.0) { z = false; break; }
    if (!z) h->all_diff2_zero = false;
  }
  return NoError;
}
int roms_b200_set_field(roms_b200_handle h, const char* name, const double* host, size_t n) { return xfer(h, name, const_cast<double*>(host), n, true); }
int roms_b200_get_field(roms_b200_handle h, const char* name, double* host, size_t n) { return xfer(h, name, host, n, false); }

int roms_b200_field_levels(roms_b200_handle h, const char* name, int* LBk, int* nk) {
  if (!h || !name) return InputError;
  auto it = h->reg.find(name);
  if (it == h->reg.end()) return InputError;
  if (LBk) *LBk = it->second.LBk;
  if (nk) *nk = it->second.nk;
  return NoError;
}

int roms_b200_set_scoord(roms_b200_handle h, int which, const double* v, int n) {
  if (!h || !v || which < 0 || which > 3 || n < h->cfg.N + 1 || n > MAXN + 1) return InputError;
  double* dst = which == 0 ? h->fl.sc_r : which == 1 ? h->fl.Cs_r : which == 2 ? h->fl.sc_w : h->fl.Cs_w;
  CK(cudaMemcpyAsync(dst, v, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NoError;
}

int roms_b200_set_weights(roms_b200_handle h, int nfast, const double* w1, const double* w2, int n) {
  if (!h || !w1 || !w2 || nfast < 1 || n < nfast + 2) return InputError;
  h->nfast = nfast; h->w1.assign(w1, w1 + n); h->w2.assign(w2, w2 + n);
  drop_graphs(h);
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  return build_loop_tables(h);
}

int roms_b200_set_indices(roms_b200_handle h, const int* v, const double* tm) {
  if (!h || !v || !tm) return InputError;
  h->iic = v[0]; h->ntstart = v[1]; h->ntfirst = v[2]; h->nstp = v[3]; h->nnew = v[4]; h->nrhs = v[5]; h->iif = v[6]; h->indx1 = v[7];
  h->kstp = v[8]; h->krhs = v[9]; h->knew = v[10]; h->predictor = v[11]; h->exit_flag = v[12]; h->time = tm[0]; h->tdays = tm[1];
  drop_graphs(h);
  return NoError;
}
int roms_b200_get_indices(roms_b200_handle h, int* v, double* tm) {
  if (!h || !v || !tm) return InputError;
  const int w[13] = {h->iic, h->ntstart, h->ntfirst, h->nstp, h->nnew, h->nrhs, h->iif, h->indx1, h->kstp, h->krhs, h->knew, h->predictor, h->exit_flag};
  std::memcpy(v, w, sizeof(w)); tm[0] = h->time; tm[1] = h->tdays;
  return NoError;
}

int roms_b200_run_phase(roms_b200_handle h, int phase) {
  if (!h) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  int rc = run_phase_async(h, phase);
  join_halo(h);
  if (rc) return rc;
  return sync_and_check(h);
}

int roms_b200_main3d_step(roms_b200_handle h, int nsteps) {
  if (!h || nsteps < 0) return InputError;
  if (h->nfast < 1) { std::fprintf(stderr, "roms_b200: set_weights must be called before stepping\n"); return ConfigError; }
  if (h->exit_flag == FatalError) return FatalError;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) { int rc = one_step(h, false); if (rc) return rc; }
  CK(cudaEventRecord(h->ev1, h->stream));
  return NoError;
}

int roms_b200_sync(roms_b200_handle h) {
  if (!h) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  return sync_and_check(h);
}

int roms_b200_set_avg(roms_b200_handle h, int nAVG, int ntsAVG) {
  if (!h || nAVG < 0) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);
  if (nAVG > 0 && !h->fl.avgzeta) {
    Flds& f = h->fl; const int N = h->cfg.N;
    int rc = 0;
    rc |= alloc_field(h, "avgzeta", &f.avgzeta, 0, 1); rc |= alloc_field(h, "avgu2d", &f.avgu2d, 0, 1); rc |= alloc_field(h, "avgv2d", &f.avgv2d, 0, 1);
    rc |= alloc_field(h, "avgu3d", &f.avgu3d, 1, N); rc |= alloc_field(h, "avgv3d", &f.avgv3d, 1, N); rc |= alloc_field(h, "avgrho", &f.avgrho, 1, N);
    rc |= alloc_field(h, "avgw3d", &f.avgw3d, 0, N + 1); rc |= alloc_field(h, "avgwvel", &f.avgwvel, 0, N + 1);
    for (int it = 0; it < h->cfg.NT; ++it) rc |= alloc_field(h, "avgt_" + std::to_string(it), &f.avgt[it], 1, N);
    if (rc) return FatalError;
    CK(cudaStreamSynchronize(h->stream));
  }
  h->navg = nAVG; h->ntsavg = ntsAVG;
  return NoError;
}

int roms_b200_set_option(roms_b200_handle h, const char* key, double value) {
  if (!h || !key) return InputError;
  const std::string k(key);
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  if (k == "cuda_graphs") h->use_graphs = value != 0.0;
  else if (k == "step2d_exchange") { const int m = (int)value; if (m < 0 || m > 2 || h->halo) return ConfigError; h->fused_mode = m; }
  else if (k == "overlap") { if (h->halo) return ConfigError; h->opt_overlap = value != 0.0; }
  else if (k == "halo_timeout_s") h->halo_timeout_s = value;
  else if (k == "step2d_loop_kernel") h->loop_kernel = value != 0.0;
  else if (k == "fuse_phases") h->fuse_phases = value != 0.0;
  else if (k == "ghost_compute") h->ghost_compute = value != 0.0;
  else return InputError;
  drop_graphs(h);
  return NoError;
}

#ifdef LK_TRACE
// tuning aid (variant builds only, not declared in the header): the trace words of k_step2d_loop
int roms_b200_debug_loop_trace(roms_b200_handle h, unsigned long long* out64) {
  if (!h || !out64) return InputError;
  CK(cudaMemcpy(out64, h->d_loop_flags + 2048, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return NoError;
}
#endif

int roms_b200_peer_error_inject(roms_b200_handle h) {
  if (!h || !h->d_err) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  const unsigned long long one = 1ULL;
  CK(cudaMemcpyAsync(h->d_err, &one, sizeof(one), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NoError;
}

int roms_b200_last_step_ms(roms_b200_handle h, float* ms) {
  if (!h || !ms) return InputError;
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return NoError;
}

int roms_b200_diag(roms_b200_handle h, double* out12) {
  if (!h || !out12) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  int rc = run_phase_async(h, ROMS_B200_DIAG);
  if (rc) return rc;
  return finish_diag(h, out12);
}

// One step the way main3d sees it from the host: nf 2-D forcing arrays by name H2D, the step, the diag scalars D2H.
static int step_with_uploads(roms_b200_handle h, int nf, const char* const* nm, const double* const* src, size_t n2d, double* out12) {
  if (!h || !out12 || nf < 0 || nf > 16) return InputError;
  if (h->nfast < 1) return ConfigError;
  if (h->exit_flag == FatalError) return FatalError;
  CK(cudaSetDevice(h->cfg.device));
  const size_t want = (size_t)h->ni * h->nj;
  bool any = false;
  for (int q = 0; q < nf; ++q) {
    if (!src[q]) continue;
    any = true;
    auto it = h->reg.find(nm[q]);
    if (it == h->reg.end() || it->second.nk != 1 || !it->second.base) { std::fprintf(stderr, "roms_b200: '%s' is not a 2-D field of this handle\n", nm[q]); return InputError; }
  }
  if (any && n2d != want) return InputError;
  if (!h->h_pinned || h->pinned_n < (size_t)nf * want) {
    if (h->h_pinned) { CK(cudaStreamSynchronize(h->stream)); if (h->copy_stream) CK(cudaStreamSynchronize(h->copy_stream)); CK(cudaFreeHost(h->h_pinned)); h->h_pinned = nullptr; }
    const size_t cap = (size_t)std::max(nf, 3) * want;
    CK(cudaMallocHost(&h->h_pinned, cap * sizeof(double))); h->pinned_n = cap;
    CK(cudaMalloc(&h->d_stage, cap * sizeof(double))); h->allocs.push_back(h->d_stage);
  }
  const size_t dp = (size_t)h->par.P * sizeof(double), sp = (size_t)h->ni * sizeof(double);
  // upload on the copy stream, overlapped with set_massflux / rho_eos / diag of this step (step_phases_body waits for ev_forcing
  // before bulk_flux / set_vbc)
  cudaStream_t cps = h->stream;
  if (any) {
    if (!h->copy_stream) {
      CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->ev_forcing, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_step_in, cudaEventDisableTiming));
      drop_graphs(h);                       // the captured steps must contain the wait node
    }
    cps = h->copy_stream;
    CK(cudaEventRecord(h->ev_step_in, h->stream));          // earlier asynchronous steps may still read the forcing arrays
    CK(cudaStreamWaitEvent(cps, h->ev_step_in, 0));
  }
  for (int q = 0; q < nf; ++q) {
    if (!src[q]) continue;
    const double* stage = src[q];
    if (!is_registered(h, src[q], want * sizeof(double))) {         // pageable caller memory: stage through the pinned buffer
      double* st = h->h_pinned + q * want;
      std::memcpy(st, src[q], want * sizeof(double));
      stage = st;
    }
    double* dev = h->reg[nm[q]].base + h->ioff + (h->b.LBi - h->LBi_dev);
    // one dense DMA over PCIe, then the re-pitch on the device (a row-by-row 2-D copy from host memory is several times slower)
    CK(cudaMemcpyAsync(h->d_stage + q * want, stage, want * sizeof(double), cudaMemcpyHostToDevice, cps));
    CK(cudaMemcpy2DAsync(dev, dp, h->d_stage + q * want, sp, sp, (size_t)h->nj, cudaMemcpyDeviceToDevice, cps));
  }
  if (cps != h->stream) CK(cudaEventRecord(h->ev_forcing, cps));
  // No halo exchange here, also in a ring: the host's arrays span LBi:UBi, i.e. they carry their two ghost columns (the
  // reference's set_data ends with mp_exchange2d), and no routine of the path reads a surface forcing field further out than
  // one column (sustr(i+1) / bustr(i+1) in lmd_skpp, the atmosphere at Istr-1 in bulk_flux).  The device's third western ghost
  // column, which only the 3-column exchanges of the advected state fill, is never read for these fields.
  int rc = one_step(h, true);
  if (rc) return rc;
  return finish_diag(h, out12);
}

int roms_b200_step_forced(roms_b200_handle h, const double* sustr, const double* svstr, const double* stflux_temp, size_t n2d, double* out12) {
  if (!h) return InputError;
  const std::string st = "stflux_" + std::to_string(h->cfg.itemp - 1);
  const char* nm[3] = {"sustr", "svstr", st.c_str()};
  const double* src[3] = {sustr, svstr, stflux_temp};
  return step_with_uploads(h, 3, nm, src, n2d, out12);
}

int roms_b200_step_fields(roms_b200_handle h, int nfields, const char* const* names, const double* const* arrays, size_t n2d, double* out12) {
  if (!h || (nfields > 0 && (!names || !arrays))) return InputError;
  for (int q = 0; q < nfields; ++q) if (!names[q]) return InputError;
  return step_with_uploads(h, nfields, names, arrays, n2d, out12);
}

int roms_b200_register_host(roms_b200_handle h, void* p, size_t bytes) {
  if (!h || !p || !bytes) return InputError;
  CK(cudaSetDevice(h->cfg.device));
  if (is_registered(h, p, bytes)) return NoError;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return FatalError; }
  h->host_pinned.emplace_back(p, bytes);
  return NoError;
}
int roms_b200_unregister_host(roms_b200_handle h, void* p) {
  if (!h || !p) return InputError;
  for (size_t q = 0; q < h->host_pinned.size(); ++q) {
    if (h->host_pinned[q].first == p) {
      cudaStreamSynchronize(h->stream);
      cudaHostUnregister(p);
      h->host_pinned.erase(h->host_pinned.begin() + q);
      return NoError;
    }
  }
  return InputError;
}

int roms_b200_profile_enable(roms_b200_handle h, int on) {
  if (!h) return InputError;
  if (on < 0 || on > 2) return InputError;
  h->profile = on; std::memset(h->phase_ms, 0, sizeof(h->phase_ms));
  return NoError;
}
int roms_b200_profile_get(roms_b200_handle h, double* ms32, long long* launches) {
  if (!h || !ms32) return InputError;
  std::memcpy(ms32, h->phase_ms, sizeof(h->phase_ms));
  if (launches) *launches = h->launches;
  return NoError;
}
long long roms_b200_launch_count(roms_b200_handle h) { return h ? h->launches : -1; }

}  // extern "C"

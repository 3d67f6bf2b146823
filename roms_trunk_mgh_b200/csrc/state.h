// Internal (not part of the C-ABI): the device-resident state behind a roms_b200_handle.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/roms_b200.h"
#include "dev.cuh"

namespace rbi {

// tile index sets: ROMS/Utility/get_bounds.F (tile_bounds_2d :933-1007, var_bounds :1009-1853, get_bounds :60-258)
struct Bounds {
  int tile, Itile, Jtile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS, JmaxS;
  int Istr, IstrB, IstrP, IstrR, IstrT, IstrM, IstrU, Iend, IendB, IendP, IendR, IendT;
  int Jstr, JstrB, JstrP, JstrR, JstrT, JstrM, JstrV, Jend, JendB, JendP, JendR, JendT;
  int Istrm3, Istrm2, Istrm1, IstrUm2, IstrUm1, Iendp1, Iendp2, Iendp2i, Iendp3;
  int Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1, Jendp1, Jendp2, Jendp2i, Jendp3;
  int west, east, south, north;
};

struct FieldInfo { double** slot; int LBk, nk; double* base; };

struct Halo;   // NCCL ring context (api_nccl.cu)

}  // namespace rbi

struct roms_b200_state {
  roms_b200_config cfg;
  rbi::Bounds b;          // the reference's bounds of this tile (host-facing array extents LBi:UBi, LBj:UBj)
  rb::Par par;
  rb::Flds fl;
  bool lazy = false;      // fields get device storage on their first upload (per-routine _tile calls, api_tile.cu)
  int itile = 0;          // the xi-column of tiles this handle owns (cfg.tile % cfg.NtileI)
  int ni, nj, ioff;       // host array extents; device origin shift
  int LBi_dev, ni_dev;    // device arrays carry 3 west ghost columns on every tile (the fused step2d kernel needs Drhs(i-3))
  std::map<std::string, rbi::FieldInfo> reg;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // stepping state (mod_stepping.F)
  int iic = 1, ntstart = 1, ntfirst = 1, nstp = 1, nnew = 1, nrhs = 1, iif = 1, indx1 = 1, kstp = 1, krhs = 1, knew = 1, predictor = 0, exit_flag = 0;
  double time = 0.0, tdays = 0.0;
  int nfast = 0;
  std::vector<double> w1, w2;
  double dtfast = 0.0;
  // Sticky device error word (allocated with the state; [0] != 0: a halo wait of a peer-memory exchange gave up, so ghost
  // columns hold garbage).  Every host synchronisation point reads it back and turns it into exit_flag 8 (check_device_error).
  unsigned long long* d_err = nullptr; unsigned long long* h_err = nullptr;
  double halo_timeout_s = 30.0;     // how long a kernel waits for a neighbour's halo before it gives up (roms_b200_set_option)
  // AVERAGES (set_avg.F): window length in steps (0: off) and first step, roms_b200_set_avg
  int navg = 0, ntsavg = 1;
  // diag
  double* d_diag_partial = nullptr; double* d_diag_out = nullptr; double* h_diag_out = nullptr;
  double* h_pinned = nullptr; size_t pinned_n = 0; double* d_stage = nullptr;
  std::vector<std::pair<void*, size_t>> host_pinned;    // caller memory pinned with roms_b200_register_host
  // step_forced: this step's surface forcing is uploaded on its own stream while the first phases of the step (which do not
  // read it) already run; the step waits for ev_forcing just before set_vbc
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_forcing = nullptr, ev_step_in = nullptr;
  // profiling.  profile = 1: every phase bracketed by CUDA events and synchronised (plain stream launches, halos joined per
  // phase); profile = 2: one event between consecutive phases on the main stream, recorded INSIDE the captured step graph
  // (external event-record nodes), so the split describes the configuration that is actually timed and the phase times
  // add up to the step time by construction.
  int profile = 0; double phase_ms[32]; long long launches = 0;
  cudaEvent_t ev_ph[40] = {}; int ev_ph_phase[40] = {}; int n_ev_ph = 0;
  bool all_diff2_zero = true;
  // multi-GPU: ring context, and the second (high-priority) stream on which halo exchanges run while the tile interior is
  // being computed on `stream` (api.cu launch_with_halo)
  rbi::Halo* halo = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_edge = nullptr, ev_halo = nullptr, ev_main = nullptr;
  int overlap = 0;        // edge-first two-stream schedule active (set at attach from opt_overlap)
  int opt_overlap = 1;    // roms_b200_set_option("overlap")
  bool edge_pending = false, halo_pending = false;   // main stream has not yet waited for the latest ev_edge / ev_halo
  // persistent barotropic-loop kernel (k_step2d_loop.cu): per-CTA completion flags, flag base, and the call tables for a loop
  // that starts with indx1 = 1 / 2 (built by roms_b200_set_weights; nullptr: per-call launches)
  unsigned long long* d_loop_flags = nullptr; unsigned long long* d_loop_base = nullptr;
  rb::LoopStep* d_loop_tab[3] = {nullptr, nullptr, nullptr};
  int loop_kernel = 1;    // roms_b200_set_option("step2d_loop_kernel")
  int fuse_phases = 1;    // roms_b200_set_option("fuse_phases"): cross-routine fusions of the whole-step path
  bool ghost_compute = true;   // roms_b200_set_option("ghost_compute"): ring tiles compute the ghost columns of the cheap phases (api.cu launch_ghosts)
  // CUDA graphs of whole time steps, keyed by the stepping state at the start of the step (api.cu one_step)
  int fused_mode = 2;     // step2d halo exchange inside the kernels: 0 off, 1 one launch per sub-step, 2 edge / interior split launches
  bool in_step = false;   // inside step_phases (cross-routine fusions are only legal there: run_phase keeps routine granularity)
  int use_graphs = 1;
  std::map<int, void*> graphs;     // key -> rbi::StepGraph*
};

namespace rbi {
// One captured time step (CUDA graph) and the host-side stepping state it leaves behind.
struct StepGraph { cudaGraphExec_t exec; int indx1, iif, kstp, krhs, knew, predictor; long long launches; int n_marks; };
}  // namespace rbi

namespace rbi {
// Halo exchange of the named fields along the xi ring (mp_exchange2d/3d/4d semantics); no-op without an attached comm.
int halo_exchange(roms_b200_state* h, const std::vector<std::string>& names, cudaStream_t s);
// read the sticky device error word back (after the stream has been synchronised): 0, or 8 with exit_flag set
int check_device_error(roms_b200_state* h);
// mailboxes of the exchange fused into the step2d kernel (false: not available, use halo_exchange)
bool fused_xchg_fill(roms_b200_state* h, rb::Xchg& x);
// cross-tile reduction of the 16-double diag buffer: [0..2] sum, [3..12] max
int halo_reduce_diag(roms_b200_state* h);
void halo_destroy(roms_b200_state* h);
// forget the captured time-step graphs (anything they bake in has changed: weights, stepping indices, the ring)
void drop_graphs(roms_b200_state* h);
}  // namespace rbi

// Multi-GPU halo layer: the B200 equivalent of mp_exchange2d/3d/4d (ROMS/Utility/mp_exchange.F:1413-2128) for a ring
// of tiles along xi (NtileI x 1; east-west periodic so tile 0 and tile NtileI-1 are neighbours, mp_exchange.F:155-187).
// Per exchange: one pack kernel (all fields, all rows incl. the closed-wall rows, all levels; 3 columns eastward, 2
// westward), one ncclGroup with two sends and two receives over NVLink, one unpack kernel -- all on the compute stream.
// NCCL is bound at run time (dlopen) so the single-GPU library has no NCCL dependency and, under torchrun, shares the
// NCCL that torch already loaded.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "state.h"

namespace rbi {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
  if (!a.lib) { std::fprintf(stderr, "roms_b200: cannot load NCCL (%s)\n", dlerror()); return a; }
#define SYM(field, name) *(void**)(&a.field) = dlsym(a.lib, name)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather"); SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.Send && a.Recv && a.AllReduce && a.AllGather && a.GroupStart && a.GroupEnd;
  return a;
}

constexpr int MAXF = 16;
constexpr int NW = 3, NE = 2;      // west / east ghost columns carried by the device arrays

struct Halo {
  ncclComm_t comm = nullptr; bool own_comm = false;
  int rank = 0, nranks = 1, east = 0, west = 0;
  double *sendE = nullptr, *sendW = nullptr, *recvW = nullptr, *recvE = nullptr;   // device buffers
  size_t cap = 0;                                                                  // doubles per buffer (per ghost column count 1)
  // NVLink peer path: every rank owns a mailbox (header + two parity slots of receive buffers) that its neighbours map
  // through CUDA IPC and write into directly, so an exchange is ONE kernel (remote stores + flag, flag wait + local
  // copy: k_halo_xchg) -- no NCCL rendezvous on the critical path of the barotropic sub-steps.
  double* box = nullptr; double* boxW = nullptr; double* boxE = nullptr;           // mine / west neighbour's / east neighbour's
  void* mapW = nullptr; void* mapE = nullptr;                                      // IPC mappings to close
  size_t box_cap = 0;                                                              // doubles per ghost column a slot can hold
  bool peer_on = false;
  // second mailbox in the same allocation (after the first): the exchange fused into the barotropic sub-step kernel
  // (dev.cuh Xchg, k_step2d.cu); off2 = its offset in doubles, identical on every rank
  size_t off2 = 0;
  bool fused_on = false;
  double* diag_all = nullptr;                                                      // 16 doubles per tile (halo_reduce_diag)
};

constexpr int BOX_HDR = 64;                       // header doubles: [2] epoch [4] block counter [6] error
__host__ __device__ inline size_t box_slot(size_t cap) { return 2 * cap * (NW + NE); }   // 16-byte line per double
__host__ __device__ inline size_t box_doubles(size_t cap) { return BOX_HDR + 2 * box_slot(cap); }

struct FieldTab { double* p[MAXF]; int k0[MAXF]; int nk[MAXF]; int off[MAXF]; int n; };   // off = plane offset (in planes) in the buffer

// pack columns [ic, ic+nc) of every (field, level, row) into buf[(plane*nj + j)*nc + c]
__global__ void k_pack(FieldTab t, int P, int PL, int nj, int ic, int nc, int total, double* __restrict__ buf, int unpack) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % nc; const int r = idx / nc; const int j = r % nj; const int plane = r / nj;
  int fi = 0;
  while (fi + 1 < t.n && plane >= t.off[fi + 1]) ++fi;
  const int k = t.k0[fi] + (plane - t.off[fi]);
  double* a = t.p[fi] + (ic + c) + j * P + k * PL;
  if (unpack) *a = buf[idx]; else buf[idx] = *a;
}

__device__ __forceinline__ double* field_elem(const FieldTab& t, int plane, int j, int i, int P, int PL) {
  int fi = 0;
  while (fi + 1 < t.n && plane >= t.off[fi + 1]) ++fi;
  return t.p[fi] + i + j * P + (t.k0[fi] + (plane - t.off[fi])) * PL;
}

// One exchange on the peer path is ONE kernel and uses no fences or separate flags: every double travels as a 16-byte
// line {lo, tag, hi, tag} (tag = low 32 bits of the exchange epoch, the flag-in-data scheme of NCCL's LL protocol, which
// only relies on 8-byte store atomicity over NVLink).  A thread (1) stores its outgoing line straight into the neighbour's
// mailbox, then (2) spins on the line the neighbour stores into the own mailbox at the same position until both tags show
// this epoch and copies the value into the ghost column.  Latency = one NVLink store flight.  Nothing a thread waits for
// depends on another thread of this kernel (the neighbour's stores belong to its own, earlier-ordered work), so the spin
// cannot deadlock.  Two parity slots suffice: a neighbour can only send epoch e+2 after it has received my epoch e+1, which
// I send after my epoch-e kernel (including its unpack) has completed.  (ll_store / ll_load: dev.cuh.)
using rb::ll_load; using rb::ll_store;
__global__ void __launch_bounds__(256) k_halo_xchg(FieldTab t, int P, int PL, int nj, int Istr, int Iend, int totE, int totW, size_t cap,
                                                   unsigned long long* hdr, double* box, double* boxE, double* boxW, unsigned long long* err, long long timeout_ns) {
  const unsigned long long e = hdr[2] + 1;                  // epoch of this exchange (bumped by the last CTA below)
  const unsigned tag = (unsigned)e;
  const size_t base = BOX_HDR + (e & 1) * box_slot(cap);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool east = idx < totE, west = !east && idx < totE + totW;
  const int q = east ? idx : idx - totE;
  const int nc = east ? NW : NE;
  const int c = q % nc, r = q / nc, j = r % nj, plane = r / nj;
  const size_t slotW = base + 2 * (size_t)q;                          // line of element q in the "from the west" half
  const size_t slotE = base + 2 * (cap * NW + (size_t)q);             // ... in the "from the east" half
  if (east) ll_store(boxE + slotW, *field_elem(t, plane, j, Iend - NW + 1 + c, P, PL), tag);   // my last NW columns -> east neighbour's west ghosts
  else if (west) ll_store(boxW + slotE, *field_elem(t, plane, j, Istr + c, P, PL), tag);       // my first NE columns -> west neighbour's east ghosts
  if (east || west) {
    const double* line = box + (east ? slotW : slotE);
    double v;
    rb::ll_wait(line, tag, v, timeout_ns, err);
    if (east) *field_elem(t, plane, j, Istr - NW + c, P, PL) = v;     // from the west neighbour
    else *field_elem(t, plane, j, Iend + 1 + c, P, PL) = v;           // from the east neighbour
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&hdr[4], 1ULL) == gridDim.x - 1) { hdr[4] = 0; hdr[2] = e; }
  }
}

int halo_exchange(roms_b200_state* h, const std::vector<std::string>& names, cudaStream_t s) {
  Halo* H = h->halo;
  if (!H || names.empty()) return 0;
  NcclApi& N = nccl();
  FieldTab t; t.n = 0; int planes = 0;
  for (const std::string& nm : names) {
    auto it = h->reg.find(nm);
    if (it == h->reg.end() || t.n >= MAXF) { std::fprintf(stderr, "roms_b200: halo_exchange: bad field '%s'\n", nm.c_str()); return 2; }
    t.p[t.n] = *it->second.slot; t.k0[t.n] = it->second.LBk; t.nk[t.n] = it->second.nk; t.off[t.n] = planes;
    planes += it->second.nk; ++t.n;
  }
  const int nj = h->nj, P = h->par.P, PL = h->par.PL;
  const size_t need = (size_t)planes * nj;                       // doubles per ghost column
  if (H->peer_on && need <= H->box_cap) {
    const int totE = (int)need * NW, totW = (int)need * NE, tot = totE + totW;
    unsigned long long* hdr = (unsigned long long*)H->box;
    k_halo_xchg<<<(tot + 255) / 256, 256, 0, s>>>(t, P, PL, nj, h->b.Istr, h->b.Iend, totE, totW, H->box_cap, hdr, H->box, H->boxE, H->boxW, h->d_err,
                                                   (long long)(h->halo_timeout_s * 1e9));
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : 8;
  }
  if (need > H->cap) {
    for (double** b : {&H->sendE, &H->sendW, &H->recvW, &H->recvE}) { if (*b) cudaFree(*b); *b = nullptr; }
    H->cap = need + need / 4;
    if (cudaMalloc(&H->sendE, H->cap * NW * sizeof(double)) != cudaSuccess || cudaMalloc(&H->recvW, H->cap * NW * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&H->sendW, H->cap * NE * sizeof(double)) != cudaSuccess || cudaMalloc(&H->recvE, H->cap * NE * sizeof(double)) != cudaSuccess)
      return 8;
  }
  const int Istr = h->b.Istr, Iend = h->b.Iend;
  const int totE = (int)need * NW, totW = (int)need * NE;
  // eastward message: my last NW interior columns -> east neighbour's west ghosts; westward: my first NE columns
  k_pack<<<(totE + 255) / 256, 256, 0, s>>>(t, P, PL, nj, Iend - NW + 1, NW, totE, H->sendE, 0);
  k_pack<<<(totW + 255) / 256, 256, 0, s>>>(t, P, PL, nj, Istr, NE, totW, H->sendW, 0);
  int rc = 0;
  rc |= N.GroupStart();
  rc |= N.Send(H->sendE, (size_t)totE, ncclFloat64, H->east, H->comm, s);
  rc |= N.Recv(H->recvW, (size_t)totE, ncclFloat64, H->west, H->comm, s);
  rc |= N.Send(H->sendW, (size_t)totW, ncclFloat64, H->west, H->comm, s);
  rc |= N.Recv(H->recvE, (size_t)totW, ncclFloat64, H->east, H->comm, s);
  rc |= N.GroupEnd();
  if (rc) { std::fprintf(stderr, "roms_b200: NCCL error in halo_exchange\n"); return 8; }
  k_pack<<<(totE + 255) / 256, 256, 0, s>>>(t, P, PL, nj, Istr - NW, NW, totE, H->recvW, 1);
  k_pack<<<(totW + 255) / 256, 256, 0, s>>>(t, P, PL, nj, Iend + 1, NE, totW, H->recvE, 1);
  h->launches += 4;
  return cudaGetLastError() == cudaSuccess ? 0 : 8;
}

// Mailbox pointers of the fused sub-step exchange; false when the ring is not on the NVLink peer path (or ROMS_B200_FUSED_XCHG=0).
bool fused_xchg_fill(roms_b200_state* h, rb::Xchg& x) {
  Halo* H = h->halo;
  if (!H || !H->peer_on || !H->fused_on || !H->off2) return false;
  std::memset(&x, 0, sizeof(x));
  x.Istr = h->b.Istr; x.Iend = h->b.Iend; x.nj = h->nj;
  x.err = h->d_err; x.timeout_ns = (long long)(h->halo_timeout_s * 1e9);
  x.box = H->box + H->off2; x.boxE = H->boxE + H->off2; x.boxW = H->boxW + H->off2;
  return true;
}

// Cross-tile reduction of the diag scalars.  diag.F reports the Courant components AT the location of the largest
// Courant number (:388-404, a MAXLOC-style reduction across tiles in the distributed build), so the tiles' 16-double
// records are gathered and reduced by one thread in tile order: sums ke, pe, volume in ascending tile number (deterministic),
// (maxC, Cu, Cv, Cw) from the tile that owns the largest C (lowest tile on ties), plain maxima for the rest.
__global__ void k_diag_ring_final(const double* __restrict__ all, int nranks, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ke = 0.0, pe = 0.0, vol = 0.0, mC = 0.0, mCu = 0.0, mCv = 0.0, mCw = 0.0;
  double mx[6] = {0.0, -1.0e37, 0.0, 0.0, 0.0, 0.0};           // max speed, max rho, umax, vmax, ubarmax, vbarmax ([7..12])
  for (int r = 0; r < nranks; ++r) {
    const double* d = all + 16 * r;
    vol = vol + d[2]; pe = pe + d[1]; ke = ke + d[0];
    if (d[3] > mC) { mC = d[3]; mCu = d[4]; mCv = d[5]; mCw = d[6]; }
    for (int q = 0; q < 6; ++q) mx[q] = rb::dmax(mx[q], d[7 + q]);
  }
  out[0] = ke; out[1] = pe; out[2] = vol; out[3] = mC; out[4] = mCu; out[5] = mCv; out[6] = mCw;
  for (int q = 0; q < 6; ++q) out[7 + q] = mx[q];
}

int halo_reduce_diag(roms_b200_state* h) {
  Halo* H = h->halo;
  if (!H) return 0;
  NcclApi& N = nccl();
  if (!H->diag_all && cudaMalloc(&H->diag_all, (size_t)16 * H->nranks * sizeof(double)) != cudaSuccess) return 8;
  if (N.AllGather(h->d_diag_out, H->diag_all, 16, ncclFloat64, H->comm, h->stream)) return 8;
  k_diag_ring_final<<<1, 32, 0, h->stream>>>(H->diag_all, H->nranks, h->d_diag_out);
  h->launches += 1;
  return cudaGetLastError() == cudaSuccess ? 0 : 8;
}

void drop_graphs(roms_b200_state* h) {
  for (auto& kv : h->graphs) { StepGraph* g = (StepGraph*)kv.second; if (g) { cudaGraphExecDestroy(g->exec); delete g; } }
  h->graphs.clear();
}

void halo_destroy(roms_b200_state* h) {
  Halo* H = h->halo;
  if (!H) return;
  for (double* b : {H->sendE, H->sendW, H->recvW, H->recvE}) if (b) cudaFree(b);
  if (H->mapW) cudaIpcCloseMemHandle(H->mapW);
  if (H->mapE && H->mapE != H->mapW) cudaIpcCloseMemHandle(H->mapE);
  if (H->box) cudaFree(H->box);
  if (H->diag_all) cudaFree(H->diag_all);
  if (H->own_comm && H->comm && nccl().CommDestroy) nccl().CommDestroy(H->comm);
  delete H;
  h->halo = nullptr;
  if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); cudaEventDestroy(h->ev_edge); cudaEventDestroy(h->ev_halo); cudaEventDestroy(h->ev_main); h->comm_stream = nullptr; }
}

}  // namespace rbi

using namespace rbi;

extern "C" {

int roms_b200_nccl_unique_id(char* out128) {
  NcclApi& N = nccl();
  if (!N.ok || !out128) return 8;
  ncclUniqueId id;
  if (N.GetUniqueId(&id)) return 8;
  std::memcpy(out128, id.internal, 128);
  return 0;
}

int roms_b200_nccl_init_rank(const char* id128, int rank, int nranks, void** comm_out) {
  NcclApi& N = nccl();
  if (!N.ok || !id128 || !comm_out) return 8;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  if (N.CommInitRank(&c, nranks, id, rank)) return 8;
  *comm_out = c;
  return 0;
}

// Attach the ring communicator and fill every ghost column of every field (call after the uploads).  Collective.
int roms_b200_attach_nccl(roms_b200_handle h, void* nccl_comm, int rank, int nranks) {
  if (!h || !nccl_comm) return 2;
  if (nranks != h->cfg.NtileI || rank != h->itile) { std::fprintf(stderr, "roms_b200: rank/tile mismatch in attach_nccl\n"); return 5; }
  if (!nccl().ok) return 8;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 8;
  Halo* H = new Halo();
  H->comm = (ncclComm_t)nccl_comm; H->rank = rank; H->nranks = nranks;
  H->east = (rank + 1) % nranks; H->west = (rank + nranks - 1) % nranks;
  drop_graphs(h);
  h->halo = H;
  if (!h->comm_stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi) != cudaSuccess) return 8;
    cudaEventCreateWithFlags(&h->ev_edge, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming);
  }
  h->overlap = h->opt_overlap;
  std::vector<std::string> batch;
  for (auto& kv : h->reg) {
    if (kv.first == "P3") continue;
    batch.push_back(kv.first);
    if ((int)batch.size() == 8) { int rc = halo_exchange(h, batch, h->stream); if (rc) return rc; batch.clear(); }
  }
  if (!batch.empty()) { int rc = halo_exchange(h, batch, h->stream); if (rc) return rc; }
  return cudaStreamSynchronize(h->stream) == cudaSuccess ? 0 : 8;
}

// ---- NVLink peer path (optional; the NCCL path stays the fall-back) ---------------------------------------------
// 1. every rank: roms_b200_peer_export -> 64-byte CUDA IPC handle of its mailbox; 2. exchange the handles (host side);
// 3. roms_b200_peer_attach(west's, east's); 4. agree collectively that every rank succeeded; 5. roms_b200_peer_enable.
int roms_b200_peer_export(roms_b200_handle h, char* out64) {
  if (!h || !h->halo || !out64) return 2;
  Halo* H = h->halo;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 8;
  if (!H->box) {
    H->box_cap = (size_t)(4 * (h->cfg.N + 1) + 8) * h->nj;        // the largest per-step exchange (step3d_uv: 4N + 4 planes)
    H->off2 = (box_doubles(H->box_cap) + 15) / 16 * 16;
    const size_t bytes = (H->off2 + rb::xbox_doubles(h->nj)) * sizeof(double);
    if (cudaMalloc(&H->box, bytes) != cudaSuccess) { cudaGetLastError(); return 8; }
    if (cudaMemset(H->box, 0, bytes) != cudaSuccess) return 8;
  }
  cudaIpcMemHandle_t mh;
  if (cudaIpcGetMemHandle(&mh, H->box) != cudaSuccess) { cudaGetLastError(); return 8; }
  static_assert(sizeof(mh) == 64, "CUDA IPC handle size");
  std::memcpy(out64, &mh, 64);
  return 0;
}

int roms_b200_peer_attach(roms_b200_handle h, const char* west64, const char* east64) {
  if (!h || !h->halo || !west64 || !east64) return 2;
  Halo* H = h->halo;
  if (!H->box) return 5;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 8;
  cudaIpcMemHandle_t mw, me;
  std::memcpy(&mw, west64, 64); std::memcpy(&me, east64, 64);
  if (cudaIpcOpenMemHandle(&H->mapW, mw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); H->mapW = nullptr; return 8; }
  if (std::memcmp(west64, east64, 64) == 0) H->mapE = H->mapW;        // two ranks: both neighbours are the same tile
  else if (cudaIpcOpenMemHandle(&H->mapE, me, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); H->mapE = nullptr; return 8; }
  H->boxW = (double*)H->mapW; H->boxE = (double*)H->mapE;
  return 0;
}

int roms_b200_peer_enable(roms_b200_handle h, int on) {
  if (!h || !h->halo) return 2;
  Halo* H = h->halo;
  if (on && !(H->box && H->boxW && H->boxE)) return 5;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 8;
  cudaStreamSynchronize(h->stream);
  drop_graphs(h);
  H->peer_on = on != 0;
  H->fused_on = H->peer_on && h->fused_mode != 0;
  return 0;
}

// 1 if a peer exchange timed out waiting for a neighbour (results are then invalid; every synchronising entry point
// also returns 8 from then on), else 0
int roms_b200_peer_error(roms_b200_handle h) {
  if (!h || !h->d_err) return 0;
  unsigned long long w = 0;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 1;
  cudaMemcpy(&w, h->d_err, sizeof(w), cudaMemcpyDeviceToHost);
  return w != 0;
}

}  // extern "C"

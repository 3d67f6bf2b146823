// Multi-GPU halo layer: the B200 equivalent of mp_exchange2d/3d/4d (ROMS/Utility/mp_exchange.F:1413-2128) for a ring
// of tiles along xi (NtileI x 1; east-west periodic so tile 0 and tile NtileI-1 are neighbours, mp_exchange.F:155-187).
// Per exchange: one pack kernel (all fields, all rows incl. the closed-wall rows, all levels; 3 columns eastward, 2
// westward), one ncclGroup with two sends and two receives over NVLink, one unpack kernel -- all on the compute stream.
// NCCL is bound at run time (dlopen) so the single-GPU library has no NCCL dependency and, under torchrun, shares the
// NCCL that torch already loaded.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
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
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce"); SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.Send && a.Recv && a.AllReduce && a.GroupStart && a.GroupEnd;
  return a;
}

constexpr int MAXF = 16;
constexpr int NW = 3, NE = 2;      // west / east ghost columns carried by the device arrays

struct Halo {
  ncclComm_t comm = nullptr; bool own_comm = false;
  int rank = 0, nranks = 1, east = 0, west = 0;
  double *sendE = nullptr, *sendW = nullptr, *recvW = nullptr, *recvE = nullptr;   // device buffers
  size_t cap = 0;                                                                  // doubles per buffer (per ghost column count 1)
};

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

int halo_exchange(roms_b200_state* h, const std::vector<std::string>& names) {
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
  if (need > H->cap) {
    for (double** b : {&H->sendE, &H->sendW, &H->recvW, &H->recvE}) { if (*b) cudaFree(*b); *b = nullptr; }
    H->cap = need + need / 4;
    if (cudaMalloc(&H->sendE, H->cap * NW * sizeof(double)) != cudaSuccess || cudaMalloc(&H->recvW, H->cap * NW * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&H->sendW, H->cap * NE * sizeof(double)) != cudaSuccess || cudaMalloc(&H->recvE, H->cap * NE * sizeof(double)) != cudaSuccess)
      return 8;
  }
  cudaStream_t s = h->stream;
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

int halo_reduce_diag(roms_b200_state* h) {
  Halo* H = h->halo;
  if (!H) return 0;
  NcclApi& N = nccl();
  int rc = N.GroupStart();
  rc |= N.AllReduce(h->d_diag_out, h->d_diag_out, 3, ncclFloat64, ncclSum, H->comm, h->stream);
  rc |= N.AllReduce(h->d_diag_out + 3, h->d_diag_out + 3, 10, ncclFloat64, ncclMax, H->comm, h->stream);
  rc |= N.GroupEnd();
  return rc ? 8 : 0;
}

void halo_destroy(roms_b200_state* h) {
  Halo* H = h->halo;
  if (!H) return;
  for (double* b : {H->sendE, H->sendW, H->recvW, H->recvE}) if (b) cudaFree(b);
  if (H->own_comm && H->comm && nccl().CommDestroy) nccl().CommDestroy(H->comm);
  delete H;
  h->halo = nullptr;
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
  if (nranks != h->cfg.NtileI || rank != h->cfg.tile) { std::fprintf(stderr, "roms_b200: rank/tile mismatch in attach_nccl\n"); return 5; }
  if (!nccl().ok) return 8;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return 8;
  Halo* H = new Halo();
  H->comm = (ncclComm_t)nccl_comm; H->rank = rank; H->nranks = nranks;
  H->east = (rank + 1) % nranks; H->west = (rank + nranks - 1) % nranks;
  h->halo = H;
  std::vector<std::string> batch;
  for (auto& kv : h->reg) {
    if (kv.first == "P3") continue;
    batch.push_back(kv.first);
    if ((int)batch.size() == 8) { int rc = halo_exchange(h, batch); if (rc) return rc; batch.clear(); }
  }
  if (!batch.empty()) { int rc = halo_exchange(h, batch); if (rc) return rc; }
  return cudaStreamSynchronize(h->stream) == cudaSuccess ? 0 : 8;
}

}  // extern "C"

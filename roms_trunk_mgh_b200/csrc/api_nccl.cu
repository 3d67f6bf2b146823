// Multi-GPU halo layer (mp_exchange2d/3d/4d, ROMS/Utility/mp_exchange.F:1413-2128) -- NCCL is bound at run time with
// dlopen/dlsym so that the library loads without NCCL for single-GPU use and shares torch's NCCL when bench.py runs
// under torch.distributed.
#include <cstdio>
#include "../../include/roms_b200.h"

extern "C" {

int roms_b200_attach_nccl(roms_b200_handle h, void* nccl_comm, int rank, int nranks) {
  (void)h; (void)nccl_comm; (void)rank; (void)nranks;
  std::fprintf(stderr, "roms_b200: multi-GPU halo exchange is not available in this build\n");
  return 5;
}
int roms_b200_nccl_unique_id(char* out128) { (void)out128; return 5; }
int roms_b200_nccl_init_rank(const char* id128, int rank, int nranks, void** comm_out) { (void)id128; (void)rank; (void)nranks; (void)comm_out; return 5; }

}  // extern "C"

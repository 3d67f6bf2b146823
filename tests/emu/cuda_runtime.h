// TEST INFRASTRUCTURE.  Stand-in for <cuda_runtime.h> used ONLY by tests/emu: lets g++ compile the unmodified source text of the
// barrier-free CUDA kernels (csrc/k_glue.cu, k_pre.cu, k_rhs.cu, k_physics.cu) for the host, one "thread" at a time, so that the
// index handling and the operation order of a kernel can be checked bit for bit against the oracle on a machine without a GPU.
// Nothing in the product links or includes this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static const
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
inline int cudaGetDevice(int* d) { *d = 0; return 0; }
inline long long __double_as_longlong(double v) { long long r; __builtin_memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; __builtin_memcpy(&r, &v, 8); return r; }
inline void __nanosleep(unsigned) {}
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
using std::sqrt; using std::exp; using std::log; using std::pow; using std::atan; using std::tanh; using std::fabs; using std::cos; using std::sin;
using std::fmin; using std::fmax; using std::floor;
// one kernel launch = every thread of every block, in order (valid for kernels without barriers or inter-thread communication)
template <class F>
inline void emu_launch(dim3 g, dim3 b, F body) {
  gridDim = g; blockDim = b;
  for (unsigned bz = 0; bz < g.z; ++bz)
    for (unsigned by = 0; by < g.y; ++by)
      for (unsigned bx = 0; bx < g.x; ++bx) {
        blockIdx = uint3{bx, by, bz};
        for (unsigned tz = 0; tz < b.z; ++tz)
          for (unsigned ty = 0; ty < b.y; ++ty)
            for (unsigned tx = 0; tx < b.x; ++tx) { threadIdx = uint3{tx, ty, tz}; body(); }
      }
}

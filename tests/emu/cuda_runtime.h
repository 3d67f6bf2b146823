// TEST INFRASTRUCTURE.  Stand-in for <cuda_runtime.h> used ONLY by tests/emu: lets g++ compile the unmodified source text of the
// barrier-free CUDA kernels (csrc/k_glue.cu, k_pre.cu, k_rhs.cu, k_physics.cu) for the host, one "thread" at a time, so that the
// index handling and the operation order of a kernel can be checked bit for bit against the oracle on a machine without a GPU.
// Nothing in the product links or includes this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__ static              /* one copy per kernel instance: blocks run one after the other */
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline int cudaGetDevice(int* d) { *d = 0; return 0; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return ""; }
inline double __ldcg(const double* a) { return *a; }
inline void __threadfence() {}
unsigned long long atomicAdd(unsigned long long* a, unsigned long long v);
double* emu_dynsmem();                 // the dynamic shared memory of the block that is running (`extern __shared__ double x[]`)
void __syncthreads();
inline long long __double_as_longlong(double v) { long long r; __builtin_memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; __builtin_memcpy(&r, &v, 8); return r; }
inline void __nanosleep(unsigned) {}
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
using std::sqrt; using std::exp; using std::log; using std::pow; using std::atan; using std::tanh; using std::fabs; using std::cos; using std::sin;
using std::fmin; using std::fmax; using std::floor; using std::copysign; using std::min; using std::max;
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

// The same for kernels WITH barriers: the threads of a block are real host threads (a pool of blockDim threads walks the blocks
// one after the other); __syncthreads() is a barrier among the threads of the block that have not returned yet.
struct EmuBlockBarrier {
  std::mutex m; std::condition_variable cv;
  int live = 0, waiting = 0; unsigned long gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    if (++waiting == live) { waiting = 0; ++gen; cv.notify_all(); return; }
    const unsigned long g = gen;
    cv.wait(lk, [&] { return gen != g; });
  }
  void drop() {                          // the thread has returned from the kernel
    std::unique_lock<std::mutex> lk(m);
    --live;
    if (live > 0 && waiting == live) { waiting = 0; ++gen; cv.notify_all(); }
  }
};
extern EmuBlockBarrier emu_block_barrier;
template <class F>
inline void emu_launch_mt(dim3 g, dim3 b, F body) {
  const int T = (int)(b.x * b.y * b.z);
  const long nblocks = (long)g.x * g.y * g.z;
  EmuBlockBarrier endbar;                // all T threads, between two blocks
  endbar.live = T;
  emu_block_barrier.live = T; emu_block_barrier.waiting = 0;
  auto worker = [&](int t) {
    gridDim = g; blockDim = b;
    threadIdx = uint3{(unsigned)(t % b.x), (unsigned)((t / b.x) % b.y), (unsigned)(t / (b.x * b.y))};
    for (long q = 0; q < nblocks; ++q) {
      blockIdx = uint3{(unsigned)(q % g.x), (unsigned)((q / g.x) % g.y), (unsigned)(q / ((long)g.x * g.y))};
      body();
      emu_block_barrier.drop();
      { // end of block: the last thread to arrive re-arms the block barrier for the next block
        std::unique_lock<std::mutex> lk(endbar.m);
        if (++endbar.waiting == endbar.live) {
          endbar.waiting = 0; ++endbar.gen;
          emu_block_barrier.live = T; emu_block_barrier.waiting = 0;
          endbar.cv.notify_all();
        } else {
          const unsigned long gg = endbar.gen;
          endbar.cv.wait(lk, [&] { return endbar.gen != gg; });
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
}

// Warp shuffles (k_diag_final): the 32 lanes of a warp exchange through a mailbox with a barrier among the lanes.  Valid when every
// lane of the warp executes the same sequence of shuffles (full mask), which is how the kernels use them.
struct EmuWarp {
  std::mutex m; std::condition_variable cv; int waiting = 0; unsigned long gen = 0; unsigned long long box[32];
  void sync() {
    std::unique_lock<std::mutex> lk(m);
    if (++waiting == 32) { waiting = 0; ++gen; cv.notify_all(); return; }
    const unsigned long g = gen;
    cv.wait(lk, [&] { return gen != g; });
  }
};
extern EmuWarp emu_warps[64];
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int d) {
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  EmuWarp& W = emu_warps[tid >> 5];
  const unsigned lane = tid & 31;
  unsigned long long bits = 0; __builtin_memcpy(&bits, &v, sizeof(T));
  W.box[lane] = bits;
  W.sync();
  const unsigned long long o = W.box[lane ^ (unsigned)d];
  W.sync();
  T r; __builtin_memcpy(&r, &o, sizeof(T));
  return r;
}

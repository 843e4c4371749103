// Host emulation shim for the CUDA sources of libparcop_b200 -- TEST INFRASTRUCTURE ONLY.
//
// There is no GPU in the development container, so the kernels in pyranda_b200/csrc/kernels.cu are
// also compiled, unmodified, by g++ against this header: every CUDA thread of a block becomes a
// std::thread, __syncthreads() a std::barrier, shared memory a per-block heap buffer, and the
// handful of runtime calls map to malloc / memcpy.  It exists to debug indexing and barrier
// placement of the real kernel source on the CPU (tests/test_emulated_kernels.py); it is never
// built, loaded or shipped by the product (pyranda_b200 loads libparcop_b200.so only).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)

struct double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9, cudaErrorNotSupported = 801 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void *cudaEvent_t;
typedef int cudaMemcpyKind;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, int, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) std::memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace emul {
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local std::barrier<> *t_barrier = nullptr;
inline thread_local double *t_smem = nullptr;
// thread-block clusters: the CTAs of a cluster run concurrently, see each other's shared memory
// and meet at a cluster-wide barrier
inline thread_local int t_cluster_rank = 0, t_cluster_size = 1;
inline thread_local std::barrier<> *t_cluster_barrier = nullptr;
inline thread_local double *const *t_cluster_smem = nullptr;

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::vector<double> smem(smem_bytes / sizeof(double) + 1);
        std::barrier<> bar(nthreads);
        auto worker = [&](unsigned t) {
          t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          t_blockIdx = dim3(bx, by, bz);
          t_blockDim = block;
          t_gridDim = grid;
          t_barrier = &bar;
          t_smem = smem.data();
          body();
          // a thread that leaves early must not deadlock the others
          bar.arrive_and_drop();
        };
        if (nthreads == 1) {
          worker(0);
        } else {
          std::vector<std::thread> th;
          th.reserve(nthreads);
          for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
          for (auto &x : th) x.join();
        }
      }
}
inline void launch_cluster(dim3 grid, int cl, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned c0 = 0; c0 < grid.x; c0 += (unsigned)cl) {
    std::vector<std::vector<double>> smem(cl, std::vector<double>(smem_bytes / sizeof(double) + 1));
    std::vector<double *> sptr(cl);
    for (int c = 0; c < cl; ++c) sptr[c] = smem[c].data();
    std::vector<std::unique_ptr<std::barrier<>>> bars;
    for (int c = 0; c < cl; ++c) bars.emplace_back(new std::barrier<>(nthreads));
    std::barrier<> cbar(nthreads * cl);
    auto worker = [&](int c, unsigned t) {
      t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      t_blockIdx = dim3(c0 + c, 0, 0);
      t_blockDim = block;
      t_gridDim = grid;
      t_barrier = bars[c].get();
      t_smem = sptr[c];
      t_cluster_rank = c;
      t_cluster_size = cl;
      t_cluster_barrier = &cbar;
      t_cluster_smem = sptr.data();
      body();
      bars[c]->arrive_and_drop();
      cbar.arrive_and_drop();
      t_cluster_rank = 0;
      t_cluster_size = 1;
    };
    std::vector<std::thread> th;
    th.reserve(nthreads * cl);
    for (int c = 0; c < cl; ++c)
      for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(worker, c, t);
    for (auto &x : th) x.join();
  }
}
}  // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::t_blockDim)
#define gridDim (emul::t_gridDim)
inline void __syncthreads() { emul::t_barrier->arrive_and_wait(); }
// every warp of a block executes the same sequence of warp barriers in these kernels, so a block
// barrier is a valid (over-synchronising) stand-in
inline void __syncwarp() { emul::t_barrier->arrive_and_wait(); }
template <typename T> inline T __ldg(const T *p) { return *p; }
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::min;
using std::max;

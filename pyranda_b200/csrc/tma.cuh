// TMA (cp.async.bulk.tensor) and mbarrier wrappers for the pipelined sweep kernels, sm_100a.
//
// A y/z sweep tile is NL lines x m rows of a Fortran-ordered field: in memory m pieces of NL*8
// bytes at the row stride.  One thread describes it to the TMA unit as a handful of 3-D boxes of a
// tensor map {nfast, m, nouter}; the unit writes the rows densely into shared memory and signals an
// mbarrier with the byte count, so the whole next tile is in flight while every thread of the CTA
// is busy with the recurrences of the current one.
//
// Host side: the tensor map is encoded per launch (it embeds the field's base pointer) through the
// driver entry point obtained from the runtime, so the library links against cudart only.
#pragma once
#ifndef PB_EMULATE
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

namespace pb {

typedef CUtensorMap TileMap;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// Blocks until the phase with the given parity has completed.  The spin is bounded: a tile that
// never lands (a wrong byte count) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 22)) __trap();
  }
}

// one box of a 3-D tensor map -> dense shared memory, completion counted on `bar`
__device__ __forceinline__ void tma_load_3d(void *dst, const TileMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
      : "memory");
}

// one box shared memory -> global (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(src))
               : "memory");
}
// the same box added to global memory (fp64 add performed by the memory system): accumulating epilogues
__device__ __forceinline__ void tma_reduce_add_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(src))
               : "memory");
}
// the same box max-reduced into global memory as unsigned 64-bit integers (the tensor map must be
// encoded as UINT64): for non-negative doubles the IEEE bit pattern orders like the value, so this is
// max(out, val) for the ring detector without reading the old output into the SM
__device__ __forceinline__ void tma_reduce_max_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.max.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(src))
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until the stores of this thread have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// makes shared-memory writes of this thread visible to the async proxy (before a TMA store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Ampere-style asynchronous 16-byte copies global -> shared (LDGSTS): used by the x sweep, whose
// padded tile layout a tensor map cannot express.
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Encodes {d0, d1, d2} doubles with byte strides {8, s1, s2} and a box {b0, b1, b2}.
// Returns false when the driver entry point is missing or the geometry is not expressible
// (strides must be multiples of 16 bytes, the base 16-byte aligned): the caller falls back to the
// register kernels.
inline bool encode_tile_map(TileMap *map, const double *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                            uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool swizzle128 = false,
                            bool as_u64 = false) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<encode_fn>(p);
  }();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (s1_bytes & 15) || (s2_bytes & 15) || b0 > 256 || b1 > 256 || b2 > 256) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  static const int promo = getenv("PB_TMA_PROMO") ? atoi(getenv("PB_TMA_PROMO")) : 2;
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  return fn(map, as_u64 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, pr,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pb
#else
// Host emulation: a "tensor map" is the plain geometry and a load is a synchronous strided copy.
#include <cstdint>
#include <cstring>
namespace pb {
struct TileMap {
  const double *base;
  uint64_t d0, d1, d2, s1, s2;  // strides in doubles
  uint32_t b0, b1, b2;
  bool swz;  // 128-byte swizzle: the 16-byte chunk index of a 128-byte row is XORed with (row & 7)
};
inline size_t emul_box_index(const TileMap *map, uint32_t i, uint32_t j, uint32_t k) {
  size_t e = ((size_t)k * map->b1 + j) * map->b0 + i;
  if (map->swz) {
    const size_t row = e / 16, col = e % 16;  // b0 == 16 doubles
    e = row * 16 + ((((col / 2) ^ (row & 7)) * 2) + (col & 1));
  }
  return e;
}
inline void cp_async16(void *dst, const void *src) { std::memcpy(dst, src, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
inline void cp_async_wait_but_one() {}
inline void mbar_init(uint64_t *, uint32_t) {}
inline void mbar_expect_tx(uint64_t *, uint32_t) {}
inline void mbar_wait(uint64_t *, uint32_t) {}
inline void tma_load_3d(void *dst, const TileMap *map, int c0, int c1, int c2, uint64_t *) {
  double *d = static_cast<double *>(dst);
  for (uint32_t k = 0; k < map->b2; ++k)
    for (uint32_t j = 0; j < map->b1; ++j)
      for (uint32_t i = 0; i < map->b0; ++i) {
        const long x = (long)c0 + i, y = (long)c1 + j, z = (long)c2 + k;
        const bool in = x >= 0 && y >= 0 && z >= 0 && x < (long)map->d0 && y < (long)map->d1 && z < (long)map->d2;
        d[emul_box_index(map, i, j, k)] = in ? map->base[x + y * (long)map->s1 + z * (long)map->s2] : 0.0;
      }
}
inline void tma_store_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  const double *s = static_cast<const double *>(src);
  double *base = const_cast<double *>(map->base);
  for (uint32_t k = 0; k < map->b2; ++k)
    for (uint32_t j = 0; j < map->b1; ++j)
      for (uint32_t i = 0; i < map->b0; ++i) {
        const long x = (long)c0 + i, y = (long)c1 + j, z = (long)c2 + k;
        const bool in = x >= 0 && y >= 0 && z >= 0 && x < (long)map->d0 && y < (long)map->d1 && z < (long)map->d2;
        if (in) base[x + y * (long)map->s1 + z * (long)map->s2] = s[emul_box_index(map, i, j, k)];
      }
}
inline void tma_reduce_add_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  const double *s = static_cast<const double *>(src);
  double *base = const_cast<double *>(map->base);
  for (uint32_t k = 0; k < map->b2; ++k)
    for (uint32_t j = 0; j < map->b1; ++j)
      for (uint32_t i = 0; i < map->b0; ++i) {
        const long x = (long)c0 + i, y = (long)c1 + j, z = (long)c2 + k;
        const bool in = x >= 0 && y >= 0 && z >= 0 && x < (long)map->d0 && y < (long)map->d1 && z < (long)map->d2;
        if (in) base[x + y * (long)map->s1 + z * (long)map->s2] += s[emul_box_index(map, i, j, k)];
      }
}
inline void tma_reduce_max_3d(const TileMap *map, int c0, int c1, int c2, const void *src) {
  const double *s = static_cast<const double *>(src);
  double *base = const_cast<double *>(map->base);
  for (uint32_t k = 0; k < map->b2; ++k)
    for (uint32_t j = 0; j < map->b1; ++j)
      for (uint32_t i = 0; i < map->b0; ++i) {
        const long x = (long)c0 + i, y = (long)c1 + j, z = (long)c2 + k;
        const bool in = x >= 0 && y >= 0 && z >= 0 && x < (long)map->d0 && y < (long)map->d1 && z < (long)map->d2;
        if (in) {
          double &o = base[x + y * (long)map->s1 + z * (long)map->s2];
          const double val = s[emul_box_index(map, i, j, k)];
          o = val > o ? val : o;
        }
      }
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() {}
inline void fence_async_smem() {}
inline bool encode_tile_map(TileMap *map, const double *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                            uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool swizzle128 = false,
                            bool = false) {
  if ((s1_bytes & 15) || (s2_bytes & 15) || b0 > 256 || b1 > 256 || b2 > 256 || (swizzle128 && b0 != 16)) return false;
  *map = TileMap{base, d0, d1, d2, s1_bytes / 8, s2_bytes / 8, b0, b1, b2, swizzle128};
  return true;
}
}  // namespace pb
#endif

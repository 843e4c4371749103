// Hand-written sm_100a kernels of libparcop_b200.
//
// A directional compact operator (reference: eval_compact_op1{x,y,z}_{d1,r3,r4},
// pyranda/parcop/compact_d1.f90:38-998, compact_r3.f90:37-572, compact_r4.f90:37-877, plus the
// batched solves pentadiagonal.f90:629-825 and the metric scale compact_operators.f90:41-46) is
// ONE kernel here: stencil right-hand side, pentadiagonal solve, scale / filter add-back and the
// composite epilogue are fused so a sweep moves 16 B/point through HBM (read v, write dv).
//
// The solve is partitioned: each grid line is cut into P chunks of C rows; a thread owns one
// (line, chunk), runs the chunk-local LU recurrences sequentially, and the chunks are stitched
// with a dense pre-inverted interface system (tables.cpp).  The forward-eliminated values live in
// shared memory (the whole line tile stays on chip between the two substitution sweeps), so the
// field is read once and written once.
//
//   y / z sweeps: a tile is NL consecutive x-lines (one 128-byte segment per row); threads of a
//   half-warp are adjacent lines, so global loads/stores are fully coalesced and shared-memory
//   accesses conflict free.  v is streamed from global memory straight into a register window.
//   x sweep: lines are unit stride; a tile of NLX lines is staged through shared memory with
//   coalesced loads (row pitch odd => conflict-free column access), solved in place, and written
//   back coalesced.
#include "sweeps.cuh"

#include <mutex>
#include <vector>

namespace pb {

std::atomic<long> g_launches{0};
long launch_count() { return g_launches.load(); }
std::atomic<long> g_pipe_launches{0};
long pipe_launch_count() { return g_pipe_launches.load(); }
std::atomic<long> g_ring_launches{0};
long ring_launch_count() { return g_ring_launches.load(); }
int g_reg_kernels = 1;
int g_pipe_kernels = 1;
// 0: never, 1: where the one-CTA pipelined kernel does not fit (long lines, short slabs, z-slab rings), 2: wherever it fits
int g_ring_kernels = getenv("PB_RING") ? atoi(getenv("PB_RING")) : 1;
static int g_ring_lines = getenv("PB_RING_LINES") ? atoi(getenv("PB_RING_LINES")) : 0;
void set_ring_kernels(int mode, int lines) { if (mode >= 0) g_ring_kernels = mode; if (lines >= 0) g_ring_lines = lines; }
// 0: shared-memory kernels, 1: register kernels, 2: register kernels + TMA-pipelined persistent kernels
void set_reg_kernels(int on) { g_reg_kernels = on >= 1; g_pipe_kernels = on >= 2; }
int sm_count() {
#ifdef PB_EMULATE
  return 1;  // two persistent CTAs: the tile loop and the prefetch hand-over are exercised
#else
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
#endif
}
int max_active_clusters_cached(const void *fn, int cl, size_t smem) {
#ifdef PB_EMULATE
  (void)fn; (void)cl; (void)smem;
  return 2;  // two clusters: the tile loop and the prefetch hand-over are exercised
#else
  struct Key { const void *fn; int cl; size_t smem; int n; };
  static std::vector<Key> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (const Key &k : cache)
    if (k.fn == fn && k.cl == cl && k.smem == smem) return k.n;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(cl * 2 * sm_count()));
  cfg.blockDim = dim3(kBlockThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache.push_back(Key{fn, cl, smem, n});
  return n;
#endif
}

static int g_yz_lines = 32;
static int g_x_lines = 32;
void set_yz_lines(int nl) { if (nl == 8 || nl == 16 || nl == 32) g_yz_lines = nl; }
void set_x_lines(int nl) { if (nl == 8 || nl == 16 || nl == 32) g_x_lines = nl; }


// instantiated in sweeps_d1.cu / sweeps_r3.cu / sweeps_r4.cu / sweeps_r4v.cu
extern template cudaError_t launch_yz_f<F_D1, false>(int, const SweepDev &, const double *, double *, const double *, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_yz_f<F_R3, false>(int, const SweepDev &, const double *, double *, const double *, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_yz_f<F_R4, false>(int, const SweepDev &, const double *, double *, const double *, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_yz_f<F_R4, true>(int, const SweepDev &, const double *, double *, const double *, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_x_f<F_D1, false>(int, const SweepDev &, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_x_f<F_R3, false>(int, const SweepDev &, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_x_f<F_R4, false>(int, const SweepDev &, const double *, double *, const EpiArgs &, cudaStream_t);
extern template cudaError_t launch_x_f<F_R4, true>(int, const SweepDev &, const double *, double *, const EpiArgs &, cudaStream_t);

extern template cudaError_t launch_ring_f<F_D1, false>(int, const SweepDev &, const double *, double *, const double *, const double *, const XRing *, cudaStream_t);
extern template cudaError_t launch_ring_f<F_R3, false>(int, const SweepDev &, const double *, double *, const double *, const double *, const XRing *, cudaStream_t);
extern template cudaError_t launch_ring_f<F_R4, false>(int, const SweepDev &, const double *, double *, const double *, const double *, const XRing *, cudaStream_t);
extern template cudaError_t launch_ring_f<F_R4, true>(int, const SweepDev &, const double *, double *, const double *, const double *, const XRing *, cudaStream_t);

// lines per tile of the ring kernel: 256 / lines chunks per CTA must divide the line's chunks into 1, 2, 4 or 8 CTAs
static int ring_lines(int P, int want) {
  // one CTA per line tile when the line has 4 / 8 / 16 chunks (64 / 32 / 16 lines); longer lines: 32-line tiles
  // over a cluster (256-byte rows measured best at 1024 points)
  const int one = P == 4 ? 64 : (P == 8 ? 32 : (P == 16 ? 16 : 0));
  const int cand[5] = {want, one, 32, 16, 64};
  for (int k = 0; k < 5; ++k) {
    const int nl = cand[k];
    if (nl != 16 && nl != 32 && nl != 64) continue;
    const int pl = kBlockThreads / nl;
    if (P % pl) continue;
    const int cl = P / pl;
    if (cl == 1 || cl == 2 || cl == 4 || cl == 8) return nl;
  }
  return 0;
}

cudaError_t launch_sweep_ring(int fam, int lines, const SweepDev &a, const double *v, double *out, const double *halo_lo,
                              const double *halo_hi, const XRing *xr, const EpiArgs &epi, cudaStream_t st) {
  if (!a.implicit || a.C != 32) return cudaErrorNotSupported;
  lines = ring_lines(a.P, lines > 0 ? lines : g_ring_lines);
  if (!lines) return cudaErrorNotSupported;
  SweepDev b = a;  // composite epilogues ride on the TMA stores (reduce-add, |.| s^2, reduce-max)
  b.acc = 0; b.ring = 0;
  if (epi.mode == EPI_ACC) b.acc = 1;
  else if (epi.mode == EPI_RING_SET || epi.mode == EPI_RING_MAX) {
    if (epi.field != nullptr || fam != F_R4 || a.add_v) return cudaErrorNotSupported;
    b.ring = 1; b.ring_s2 = epi.s2; b.acc = epi.mode == EPI_RING_MAX ? 2 : 0;
  }
  switch (fam) {
    case F_D1: return launch_ring_f<F_D1, false>(lines, b, v, out, halo_lo, halo_hi, xr, st);
    case F_R3: return launch_ring_f<F_R3, false>(lines, b, v, out, halo_lo, halo_hi, xr, st);
    default:
      return a.add_v ? launch_ring_f<F_R4, true>(lines, b, v, out, halo_lo, halo_hi, xr, st)
                     : launch_ring_f<F_R4, false>(lines, b, v, out, halo_lo, halo_hi, xr, st);
  }
}

// lines per tile: as many as fit a 256-thread block (lines * chunks) and ~64 KB of shared memory
static int pick_lines(int want, int P, size_t row_bytes) {
  int lines = (want == 8 || want == 16 || want == 32) ? want : 32;
  while (lines > 8 && (lines * P > kBlockThreads || (size_t)lines * row_bytes > 72 * 1024)) lines /= 2;
  return lines;
}

cudaError_t launch_sweep_yz(int fam, int lines, const SweepDev &a, const double *v, double *out,
                            const double *halo_lo, const double *halo_hi, double *iface,
                            const EpiArgs &epi, cudaStream_t st) {
  if (a.P * 8 > kBlockThreads) return cudaErrorInvalidConfiguration;
  if (g_pipe_kernels && g_ring_kernels && a.implicit && a.C == 32 && iface == nullptr && lines == 0) {
    // long lines (clusters), short slabs (64-line tiles): the one-CTA pipelined kernel covers 256- and 512-point lines
    const bool one_cta_fits = a.P == 8 || a.P == 16;
    if (g_ring_kernels >= 2 || !one_cta_fits) {
      const cudaError_t err = launch_sweep_ring(fam, 0, a, v, out, halo_lo, halo_hi, nullptr, epi, st);
      if (err != cudaErrorNotSupported) return err;
    }
  }
  lines = pick_lines(lines > 0 ? lines : g_yz_lines, a.P, a.implicit ? (size_t)a.m * sizeof(double) : 0);
  switch (fam) {
    case F_D1: return launch_yz_f<F_D1, false>(lines, a, v, out, halo_lo, halo_hi, iface, epi, st);
    case F_R3: return launch_yz_f<F_R3, false>(lines, a, v, out, halo_lo, halo_hi, iface, epi, st);
    default:
      return a.add_v ? launch_yz_f<F_R4, true>(lines, a, v, out, halo_lo, halo_hi, iface, epi, st)
                     : launch_yz_f<F_R4, false>(lines, a, v, out, halo_lo, halo_hi, iface, epi, st);
  }
}

cudaError_t launch_sweep_x(int fam, int lines, const SweepDev &a, const double *v, double *out,
                           const EpiArgs &epi, cudaStream_t st) {
  if (a.P * 8 > kBlockThreads) return cudaErrorInvalidConfiguration;
  lines = pick_lines(lines > 0 ? lines : g_x_lines, a.P, (size_t)(a.m | 1) * sizeof(double));
  switch (fam) {
    case F_D1: return launch_x_f<F_D1, false>(lines, a, v, out, epi, st);
    case F_R3: return launch_x_f<F_R3, false>(lines, a, v, out, epi, st);
    default:
      return a.add_v ? launch_x_f<F_R4, true>(lines, a, v, out, epi, st) : launch_x_f<F_R4, false>(lines, a, v, out, epi, st);
  }
}

static inline int ew_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = 148L * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}
#define PB_GRID_STRIDE(t, n) for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < (n); t += (long)gridDim.x * blockDim.x)

// ---- z-slab helpers ------------------------------------------------------------------------------
__global__ void pack_planes_kernel(const double *__restrict__ v, long plane, int m, int h,
                                   double *__restrict__ send_lo, double *__restrict__ send_hi) {
  const long n = plane * h;
  PB_GRID_STRIDE(t, n) {
    send_lo[t] = v[t];                          // first h planes   (compact_d1.f90:723)
    send_hi[t] = v[(long)(m - h) * plane + t];  // last h planes    (compact_d1.f90:722)
  }
}
cudaError_t launch_pack_planes(const double *v, long plane, int m, int h, double *send_lo, double *send_hi, cudaStream_t st) {
  const long n = plane * h;
  PB_LAUNCH(pack_planes_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, v, plane, m, h, send_lo, send_hi);
  ++g_launches;
  return cudaGetLastError();
}

// reduced interface solve (dense pre-inverted rows) + rank-level spike correction
// (compact_d1.f90:892-928, compact_r4.f90:820-...).  The local pass has already written
// scale * x_local (+ v); this adds  -scale * RC[row] . g  on the rows near the two slab faces where
// the spike columns are not negligible (zone_lo rows from the bottom, zone_hi from the top).
// One thread per line of the xy plane and a block of zone rows (blockIdx.y).
__global__ void z_finish_kernel(double *__restrict__ out, long plane, int m, const double4 *__restrict__ RC,
                                const double *__restrict__ GR, int np, unsigned long long rank_mask, int zone_lo,
                                int zone_hi, const double *__restrict__ iface_all, double scale, int rows_per_block) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= plane) return;
  double g[4] = {0.0, 0.0, 0.0, 0.0};
  const int n4 = 4 * np;
  for (int r = 0; r < np; ++r) {
    if (!((rank_mask >> r) & 1ull)) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double d = __ldg(iface_all + ((long)r * 4 + j) * plane + li);
#pragma unroll
      for (int c = 0; c < 4; ++c) g[c] += __ldg(GR + c * n4 + 4 * r + j) * d;
    }
  }
  // zone rows are numbered 0 .. nz-1: first the bottom zone, then the top zone
  const int lo = min(zone_lo, m), hi = min(zone_hi, m - lo);
  const int nz = lo + hi;
  const int q0 = blockIdx.y * rows_per_block;
  const int q1 = min(nz, q0 + rows_per_block);
  for (int qb = q0; qb < q1; qb += 8) {  // eight rows at a time: all loads first, then the stores
    double old[8];
    long idx[8];
    double4 c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = min(qb + k, q1 - 1);
      const int r = q < lo ? q : m - hi + (q - lo);
      idx[k] = (long)r * plane + li;
      c[k] = ldg4(RC + r);
      old[k] = out[idx[k]];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (qb + k >= q1) break;
      double corr = c[k].x * g[0];
      corr = fma(c[k].y, g[1], corr);
      corr = fma(c[k].z, g[2], corr);
      corr = fma(c[k].w, g[3], corr);
      out[idx[k]] = fma(-scale, corr, old[k]);
    }
  }
}
cudaError_t launch_z_finish(double *out, long plane, int m, const double4 *RC, const double *GR, int np,
                            unsigned long long rank_mask, int zone_lo, int zone_hi, const double *iface_all, double scale,
                            cudaStream_t st) {
  const int rows_per_block = 32;
  const int lo = zone_lo < m ? zone_lo : m, hi = zone_hi < m - lo ? zone_hi : m - lo;
  const int nz = lo + hi;
  if (nz <= 0) return cudaSuccess;
  dim3 grid((unsigned)((plane + 127) / 128), (unsigned)((nz + rows_per_block - 1) / rows_per_block));
  PB_LAUNCH(z_finish_kernel, grid, dim3(128), 0, st, out, plane, m, RC, GR, np, rank_mask, zone_lo, zone_hi, iface_all, scale,
            rows_per_block);
  ++g_launches;
  return cudaGetLastError();
}

// ---- peer exchange --------------------------------------------------------------------------------
// One launch that writes this rank's halo planes / interface values straight into the neighbours'
// memory over NVLink, raises the neighbours' flags and waits for its own: the whole MPI_Sendrecv /
// mpi_allgather step of compact_d1.f90:719-735,890 without a library call.  Every thread fences its
// peer stores at system scope; the last block to finish publishes `epoch` in each neighbour's flag
// word and then spins (bounded) until every neighbour has published the same epoch here, so kernels
// queued behind this one see the neighbours' data.
#ifndef PB_EMULATE
__global__ void __launch_bounds__(256) peer_exchange_kernel(const __grid_constant__ PeerExchange x) {
  for (int c = 0; c < x.ncopies; ++c) {
    const double2 *s = reinterpret_cast<const double2 *>(x.src[c]);
    double2 *d = reinterpret_cast<double2 *>(x.dst[c]);
    const long n2 = (long)(x.bytes[c] / sizeof(double2));
    const long stride = (long)gridDim.x * blockDim.x;
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    for (; t + 3 * stride < n2; t += 4 * stride) {  // four independent 16-byte loads in flight per thread
      const double2 v0 = s[t], v1 = s[t + stride], v2 = s[t + 2 * stride], v3 = s[t + 3 * stride];
      d[t] = v0; d[t + stride] = v1; d[t + 2 * stride] = v2; d[t + 3 * stride] = v3;
    }
    for (; t < n2; t += stride) d[t] = s[t];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(x.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    *x.counter = 0;
    __threadfence_system();
    for (int p = 0; p < x.npeers; ++p) *reinterpret_cast<volatile unsigned long long *>(x.remote_flag[p]) = x.epoch;
  }
  if ((int)threadIdx.x < x.npeers) {
    const volatile unsigned long long *f = x.local_flag[threadIdx.x];
    unsigned long long spin = 0;
    while (*f < x.epoch)
      if (++spin > (1ull << 31)) __trap();  // a neighbour that never arrives must not hang the GPU
    __threadfence_system();
  }
}
#endif
cudaError_t launch_peer_exchange(const PeerExchange &x, cudaStream_t st) {
#ifdef PB_EMULATE
  (void)x; (void)st;
  return cudaErrorNotSupported;
#else
  long most = 0;
  for (int c = 0; c < x.ncopies; ++c) most = x.bytes[c] > (size_t)most ? (long)x.bytes[c] : most;
  long blocks = (most / 16 + 255) / 256 / 4;
  blocks = blocks < 1 ? 1 : (blocks > 2L * sm_count() ? 2L * sm_count() : blocks);
  peer_exchange_kernel<<<(unsigned)blocks, 256, 0, st>>>(x);
  ++g_launches;
  return cudaGetLastError();
#endif
}

// ---- pointwise -----------------------------------------------------------------------------------

// pyranda.py:800-804: PHI = dt*F + A*PHI ; U = U + B*PHI
__global__ void rk4_stage_kernel(long n, double dt, double A, double B, const double *__restrict__ F,
                                 double *__restrict__ PHI, double *__restrict__ U) {
  PB_GRID_STRIDE(t, n) {  // every product rounded before it is added, as numpy evaluates the reference's lines
#ifdef PB_EMULATE
    const double tmp1 = A * PHI[t];
    const double phi = dt * F[t] + tmp1;
    PHI[t] = phi;
    const double tmp2 = B * phi;
    U[t] = U[t] + tmp2;
#else
    const double tmp1 = __dmul_rn(A, PHI[t]);
    const double phi = __dadd_rn(__dmul_rn(dt, F[t]), tmp1);
    PHI[t] = phi;
    const double tmp2 = __dmul_rn(B, phi);
    U[t] = __dadd_rn(U[t], tmp2);
#endif
  }
}
cudaError_t launch_rk4_stage(long n, double dt, double A, double B, const double *F, double *PHI, double *U, cudaStream_t st) {
  PB_LAUNCH(rk4_stage_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, dt, A, B, F, PHI, U);
  ++g_launches;
  return cudaGetLastError();
}

__global__ void copy_kernel(long n, const double *__restrict__ a, double *__restrict__ o) { PB_GRID_STRIDE(t, n) o[t] = a[t]; }
__global__ void fill_kernel(long n, double val, double *__restrict__ o) { PB_GRID_STRIDE(t, n) o[t] = val; }
__global__ void mul_kernel(long n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ o) { PB_GRID_STRIDE(t, n) o[t] = a[t] * b[t]; }
__global__ void div_kernel(long n, const double *a, const double *b, double *o) { PB_GRID_STRIDE(t, n) o[t] = a[t] / b[t]; }
// ringV with per-point length scales (operators.f90:680-683): out = a*b, or max(out, a*b)
__global__ void mul_max_kernel(long n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ o, int first) {
  PB_GRID_STRIDE(t, n) { const double r = a[t] * b[t]; o[t] = first ? r : fmax(o[t], r); }
}
cudaError_t launch_mul_max(long n, const double *a, const double *b, double *out, int first, cudaStream_t st) { PB_LAUNCH(mul_max_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, a, b, out, first); ++g_launches; return cudaGetLastError(); }
cudaError_t launch_copy(long n, const double *a, double *out, cudaStream_t st) { PB_LAUNCH(copy_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, a, out); ++g_launches; return cudaGetLastError(); }
cudaError_t launch_fill(long n, double val, double *out, cudaStream_t st) { PB_LAUNCH(fill_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, val, out); ++g_launches; return cudaGetLastError(); }
cudaError_t launch_mul(long n, const double *a, const double *b, double *out, cudaStream_t st) { PB_LAUNCH(mul_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, a, b, out); ++g_launches; return cudaGetLastError(); }
cudaError_t launch_div(long n, const double *a, const double *b, double *out, cudaStream_t st) { PB_LAUNCH(div_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, a, b, out); ++g_launches; return cudaGetLastError(); }

struct CP9 { const double *p[9]; };
struct P9 { double *p[9]; };

// operators.f90:79-81
__global__ void contra_kernel(long n, const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
                              CP9 M, const double *__restrict__ det, double *__restrict__ fA, double *__restrict__ fB, double *__restrict__ fC) {
  PB_GRID_STRIDE(t, n) {
    const double x = fx[t], y = fy[t], zz = fz[t], d = det[t];
    fA[t] = (x * M.p[0][t] + y * M.p[1][t] + zz * M.p[2][t]) * d;
    fB[t] = (x * M.p[3][t] + y * M.p[4][t] + zz * M.p[5][t]) * d;
    fC[t] = (x * M.p[6][t] + y * M.p[7][t] + zz * M.p[8][t]) * d;
  }
}
cudaError_t launch_contra(long n, const double *fx, const double *fy, const double *fz, const double *const *metric9,
                          const double *det, double *fA, double *fB, double *fC, cudaStream_t st) {
  CP9 M;
  for (int k = 0; k < 9; ++k) M.p[k] = metric9[k];
  PB_LAUNCH(contra_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, fx, fy, fz, M, det, fA, fB, fC);
  ++g_launches;
  return cudaGetLastError();
}

// operators.f90:204-209; metric9 = dAdx dAdy dAdz dBdx dBdy dBdz dCdx dCdy dCdz
__global__ void grad_contract_kernel(long n, CP9 M, double *__restrict__ gx, double *__restrict__ gy, double *__restrict__ gz) {
  PB_GRID_STRIDE(t, n) {
    const double a = gx[t], b = gy[t], c = gz[t];
    gx[t] = a * M.p[0][t] + b * M.p[3][t] + c * M.p[6][t];
    gy[t] = a * M.p[1][t] + b * M.p[4][t] + c * M.p[7][t];
    gz[t] = a * M.p[2][t] + b * M.p[5][t] + c * M.p[8][t];
  }
}
cudaError_t launch_grad_contract(long n, const double *const *metric9, double *gx, double *gy, double *gz, cudaStream_t st) {
  CP9 M;
  for (int k = 0; k < 9; ++k) M.p[k] = metric9[k];
  PB_LAUNCH(grad_contract_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, M, gx, gy, gz);
  ++g_launches;
  return cudaGetLastError();
}

// mesh.f90:337-358; J9 = dxdA dxdB dxdC dydA dydB dydC dzdA dzdB dzdC (already divided by dA,dB,dC)
__global__ void metrics_kernel(long n, CP9 J, double dA, double dB, double dC, P9 I, double *__restrict__ det,
                               double *__restrict__ d1, double *__restrict__ d2, double *__restrict__ d3,
                               double *__restrict__ cellvol, double *__restrict__ gridlen) {
  PB_GRID_STRIDE(t, n) {
    const double dxdA = J.p[0][t], dxdB = J.p[1][t], dxdC = J.p[2][t];
    const double dydA = J.p[3][t], dydB = J.p[4][t], dydC = J.p[5][t];
    const double dzdA = J.p[6][t], dzdB = J.p[7][t], dzdC = J.p[8][t];
    const double dt = -dxdC * dydB * dzdA + dxdB * dydC * dzdA + dxdC * dydA * dzdB - dxdA * dydC * dzdB - dxdB * dydA * dzdC + dxdA * dydB * dzdC;
    det[t] = dt;
    I.p[0][t] = (-dydC * dzdB + dydB * dzdC) / dt;
    I.p[1][t] = (dxdC * dzdB - dxdB * dzdC) / dt;
    I.p[2][t] = (-dxdC * dydB + dxdB * dydC) / dt;
    I.p[3][t] = (dydC * dzdA - dydA * dzdC) / dt;
    I.p[4][t] = (-dxdC * dzdA + dxdA * dzdC) / dt;
    I.p[5][t] = (dxdC * dydA - dxdA * dydC) / dt;
    I.p[6][t] = (-dydB * dzdA + dydA * dzdB) / dt;
    I.p[7][t] = (dxdB * dzdA - dxdA * dzdB) / dt;
    I.p[8][t] = (-dxdB * dydA + dxdA * dydB) / dt;
    const double a = sqrt((dxdA * dA) * (dxdA * dA) + (dydA * dA) * (dydA * dA) + (dzdA * dA) * (dzdA * dA));
    const double b = sqrt((dxdB * dB) * (dxdB * dB) + (dydB * dB) * (dydB * dB) + (dzdB * dB) * (dzdB * dB));
    const double c = sqrt((dxdC * dC) * (dxdC * dC) + (dydC * dC) * (dydC * dC) + (dzdC * dC) * (dzdC * dC));
    d1[t] = a; d2[t] = b; d3[t] = c;
    cellvol[t] = a * b * c;
    gridlen[t] = fmin(a, fmin(b, c));
  }
}
cudaError_t launch_metrics(long n, const double *const *J9, double dA, double dB, double dC, double *const *inv9,
                           double *det, double *d1, double *d2, double *d3, double *cellvol, double *gridlen, cudaStream_t st) {
  CP9 J; P9 I;
  for (int k = 0; k < 9; ++k) { J.p[k] = J9[k]; I.p[k] = inv9[k]; }
  PB_LAUNCH(metrics_kernel, PB_EW_GRID(n), PB_EW_BLOCK, 0, st, n, J, dA, dB, dC, I, det, d1, d2, d3, cellvol, gridlen);
  ++g_launches;
  return cudaGetLastError();
}

// ---- reductions (pyrandaMPI.py:307-326, local part) ---------------------------------------------
template <int KIND>
__device__ __forceinline__ double red_op(double a, double b) {
  return KIND == 0 ? a + b : (KIND == 1 ? fmax(a, b) : fmin(a, b));
}
#ifdef PB_EMULATE
template <int KIND>
__global__ void reduce_kernel(long n, const double *__restrict__ v, double *__restrict__ outp) {
  double acc = (KIND == 0) ? 0.0 : ((KIND == 1) ? -INFINITY : INFINITY);
  for (long t = 0; t < n; ++t) acc = red_op<KIND>(acc, v[t]);
  outp[0] = acc;
}
#else
template <int KIND>
__global__ void reduce_kernel(long n, const double *__restrict__ v, double *__restrict__ outp) {
  __shared__ double sm[32];
  double acc = (KIND == 0) ? 0.0 : ((KIND == 1) ? -INFINITY : INFINITY);
  PB_GRID_STRIDE(t, n) acc = red_op<KIND>(acc, v[t]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc = red_op<KIND>(acc, __shfl_down_sync(0xffffffffu, acc, off));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    acc = (lane < nw) ? sm[lane] : ((KIND == 0) ? 0.0 : ((KIND == 1) ? -INFINITY : INFINITY));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = red_op<KIND>(acc, __shfl_down_sync(0xffffffffu, acc, off));
    if (lane == 0) outp[blockIdx.x] = acc;
  }
}
#endif

template <int KIND>
static void reduce_two_stage(long n, const double *v, double *partial, int blocks, double *result, cudaStream_t st) {
  auto kfn = reduce_kernel<KIND>;
#ifdef PB_EMULATE
  (void)blocks;
  PB_LAUNCH(kfn, dim3(1), dim3(1), 0, st, n, v, partial);
  PB_LAUNCH(kfn, dim3(1), dim3(1), 0, st, 1L, partial, result);
#else
  PB_LAUNCH(kfn, dim3(blocks), dim3(256), 0, st, n, v, partial);
  PB_LAUNCH(kfn, dim3(1), dim3(256), 0, st, (long)blocks, partial, result);
#endif
}

cudaError_t launch_reduce(int kind, long n, const double *v, double *partial, int nblocks, double *result, cudaStream_t st) {
  long want = (n + 1023) / 1024;
  int blocks = (int)(want < nblocks ? (want < 1 ? 1 : want) : nblocks);
  switch (kind) {
    case 0: reduce_two_stage<0>(n, v, partial, blocks, result, st); break;
    case 1: reduce_two_stage<1>(n, v, partial, blocks, result, st); break;
    default: reduce_two_stage<2>(n, v, partial, blocks, result, st); break;
  }
  g_launches += 2;
  return cudaGetLastError();
}

}  // namespace pb

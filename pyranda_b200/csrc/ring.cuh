// The "ring" y/z sweep: the pipelined persistent kernel of sweeps.cuh generalised so that the chunks
// of one grid line may live in different places.  A line is always a ring (periodic) or a row
// (bounded) of 32-row chunks whose two-value states are combined with the short Mf / Mb sums; what
// changes is where a chunk's state is found:
//   * in this CTA's shared memory                        (what sweep_yz_pipe_kernel does),
//   * in the shared memory of another CTA of the cluster  (lines longer than one CTA's tile: the
//     CTAs of a thread-block cluster each take ML = 256 * 32 / NL rows of the same NL lines and
//     read each other's states through distributed shared memory; NL = 32 keeps 256-byte rows
//     for 512- and 1024-point lines),
//   * on another GPU                                      (z-slab: the ranks' slabs are consecutive
//     pieces of the global line; the states of the chunks next to a slab face are written into the
//     neighbouring ranks' memory over NVLink as self-validating records and polled there, tile by
//     tile, while the rest of the CTA keeps working).
// Replaces, for a split z axis, the MPI_Sendrecv + mpi_allgather + redundant reduced solve + spike
// correction of compact_d1.f90:858-928 / compact_r4.f90:820-877 with ONE kernel per sweep and no
// correction pass: the result is the global factorisation's solution, row for row.
#pragma once

namespace pb {

extern int g_ring_kernels;
constexpr int kXSum = 8;  // most terms of a carried-state sum (the host checks the tables against it)
int max_active_clusters_cached(const void *fn, int cl, size_t smem);

#ifndef PB_EMULATE
__device__ __forceinline__ int cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int cl_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the double2 at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ double2 ld_cl(const double2 *p, int rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_addr(p)), "r"((uint32_t)rank));
  double2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ra) : "memory");
  return v;
}
// a chunk state as four words of {32 data bits, epoch}: every word validates itself, so the record
// needs neither a fence nor a separate flag and may land in any order
__device__ __forceinline__ void xr_store(unsigned long long *rec, double2 v, unsigned int epoch) {
  const unsigned long long e = (unsigned long long)epoch << 32;
  const unsigned long long bx = (unsigned long long)__double_as_longlong(v.x), by = (unsigned long long)__double_as_longlong(v.y);
  // one 256-bit store per record (a whole 32-byte sector crosses the link in one piece)
  asm volatile("st.relaxed.sys.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "l"((bx & 0xffffffffull) | e), "l"((bx >> 32) | e),
               "l"((by & 0xffffffffull) | e), "l"((by >> 32) | e) : "memory");
}
__device__ __forceinline__ bool xr_try_load(const unsigned long long *rec, unsigned int epoch, double2 *v) {
  unsigned long long w0, w1, w2, w3;
  asm volatile("ld.relaxed.sys.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(rec) : "memory");
  const unsigned long long e = (unsigned long long)epoch;
  if ((w0 >> 32) != e || (w1 >> 32) != e || (w2 >> 32) != e || (w3 >> 32) != e) return false;
  v->x = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  v->y = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  return true;
}
__device__ __forceinline__ void xr_pause(unsigned long long &spin) {
  if (++spin > (1ull << 24)) __trap();  // a neighbour that never arrives must not hang the GPU
  if (spin > 64) __nanosleep(64);
}
__device__ __forceinline__ void xr_fence_sys() { __threadfence_system(); }
// barrier `id` (1..15) among `count` threads (whole warps): the chunks that take states from a neighbouring rank
__device__ __forceinline__ void xr_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ unsigned int xr_count(unsigned int *c) { return atomicAdd(c, 1u); }
__device__ __forceinline__ void xr_flag_set(unsigned long long *f, unsigned long long v) { *reinterpret_cast<volatile unsigned long long *>(f) = v; }
__device__ __forceinline__ unsigned long long xr_flag_get(const volatile unsigned long long *f) { return *f; }
__device__ __forceinline__ void xr_fence_tma() { __threadfence_system(); asm volatile("fence.proxy.async;" ::: "memory"); }
#else
inline int cl_rank() { return emul::t_cluster_rank; }
inline int cl_size() { return emul::t_cluster_size; }
inline void cl_sync() { emul::t_cluster_barrier->arrive_and_wait(); }
inline double2 ld_cl(const double2 *p, int rank) {
  const char *q = reinterpret_cast<const char *>(emul::t_cluster_smem[rank]) + (reinterpret_cast<const char *>(p) - reinterpret_cast<const char *>(emul::t_smem));
  return *reinterpret_cast<const double2 *>(q);
}
inline void xr_store(unsigned long long *rec, double2 v, unsigned int epoch) {
  const unsigned long long e = (unsigned long long)epoch << 32;
  unsigned long long bx, by;
  std::memcpy(&bx, &v.x, 8); std::memcpy(&by, &v.y, 8);
  __atomic_store_n(rec + 0, (bx & 0xffffffffull) | e, __ATOMIC_RELEASE);
  __atomic_store_n(rec + 1, (bx >> 32) | e, __ATOMIC_RELEASE);
  __atomic_store_n(rec + 2, (by & 0xffffffffull) | e, __ATOMIC_RELEASE);
  __atomic_store_n(rec + 3, (by >> 32) | e, __ATOMIC_RELEASE);
}
inline bool xr_try_load(const unsigned long long *rec, unsigned int epoch, double2 *v) {
  unsigned long long w[4];
  for (int k = 0; k < 4; ++k) w[k] = __atomic_load_n(rec + k, __ATOMIC_ACQUIRE);
  for (int k = 0; k < 4; ++k)
    if ((w[k] >> 32) != (unsigned long long)epoch) return false;
  const unsigned long long bx = (w[0] & 0xffffffffull) | (w[1] << 32), by = (w[2] & 0xffffffffull) | (w[3] << 32);
  std::memcpy(&v->x, &bx, 8); std::memcpy(&v->y, &by, 8);
  return true;
}
inline void xr_pause(unsigned long long &) { std::this_thread::yield(); }  // the other "rank" is another host thread
inline void xr_fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned int xr_count(unsigned int *c) { return __atomic_fetch_add(c, 1u, __ATOMIC_ACQ_REL); }
inline void xr_flag_set(unsigned long long *f, unsigned long long v) { __atomic_store_n(f, v, __ATOMIC_RELEASE); }
inline unsigned long long xr_flag_get(const volatile unsigned long long *f) { return __atomic_load_n(const_cast<const unsigned long long *>(f), __ATOMIC_ACQUIRE); }
inline void xr_fence_tma() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#endif

// every chunk handled by this warp has constant coefficients (see warp_all_const in sweeps.cuh)
template <int NLT, int PL>
__device__ __forceinline__ bool ring_warp_all_const(const SweepDev &a, int tid, int crank, bool cc) {
#ifdef PB_EMULATE
  if (NLT >= 32) return cc;
  const int per = 32 / NLT, s0 = (tid / NLT) / per * per;
  bool all = a.has_const != 0;
  for (int w = 0; w < per; ++w) all = all && a.ctype[crank * PL + a.perm[crank * PL + s0 + w]] == 0;
  return all;
#else
  (void)a; (void)tid; (void)crank;
  return NLT < 32 ? __all_sync(0xffffffffu, cc) : cc;
#endif
}

// CLUSTER = false: one CTA per line tile (no distributed shared memory, plain block barriers)
// XRM: 0 no cross-rank exchange, 1 waiting form, 2 early form (see XRing)
// TAB: some chunk of the line reads coefficient tables (see sweep_yz_pipe_kernel)
template <int FAM, int NL, bool ADDV, bool LATE, bool RING, bool CLUSTER, int XRM, bool TAB>
__global__ void __launch_bounds__(kBlockThreads, 2)
sweep_yz_ring_kernel(const __grid_constant__ SweepDev a, const __grid_constant__ TileMap tmain, const __grid_constant__ TileMap th4,
                     const __grid_constant__ TileMap tlo, const __grid_constant__ TileMap thi,
                     const __grid_constant__ TileMap tout, const __grid_constant__ PipeGeo g, const __grid_constant__ XRing xr,
                     const double *__restrict__ v) {
  constexpr int CT = 32, H = FT<FAM>::H, HP = 4;
  constexpr bool kRingMConst = PB_MCONST && FAM == F_R4 && !CLUSTER && kXSum <= kMTerms;
  constexpr int PL = kBlockThreads / NL;    // chunks of a line per CTA
  constexpr int ML = PL * CT;               // rows per CTA
  constexpr int SR = NL >= 32 ? 8 : 16;     // rows of every chunk per round of TMA stores
  constexpr int WB = NL < 32 ? NL : 32;     // lines per store box (one warp fills it)
  constexpr int GW = NL / WB;               // store boxes per chunk and round
  PB_SHARED(S);
  const int P = a.P;                        // chunks of the line on this rank = cluster size * PL
  const int CL = CLUSTER ? cl_size() : 1, crank = CLUSTER ? cl_rank() : 0;
  double *tile = S;                                                               // [ML + 2 HP][NL]
  double *stage = S + (size_t)(ML + 2 * HP) * NL;                                 // [PL][GW][SR][WB]
  double2 *EN = reinterpret_cast<double2 *>(stage + (size_t)PL * SR * NL);         // [PL + kXExt][NL]
  double2 *ST = EN + (PL + kXExt) * NL;                                           // [PL + kXExt][NL]
  uint64_t *bar = reinterpret_cast<uint64_t *>(ST + (PL + kXExt) * NL);
  const int tid = threadIdx.x, l = tid % NL, p = a.perm[crank * PL + tid / NL], lp = crank * PL + p;
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const long ntiles = (long)tiles_i * a.nouter;
  const int type = a.ctype[lp];
  const bool cc = ring_warp_all_const<NL, PL>(a, tid, crank, a.has_const && type == 0);
  const bool lo_sp = a.phys_lo && lp == 0, hi_sp = a.phys_hi && lp == P - 1;
  const double scale = a.scale;
  const bool rd1 = g.rowdim == 1;
  // rows -4 .. -1 and ML .. ML+3 of this CTA's piece: the neighbouring CTA's rows, or at the ends of
  // the rank's line the periodic wrap rows / the neighbouring ranks' planes (none at a physical end)
  const bool halo_lo = crank > 0 || g.halo, halo_hi = crank < CL - 1 || g.halo;
  const uint32_t tx_bytes = (uint32_t)((ML + (halo_lo ? HP : 0) + (halo_hi ? HP : 0)) * NL * sizeof(double));
  const int row0 = crank * ML;

  bool halo_seen = XRM == 0 || !xr.push;
  auto issue = [&](long t) {  // one thread: arm the barrier, describe this CTA's piece of the tile to the TMA unit
    const int x0 = (int)(t % tiles_i) * NL, o = (int)(t / tiles_i);
    mbar_expect_tx(bar, tx_bytes);
    for (int b = 0; b < g.nbox; ++b) {
      const int row = row0 + b * g.box_rows;
      tma_load_3d(tile + (size_t)(HP + b * g.box_rows) * NL, &tmain, x0, rd1 ? row : o, rd1 ? o : row, bar);
    }
    if (XRM != 0 && !halo_seen && g.halo && (crank == 0 || crank == CL - 1)) {  // the neighbours' planes have landed in this rank's halo buffers
      for (int q = 0; q < xr.npeers; ++q) {
        unsigned long long spin = 0;
        while (xr_flag_get(xr.hflag_local[q]) < xr.hepoch) xr_pause(spin);
      }
      xr_fence_tma();
      halo_seen = true;
    }
    if (halo_lo) {
      if (crank > 0) tma_load_3d(tile, &th4, x0, rd1 ? row0 - HP : o, rd1 ? o : row0 - HP, bar);
      else tma_load_3d(tile, &tlo, x0, rd1 ? g.lo_row : o, rd1 ? o : g.lo_row, bar);
    }
    if (halo_hi) {
      if (crank < CL - 1) tma_load_3d(tile + (size_t)(HP + ML) * NL, &th4, x0, rd1 ? row0 + ML : o, rd1 ? o : row0 + ML, bar);
      else tma_load_3d(tile + (size_t)(HP + ML) * NL, &thi, x0, rd1 ? g.hi_row : o, rd1 ? o : g.hi_row, bar);
    }
  };
  auto sync_all = [&]() {
    if (CLUSTER && CL > 1) cl_sync();
    else __syncthreads();
  };
  // state of chunk q of this rank's line, q outside [0, P): around the ring, or a neighbouring rank's
  auto en_get = [&](int q) -> double2 {
    if (q < 0) {
      if (XRM != 0 && !a.wrap) return EN[(PL - q - 1) * NL + l];
      q += P;
    }
    if constexpr (!CLUSTER) return EN[q * NL + l];
    const int owner = q / PL, idx = q - owner * PL;
    return owner == crank ? EN[idx * NL + l] : ld_cl(EN + idx * NL + l, owner);
  };
  auto st_get = [&](int q) -> double2 {
    if (q >= P) {
      if (XRM != 0 && !a.wrap) return ST[(PL + q - P) * NL + l];
      q -= P;
    }
    if constexpr (!CLUSTER) return ST[q * NL + l];
    const int owner = q / PL, idx = q - owner * PL;
    return owner == crank ? ST[idx * NL + l] : ld_cl(ST + idx * NL + l, owner);
  };

#ifndef PB_EMULATE
  // early form: the warps that hold a chunk taking forward states from below / backward states from above
  // fetch one record per thread into shared memory and meet at their own barrier
  const bool pollF = XRM == 2 && lp < xr.need_f, pollB = XRM == 2 && lp >= P - xr.need_b;
  const bool warpF = XRM == 2 && __any_sync(0xffffffffu, pollF), warpB = XRM == 2 && __any_sync(0xffffffffu, pollB);
  const int nbarF = XRM == 2 ? __syncthreads_count(warpF) : 0, nbarB = XRM == 2 ? __syncthreads_count(warpB) : 0;
#endif
  if (tid == 0) mbar_init(bar, 1);
  if (XRM != 0 && xr.push) {  // compact_d1.f90:719-735 without a separate pass: my first / last planes into the neighbours' halo buffers
    for (int c = 0; c < 2; ++c) {
      if (xr.push_dst[c] == nullptr) continue;
      const double2 *src = reinterpret_cast<const double2 *>(xr.push_src[c]);
      double2 *dst = reinterpret_cast<double2 *>(xr.push_dst[c]);
      const long n2 = xr.push_n / 2, stride = (long)gridDim.x * kBlockThreads;
      long i = (long)blockIdx.x * kBlockThreads + tid;
      for (; i + 3 * stride < n2; i += 4 * stride) {
        const double2 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
      }
      for (; i < n2; i += stride) dst[i] = src[i];
    }
    xr_fence_sys();
    __syncthreads();
    if (tid == 0 && xr_count(xr.hcounter) == gridDim.x - 1) {
      *xr.hcounter = 0;
      xr_fence_sys();
      for (int q = 0; q < xr.npeers; ++q) xr_flag_set(xr.hflag_remote[q], xr.hepoch);
    }
  }
  sync_all();
  const long ncl = gridDim.x / CL;
  long t = blockIdx.x / CL;
  if (tid == 0 && t < ntiles) issue(t);
#ifdef PB_EMULATE
  __syncthreads();
#else
  if (a.skew_ns > 0 && blockIdx.x >= gridDim.x / 2)
    for (int w = 0; w < a.skew_ns; w += 500) __nanosleep(500);
#endif
  uint32_t parity = 0;
  const double *tw = tile + (size_t)(p * CT + HP - H) * NL + l;

  for (; t < ntiles; t += ncl) {
    const int ti = (int)(t % tiles_i), o = (int)(t / tiles_i);
    const long line0 = (long)ti * NL + (long)o * a.ostride;  // z sweeps: index of the tile's first line in the xy plane
    double rl[CT];

    mbar_wait(bar, parity);
    parity ^= 1;
#ifdef PB_EMULATE
    __syncthreads();  // the emulated load is a synchronous copy by thread 0
#endif
    {  // ---- A: rhs + forward recurrence (zero incoming state), chunk pulled out of the tile ----
      const double2 *luf = a.luf + (size_t)type * CT;
      const double l2c = a.cst[0], l1c = a.cst[1];
      double rm1 = 0.0, rm2 = 0.0;
      stream_chunk_tile<FAM, CT, NL>(a, tw, lo_sp, hi_sp, [&](auto lrc, double rhs, double) {
        constexpr int lr = decltype(lrc)::value;
        double2 c;
        if (cc) c = make_double2(l2c, l1c);
        else c = __ldg(luf + lr);
        double x = fma(-c.x, rm2, rhs);
        x = fma(-c.y, rm1, x);
        rl[lr] = x;
        rm2 = rm1;
        rm1 = x;
      });
      EN[p * NL + l] = make_double2(rm1, rm2);
      if constexpr (XRM != 0) {
        const bool lv = ti * NL + l < a.nfast;
        const int e = P - 1 - lp;  // chunks from the top of the slab
        for (int k = 0; k < xr.nup; ++k)
          if (e < xr.cnt_up[k] && lv && !(xr.nopoll & 2)) xr_store(xr.en_out[k] + ((long)(k * P + e) * xr.plane + line0 + l) * 4, make_double2(rm1, rm2), xr.epoch);
        const int need = XRM == 2 ? 0 : xr.need_f - crank * PL;  // waiting form: forward end states of the chunks below this slab
        for (int idx = tid; idx < need * NL; idx += kBlockThreads) {  // 64-line tiles have more records than threads
          const int e2 = idx / NL, ll = idx - e2 * NL;
          double2 v = make_double2(0.0, 0.0);
          if (ti * NL + ll < a.nfast && !(xr.nopoll & 1)) {
            const unsigned long long *rec = xr.en_in + ((long)e2 * xr.plane + line0 + ll) * 4;
            unsigned long long spin = 0;
            while (!xr_try_load(rec, xr.epoch, &v)) xr_pause(spin);
          }
          EN[(PL + e2) * NL + ll] = v;
        }
      }
    }
    sync_all();  // the tile buffer is free (unless the add-back still reads it), the forward states are visible
    if (!(ADDV && LATE) && tid == 0 && t + ncl < ntiles) issue(t + ncl);

    {  // ---- B: add the carried forward state, backward recurrence (zero incoming state) ----
      double2 st = make_double2(0.0, 0.0);
      {
        int nf = a.nf[lp];
        if (XRM == 2 && nf > lp) nf = lp;  // early form: this rank's chunks now, the ranks below when their states have arrived
        const double4 *Mp = a.Mf + (size_t)lp * a.mstride;
        // unrolled with a guard: the loads of all terms leave together (as a counted loop the sum exposed one
        // shared + one table load latency per term: 11 % of the filter kernel's samples, profiles/r2_xr_kernel_single_gpu_ncu.txt)
        // (cluster form: the states come through distributed shared memory and the counted loop measured faster,
        // 1024^3 ddy 3.5 vs 4.4 ms)
        if constexpr (CLUSTER) {
          for (int j = 1; j <= nf; ++j) {
            const double2 en = en_get(lp - j);
            const double4 M = ldg4(Mp + j);
            st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
            st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
          }
        } else {
#pragma unroll
          for (int j = 1; j <= kXSum; ++j) {
            if (j <= nf) {
              const double2 en = en_get(lp - j);
              // nine-point family on periodic lines: the products are kernel parameters (see sweep_yz_pipe_kernel)
              const double4 M = (kRingMConst && a.mconst) ? a.Mf0[j - 1] : ldg4(Mp + j);
              st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
              st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
            }
          }
        }
      }
      double x1 = 0.0, x2 = 0.0;
      if (cc) {
        const double ip = a.cst[2], u1 = a.cst[2] * a.cst[3], u2 = a.cst[2] * a.cst[4];
        static_for<0, CT>([&](auto jc) {
          constexpr int r = CT - 1 - decltype(jc)::value;
          double x = rl[r];
          x = fma(a.phi0[r].x, st.x, x);
          x = fma(a.phi0[r].y, st.y, x);
          x = fma(-u2, x2, x * ip);
          x = fma(-u1, x1, x);
          rl[r] = (ADDV && LATE) ? fma(x, scale, tw[(r + H) * NL]) : x;
          x2 = x1;
          x1 = x;
        });
      } else {
        if constexpr (TAB) {
          // table chunks: coefficients kTabAhead rows ahead of the chain (see sweep_yz_pipe_kernel)
          const double2 *ph = a.phi + (size_t)type * CT;
          const double4 *lub = a.lub + (size_t)type * CT;
          double2 pf[kTabAhead];
          double pc[kTabAhead][3];
#pragma unroll
          for (int k = 0; k < kTabAhead; ++k) {
            pf[k] = __ldg(ph + (CT - 1 - k));
            const double4 c4 = ldg4(lub + (CT - 1 - k));
            pc[k][0] = c4.x; pc[k][1] = c4.y; pc[k][2] = c4.z;
          }
          static_for<0, CT>([&](auto jc) {
            constexpr int j = decltype(jc)::value, r = CT - 1 - j, slot = j % kTabAhead;
            const double2 f = pf[slot];
            const double cx = pc[slot][0], cy = pc[slot][1], cz = pc[slot][2];
            if constexpr (r - kTabAhead >= 0) {
              pf[slot] = __ldg(ph + (r - kTabAhead));
              const double4 c4 = ldg4(lub + (r - kTabAhead));
              pc[slot][0] = c4.x; pc[slot][1] = c4.y; pc[slot][2] = c4.z;
            }
            double x = rl[r];
            x = fma(f.x, st.x, x);
            x = fma(f.y, st.y, x);
            x = fma(-cz, x2, x * cx);  // lub = {1/pivot, u1/pivot, u2/pivot}: one operation on the chain through x1
            x = fma(-cy, x1, x);
            rl[r] = (ADDV && LATE) ? fma(x, scale, tw[(r + H) * NL]) : x;
            x2 = x1;
            x1 = x;
          });
        } else {
          const double2 *ph = a.phi + (size_t)type * CT;
          const double4 *lub = a.lub + (size_t)type * CT;
          static_for<0, CT>([&](auto jc) {
            constexpr int r = CT - 1 - decltype(jc)::value;
            const double2 f = __ldg(ph + r);
            const double4 c = ldg4(lub + r);
            double x = rl[r];
            x = fma(f.x, st.x, x);
            x = fma(f.y, st.y, x);
            x = fma(-c.z, x2, x * c.x);  // lub = {1/pivot, u1/pivot, u2/pivot}: one operation on the chain through x1
            x = fma(-c.y, x1, x);
            rl[r] = (ADDV && LATE) ? fma(x, scale, tw[(r + H) * NL]) : x;
            x2 = x1;
            x1 = x;
          });
        }
      }
      if constexpr (XRM != 0) {  // my backward start state to the ranks below (early form: before the states from below have arrived;
                    // the receiver adds what its own forward states change in it)
        const bool lv = ti * NL + l < a.nfast;
        for (int k = 0; k < xr.ndn; ++k)
          if (lp < xr.cnt_dn[k] && lv && !(xr.nopoll & 2)) xr_store(xr.st_out[k] + ((long)(k * P + lp) * xr.plane + line0 + l) * 4, make_double2(x1, x2), xr.epoch);
      }
#ifndef PB_EMULATE
      if (XRM == 2 && warpF) {  // one record per thread of the chunks next to the lower face, then everybody who needs them reads shared memory
        if (pollF) {
          double2 en = make_double2(0.0, 0.0);
          if (ti * NL + l < a.nfast && !(xr.nopoll & 1)) {
            const unsigned long long *rec = xr.en_in + ((long)lp * xr.plane + line0 + l) * 4;
            unsigned long long spin = 0;
            while (!xr_try_load(rec, xr.epoch, &en)) xr_pause(spin);
          }
          EN[(PL + lp) * NL + l] = en;
        }
        __syncwarp();
        xr_bar(1, nbarF);
      }
#endif
      if (XRM == 2 && a.nf[lp] > lp) {  // the forward states from below: by now they have usually landed
        double2 sg2 = make_double2(0.0, 0.0);
        if (ti * NL + l < a.nfast && !(xr.nopoll & 1)) {
          const double4 *Mp = a.Mf + (size_t)lp * a.mstride;
          const int nfl = a.nf[lp];
          for (int j = lp + 1; j <= nfl; ++j) {
            double2 en;
#ifdef PB_EMULATE
            const unsigned long long *rec = xr.en_in + ((long)(j - lp - 1) * xr.plane + line0 + l) * 4;
            unsigned long long spin = 0;
            while (!xr_try_load(rec, xr.epoch, &en)) xr_pause(spin);
#else
            if (warpF) {
              en = EN[(PL + j - lp - 1) * NL + l];
            } else {  // a consumer outside the fetching warps (cannot happen for equal state counts per chunk): its own fetch
              const unsigned long long *rec = xr.en_in + ((long)(j - lp - 1) * xr.plane + line0 + l) * 4;
              unsigned long long spin = 0;
              while (!xr_try_load(rec, xr.epoch, &en)) xr_pause(spin);
            }
#endif
            const double4 M = ldg4(Mp + j);
            sg2.x = fma(M.y, en.y, fma(M.x, en.x, sg2.x));
            sg2.y = fma(M.w, en.y, fma(M.z, en.x, sg2.y));
          }
        }
        const double sc = (ADDV && LATE) ? scale : 1.0;  // the late add-back already holds scale * x + v
        const double2 *ch = a.chi + (size_t)type * CT;
        static_for<0, CT>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const double2 c = cc ? a.chi0[r] : __ldg(ch + r);
          const double dx = fma(c.y, sg2.y, c.x * sg2.x);
          rl[r] = fma(dx, sc, rl[r]);
          if (r == 0) x1 += dx;
          if (r == 1) x2 += dx;
        });
      }
      ST[p * NL + l] = make_double2(x1, x2);
      if constexpr (XRM == 1) {
        const int need = xr.need_b - (CL - 1 - crank) * PL;  // waiting form: backward start states of the chunks above this slab
        for (int idx = tid; idx < need * NL; idx += kBlockThreads) {
          const int e2 = idx / NL, ll = idx - e2 * NL;
          double2 v = make_double2(0.0, 0.0);
          if (ti * NL + ll < a.nfast && !(xr.nopoll & 1)) {
            const unsigned long long *rec = xr.st_in + ((long)e2 * xr.plane + line0 + ll) * 4;
            unsigned long long spin = 0;
            while (!xr_try_load(rec, xr.epoch, &v)) xr_pause(spin);
          }
          ST[(PL + e2) * NL + ll] = v;
        }
      }
    }
    sync_all();
    if (ADDV && LATE && tid == 0 && t + ncl < ntiles) issue(t + ncl);

    {  // ---- D: add the carried backward state, scale / add-back, TMA stores ----
      double2 tb = make_double2(0.0, 0.0);
      {
        int nb = a.nb[lp];
        const double4 *Mp = a.Mb + (size_t)lp * a.mstride;
#ifndef PB_EMULATE
        if (XRM == 2 && warpB) {  // one record per thread of the chunks next to the upper face
          if (pollB) {
            double2 sv = make_double2(0.0, 0.0);
            if (ti * NL + l < a.nfast && !(xr.nopoll & 1)) {
              const unsigned long long *rec = xr.st_in + ((long)(lp - (P - xr.need_b)) * xr.plane + line0 + l) * 4;
              unsigned long long spin = 0;
              while (!xr_try_load(rec, xr.epoch, &sv)) xr_pause(spin);
            }
            ST[(PL + lp - (P - xr.need_b)) * NL + l] = sv;
          }
          __syncwarp();
          xr_bar(2, nbarB);
        }
#endif
        if (XRM == 2 && nb > P - 1 - lp) {  // the chunks above this slab: their states as sent, plus what this rank's forward states add to them
          if (ti * NL + l < a.nfast && !(xr.nopoll & 1)) {
            for (int j = P - lp; j <= nb; ++j) {
              double2 sv;
#ifdef PB_EMULATE
              const unsigned long long *rec = xr.st_in + ((long)(lp + j - P) * xr.plane + line0 + l) * 4;
              unsigned long long spin = 0;
              while (!xr_try_load(rec, xr.epoch, &sv)) xr_pause(spin);
#else
              if (warpB) {
                sv = ST[(PL + lp + j - P) * NL + l];
        } else {
                const unsigned long long *rec = xr.st_in + ((long)(lp + j - P) * xr.plane + line0 + l) * 4;
                unsigned long long spin = 0;
                while (!xr_try_load(rec, xr.epoch, &sv)) xr_pause(spin);
              }
#endif
              const double4 M = ldg4(Mp + j);
              tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
              tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
            }
            const double4 *Bp = xr.Bc + (size_t)(lp - (P - xr.need_b)) * xr.bc_n;
            for (int c = 0; c < xr.bc_n; ++c) {
              const double2 en = en_get(P - xr.bc_n + c);
              const double4 M = ldg4(Bp + c);
              tb.x = fma(M.y, en.y, fma(M.x, en.x, tb.x));
              tb.y = fma(M.w, en.y, fma(M.z, en.x, tb.y));
            }
          }
          nb = P - 1 - lp;
        }
        if constexpr (CLUSTER) {
          for (int j = 1; j <= nb; ++j) {
            const double2 sv = st_get(lp + j);
            const double4 M = ldg4(Mp + j);
            tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
            tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
          }
        } else {
#pragma unroll
          for (int j = 1; j <= kXSum; ++j) {
            if (j <= nb) {
              const double2 sv = st_get(lp + j);
              const double4 M = (kRingMConst && a.mconst) ? a.Mb0[j - 1] : ldg4(Mp + j);
              tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
              tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
            }
          }
        }
      }
      double *sg = stage + (size_t)((p * GW + l / WB) * SR) * WB + (l % WB);
      // add-back that is not LATE: the input left the tile when the prefetch started, it is read again (from L2)
      const double *pv = v;
      if (ADDV && !LATE) {
        int xi = ti * NL + l;
        if (xi >= a.nfast) xi = a.nfast - 1;
        pv = v + (long)xi + (long)o * a.ostride + (long)(lp * CT) * a.rstride;
      }
      auto rowD = [&](auto rc, double gx, double gy, double xl) {
        constexpr int r = decltype(rc)::value;
        double val;
        if (ADDV && LATE) {  // xl already holds scale * x_local + v
          val = fma(gx * scale, tb.x, xl);
          val = fma(gy * scale, tb.y, val);
        } else {
          double x = fma(gx, tb.x, xl);
          x = fma(gy, tb.y, x);
          val = x * scale;
          if (ADDV) val += __ldg(pv + (long)r * a.rstride);
        }
        if constexpr (RING) val = fabs(val) * a.ring_s2;
        sg[(r % SR) * WB] = val;
      };
      const double2 *ps = a.psi + (size_t)type * CT;
      static_for<0, CT / SR>([&](auto gc) {
        constexpr int r0 = decltype(gc)::value * SR;
        if ((tid & 31) == 0) tma_store_wait_read();  // this warp's previous stores have left its part of the stage
        __syncwarp();
        if (cc) {
          static_for<r0, r0 + SR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            rowD(rc, a.psi0[r].x, a.psi0[r].y, rl[r]);
          });
        } else {
          constexpr int NQ = TAB ? SR : 1;  // table instantiation: the rows of the group at once
          double2 gq[NQ];
          if constexpr (TAB) {
#pragma unroll
            for (int k = 0; k < SR; ++k) gq[k] = __ldg(ps + r0 + k);
          }
          static_for<r0, r0 + SR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (!TAB) gq[0] = __ldg(ps + r);
            rowD(rc, gq[TAB ? r - r0 : 0].x, gq[TAB ? r - r0 : 0].y, rl[r]);
          });
        }
        fence_async_smem();
        __syncwarp();
        if ((tid & 31) == 0) {  // every warp hands the rows of its own chunks to the TMA unit: no block-wide barrier
#pragma unroll
          for (int w = 0; w < (NL < 32 ? 32 / NL : 1); ++w) {
            const int q = NL < 32 ? a.perm[crank * PL + tid / NL + w] : p;
            const int x0 = ti * NL + (NL < 32 ? 0 : (l / WB) * WB);
            const int row = (crank * PL + q) * CT + r0;
            const int c1 = rd1 ? row : o, c2 = rd1 ? o : row;
            const double *src = stage + (size_t)((q * GW + (NL < 32 ? 0 : l / WB)) * SR) * WB;
            if constexpr (RING) {
              if (a.acc == 2) tma_reduce_max_3d(&tout, x0, c1, c2, src);
              else tma_store_3d(&tout, x0, c1, c2, src);
            } else {
              if (a.acc) tma_reduce_add_3d(&tout, x0, c1, c2, src);
              else tma_store_3d(&tout, x0, c1, c2, src);
            }
          }
          tma_store_commit();
        }
      });
    }
  }
  if ((tid & 31) == 0) tma_store_wait_read();
  if (CLUSTER && CL > 1) cl_sync();  // nobody leaves while a neighbour may still read its states
}

// Host side: tensor maps, per-CTA chunk order, cluster launch.  Returns cudaErrorNotSupported when
// the geometry does not fit (the caller falls back to the other kernels).
template <int FAM, int NL, bool ADDV>
static cudaError_t launch_yz_ring(const SweepDev &a0, const double *v, double *out, const double *hlo, const double *hhi,
                                  const XRing *xrp, cudaStream_t st) {
  constexpr int H = FT<FAM>::H, PL = kBlockThreads / NL, ML = PL * 32, SR = NL >= 32 ? 8 : 16, WB = NL < 32 ? NL : 32;
  const int m = a0.m;
  if (a0.C != 32 || !a0.implicit || a0.P % PL != 0 || m != 32 * a0.P || (hlo == nullptr) != (hhi == nullptr)) return cudaErrorNotSupported;
  const int CL = a0.P / PL;
  if (CL != 1 && CL != 2 && CL != 4 && CL != 8) return cudaErrorNotSupported;
  SweepDev a = a0;
  if (a.mstride == 0) a.mstride = a.P + 1;
  for (int q = 0; q < a.P; ++q)
    if (a.nf[q] > kXSum || a.nb[q] > kXSum) return cudaErrorNotSupported;
  for (int c = 0; c < CL; ++c) {  // chunk order inside every CTA: table chunks first, so they share warps
    int k = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int q = 0; q < PL; ++q) {
        const bool is_const = a.has_const && a.ctype[c * PL + q] == 0;
        if ((pass == 0) != is_const) a.perm[c * PL + k++] = (unsigned char)q;
      }
  }
  const bool ysweep = a.rstride < a.ostride;
  const uint64_t ax = (uint64_t)a.nfast;
  const uint64_t d1 = ysweep ? (uint64_t)m : (uint64_t)a.nouter, d2 = ysweep ? (uint64_t)a.nouter : (uint64_t)m;
  const uint64_t s1 = (uint64_t)(ysweep ? a.rstride : a.ostride) * 8, s2 = (uint64_t)(ysweep ? a.ostride : a.rstride) * 8;
  PipeGeo g;
  g.rowdim = ysweep ? 1 : 2;
  g.box_rows = ML < 256 ? ML : 256;
  g.nbox = ML / g.box_rows;
  g.halo = 0; g.lo_row = 0; g.hi_row = 0;
  TileMap tmain, th4, tlo, thi, tout;
  if (!encode_tile_map(&tmain, v, ax, d1, d2, s1, s2, NL, ysweep ? g.box_rows : 1, ysweep ? 1 : g.box_rows)) return cudaErrorNotSupported;
  if (!encode_tile_map(&th4, v, ax, d1, d2, s1, s2, NL, ysweep ? 4 : 1, ysweep ? 1 : 4)) return cudaErrorNotSupported;
  tlo = th4; thi = th4;
  if (!encode_tile_map(&tout, out, ax, d1, d2, s1, s2, WB, ysweep ? SR : 1, ysweep ? 1 : SR, false, a.acc == 2)) return cudaErrorNotSupported;
  if (hlo != nullptr) {  // z-slab halo planes received from the neighbours: {ax, ay, H} each
    if (ysweep) return cudaErrorNotSupported;
    if (!encode_tile_map(&tlo, hlo, ax, d1, H, s1, s2, NL, 1, 4) || !encode_tile_map(&thi, hhi, ax, d1, H, s1, s2, NL, 1, 4))
      return cudaErrorNotSupported;
    g.halo = 1; g.lo_row = H - 4; g.hi_row = 0;
  } else if (a.wrap) {  // periodic: rows m-4..m-1 and 0..3 of the field itself
    g.halo = 1; g.lo_row = m - 4; g.hi_row = 0;
  }
  XRing xr;
  memset(&xr, 0, sizeof(xr));
  if (xrp != nullptr) {
    xr = *xrp;
    if (xr.on && (ysweep || xr.need_f > kXExt || xr.need_b > kXExt)) return cudaErrorNotSupported;
  }
  const size_t smem = ((size_t)(ML + 8) * NL + (size_t)PL * SR * NL) * sizeof(double) + 2 * (size_t)(PL + kXExt) * NL * sizeof(double2) + 16;
  static const bool late = getenv("PB_ADDV_LATE") ? atoi(getenv("PB_ADDV_LATE")) != 0 : true;
  void (*kfn)(SweepDev, TileMap, TileMap, TileMap, TileMap, TileMap, PipeGeo, XRing, const double *) = nullptr;
  int slot = 0;
  const bool cl1 = CL == 1;
  const int xrm = xr.on ? (xr.early ? 2 : 1) : 0;
  if (xrm != 0 && !cl1) return cudaErrorNotSupported;  // a z-slab line is one CTA's tile (128, 256 or 512 planes)
  // filters: the add-back from the tile (LATE) keeps the tile until the backward pass, so the next tile is
  // requested one phase later; re-reading the input from L2 instead (request after the forward pass) measured
  // slower on one GPU and across ranks (profiles/r2_xr_variants_2gpu.log)
  bool tab = !a.has_const;  // any chunk on the table path?
  for (int q = 0; q < a.P; ++q) tab = tab || a.ctype[q] != 0;
#define PB_RING_PICK1(LATEV, RINGV, TABV)                                                                         \
  (cl1 ? (xrm == 0 ? sweep_yz_ring_kernel<FAM, NL, ADDV, LATEV, RINGV, false, 0, TABV>                            \
                   : xrm == 1 ? sweep_yz_ring_kernel<FAM, NL, ADDV, LATEV, RINGV, false, 1, TABV>                 \
                              : sweep_yz_ring_kernel<FAM, NL, ADDV, LATEV, RINGV, false, 2, TABV>)                \
       : sweep_yz_ring_kernel<FAM, NL, ADDV, LATEV, RINGV, true, 0, TABV>)
#define PB_RING_PICK(LATEV, RINGV) (tab ? PB_RING_PICK1(LATEV, RINGV, true) : PB_RING_PICK1(LATEV, RINGV, false))
  if (a.ring) {
    if constexpr (!ADDV && FAM == F_R4) kfn = PB_RING_PICK(false, true);
    slot = 1;
  } else if (ADDV) {
    if (!late) return cudaErrorNotSupported;
    if constexpr (ADDV) kfn = PB_RING_PICK(true, false);
  } else {
    if constexpr (!ADDV) kfn = PB_RING_PICK(false, false);
  }
#undef PB_RING_PICK
#undef PB_RING_PICK1
  if (kfn == nullptr) return cudaErrorNotSupported;
  static bool configured[16] = {false, false, false, false, false, false, false, false, false, false, false, false, false, false, false, false};
  slot = 2 * (4 * slot + (cl1 ? xrm : 3)) + (tab ? 1 : 0);
  if (!configured[slot]) {
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured[slot] = true;
  }
  const long ntiles = (long)((a.nfast + NL - 1) / NL) * a.nouter;
  long ncl = CL == 1 ? 2L * sm_count() : (long)max_active_clusters_cached(reinterpret_cast<const void *>(kfn), CL, smem);
  if (ncl < 1) return cudaErrorNotSupported;
  if (ntiles < ncl) ncl = ntiles;
#ifdef PB_EMULATE
  if (xr.push) ncl = 1;  // emulated CTAs run one after the other: the flag handshake of the push needs them all resident
#endif
#ifdef PB_EMULATE
  emul::launch_cluster(dim3((unsigned)(ncl * CL)), CL, dim3(kBlockThreads), smem, [&] { kfn(a, tmain, th4, tlo, thi, tout, g, xr, v); });
#else
  static const bool plain1 = getenv("PB_RING_PLAIN_LAUNCH") ? atoi(getenv("PB_RING_PLAIN_LAUNCH")) != 0 : true;
  if (CL == 1 && plain1) {  // no cluster: an ordinary launch
    kfn<<<dim3((unsigned)ncl), dim3(kBlockThreads), smem, st>>>(a, tmain, th4, tlo, thi, tout, g, xr, v);
    ++g_launches;
    ++g_pipe_launches;
    ++g_ring_launches;
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(ncl * CL));
  cfg.blockDim = dim3(kBlockThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, kfn, a, tmain, th4, tlo, thi, tout, g, xr, v);
  if (err != cudaSuccess) return err;
#endif
  ++g_launches;
  ++g_pipe_launches;
  ++g_ring_launches;
  return cudaGetLastError();
}

template <int FAM, bool ADDV>
cudaError_t launch_ring_f(int lines, const SweepDev &a, const double *v, double *out, const double *hlo, const double *hhi,
                          const XRing *xr, cudaStream_t st) {
  if (lines == 16) return launch_yz_ring<FAM, 16, ADDV>(a, v, out, hlo, hhi, xr, st);
  if (lines == 32) return launch_yz_ring<FAM, 32, ADDV>(a, v, out, hlo, hhi, xr, st);
  if (lines == 64) return launch_yz_ring<FAM, 64, ADDV>(a, v, out, hlo, hhi, xr, st);
  return cudaErrorNotSupported;
}

}  // namespace pb

// Sweep kernels of libparcop_b200 (templates).  Included by the per-family translation units
// sweeps_*.cu, which instantiate launch_yz_f / launch_x_f, so the families compile in parallel.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "kernels.cuh"
#include "tables.hpp"
#include "tma.cuh"

#ifdef PB_EMULATE
#define PB_SHARED(S) double *S = emul::t_smem
#define PB_LAUNCH(kernel, grid, block, smem, st, ...) emul::launch(grid, block, smem, [&] { kernel(__VA_ARGS__); })
#define PB_EW_GRID(n) dim3(1)
#define PB_EW_BLOCK dim3(1)
#else
#define PB_SHARED(S) extern __shared__ __align__(1024) double S[]
#define PB_LAUNCH(kernel, grid, block, smem, st, ...) kernel<<<grid, block, smem, st>>>(__VA_ARGS__)
#define PB_EW_GRID(n) dim3(ew_blocks(n))
#define PB_EW_BLOCK dim3(256)
#endif

namespace pb {

extern std::atomic<long> g_launches;
extern std::atomic<long> g_pipe_launches;
extern std::atomic<long> g_ring_launches;
extern int g_reg_kernels;
extern int g_pipe_kernels;
int sm_count();

__device__ __forceinline__ double4 ldg4(const double4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

template <int FAM>
struct FT {
  static constexpr int H = (FAM == F_R4) ? 4 : 3;
  static constexpr int W = 2 * H + 1;
};

// ---- right-hand sides ---------------------------------------------------------------------------
// interior / halo rows.  w = v[i-H .. i+H].
// D1: compact_d1.f90:156, R3: compact_r3.f90:127-128, R4: compact_r4.f90:148-152
template <int FAM>
__device__ __forceinline__ double rhs_center(const double *w, const double *ar) {
  if (FAM == F_D1) {
    return ar[4] * (w[4] - w[2]) + ar[5] * (w[5] - w[1]) + ar[6] * (w[6] - w[0]);
  } else if (FAM == F_R3) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l != 3) s += ar[l] * (w[l] - w[3]);
    return s;
  } else {
    double s = ar[4] * w[4];
#pragma unroll
    for (int l = 0; l < 9; ++l)
      if (l != 4) s += (w[l] - w[4]) * ar[l];
    return s;
  }
}

// first four rows at a one-sided physical boundary; vv = v[0..8], b = closure rows.
// D1: compact_d1.f90:134-136 (+ row 4 by :156 with closure weights), R3: compact_r3.f90:118-120,
// R4: compact_r4.f90:123-126
template <int FAM>
__device__ __forceinline__ void rhs_lo4(const double *vv, const double (*b)[9], double *r) {
  if (FAM == F_D1) {
    // column 8 is sigma, the row sum: non-zero only below an antisymmetric plane (compact_d1.f90:124-126)
    const double v0 = vv[0];
    r[0] = b[0][8] * v0 + b[0][4] * (vv[1] - v0) + b[0][5] * (vv[2] - v0) + b[0][6] * (vv[3] - v0);
    r[1] = b[1][8] * v0 + b[1][3] * (vv[1] - v0) + b[1][4] * (vv[2] - v0) + b[1][5] * (vv[3] - v0) + b[1][6] * (vv[4] - v0);
    r[2] = b[2][8] * v0 + b[2][4] * (vv[3] - vv[1]) + b[2][5] * (vv[4] - v0) + b[2][6] * (vv[5] - v0);
    r[3] = b[3][4] * (vv[4] - vv[2]) + b[3][5] * (vv[5] - vv[1]) + b[3][6] * (vv[6] - v0);
  } else if (FAM == F_R3) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 7; ++l) {
        const int k = l - 3 + i;
        if (k >= 0 && k != i) s += b[i][l] * (vv[k] - vv[i]);
      }
      r[i] = s;
    }
  } else {
    const double v0 = vv[0];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = b[i][4 - i] * v0;
#pragma unroll
      for (int l = 5 - i; l < 9; ++l) s += b[i][l] * (vv[l - 4 + i] - v0);
      r[i] = s;
    }
  }
}

// last four rows; u = v[m-8..m-1], h = closure rows for rows m-4..m-1.
// D1: compact_d1.f90:176-178, R3: compact_r3.f90:144-146, R4: compact_r4.f90:174-177
template <int FAM>
__device__ __forceinline__ void rhs_hi4(const double *u, const double (*h)[9], double *r) {
  if (FAM == F_D1) {
    const double vm = u[7];
    r[0] = h[0][4] * (u[5] - u[3]) + h[0][5] * (u[6] - u[2]) + h[0][6] * (u[7] - u[1]);
    r[1] = h[1][8] * vm + h[1][0] * (u[2] - vm) + h[1][1] * (u[3] - vm) + h[1][2] * (u[4] - u[6]);
    r[2] = h[2][8] * vm + h[2][0] * (u[3] - vm) + h[2][1] * (u[4] - vm) + h[2][2] * (u[5] - vm) + h[2][3] * (u[6] - vm);
    r[3] = h[3][8] * vm + h[3][0] * (u[4] - vm) + h[3][1] * (u[5] - vm) + h[3][2] * (u[6] - vm);
  } else if (FAM == F_R3) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = 4 + q;  // index of this row in u
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 7; ++l) {
        const int k = c - 3 + l;
        if (k < 8 && k != c) s += h[q][l] * (u[k] - u[c]);
      }
      r[q] = s;
    }
  } else {
    const double vm = u[7];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double s = h[q][7 - q] * vm;
#pragma unroll
      for (int l = 0; l < 7 - q; ++l) s += h[q][l] * (u[q + l] - vm);
      r[q] = s;
    }
  }
}

__device__ __forceinline__ void epi_store(double *__restrict__ out, long idx, double val, const EpiArgs &epi) {
  switch (epi.mode) {
    case EPI_STORE: out[idx] = val; break;
    case EPI_ACC: out[idx] += val; break;
    default: {
      double sc = epi.s2;
      if (epi.field) { const double f = epi.field[idx]; sc = f * f; }
      double r = fabs(val) * sc;
      if (epi.mode == EPI_RING_MAX) r = fmax(r, out[idx]);
      out[idx] = r;
    }
  }
}

// epilogue on a value whose previous output `old` was fetched ahead of time (pipelined kernels)
__device__ __forceinline__ bool epi_needs_old(const EpiArgs &epi) { return epi.mode == EPI_ACC || epi.mode == EPI_RING_MAX; }
__device__ __forceinline__ double epi_value(double val, double old, long idx, const EpiArgs &epi) {
  switch (epi.mode) {
    case EPI_STORE: return val;
    case EPI_ACC: return old + val;
    default: {
      double sc = epi.s2;
      if (epi.field) { const double f = __ldg(epi.field + idx); sc = f * f; }
      const double r = fabs(val) * sc;
      return epi.mode == EPI_RING_MAX ? fmax(r, old) : r;
    }
  }
}

template <bool PLAIN>
__device__ __forceinline__ void put(double *__restrict__ out, long idx, double val, const EpiArgs &epi) {
  if (PLAIN) out[idx] = val;
  else epi_store(out, idx, val, epi);
}

// compile-time loop: F(k) is called with k as a template argument
template <int K, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (K < N) {
    f(std::integral_constant<int, K>{});
    static_for<K + 1, N>(f);
  }
}

// ring access: at step K of a 16-row block the window element j (row - H + j) lives in slot (K+j)&15
template <int FAM, int K>
__device__ __forceinline__ double rhs_ring(const double *g, const double *ar) {
#define WR(j) g[(K + (j)) & 15]
  if (FAM == F_D1) {
    return ar[4] * (WR(4) - WR(2)) + ar[5] * (WR(5) - WR(1)) + ar[6] * (WR(6) - WR(0));
  } else if (FAM == F_R3) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l != 3) s += ar[l] * (WR(l) - WR(3));
    return s;
  } else {
    double s = ar[4] * WR(4);
#pragma unroll
    for (int l = 0; l < 9; ++l)
      if (l != 4) s += (WR(l) - WR(4)) * ar[l];
    return s;
  }
#undef WR
}

// Streams one chunk of one grid line through a 16-slot register ring (stencil window + prefetch)
// and hands every row's right-hand side to `emit(local_row, rhs, centre_value)`.
// When C % 16 == 0 every chunk runs the same branch-free block loop: the first chunk only starts
// from a different pointer (periodic wrap rows or the lower halo planes), the last chunk switches
// its load pointer once (wrap rows / upper halo planes), and the one-sided closure rows override
// the right-hand side of the first / last four rows.  Other chunk lengths use the checked path.
template <int FAM, class LDC, class EMIT>
__device__ __forceinline__ void stream_chunk(const SweepDev &a, const double *__restrict__ vp, long rs, int p,
                                             const double *__restrict__ lo_rows, const double *__restrict__ hi_rows,
                                             LDC &&ldc, EMIT &&emit) {
  constexpr int H = FT<FAM>::H;
  const int m = a.m, C = a.C, P = a.P;
  const int s = p * C;
  const bool first = p == 0, last = p == P - 1;
  const bool lo_sp = a.phys_lo && first, hi_sp = a.phys_hi && last;
  double ring[16];
  if ((C & 15) == 0) {
    // rows s-H .. s-1: previous chunk, periodic wrap, lower halo planes, or (closure) unused
    const double *pl = first ? (a.wrap ? vp + (long)(m - H) * rs : (lo_rows ? lo_rows : vp)) : vp + (long)(s - H) * rs;
#pragma unroll
    for (int j = 0; j < H; ++j) { ring[j] = __ldg(pl); pl += rs; }
    if (first) pl = vp;
#pragma unroll
    for (int j = H; j < 16; ++j) { ring[j] = __ldg(pl); pl += rs; }
    double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
    if (lo_sp) {
      double vv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) vv[q] = ring[(q + H) & 15];
      rhs_lo4<FAM>(vv, a.arb_lo, rlo);
    }
    // rows m .. m+H-1 of the last chunk: periodic wrap / upper halo planes / (closure) unused
    const double *pend = a.wrap ? vp : (hi_rows ? hi_rows : vp);
    for (int b = 0; b < C; b += 16) {
      const bool lastblk = last && b == C - 16;
      static_for<0, 16>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        double rhs = rhs_ring<FAM, k>(ring, a.ari);
        const double vc = ring[(k + H) & 15];
        if (k < 4) {
          if (lo_sp && b == 0) rhs = rlo[k];
        }
        if (k == 12) {
          if (hi_sp && lastblk) {
            double u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = ring[(q + H + 8) & 15];
            rhs_hi4<FAM>(u, a.arb_hi, rhi);
          }
        }
        if (k >= 12) {
          if (hi_sp && lastblk) rhs = rhi[k - 12];
        }
        if (k == H) {
          if (lastblk) pl = pend;  // the next row to load is row m
        }
        if (!(lastblk && k >= 2 * H)) {  // rows past m + H - 1 are never used (and a halo buffer holds only H planes)
          ring[k] = __ldg(pl);
          pl += rs;
        }
        emit(b + k, rhs, vc);
      });
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) ring[j] = ldc(s - H + j);
  if (lo_sp) {  // one-sided closure rows 0..3
    double vv[9], r4[4];
#pragma unroll
    for (int k = 0; k < 9; ++k) vv[k] = ring[(k + H) & 15];
    rhs_lo4<FAM>(vv, a.arb_lo, r4);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(k, r4[k], vv[k]);
  }
  for (int b = 0; b < C; b += 16) {
    static_for<0, 16>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int lr = b + k, row = s + lr;
      if (lr < C && !(lo_sp && lr < 4) && !(hi_sp && row >= m - 4))
        emit(lr, rhs_ring<FAM, k>(ring, a.ari), ring[(k + H) & 15]);
      ring[k] = ldc(row - H + 16);
    });
  }
  if (hi_sp) {  // one-sided closure rows m-4..m-1
    double u[8], r4[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = __ldg(vp + (long)(m - 8 + k) * rs);
    rhs_hi4<FAM>(u, a.arb_hi, r4);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(C - 4 + k, r4[k], u[4 + k]);
  }
}

// ---- y / z sweep (implicit operators) ------------------------------------------------------------
// Phases per tile:  A  rhs + forward recurrence over the chunk (zero incoming state)  -> S
//                   F  serial scan over chunks: true forward state entering each chunk -> SF
//                   B  add phi * state, backward recurrence (zero incoming state)       -> S
//                   T  serial scan (reverse): true backward state; periodic: y = K z_R  -> TB, YW
//                   D  add psi * state, Woodbury corner correction, scale / add-back    -> global
template <int FAM, int NL, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(kBlockThreads, 3)
sweep_yz_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v,
                double *__restrict__ out, const double *__restrict__ halo_lo,
                const double *__restrict__ halo_hi, double *__restrict__ iface,
                const __grid_constant__ EpiArgs epi) {
  constexpr int H = FT<FAM>::H;
  PB_SHARED(S);  // [m][NL] recurrence values, then scan states
  const int m = a.m, C = a.C, P = a.P;
  double2 *EN = reinterpret_cast<double2 *>(S + (size_t)m * NL);  // [P][NL] local end values of the forward pass
  double2 *ST = EN + P * NL;                                       // [P][NL] local start values of the backward pass
  const int tid = threadIdx.x, l = tid % NL, p = tid / NL;
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const int ti = blockIdx.x % tiles_i, o = blockIdx.x / tiles_i;
  int i0 = ti * NL + l;
  const bool valid = i0 < a.nfast;
  if (!valid) i0 = a.nfast - 1;
  const long rs = a.rstride;
  const long base = (long)i0 + (long)o * a.ostride;
  const double *vp = v + base;
  const int s = p * C;
  const int type = a.ctype[p];
  const bool cc = a.has_const && type == 0;
  const double scale = a.scale;

  auto ldc = [&](int r) -> double {  // row outside [0,m): periodic wrap or neighbour halo planes
    if (r < 0) {
      if (a.wrap) return __ldg(vp + (long)(r + m) * rs);
      return halo_lo ? __ldg(halo_lo + base + (long)(r + H) * rs) : 0.0;
    }
    if (r >= m) {
      if (r >= m + H) return 0.0;  // look-ahead past the stencil: never used, and a halo buffer holds only H planes
      if (a.wrap) return __ldg(vp + (long)(r - m) * rs);
      return halo_hi ? __ldg(halo_hi + base + (long)(r - m) * rs) : 0.0;
    }
    return __ldg(vp + (long)r * rs);
  };

  // ---- A: forward elimination, pentadiagonal.f90:639-642 in pull form ----
  {
    double rm1 = 0.0, rm2 = 0.0;
    double *sp = S + (size_t)s * NL + l;
    if (cc) {
      const double l2c = a.cst[0], l1c = a.cst[1];
      stream_chunk<FAM>(a, vp, rs, p, halo_lo ? halo_lo + base : nullptr, halo_hi ? halo_hi + base : nullptr, ldc, [&](int lr, double rhs, double) {
        double t = fma(-l2c, rm2, rhs);
        t = fma(-l1c, rm1, t);
        sp[lr * NL] = t;
        rm2 = rm1;
        rm1 = t;
      });
    } else {
      const double2 *luf = a.luf + (size_t)type * C;
      stream_chunk<FAM>(a, vp, rs, p, halo_lo ? halo_lo + base : nullptr, halo_hi ? halo_hi + base : nullptr, ldc, [&](int lr, double rhs, double) {
        const double2 c = __ldg(luf + lr);
        double t = fma(-c.x, rm2, rhs);
        t = fma(-c.y, rm1, t);
        sp[lr * NL] = t;
        rm2 = rm1;
        rm1 = t;
      });
    }
    EN[p * NL + l] = make_double2(rm1, rm2);  // r'_loc[e-1], r'_loc[e-2]
  }
  __syncthreads();

  // ---- B: back substitution (pentadiagonal.f90:643-647) ----
  {
    // true forward state entering this chunk: short weighted sum over the chunks before it
    double2 st = make_double2(0.0, 0.0);
    {
      const int nf = a.nf[p];
      const double4 *Mp = a.Mf + (size_t)p * (P + 1);
      for (int j = 1; j <= nf; ++j) {
        int q = p - j;
        if (q < 0) q += P;  // periodic line: the ring of chunks
        const double2 en = EN[q * NL + l];
        const double4 M = ldg4(Mp + j);
        st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
        st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
      }
    }
    const double2 *ph = a.phi + (size_t)type * C + (C - 1);
    double *sp = S + (size_t)(s + C - 1) * NL + l;
    double x1 = 0.0, x2 = 0.0;
    if (cc) {
      const double ip = a.cst[2], u1 = a.cst[2] * a.cst[3], u2 = a.cst[2] * a.cst[4];
      auto rowB = [&](double2 f, int j) {
        double t = sp[-(j * NL)];
        t = fma(f.x, st.x, t);
        t = fma(f.y, st.y, t);
        t = fma(-u2, x2, t * ip);
        t = fma(-u1, x1, t);
        sp[-(j * NL)] = t;
        x2 = x1;
        x1 = t;
      };
      if (a.cparam && C == 32) {
        static_for<0, 32>([&](auto jc) { constexpr int j = decltype(jc)::value; rowB(a.phi0[31 - j], j); });
      } else if (a.cparam && C == 16) {
        static_for<0, 16>([&](auto jc) { constexpr int j = decltype(jc)::value; rowB(a.phi0[15 - j], j); });
      } else {
#pragma unroll 8
        for (int r = 0; r < C; ++r) rowB(__ldg(ph - r), r);
      }
    } else {
      const double4 *lub = a.lub + (size_t)type * C + (C - 1);
#pragma unroll 4
      for (int r = 0; r < C; ++r) {
        const double2 f = __ldg(ph - r);
        const double4 c = ldg4(lub - r);
        double t = sp[-(r * NL)];
        t = fma(f.x, st.x, t);
        t = fma(f.y, st.y, t);
        t = fma(-c.z, x2, t * c.x);  // lub = {1/pivot, u1/pivot, u2/pivot}: one operation on the chain through x1
        t = fma(-c.y, x1, t);
        sp[-(r * NL)] = t;
        x2 = x1;
        x1 = t;
      }
    }
    ST[p * NL + l] = make_double2(x1, x2);  // x_loc[s], x_loc[s+1]
  }
  __syncthreads();

  // ---- D: carried state, corner correction, metric scale (compact_operators.f90:43), filter
  //         add-back (compact_r4.f90:226-232) and the composite epilogue, straight to global ----
  {
    // true backward state entering chunk q: short weighted sum over the chunks after it
    auto tin = [&](int q) -> double2 {
      double2 t = make_double2(0.0, 0.0);
      const int nb = a.nb[q];
      const double4 *Mp = a.Mb + (size_t)q * (P + 1);
      for (int j = 1; j <= nb; ++j) {
        int qq = q + j;
        if (qq >= P) qq -= P;
        const double2 sv = ST[qq * NL + l];
        const double4 M = ldg4(Mp + j);
        t.x = fma(M.y, sv.y, fma(M.x, sv.x, t.x));
        t.y = fma(M.w, sv.y, fma(M.z, sv.x, t.y));
      }
      return t;
    };
    const double2 tb = tin(p);
    const double2 *ps = a.psi + (size_t)type * C;
    const double *sp = S + (size_t)s * NL + l;
    const double *pv = vp + (long)s * rs;
    long oidx = base + (long)s * rs;
    double *po = out + oidx;
    auto rowD = [&](double2 g, int r) {
      double x = sp[r * NL];
      x = fma(g.x, tb.x, x);
      x = fma(g.y, tb.y, x);
      double val = x * scale;
      if (ADDV) val += __ldg(pv);
      if (PLAIN) {
        if (valid) *po = val;
        po += rs;
      } else {
        if (valid) epi_store(out, oidx, val, epi);
        oidx += rs;
      }
      pv += rs;
    };
    if (cc && a.cparam && C == 32) {
      static_for<0, 32>([&](auto rc) { constexpr int r = decltype(rc)::value; rowD(a.psi0[r], r); });
    } else if (cc && a.cparam && C == 16) {
      static_for<0, 16>([&](auto rc) { constexpr int r = decltype(rc)::value; rowD(a.psi0[r], r); });
    } else {
#pragma unroll 8
      for (int r = 0; r < C; ++r) rowD(__ldg(ps + r), r);
    }
    if (iface != nullptr && valid && (p == 0 || p == P - 1)) {
      // z-slab: publish this rank's 4 interface values, unscaled (compact_d1.f90:858-878)
      const long plane = (long)a.nfast * a.nouter;
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // q = 0, 1: first two rows (chunk 0); q = 2, 3: last two rows (chunk P-1)
        if (q < 2 ? p != 0 : p != P - 1) continue;
        const int lr = q < 2 ? q : C - 4 + q;
        const double2 g = __ldg(ps + lr);
        double x = sp[lr * NL];
        x = fma(g.x, tb.x, x);
        x = fma(g.y, tb.y, x);
        iface[(long)q * plane + base] = (q < 2 ? a.phys_lo : a.phys_hi) ? 0.0 : x;
      }
    }
  }
}

// ---- y / z sweep (explicit operators: the Gaussian filter) ---------------------------------------
template <int FAM, int NL, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(kBlockThreads, 3)
explicit_yz_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v,
                   double *__restrict__ out, const double *__restrict__ halo_lo,
                   const double *__restrict__ halo_hi, const __grid_constant__ EpiArgs epi) {
  constexpr int H = FT<FAM>::H;
  const int m = a.m, C = a.C;
  const int tid = threadIdx.x, l = tid % NL, p = tid / NL;
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const int ti = blockIdx.x % tiles_i, o = blockIdx.x / tiles_i;
  int i0 = ti * NL + l;
  const bool valid = i0 < a.nfast;
  if (!valid) i0 = a.nfast - 1;
  const long rs = a.rstride;
  const long base = (long)i0 + (long)o * a.ostride;
  const double *vp = v + base;
  const double scale = a.scale;
  auto ldc = [&](int r) -> double {
    if (r < 0) {
      if (a.wrap) return __ldg(vp + (long)(r + m) * rs);
      return halo_lo ? __ldg(halo_lo + base + (long)(r + H) * rs) : 0.0;
    }
    if (r >= m) {
      if (r >= m + H) return 0.0;  // look-ahead past the stencil: never used, and a halo buffer holds only H planes
      if (a.wrap) return __ldg(vp + (long)(r - m) * rs);
      return halo_hi ? __ldg(halo_hi + base + (long)(r - m) * rs) : 0.0;
    }
    return __ldg(vp + (long)r * rs);
  };
  const long obase = base + (long)(p * C) * rs;
  stream_chunk<FAM>(a, vp, rs, p, halo_lo ? halo_lo + base : nullptr, halo_hi ? halo_hi + base : nullptr, ldc, [&](int lr, double rhs, double vc) {  // compact_r4.f90:209-218
    double val = rhs * scale;
    if (ADDV) val += vc;
    if (valid) put<PLAIN>(out, obase + (long)lr * rs, val, epi);
  });
}

// ---- x sweep (explicit operators) ----------------------------------------------------------------
// Unit-stride lines need no tile: a thread produces two neighbouring points from five aligned
// 16-byte loads (v[x-4 .. x+5]; the overlap between threads is served by L1), so loads and stores are
// 128-bit and fully coalesced.  Line ends: periodic wrap or the one-sided closure rows.
constexpr int kExplicitXLines = 8;  // 32 KB in flight per block at m = 512

template <int FAM, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(256, 3)
explicit_x_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v, double *__restrict__ out,
                  const __grid_constant__ EpiArgs epi) {
  constexpr int H = FT<FAM>::H, NLB = kExplicitXLines;  // lines per block iteration
  PB_SHARED(S);  // 2 x [NLB][m + 8]: column c holds x = c - 4; the two halves alternate (prefetch)
  const int m = a.m, half = m >> 1, LP = m + 8;
  const double scale = a.scale;
  auto stage = [&](long L0, double *buf) {  // asynchronous 16-byte copies: the whole group of lines in flight at once
#pragma unroll
    for (int ll = 0; ll < NLB; ++ll) {
      long L = L0 + ll;
      if (L >= a.nfast) L = a.nfast - 1;
      const double *lp = v + L * (long)m;
      for (int j = threadIdx.x; j < half + 4; j += blockDim.x) {
        int xs = 2 * j - 4;  // source x of this piece; the first / last two pieces are the wrap halo
        if (xs < 0) xs = a.wrap ? xs + m : 0;
        if (xs >= m) xs = a.wrap ? xs - m : m - 2;
        cp_async16(buf + ll * LP + 2 * j, lp + xs);
      }
    }
    cp_async_commit();
  };
  const long step = (long)gridDim.x * NLB;
  long L0 = (long)blockIdx.x * NLB;
  int b = 0;
  if (L0 < a.nfast) stage(L0, S);
  for (; L0 < a.nfast; L0 += step, b ^= 1) {
    const double *cur = S + (size_t)b * NLB * LP;
    if (L0 + step < a.nfast) {
      stage(L0 + step, S + (size_t)(b ^ 1) * NLB * LP);
      cp_async_wait_but_one();
    } else {
      cp_async_wait_all();
    }
    __syncthreads();
#pragma unroll 2
    for (int q = threadIdx.x; q < NLB * half; q += blockDim.x) {
      const int ll = q / half, j = q - ll * half, x = 2 * j;
      const long L = L0 + ll;
      const double *sl = cur + ll * LP;
      double w[10];  // v[x-4 .. x+5]
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const double2 p2 = *reinterpret_cast<const double2 *>(sl + x + 2 * t);
        w[2 * t] = p2.x;
        w[2 * t + 1] = p2.y;
      }
      double r0 = rhs_center<FAM>(w + (4 - H), a.ari), r1 = rhs_center<FAM>(w + (5 - H), a.ari);
      if (j < 2 || j >= half - 2) {
        if (a.phys_lo && x < 4) {  // rows 0..3: closure weights on v[0..8]
          double r4[4];
          rhs_lo4<FAM>(sl + 4, a.arb_lo, r4);
          r0 = x == 0 ? r4[0] : r4[2];
          r1 = x == 0 ? r4[1] : r4[3];
        }
        if (a.phys_hi && x >= m - 4) {  // rows m-4..m-1: closure weights on v[m-8..m-1]
          double r4[4];
          rhs_hi4<FAM>(sl + 4 + m - 8, a.arb_hi, r4);
          r0 = x == m - 4 ? r4[0] : r4[2];
          r1 = x == m - 4 ? r4[1] : r4[3];
        }
      }
      double2 val = make_double2(r0 * scale, r1 * scale);
      if (ADDV) { val.x += w[4]; val.y += w[5]; }
      if (L < a.nfast) {
        const long idx = L * (long)m + x;
        if (PLAIN) {
          *reinterpret_cast<double2 *>(out + idx) = val;
        } else {
          epi_store(out, idx, val.x, epi);
          epi_store(out, idx + 1, val.y, epi);
        }
      }
    }
    __syncthreads();  // everyone is done with `cur` before it is staged again
  }
}

// ---- x sweep -------------------------------------------------------------------------------------
// Same algorithm on a tile of NLX unit-stride lines staged through shared memory (row pitch odd):
// coalesced tile load, in-place recurrences, coalesced write-back.
template <int FAM, int NLX, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(kBlockThreads, 3)
sweep_x_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v,
               double *__restrict__ out, const __grid_constant__ EpiArgs epi) {
  constexpr int H = FT<FAM>::H;
  PB_SHARED(S);  // [NLX][LD], then scan states
  const int m = a.m, LD = m | 1, C = a.C, P = a.P;
  double2 *EN = reinterpret_cast<double2 *>(S + (((size_t)NLX * LD + 1) & ~(size_t)1));  // [P][NLX]
  double2 *ST = EN + P * NLX;  // [P][NLX]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  const long nlines = a.nfast;
  const long L0 = (long)blockIdx.x * NLX;
  const bool implicit = a.implicit != 0;

  for (int ll = wid; ll < NLX; ll += nw) {  // stage the tile, coalesced
    long L = L0 + ll;
    if (L >= nlines) L = nlines - 1;
    const double *src = v + L * (long)m;
    double *dst = S + ll * LD;
    for (int ii = lane; ii < m; ii += 32) dst[ii] = __ldg(src + ii);
  }
  __syncthreads();

  const int l = tid % NLX, p = tid / NLX;
  const bool active = p < P;  // blockDim is rounded up to a warp multiple
  const int s = p * C;
  double *Sl = S + l * LD + s;  // this thread's chunk
  const int type = active ? a.ctype[p] : 0;
  const bool cc = a.has_const && type == 0;
  const bool lo_sp = a.phys_lo && p == 0, hi_sp = a.phys_hi && p == P - 1;

  // rows of the neighbouring chunks this thread's stencil needs, read before anyone overwrites them
  double hv[H], tv[H], u[8];
  if (active) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      int r = s - H + k;
      if (r < 0) r += m;
      hv[k] = lo_sp ? 0.0 : S[l * LD + r];
      r = s + C + k;
      if (r >= m) r -= m;
      tv[k] = hi_sp ? 0.0 : S[l * LD + r];
    }
    if (hi_sp) {
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = Sl[C - 8 + k];
    }
  }
  __syncthreads();

  if (active) {  // ---- A ----
    double ring[16];
    double rm1 = 0.0, rm2 = 0.0;
    const double2 *luf = a.luf + (size_t)type * C;
    const double l2c = a.cst[0], l1c = a.cst[1];
    auto emit = [&](int lr, double rhs) {
      if (implicit) {
        double2 c;
        if (cc) c = make_double2(l2c, l1c);
        else c = __ldg(luf + lr);
        double t = fma(-c.x, rm2, rhs);
        t = fma(-c.y, rm1, t);
        Sl[lr] = t;
        rm2 = rm1;
        rm1 = t;
      } else {
        Sl[lr] = rhs;
      }
    };
#pragma unroll
    for (int j = 0; j < H; ++j) ring[j] = hv[j];
#pragma unroll
    for (int j = H; j < 16; ++j) ring[j] = Sl[j - H];
    if ((C & 15) == 0) {
      double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
      if (lo_sp) {
        double vv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) vv[q] = ring[(q + H) & 15];
        rhs_lo4<FAM>(vv, a.arb_lo, rlo);
      }
      if (hi_sp) rhs_hi4<FAM>(u, a.arb_hi, rhi);
      for (int b = 0; b < C - 16; b += 16) {
        static_for<0, 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          double rhs = rhs_ring<FAM, k>(ring, a.ari);
          if (k < 4) {
            if (lo_sp && b == 0) rhs = rlo[k];
          }
          ring[k] = Sl[b + k - H + 16];
          emit(b + k, rhs);
        });
      }
      static_for<0, 16>([&](auto kc) {  // last block: the look-ahead rows come from the next chunk
        constexpr int k = decltype(kc)::value;
        double rhs = rhs_ring<FAM, k>(ring, a.ari);
        if (k < 4) {
          if (lo_sp && C == 16) rhs = rlo[k];
        }
        if (k >= 12) {
          if (hi_sp) rhs = rhi[k - 12];
        }
        if (k < H) ring[k] = Sl[C + k - H];
        else if (k < 2 * H) ring[k] = tv[k - H];
        emit(C - 16 + k, rhs);
      });
    } else {
      if (lo_sp) {
        double vv[9], r4[4];
#pragma unroll
        for (int k = 0; k < 9; ++k) vv[k] = ring[(k + H) & 15];
        rhs_lo4<FAM>(vv, a.arb_lo, r4);
#pragma unroll
        for (int k = 0; k < 4; ++k) emit(k, r4[k]);
      }
      for (int b = 0; b < C; b += 16) {
        static_for<0, 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          const int lr = b + k;
          double rhs = 0.0;
          const bool doit = lr < C && !(lo_sp && lr < 4) && !(hi_sp && lr >= C - 4);
          if (doit) rhs = rhs_ring<FAM, k>(ring, a.ari);
          const int r2 = lr - H + 16;
          double nx = 0.0;
          if (r2 < C) nx = Sl[r2];
          else {
#pragma unroll
            for (int q = 0; q < H; ++q)
              if (r2 - C == q) nx = tv[q];
          }
          ring[k] = nx;
          if (doit) emit(lr, rhs);
        });
      }
      if (hi_sp) {
        double r4[4];
        rhs_hi4<FAM>(u, a.arb_hi, r4);
#pragma unroll
        for (int k = 0; k < 4; ++k) emit(C - 4 + k, r4[k]);
      }
    }
    if (implicit) EN[p * NLX + l] = make_double2(rm1, rm2);
  }
  if (implicit) {
    __syncthreads();
    if (active) {  // ---- B ----
      double2 st = make_double2(0.0, 0.0);
      {
        const int nf = a.nf[p];
        const double4 *Mp = a.Mf + (size_t)p * (P + 1);
        for (int j = 1; j <= nf; ++j) {
          int q = p - j;
          if (q < 0) q += P;
          const double2 en = EN[q * NLX + l];
          const double4 M = ldg4(Mp + j);
          st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
          st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
        }
      }
      const double2 *ph = a.phi + (size_t)type * C;
      double x1 = 0.0, x2 = 0.0;
      if (cc) {
        const double ip = a.cst[2], u1 = a.cst[3], u2 = a.cst[4];
#pragma unroll 8
        for (int r = C - 1; r >= 0; --r) {
          const double2 f = __ldg(ph + r);
          double t = Sl[r];
          t = fma(f.x, st.x, t);
          t = fma(f.y, st.y, t);
          t = fma(-u1, x1, t);
          t = fma(-u2, x2, t);
          t *= ip;
          Sl[r] = t;
          x2 = x1;
          x1 = t;
        }
      } else {
        const double4 *lub = a.lub + (size_t)type * C;
#pragma unroll 4
        for (int r = C - 1; r >= 0; --r) {
          const double2 f = __ldg(ph + r);
          const double4 c = ldg4(lub + r);
          double t = Sl[r];
          t = fma(f.x, st.x, t);
          t = fma(f.y, st.y, t);
          t = fma(-c.y, x1, t);
          t = fma(-c.z, x2, t);
          t *= c.x;
          Sl[r] = t;
          x2 = x1;
          x1 = t;
        }
      }
      ST[p * NLX + l] = make_double2(x1, x2);
    }
    __syncthreads();
    if (active) {  // ---- D (in place; the coalesced write-back follows) ----
      auto tin = [&](int q) -> double2 {
        double2 t = make_double2(0.0, 0.0);
        const int nb = a.nb[q];
        const double4 *Mp = a.Mb + (size_t)q * (P + 1);
        for (int j = 1; j <= nb; ++j) {
          int qq = q + j;
          if (qq >= P) qq -= P;
          const double2 sv = ST[qq * NLX + l];
          const double4 M = ldg4(Mp + j);
          t.x = fma(M.y, sv.y, fma(M.x, sv.x, t.x));
          t.y = fma(M.w, sv.y, fma(M.z, sv.x, t.y));
        }
        return t;
      };
      const double2 tb = tin(p);
      const double2 *ps = a.psi + (size_t)type * C;
      auto rowDx = [&](double2 g, int r) {
        double x = Sl[r];
        x = fma(g.x, tb.x, x);
        x = fma(g.y, tb.y, x);
        Sl[r] = x;
      };
      if (cc && a.cparam && C == 32) {
        static_for<0, 32>([&](auto rc) { constexpr int r = decltype(rc)::value; rowDx(a.psi0[r], r); });
      } else if (cc && a.cparam && C == 16) {
        static_for<0, 16>([&](auto rc) { constexpr int r = decltype(rc)::value; rowDx(a.psi0[r], r); });
      } else {
#pragma unroll 8
        for (int r = 0; r < C; ++r) rowDx(__ldg(ps + r), r);
      }
    }
  }
  __syncthreads();

  const double scale = a.scale;  // write back, coalesced, with scale / add-back / epilogue
  for (int ll = wid; ll < NLX; ll += nw) {
    const long L = L0 + ll;
    if (L >= nlines) break;
    const double *src = S + ll * LD;
    for (int ii = lane; ii < m; ii += 32) {
      const long idx = L * (long)m + ii;
      double val = src[ii] * scale;
      if (ADDV) val += __ldg(v + idx);
      put<PLAIN>(out, idx, val, epi);
    }
  }
}

// ==== register-resident sweeps ====================================================================
// When the chunk length is 16 or 32 the whole chunk of a thread (its forward-eliminated rows, then
// its solution) stays in registers between the passes: shared memory only carries the two-value
// states exchanged between chunks.  y/z: one global read and one global write per point and
// nothing else on the load/store pipe; x: the tile is staged once and written back once.

// like stream_chunk's block loop with every row index known at compile time
template <int FAM, int CT, class EMIT>
__device__ __forceinline__ void stream_chunk_static(const SweepDev &a, const double *__restrict__ vp, long rs, int p,
                                                    const double *__restrict__ lo_rows, const double *__restrict__ hi_rows,
                                                    EMIT &&emit) {
  constexpr int H = FT<FAM>::H;
  const int m = a.m, P = a.P;
  const int s = p * CT;
  const bool first = p == 0, last = p == P - 1;
  const bool lo_sp = a.phys_lo && first, hi_sp = a.phys_hi && last;
  double ring[16];
  const double *pl = first ? (a.wrap ? vp + (long)(m - H) * rs : (lo_rows ? lo_rows : vp)) : vp + (long)(s - H) * rs;
#pragma unroll
  for (int j = 0; j < H; ++j) { ring[j] = __ldg(pl); pl += rs; }
  if (first) pl = vp;
#pragma unroll
  for (int j = H; j < 16; ++j) { ring[j] = __ldg(pl); pl += rs; }
  double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
  if (lo_sp) {
    double vv[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) vv[q] = ring[(q + H) & 15];
    rhs_lo4<FAM>(vv, a.arb_lo, rlo);
  }
  const double *pend = a.wrap ? vp : (hi_rows ? hi_rows : vp);
  static_for<0, CT / 16>([&](auto bc) {
    constexpr int b = decltype(bc)::value * 16;
    constexpr bool lastblk = b == CT - 16;
    static_for<0, 16>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      double rhs = rhs_ring<FAM, k>(ring, a.ari);
      const double vc = ring[(k + H) & 15];
      if (b == 0 && k < 4) {
        if (lo_sp) rhs = rlo[k];
      }
      if (lastblk && k == 12) {
        if (hi_sp) {
          double u[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) u[q] = ring[(q + H + 8) & 15];
          rhs_hi4<FAM>(u, a.arb_hi, rhi);
        }
      }
      if (lastblk && k >= 12) {
        if (hi_sp) rhs = rhi[k - 12];
      }
      if (lastblk && k == H) {
        if (last) pl = pend;  // the next row to load is row m
      }
      if (!(lastblk && k >= 2 * H)) {  // rows past s + CT + H - 1 are never used
        ring[k] = __ldg(pl);
        pl += rs;
      }
      emit(std::integral_constant<int, b + k>{}, rhs, vc);
    });
  });
}

template <int FAM, int CT, int NL, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(kBlockThreads, 2)
sweep_yz_reg_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v,
                    double *__restrict__ out, const double *__restrict__ halo_lo,
                    const double *__restrict__ halo_hi, double *__restrict__ iface,
                    const __grid_constant__ EpiArgs epi) {
  PB_SHARED(S);
  const int P = a.P;
  double2 *EN = reinterpret_cast<double2 *>(S);  // [P][NL] local end values of the forward pass
  double2 *ST = EN + P * NL;                      // [P][NL] local start values of the backward pass
  const int tid = threadIdx.x, l = tid % NL, slot = tid / NL, p = slot < a.P ? a.perm[slot] : slot;
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const int ti = blockIdx.x % tiles_i, o = blockIdx.x / tiles_i;
  int i0 = ti * NL + l;
  const bool valid = i0 < a.nfast;
  if (!valid) i0 = a.nfast - 1;
  const long rs = a.rstride;
  const long base = (long)i0 + (long)o * a.ostride;
  const double *vp = v + base;
  const int s = p * CT;
  const int type = a.ctype[p];
  const bool cc = a.has_const && type == 0;
  const double scale = a.scale;
  double rl[CT];

  {  // ---- A: rhs + forward recurrence (zero incoming state) ----
    const double2 *luf = a.luf + (size_t)type * CT;
    const double l2c = a.cst[0], l1c = a.cst[1];
    double rm1 = 0.0, rm2 = 0.0;
    stream_chunk_static<FAM, CT>(a, vp, rs, p, halo_lo ? halo_lo + base : nullptr, halo_hi ? halo_hi + base : nullptr,
                                 [&](auto lrc, double rhs, double) {
                                   constexpr int lr = decltype(lrc)::value;
                                   double2 c;
                                   if (cc) c = make_double2(l2c, l1c);
                                   else c = __ldg(luf + lr);
                                   double t = fma(-c.x, rm2, rhs);
                                   t = fma(-c.y, rm1, t);
                                   rl[lr] = t;
                                   rm2 = rm1;
                                   rm1 = t;
                                 });
    EN[p * NL + l] = make_double2(rm1, rm2);
  }
  __syncthreads();

  {  // ---- B: add the carried forward state, backward recurrence (zero incoming state) ----
    double2 st = make_double2(0.0, 0.0);
    {
      const int nf = a.nf[p];
      const double4 *Mp = a.Mf + (size_t)p * (P + 1);
      for (int j = 1; j <= nf; ++j) {
        int q = p - j;
        if (q < 0) q += P;
        const double2 en = EN[q * NL + l];
        const double4 M = ldg4(Mp + j);
        st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
        st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
      }
    }
    double x1 = 0.0, x2 = 0.0;
    if (cc) {
      const double ip = a.cst[2], u1 = a.cst[2] * a.cst[3], u2 = a.cst[2] * a.cst[4];
      static_for<0, CT>([&](auto jc) {
        constexpr int r = CT - 1 - decltype(jc)::value;
        double t = rl[r];
        t = fma(a.phi0[r].x, st.x, t);
        t = fma(a.phi0[r].y, st.y, t);
        t = fma(-u2, x2, t * ip);
        t = fma(-u1, x1, t);
        rl[r] = t;
        x2 = x1;
        x1 = t;
      });
    } else {
      const double2 *ph = a.phi + (size_t)type * CT;
      const double4 *lub = a.lub + (size_t)type * CT;
      static_for<0, CT>([&](auto jc) {
        constexpr int r = CT - 1 - decltype(jc)::value;
        const double2 f = __ldg(ph + r);
        const double4 c = ldg4(lub + r);
        double t = rl[r];
        t = fma(f.x, st.x, t);
        t = fma(f.y, st.y, t);
        t = fma(-c.z, x2, t * c.x);  // lub = {1/pivot, u1/pivot, u2/pivot}: one operation on the chain through x1
        t = fma(-c.y, x1, t);
        rl[r] = t;
        x2 = x1;
        x1 = t;
      });
    }
    ST[p * NL + l] = make_double2(x1, x2);
  }
  __syncthreads();

  {  // ---- D: add the carried backward state, scale / add-back / epilogue, store ----
    double2 tb = make_double2(0.0, 0.0);
    {
      const int nb = a.nb[p];
      const double4 *Mp = a.Mb + (size_t)p * (P + 1);
      for (int j = 1; j <= nb; ++j) {
        int q = p + j;
        if (q >= P) q -= P;
        const double2 sv = ST[q * NL + l];
        const double4 M = ldg4(Mp + j);
        tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
        tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
      }
    }
    const double *pv = vp + (long)s * rs;
    long oidx = base + (long)s * rs;
    double *po = out + oidx;
    auto rowD = [&](double gx, double gy, double xl) -> double {
      double x = fma(gx, tb.x, xl);
      x = fma(gy, tb.y, x);
      double val = x * scale;
      if (ADDV) val += __ldg(pv);
      if (PLAIN) {
        if (valid) *po = val;
        po += rs;
      } else {
        if (valid) epi_store(out, oidx, val, epi);
        oidx += rs;
      }
      pv += rs;
      return x;
    };
    double xi[4] = {0.0, 0.0, 0.0, 0.0};  // first / last two solved values (z-slab interface)
    if (cc) {
      static_for<0, CT>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const double x = rowD(a.psi0[r].x, a.psi0[r].y, rl[r]);
        if (r < 2) xi[r] = x;
        if (r >= CT - 2) xi[r - (CT - 4)] = x;
      });
    } else {
      const double2 *ps = a.psi + (size_t)type * CT;
      static_for<0, CT>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const double2 g = __ldg(ps + r);
        const double x = rowD(g.x, g.y, rl[r]);
        if (r < 2) xi[r] = x;
        if (r >= CT - 2) xi[r - (CT - 4)] = x;
      });
    }
    if (iface != nullptr && valid) {  // z-slab: this rank's 4 interface values, unscaled (compact_d1.f90:858-878)
      const long plane = (long)a.nfast * a.nouter;
      if (p == 0) {
        iface[base] = a.phys_lo ? 0.0 : xi[0];
        iface[plane + base] = a.phys_lo ? 0.0 : xi[1];
      }
      if (p == P - 1) {
        iface[2 * plane + base] = a.phys_hi ? 0.0 : xi[2];
        iface[3 * plane + base] = a.phys_hi ? 0.0 : xi[3];
      }
    }
  }
}

// x sweep, register resident: tile staged through shared memory for coalescing only
template <int FAM, int CT, int NLX, bool PLAIN, bool ADDV>
__global__ void __launch_bounds__(kBlockThreads, 2)
sweep_x_reg_kernel(const __grid_constant__ SweepDev a, const double *__restrict__ v,
                   double *__restrict__ out, const __grid_constant__ EpiArgs epi) {
  constexpr int H = FT<FAM>::H;
  PB_SHARED(S);  // [NLX][LD] tile, then the exchanged states
  const int m = a.m, LD = m | 1, P = a.P;
  double2 *EN = reinterpret_cast<double2 *>(S + (((size_t)NLX * LD + 1) & ~(size_t)1));  // [P][NLX]
  double2 *ST = EN + P * NLX;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  const long nlines = a.nfast;
  const long L0 = (long)blockIdx.x * NLX;

  for (int ll = wid; ll < NLX; ll += nw) {  // stage the tile, coalesced
    long L = L0 + ll;
    if (L >= nlines) L = nlines - 1;
    const double *src = v + L * (long)m;
    double *dst = S + ll * LD;
    for (int ii = lane; ii < m; ii += 32) dst[ii] = __ldg(src + ii);
  }
  __syncthreads();

  const int l = tid % NLX, slot = tid / NLX, p = slot < a.P ? a.perm[slot] : slot;
  const bool active = p < P;
  const int s = p * CT;
  double *Sl = S + l * LD;  // this thread's line
  const int type = active ? a.ctype[p] : 0;
  const bool cc = a.has_const && type == 0;
  const bool first = p == 0, last = p == P - 1;
  const bool lo_sp = a.phys_lo && first, hi_sp = a.phys_hi && last;
  double rl[CT];

  if (active) {  // ---- A: the tile is read-only in this phase, neighbours' rows are read in place ----
    double ring[16];
    const double2 *luf = a.luf + (size_t)type * CT;
    const double l2c = a.cst[0], l1c = a.cst[1];
    double rm1 = 0.0, rm2 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int r = s - H + j;
      if (r < 0) r += m;  // periodic wrap (closure: value unused)
      ring[j] = Sl[r];
    }
    double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
    if (lo_sp) {
      double vv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) vv[q] = ring[(q + H) & 15];
      rhs_lo4<FAM>(vv, a.arb_lo, rlo);
    }
    static_for<0, CT / 16>([&](auto bc) {
      constexpr int b = decltype(bc)::value * 16;
      constexpr bool lastblk = b == CT - 16;
      static_for<0, 16>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        constexpr int lr = b + k;
        double rhs = rhs_ring<FAM, k>(ring, a.ari);
        if (b == 0 && k < 4) {
          if (lo_sp) rhs = rlo[k];
        }
        if (lastblk && k == 12) {
          if (hi_sp) {
            double u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = ring[(q + H + 8) & 15];
            rhs_hi4<FAM>(u, a.arb_hi, rhi);
          }
        }
        if (lastblk && k >= 12) {
          if (hi_sp) rhs = rhi[k - 12];
        }
        if (!(lastblk && k >= 2 * H)) {
          int r = s + lr - H + 16;
          if (lastblk && k >= H) {  // rows of the next chunk; the last chunk wraps around
            if (r >= m) r -= m;
          }
          ring[k] = Sl[r];
        }
        double2 c;
        if (cc) c = make_double2(l2c, l1c);
        else c = __ldg(luf + lr);
        double t = fma(-c.x, rm2, rhs);
        t = fma(-c.y, rm1, t);
        rl[lr] = t;
        rm2 = rm1;
        rm1 = t;
      });
    });
    EN[p * NLX + l] = make_double2(rm1, rm2);
  }
  __syncthreads();

  if (active) {  // ---- B ----
    double2 st = make_double2(0.0, 0.0);
    {
      const int nf = a.nf[p];
      const double4 *Mp = a.Mf + (size_t)p * (P + 1);
      for (int j = 1; j <= nf; ++j) {
        int q = p - j;
        if (q < 0) q += P;
        const double2 en = EN[q * NLX + l];
        const double4 M = ldg4(Mp + j);
        st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
        st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
      }
    }
    double x1 = 0.0, x2 = 0.0;
    if (cc) {
      const double ip = a.cst[2], u1 = a.cst[3], u2 = a.cst[4];
      static_for<0, CT>([&](auto jc) {
        constexpr int r = CT - 1 - decltype(jc)::value;
        double t = rl[r];
        t = fma(a.phi0[r].x, st.x, t);
        t = fma(a.phi0[r].y, st.y, t);
        t = fma(-u1, x1, t);
        t = fma(-u2, x2, t);
        t *= ip;
        rl[r] = t;
        x2 = x1;
        x1 = t;
      });
    } else {
      const double2 *ph = a.phi + (size_t)type * CT;
      const double4 *lub = a.lub + (size_t)type * CT;
      static_for<0, CT>([&](auto jc) {
        constexpr int r = CT - 1 - decltype(jc)::value;
        const double2 f = __ldg(ph + r);
        const double4 c = ldg4(lub + r);
        double t = rl[r];
        t = fma(f.x, st.x, t);
        t = fma(f.y, st.y, t);
        t = fma(-c.y, x1, t);
        t = fma(-c.z, x2, t);
        t *= c.x;
        rl[r] = t;
        x2 = x1;
        x1 = t;
      });
    }
    ST[p * NLX + l] = make_double2(x1, x2);
  }
  __syncthreads();

  if (active) {  // ---- D: solution back into the tile ----
    double2 tb = make_double2(0.0, 0.0);
    {
      const int nb = a.nb[p];
      const double4 *Mp = a.Mb + (size_t)p * (P + 1);
      for (int j = 1; j <= nb; ++j) {
        int q = p + j;
        if (q >= P) q -= P;
        const double2 sv = ST[q * NLX + l];
        const double4 M = ldg4(Mp + j);
        tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
        tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
      }
    }
    double *So = Sl + s;
    if (cc) {
      static_for<0, CT>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        double x = fma(a.psi0[r].x, tb.x, rl[r]);
        So[r] = fma(a.psi0[r].y, tb.y, x);
      });
    } else {
      const double2 *ps = a.psi + (size_t)type * CT;
      static_for<0, CT>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const double2 g = __ldg(ps + r);
        double x = fma(g.x, tb.x, rl[r]);
        So[r] = fma(g.y, tb.y, x);
      });
    }
  }
  __syncthreads();

  const double scale = a.scale;  // write back, coalesced, with scale / add-back / epilogue
  for (int ll = wid; ll < NLX; ll += nw) {
    const long L = L0 + ll;
    if (L >= nlines) break;
    const double *src = S + ll * LD;
    for (int ii = lane; ii < m; ii += 32) {
      const long idx = L * (long)m + ii;
      double val = src[ii] * scale;
      if (ADDV) val += __ldg(v + idx);
      put<PLAIN>(out, idx, val, epi);
    }
  }
}


// true when every chunk handled by this warp has constant coefficients: a warp that mixes a table
// chunk and a constant chunk (16-line tiles, odd number of table chunks) runs the table path for
// both halves instead of the two paths one after the other
template <int NLT>
__device__ __forceinline__ bool warp_all_const(const SweepDev &a, int tid, bool cc) {
#ifdef PB_EMULATE
  if (NLT >= 32) return cc;
  const int per = 32 / NLT, s0 = (tid / NLT) / per * per;
  bool all = a.has_const != 0;
  for (int w = 0; w < per; ++w) all = all && a.ctype[a.perm[s0 + w]] == 0;
  return all;
#else
  return NLT < 32 ? __all_sync(0xffffffffu, cc) : cc;
#endif
}

// chunk of the other half of a warp (tiles of 16 lines put two chunks into one warp)
template <int NLT>
__device__ __forceinline__ int warp_other_chunk(const SweepDev &a, int tid, int p) {
#ifdef PB_EMULATE
  return NLT == 16 ? a.perm[((tid / NLT) & ~1) + 1] : p;
#else
  return NLT == 16 ? __shfl_sync(0xffffffffu, p, 16) : p;
#endif
}

// carried-state sum of a line whose chunks share one row of transfer products (SweepDev::mconst): the same terms
// in the same order as the table loop, with the products as constant-bank operands of the fmas
template <int NLT, bool FWD, int NT = kMTerms>
__device__ __forceinline__ double2 state_sum_const(const SweepDev &a, const double2 *E, int p, int P, int l) {
  double2 s = make_double2(0.0, 0.0);
  const int n = FWD ? a.nf0 : a.nb0;
  static_for<1, NT + 1>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    if (j <= n) {
      int q = FWD ? p - j : p + j;
      if (FWD) { if (q < 0) q += P; }
      else { if (q >= P) q -= P; }
      const double2 e = E[q * NLT + l];
      const double4 M = FWD ? a.Mf0[j - 1] : a.Mb0[j - 1];
      s.x = fma(M.y, e.y, fma(M.x, e.x, s.x));
      s.y = fma(M.w, e.y, fma(M.z, e.x, s.y));
    }
  });
  return s;
}

#ifndef PB_MCONST
#define PB_MCONST 1
#endif
#ifndef PB_MCONST_ALL  // experiment: the seven-point families too, with a four-term bound
#define PB_MCONST_ALL 0
#endif

// ---- pipelined sweeps: persistent CTAs, next tile prefetched by TMA --------------------------------
// The register kernels above only have loads in flight during their forward phase.  Here a CTA
// walks over tiles; as soon as every thread has pulled its chunk of the current tile out of shared
// memory (forward phase), one thread asks the TMA unit for the whole next tile, which lands while
// the CTA runs the backward recurrence and the store phase.  The tile buffer holds rows -4 .. m+3
// (periodic wrap rows or z-slab halo planes are extra boxes), so every chunk runs the same code with
// compile-time shared-memory offsets.

#ifndef PB_TAB_AHEAD
#define PB_TAB_AHEAD 4
#endif
constexpr int kTabAhead = PB_TAB_AHEAD;  // rows of table coefficients in flight in the backward pass of a table chunk

struct PipeGeo {
  int rowdim;          // tensor-map dimension the sweep runs along (1: y, 2: z)
  int nbox, box_rows;  // main boxes per tile
  int halo;            // 0: none (one-sided closures), 1: extra boxes for rows -4..-1 and m..m+3
  int lo_row, hi_row;  // row coordinates of those boxes in their maps
};

template <int FAM, int CT, int NL, class EMIT>
__device__ __forceinline__ void stream_chunk_tile(const SweepDev &a, const double *__restrict__ tw, bool lo_sp, bool hi_sp,
                                                  EMIT &&emit) {
  // tw points at row s-H of this thread's line inside the tile (row pitch NL)
  constexpr int H = FT<FAM>::H;
  double ring[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) ring[j] = tw[j * NL];
  double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
  if (lo_sp) {
    double vv[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) vv[q] = ring[(q + H) & 15];
    rhs_lo4<FAM>(vv, a.arb_lo, rlo);
  }
  static_for<0, CT / 16>([&](auto bc) {
    constexpr int b = decltype(bc)::value * 16;
    constexpr bool lastblk = b == CT - 16;
    static_for<0, 16>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      double rhs = rhs_ring<FAM, k>(ring, a.ari);
      const double vc = ring[(k + H) & 15];
      if (b == 0 && k < 4) {
        if (lo_sp) rhs = rlo[k];
      }
      if (lastblk && k == 12) {
        if (hi_sp) {
          double u[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) u[q] = ring[(q + H + 8) & 15];
          rhs_hi4<FAM>(u, a.arb_hi, rhi);
        }
      }
      if (lastblk && k >= 12) {
        if (hi_sp) rhs = rhi[k - 12];
      }
      if (!(lastblk && k >= 2 * H)) ring[k] = tw[(b + k + 16) * NL];  // rows past s + CT + H - 1 are never used
      emit(std::integral_constant<int, b + k>{}, rhs, vc);
    });
  });
}

// RING: the plain-store path writes |val| s^2 (ring detector with a constant length scale)
// TAB: the line has chunks that read coefficient tables (first / last chunks of a bounded line; periodic lines
// and interior z-slabs have none).  The table instantiation keeps the coefficient loads ahead of the recurrence
// with a register ring; the other one is the kernel as tuned on periodic lines.
template <int FAM, int NL, bool PLAIN, bool ADDV, bool LATE, bool RING, bool TAB>
__global__ void __launch_bounds__(kBlockThreads, 2)
sweep_yz_pipe_kernel(const __grid_constant__ SweepDev a, const __grid_constant__ TileMap tmain,
                     const __grid_constant__ TileMap tlo, const __grid_constant__ TileMap thi,
                     const __grid_constant__ TileMap tout, const __grid_constant__ PipeGeo g, const double *__restrict__ v, double *__restrict__ out,
                     double *__restrict__ iface, const __grid_constant__ EpiArgs epi) {
  constexpr int CT = 32, H = FT<FAM>::H, HP = 4;
  PB_SHARED(S);
  const int P = a.P, m = a.m;
  double *tile = S;                                                         // [m + 2 HP][NL]
  double *stage = S + (size_t)(m + 2 * HP) * NL;                            // [P][16][NL]: 16 rows of every chunk on their way out
  double2 *EN = reinterpret_cast<double2 *>(stage + (size_t)P * 16 * NL);    // [P][NL]
  double2 *ST = EN + P * NL;                                                // [P][NL]
  uint64_t *bar = reinterpret_cast<uint64_t *>(ST + P * NL);
  const int tid = threadIdx.x, l = tid % NL, p = a.perm[tid / NL];
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const long ntiles = (long)tiles_i * a.nouter;
  const long rs = a.rstride;
  const int s = p * CT;
  const int type = a.ctype[p];
  // lines without table chunks: the second / eighth derivative kernels run 6-10 % faster with the table code
  // compiled out, the first derivative and the filter kernels 7-9 % slower (register allocation at the 128
  // limit; gpurun logs r2ab / r2ac in profiles/r2_table_path_variants.log) -- so only the former drop it
  constexpr bool kNoTables = !TAB && FAM != F_D1 && !ADDV;
  // transfer products of the carried-state sums as constant-bank operands (SweepDev::Mf0 / Mb0) instead of loads:
  // the nine-point family gains 7-9 % (filter 1.44 -> 1.32 ms at 512^3), the first / second derivative kernels lose
  // 4-10 % with the same change (gpurun r2bb, profiles/r2_const_products_ab.log), so only the former take it
  constexpr bool kMConst = PB_MCONST && (FAM == F_R4 || PB_MCONST_ALL) && !TAB;
  constexpr int kMT = FAM == F_R4 ? kMTerms : 4;
  const bool cc = kNoTables ? true : warp_all_const<NL>(a, tid, a.has_const && type == 0);
  const bool lo_sp = a.phys_lo && p == 0, hi_sp = a.phys_hi && p == P - 1;
  const double scale = a.scale;
  const uint32_t tx_bytes = (uint32_t)((m + (g.halo ? 2 * HP : 0)) * NL * sizeof(double));

  auto issue = [&](long t) {  // one thread: arm the barrier, describe the tile to the TMA unit
    const int x0 = (int)(t % tiles_i) * NL, o = (int)(t / tiles_i);
    mbar_expect_tx(bar, tx_bytes);
    for (int b = 0; b < g.nbox; ++b) {
      const int row = b * g.box_rows;
      tma_load_3d(tile + (size_t)(HP + row) * NL, &tmain, x0, g.rowdim == 1 ? row : o, g.rowdim == 1 ? o : row, bar);
    }
    if (g.halo) {
      tma_load_3d(tile, &tlo, x0, g.rowdim == 1 ? g.lo_row : o, g.rowdim == 1 ? o : g.lo_row, bar);
      tma_load_3d(tile + (size_t)(HP + m) * NL, &thi, x0, g.rowdim == 1 ? g.hi_row : o, g.rowdim == 1 ? o : g.hi_row, bar);
    }
  };

  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  long t = blockIdx.x;
  if (tid == 0 && t < ntiles) issue(t);
#ifdef PB_EMULATE
  __syncthreads();
#else
  if (a.skew_ns > 0 && blockIdx.x >= gridDim.x / 2)
    for (int w = 0; w < a.skew_ns; w += 500) __nanosleep(500);
#endif
  uint32_t parity = 0;
  const double *tw = tile + (size_t)(s + HP - H) * NL + l;

  for (; t < ntiles; t += gridDim.x) {
    const int ti = (int)(t % tiles_i), o = (int)(t / tiles_i);
    int i0 = ti * NL + l;
    const bool valid = i0 < a.nfast;
    if (!valid) i0 = a.nfast - 1;
    const long base = (long)i0 + (long)o * a.ostride;
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
    }
    __syncthreads();  // the tile buffer is free (unless the add-back still reads it), EN is visible
    if (!(ADDV && LATE) && tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);

    double xloc[4] = {0.0, 0.0, 0.0, 0.0};  // late add-back: local solution of the four interface rows
    {  // ---- B: add the carried forward state, backward recurrence (zero incoming state) ----
      double2 st = make_double2(0.0, 0.0);
      if constexpr (kMConst) {  // the launcher sends lines without a shared row of products to the TAB kernel
        st = state_sum_const<NL, true, kMT>(a, EN, p, P, l);
      } else {
        const int nf = a.nf[p];
        const double4 *Mp = a.Mf + (size_t)p * (P + 1);
        for (int j = 1; j <= nf; ++j) {
          int q = p - j;
          if (q < 0) q += P;
          const double2 en = EN[q * NL + l];
          const double4 M = ldg4(Mp + j);
          st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
          st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
        }
      }
      double x1 = 0.0, x2 = 0.0;
      if (cc) {
        // x = ip (t - u1 x1 - u2 x2) with only one fma on the chain through x1
        const double ip = a.cst[2], u1 = a.cst[2] * a.cst[3], u2 = a.cst[2] * a.cst[4];
        static_for<0, CT>([&](auto jc) {
          constexpr int r = CT - 1 - decltype(jc)::value;
          double x = rl[r];
          x = fma(a.phi0[r].x, st.x, x);
          x = fma(a.phi0[r].y, st.y, x);
          x = fma(-u2, x2, x * ip);
          x = fma(-u1, x1, x);
          rl[r] = (ADDV && LATE) ? fma(x, scale, tw[(r + H) * NL]) : x;
          if (ADDV && LATE) {
            if (r < 2) xloc[r] = x;
            if (r >= CT - 2) xloc[r - (CT - 4)] = x;
          }
          x2 = x1;
          x1 = x;
        });
      } else {
        if constexpr (TAB) {
          // table chunks: the coefficients of row r - kTabAhead are requested while row r is solved (a row of the
          // chain is shorter than an L1 hit; left to the compiler the loads sat one row ahead and this warp took
          // 4.7x a constant warp through this phase, profiles/r2_bounded_table_path_ncu.txt)
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
            if (ADDV && LATE) {
              if (r < 2) xloc[r] = x;
              if (r >= CT - 2) xloc[r - (CT - 4)] = x;
            }
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
            if (ADDV && LATE) {
              if (r < 2) xloc[r] = x;
              if (r >= CT - 2) xloc[r - (CT - 4)] = x;
            }
            x2 = x1;
            x1 = x;
          });
        }
      }
      ST[p * NL + l] = make_double2(x1, x2);
    }
    // filters add the input back: it left the tile when the prefetch started, so it is re-read from
    // L2 through a 16-row register window whose first half is requested before the barrier
    const double *pv = v + base + (long)s * rs;
    double vr[16];
    if (ADDV && !LATE) {
#pragma unroll
      for (int k = 0; k < 16; ++k) vr[k] = __ldg(pv + k * rs);
    }
    // accumulating epilogues (div, laplacian, ring) read the previous output: same 16-row window
    const double *pold = out + base + (long)s * rs;
    const bool need_old = !PLAIN && epi_needs_old(epi);
    double ow[16];
    if (!PLAIN) {
#pragma unroll
      for (int k = 0; k < 16; ++k) ow[k] = need_old ? pold[k * rs] : 0.0;
    }
    __syncthreads();
    if (ADDV && LATE && tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);

    {  // ---- D: add the carried backward state, scale / add-back / epilogue, store ----
      double2 tb = make_double2(0.0, 0.0);
      if constexpr (kMConst) {
        tb = state_sum_const<NL, false, kMT>(a, ST, p, P, l);
      } else {
        const int nb = a.nb[p];
        const double4 *Mp = a.Mb + (size_t)p * (P + 1);
        for (int j = 1; j <= nb; ++j) {
          int q = p + j;
          if (q >= P) q -= P;
          const double2 sv = ST[q * NL + l];
          const double4 M = ldg4(Mp + j);
          tb.x = fma(M.y, sv.y, fma(M.x, sv.x, tb.x));
          tb.y = fma(M.w, sv.y, fma(M.z, sv.x, tb.y));
        }
      }
      long oidx = base + (long)s * rs;
      auto rowD = [&](auto rc, double gx, double gy, double xl) -> double {
        constexpr int r = decltype(rc)::value;
        double x = fma(gx, tb.x, xl);
        x = fma(gy, tb.y, x);
        double val = x * scale;
        if (ADDV && LATE) {  // xl already holds scale * x_local + v
          val = fma(gx * scale, tb.x, xl);
          val = fma(gy * scale, tb.y, val);
          if (r < 2 || r >= CT - 2) {
            const double xq = xloc[r < 2 ? r : r - (CT - 4)];
            x = fma(gy, tb.y, fma(gx, tb.x, xq));
          }
        }
        if (ADDV && !LATE) {
          val += vr[r & 15];
          if (r < CT - 16) vr[r & 15] = __ldg(pv + (long)(r + 16) * rs);
        }
        if (PLAIN) {  // plain stores leave through the stage and the TMA unit, 16 rows of every chunk at a time
          if constexpr (RING) val = fabs(val) * a.ring_s2;
          stage[(size_t)(p * 16 + (r & 15)) * NL + l] = val;
        } else {
          const double o = epi_value(val, ow[r & 15], oidx, epi);
          if (r < CT - 16) {
            if (need_old) ow[r & 15] = pold[(long)(r + 16) * rs];
          }
          if (valid) out[oidx] = o;
          oidx += rs;
        }
        return x;
      };
      double xi[4] = {0.0, 0.0, 0.0, 0.0};  // first / last two solved values (z-slab interface)
      const double2 *ps = a.psi + (size_t)type * CT;
      static_for<0, CT / 16>([&](auto gc) {
        constexpr int r0 = decltype(gc)::value * 16;
        if (PLAIN) {  // this warp's previous stores have left its part of the stage
          if ((tid & 31) == 0) tma_store_wait_read();
          __syncwarp();
        }
        if (cc) {
          static_for<r0, r0 + 16>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const double x = rowD(rc, a.psi0[r].x, a.psi0[r].y, rl[r]);
            if (r < 2) xi[r] = x;
            if (r >= CT - 2) xi[r - (CT - 4)] = x;
          });
        } else {
          // plain stores of a table instantiation: the 16 rows of the group at once, nothing waits on a load in the
          // middle of the stores (the read-modify-write epilogues already hold a 16-row window of the old output)
          constexpr int NQ = (PLAIN && TAB) ? 16 : 1;
          double2 gq[NQ];
          if constexpr (PLAIN && TAB) {
#pragma unroll
            for (int k = 0; k < 16; ++k) gq[k] = __ldg(ps + r0 + k);
          }
          static_for<r0, r0 + 16>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (!(PLAIN && TAB)) gq[0] = __ldg(ps + r);
            const double x = rowD(rc, gq[(PLAIN && TAB) ? r - r0 : 0].x, gq[(PLAIN && TAB) ? r - r0 : 0].y, rl[r]);
            if (r < 2) xi[r] = x;
            if (r >= CT - 2) xi[r - (CT - 4)] = x;
          });
        }
        if (PLAIN) {  // every warp hands the rows of its own chunks to the TMA unit: no block-wide barrier
          fence_async_smem();
          __syncwarp();
          if ((tid & 31) == 0) {
            const int x0 = ti * NL;
#pragma unroll
            for (int w = 0; w < 32 / NL; ++w) {
              const int q = a.perm[tid / NL + w];
              const int row = q * CT + r0;
              const int c1 = g.rowdim == 1 ? row : o, c2 = g.rowdim == 1 ? o : row;
              if constexpr (RING) {
                if (a.acc == 2) tma_reduce_max_3d(&tout, x0, c1, c2, stage + (size_t)q * 16 * NL);
                else tma_store_3d(&tout, x0, c1, c2, stage + (size_t)q * 16 * NL);
              } else {
                if (a.acc) tma_reduce_add_3d(&tout, x0, c1, c2, stage + (size_t)q * 16 * NL);
                else tma_store_3d(&tout, x0, c1, c2, stage + (size_t)q * 16 * NL);
              }
            }
            tma_store_commit();
          }
        }
      });
      if (iface != nullptr && valid) {  // z-slab: this rank's 4 interface values, unscaled (compact_d1.f90:858-878)
        const long plane = (long)a.nfast * a.nouter;
        if (p == 0) {
          iface[base] = a.phys_lo ? 0.0 : xi[0];
          iface[plane + base] = a.phys_lo ? 0.0 : xi[1];
        }
        if (p == P - 1) {
          iface[2 * plane + base] = a.phys_hi ? 0.0 : xi[2];
          iface[3 * plane + base] = a.phys_hi ? 0.0 : xi[3];
        }
      }
    }
  }
  if (PLAIN && ((tid & 31) == 0)) tma_store_wait_read();  // covers both store modes (thread 0 is a lane 0)
}

// Builds the tensor maps of one y/z sweep and launches the persistent kernel.  Returns
// cudaErrorNotSupported when the geometry does not fit (the caller then uses the register kernels).
template <int FAM, int NL, bool PLAIN, bool ADDV>
static cudaError_t launch_yz_pipe(const SweepDev &a, const double *v, double *out, const double *hlo, const double *hhi,
                                  double *iface, const EpiArgs &epi, cudaStream_t st) {
  constexpr int H = FT<FAM>::H;
  const int m = a.m;
  if (a.C != 32 || NL * a.P != kBlockThreads || m % 256 != 0 || (hlo == nullptr) != (hhi == nullptr)) return cudaErrorNotSupported;
  // the field as {ax, ay, az}: y sweeps run along dimension 1, z sweeps along dimension 2
  const bool ysweep = a.rstride < a.ostride;
  const uint64_t ax = (uint64_t)a.nfast;
  const uint64_t d1 = ysweep ? (uint64_t)m : (uint64_t)a.nouter, d2 = ysweep ? (uint64_t)a.nouter : (uint64_t)m;
  const uint64_t s1 = (uint64_t)(ysweep ? a.rstride : a.ostride) * 8, s2 = (uint64_t)(ysweep ? a.ostride : a.rstride) * 8;
  PipeGeo g;
  g.rowdim = ysweep ? 1 : 2;
  g.box_rows = 256;
  g.nbox = m / 256;
  g.halo = 0; g.lo_row = 0; g.hi_row = 0;
  TileMap tmain, tlo, thi, tout;
  if (!encode_tile_map(&tmain, v, ax, d1, d2, s1, s2, NL, ysweep ? 256 : 1, ysweep ? 1 : 256)) return cudaErrorNotSupported;
  tlo = tmain; thi = tmain; tout = tmain;
  if (PLAIN && !encode_tile_map(&tout, out, ax, d1, d2, s1, s2, NL, ysweep ? 16 : 1, ysweep ? 1 : 16, false, a.acc == 2))
    return cudaErrorNotSupported;
  if (hlo != nullptr) {  // z-slab halo planes received from the neighbours: {ax, ay, H} each
    if (ysweep) return cudaErrorNotSupported;
    if (!encode_tile_map(&tlo, hlo, ax, d1, H, s1, s2, NL, 1, 4) || !encode_tile_map(&thi, hhi, ax, d1, H, s1, s2, NL, 1, 4))
      return cudaErrorNotSupported;
    g.halo = 1; g.lo_row = H - 4; g.hi_row = 0;
  } else if (a.wrap) {  // periodic: rows m-4..m-1 and 0..3 of the field itself
    if (!encode_tile_map(&tlo, v, ax, d1, d2, s1, s2, NL, ysweep ? 4 : 1, ysweep ? 1 : 4)) return cudaErrorNotSupported;
    thi = tlo;
    g.halo = 1; g.lo_row = m - 4; g.hi_row = 0;
  }
  const size_t smem = ((size_t)(m + 8) * NL + 16 * (size_t)a.P * NL + 4 * (size_t)a.P * NL) * sizeof(double) + 16;
  static const bool late = getenv("PB_ADDV_LATE") ? atoi(getenv("PB_ADDV_LATE")) != 0 : true;
  void (*kfn)(SweepDev, TileMap, TileMap, TileMap, TileMap, PipeGeo, const double *, double *, double *, EpiArgs) = nullptr;
  int slot;
  bool tab = !a.has_const;  // any chunk on the table path?
  for (int q = 0; q < a.P; ++q) tab = tab || a.ctype[q] != 0;
  if (PB_MCONST && FAM == F_R4 && !a.mconst) tab = true;  // the other instantiation takes its transfer products from the parameters
  if (PB_MCONST_ALL && FAM != F_R4 && (!a.mconst || a.nf0 > 4 || a.nb0 > 4)) tab = true;
#define PB_PIPE_PICK(LATEV, RINGV) (tab ? sweep_yz_pipe_kernel<FAM, NL, PLAIN, ADDV, LATEV, RINGV, true> : sweep_yz_pipe_kernel<FAM, NL, PLAIN, ADDV, LATEV, RINGV, false>)
  if (a.ring) {
    if constexpr (PLAIN && !ADDV && FAM == F_R4) kfn = PB_PIPE_PICK(false, true);
    slot = 2;
  } else if (ADDV && late) {
    kfn = PB_PIPE_PICK(true, false); slot = 1;
  } else {
    kfn = PB_PIPE_PICK(false, false); slot = 0;
  }
#undef PB_PIPE_PICK
  if (kfn == nullptr) return cudaErrorNotSupported;
  slot = 2 * slot + (tab ? 1 : 0);
  static bool configured[6] = {false, false, false, false, false, false};
  if (!configured[slot]) {
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured[slot] = true;
  }
  const long ntiles = (long)((a.nfast + NL - 1) / NL) * a.nouter;
  const long want = 2L * sm_count();
  const dim3 grid((unsigned)(ntiles < want ? ntiles : want)), block(kBlockThreads);
  PB_LAUNCH(kfn, grid, block, smem, st, a, tmain, tlo, thi, tout, g, v, out, iface, epi);
  ++g_launches;
  ++g_pipe_launches;
  return cudaGetLastError();
}


// x sweep, pipelined.  Lines are unit stride, so a tile of NLX lines is described to the TMA unit as
// boxes of 16 points x NLX lines (128-byte rows) with the hardware 128-byte swizzle: the 16-byte
// piece c of row r lands at piece c ^ (r & 7), so the eight lanes of a quarter warp -- eight lines,
// same x -- hit eight different bank groups and a thread reads its chunk with conflict-free 128-bit
// shared loads without any padding.  Box slot b of the tile holds x = 16 (b - 1) .. 16 b - 1 (slot 0
// and the last slot are the periodic wrap).  The solution leaves the same way: 16 rows of every
// chunk at a time are written, swizzled, into a stage of P boxes and one thread hands them to the
// TMA unit as stores; composite epilogues read the stage back and store with the old output.
// MODE 0: implicit operator; 1: implicit with the late add-back (filters); 2: explicit operator (stencil only)
// ACC 0: store, 1: out += val, 2: out = max(out, |val| s^2), 3: out = |val| s^2 (ring detector)
template <int FAM, int NLX, bool PLAIN, bool ADDV, int MODE, int ACC>
__global__ void __launch_bounds__(kBlockThreads, 2)
sweep_x_pipe_kernel(const __grid_constant__ SweepDev a, const __grid_constant__ TileMap tin,
                    const __grid_constant__ TileMap tout, const double *__restrict__ v, double *__restrict__ out,
                    const __grid_constant__ EpiArgs epi) {
  constexpr int CT = 32, H = FT<FAM>::H, G = 16;
  constexpr bool LATE = MODE == 1, implicit = MODE != 2;
  // who issues the TMA stores: lane 0 of every warp for its own chunks (no block-wide barrier in the
  // store phase; measured faster), or one thread for the whole block
  constexpr bool WSTORE = true;
  constexpr int PC = kBlockThreads / NLX, M = PC * CT;  // chunks per line, line length
  constexpr int BOXB = NLX * 128;                       // bytes of one box
  constexpr int NBOX = M / 16;
  PB_SHARED(S);
  constexpr int m = M, P = PC;
  char *tile = reinterpret_cast<char *>(S);                                   // [NBOX + 2] boxes
  char *stage = tile + (size_t)(NBOX + 2) * BOXB;                            // [P] boxes
  double2 *EN = reinterpret_cast<double2 *>(stage + (size_t)P * BOXB);       // [P][NLX]
  double2 *ST = EN + P * NLX;
  uint64_t *bar = reinterpret_cast<uint64_t *>(ST + P * NLX);
  const int tid = threadIdx.x, l = tid % NLX, p = a.perm[tid / NLX];
  const long nlines = a.nfast;
  const long ntiles = (nlines + NLX - 1) / NLX;
  const int type = a.ctype[p];
  const bool cc = warp_all_const<NLX>(a, tid, a.has_const && type == 0);
  const bool lo_sp = a.phys_lo && p == 0, hi_sp = a.phys_hi && p == P - 1;
  const double scale = a.scale;
  const uint32_t tx_bytes = (uint32_t)((NBOX + (a.wrap ? 2 : 0)) * BOXB);

  auto issue = [&](long t) {  // one thread: the whole tile as boxes
    const int L0 = (int)(t * NLX);
    mbar_expect_tx(bar, tx_bytes);
    for (int b = 0; b < NBOX; ++b) tma_load_3d(tile + (size_t)(b + 1) * BOXB, &tin, 16 * b, L0, 0, bar);
    if (a.wrap) {
      tma_load_3d(tile, &tin, m - 16, L0, 0, bar);
      tma_load_3d(tile + (size_t)(NBOX + 1) * BOXB, &tin, 0, L0, 0, bar);
    }
  };
  // this thread's row inside a box and the swizzled offsets of the eight 16-byte pieces of that row
  const int rowoff = l * 128, swz = (l & 7) << 4;
  const char *tb = tile + (size_t)(2 * p) * BOXB + rowoff;  // slot 2p: x = 32 p - 16 ..
  // pair j holds x = 32 p - 4 + 2 j, 32 p - 3 + 2 j  (columns 12 + 2 j of the window starting at slot 2p)
  auto pair = [&](auto jc) -> double2 {
    constexpr int col = 12 + 2 * decltype(jc)::value;
    return *reinterpret_cast<const double2 *>(tb + (col >> 4) * BOXB + ((((col & 15) >> 1) << 4) ^ swz));
  };

  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  long t = blockIdx.x;
  if (tid == 0 && t < ntiles) issue(t);
#ifdef PB_EMULATE
  __syncthreads();
#else
  if (a.skew_ns > 0 && blockIdx.x >= gridDim.x / 2)
    for (int w = 0; w < a.skew_ns; w += 500) __nanosleep(500);
#endif
  uint32_t parity = 0;

  for (; t < ntiles; t += gridDim.x) {
    const long L0 = t * NLX;
    double rl[CT];
    mbar_wait(bar, parity);
    parity ^= 1;
#ifdef PB_EMULATE
    __syncthreads();
#endif

    {  // ---- A: rhs + forward recurrence; ring slot = (x - 32 p + 4) & 15 ----
      const double2 *luf = a.luf + (size_t)type * CT;
      const double l2c = a.cst[0], l1c = a.cst[1];
      double rm1 = 0.0, rm2 = 0.0;
      double ring[16];
      static_for<0, 8>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const double2 w2 = pair(jc);
        ring[2 * j] = w2.x;
        ring[2 * j + 1] = w2.y;
      });
      double rlo[4] = {0.0, 0.0, 0.0, 0.0}, rhi[4] = {0.0, 0.0, 0.0, 0.0};
      if (lo_sp) {
        double vv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) vv[q] = ring[(q + 4) & 15];
        rhs_lo4<FAM>(vv, a.arb_lo, rlo);
      }
      static_for<0, CT>([&](auto lrc) {
        constexpr int lr = decltype(lrc)::value;
        double rhs = rhs_ring<FAM, (lr + 4 - H) & 15>(ring, a.ari);
        if (lr < 4) {
          if (lo_sp) rhs = rlo[lr];
        }
        if (lr == CT - 4) {
          if (hi_sp) {
            double u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = ring[(CT - 4 + q) & 15];
            rhs_hi4<FAM>(u, a.arb_hi, rhi);
          }
        }
        if (lr >= CT - 4) {
          if (hi_sp) rhs = rhi[lr - (CT - 4)];
        }
        if constexpr ((lr & 1) && lr + 16 < CT + 8) {  // window columns lr+15, lr+16 replace the two just retired
          const double2 w2 = pair(std::integral_constant<int, (lr + 15) / 2>{});
          ring[(lr - 1) & 15] = w2.x;
          ring[lr & 15] = w2.y;
        }
        if constexpr (!implicit) {  // explicit operators (the Gaussian filter) are complete after the stencil
          const double vc = ring[(lr + 4) & 15];
          rl[lr] = ADDV ? fma(rhs, scale, vc) : rhs * scale;
        } else {
          double2 c;
          if (cc) c = make_double2(l2c, l1c);
          else c = __ldg(luf + lr);
          double x = fma(-c.x, rm2, rhs);
          x = fma(-c.y, rm1, x);
          rl[lr] = x;
          rm2 = rm1;
          rm1 = x;
        }
      });
      if constexpr (implicit) EN[p * NLX + l] = make_double2(rm1, rm2);
    }
    __syncthreads();  // tile consumed (unless the add-back still reads it), EN visible
    if (!(ADDV && LATE) && tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);

    if constexpr (implicit) {  // ---- B ----
      double2 st = make_double2(0.0, 0.0);
      if (PB_MCONST && FAM == F_R4 && a.mconst) {  // nine-point family only, as in the y / z kernel
        st = state_sum_const<NLX, true>(a, EN, p, P, l);
      } else {
        const int nf = a.nf[p];
        const double4 *Mp = a.Mf + (size_t)p * (P + 1);
        for (int j = 1; j <= nf; ++j) {
          int q = p - j;
          if (q < 0) q += P;
          const double2 en = EN[q * NLX + l];
          const double4 M = ldg4(Mp + j);
          st.x = fma(M.y, en.y, fma(M.x, en.x, st.x));
          st.y = fma(M.w, en.y, fma(M.z, en.x, st.y));
        }
      }
      double x1 = 0.0, x2 = 0.0;
      double2 vv2 = make_double2(0.0, 0.0);  // late add-back: v[r - 1], v[r] of the tile (r odd)
      auto rowB = [&](auto rc, double f0, double f1, double ip, double u1, double u2) {
        constexpr int r = decltype(rc)::value;
        double x = rl[r];
        x = fma(f0, st.x, x);
        x = fma(f1, st.y, x);
        x = fma(-u1, x1, x);
        x = fma(-u2, x2, x);
        x *= ip;
        if constexpr (ADDV && LATE) {
          if constexpr (r & 1) vv2 = pair(std::integral_constant<int, (r + 3) / 2>{});  // x = 32 p + r - 1, 32 p + r
          rl[r] = fma(x, scale, (r & 1) ? vv2.y : vv2.x);
        } else {
          rl[r] = x;
        }
        x2 = x1;
        x1 = x;
      };
      if (cc) {
        const double ip = a.cst[2], u1 = a.cst[3], u2 = a.cst[4];
        static_for<0, CT>([&](auto jc) {
          constexpr int r = CT - 1 - decltype(jc)::value;
          rowB(std::integral_constant<int, r>{}, a.phi0[r].x, a.phi0[r].y, ip, u1, u2);
        });
      } else {
        const double2 *ph = a.phi + (size_t)type * CT;
        const double4 *lub = a.lub + (size_t)type * CT;
        static_for<0, CT>([&](auto jc) {
          constexpr int r = CT - 1 - decltype(jc)::value;
          const double2 f = __ldg(ph + r);
          const double4 c = ldg4(lub + r);
          rowB(std::integral_constant<int, r>{}, f.x, f.y, c.x, c.y, c.z);
        });
      }
      ST[p * NLX + l] = make_double2(x1, x2);
    }
    if (PLAIN && !WSTORE && tid == 0) tma_store_wait_read();  // the stage of the previous tile has been read
    __syncthreads();
    if (ADDV && LATE && tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);

    {  // ---- D: carried backward state, then 16 rows of every chunk at a time through the stage ----
      double2 tbk = make_double2(0.0, 0.0);
      if constexpr (implicit) {
        if (PB_MCONST && FAM == F_R4 && a.mconst) {
          tbk = state_sum_const<NLX, false>(a, ST, p, P, l);
        } else {
          const int nb = a.nb[p];
          const double4 *Mp = a.Mb + (size_t)p * (P + 1);
          for (int j = 1; j <= nb; ++j) {
            int q = p + j;
            if (q >= P) q -= P;
            const double2 sv = ST[q * NLX + l];
            const double4 M = ldg4(Mp + j);
            tbk.x = fma(M.y, sv.y, fma(M.x, sv.x, tbk.x));
            tbk.y = fma(M.w, sv.y, fma(M.z, sv.x, tbk.y));
          }
        }
      }
      const double2 *ps = a.psi + (size_t)type * CT;
      char *sb = stage + (size_t)p * BOXB + rowoff;
      static_for<0, CT / G>([&](auto gc) {
        constexpr int g = decltype(gc)::value;
        if (PLAIN && WSTORE) {  // this warp's previous stores have left its part of the stage
          if ((tid & 31) == 0) tma_store_wait_read();
          __syncwarp();
        } else if (g > 0) {
          if (PLAIN && tid == 0) tma_store_wait_read();
          __syncthreads();
        }
        static_for<0, G / 2>([&](auto kc) {
          constexpr int kq = decltype(kc)::value, r = g * G + 2 * kq;
          double2 q0, q1;
          if constexpr (implicit) {
            if (cc) { q0 = a.psi0[r]; q1 = a.psi0[r + 1]; }
            else { q0 = __ldg(ps + r); q1 = __ldg(ps + r + 1); }
          } else {
            q0 = make_double2(0.0, 0.0); q1 = q0;
          }
          const double sc = (ADDV && LATE) ? scale : 1.0;  // late add-back: rl already holds scale * x + v
          double xa = fma(q0.x * sc, tbk.x, rl[r]);
          xa = fma(q0.y * sc, tbk.y, xa);
          double xb = fma(q1.x * sc, tbk.x, rl[r + 1]);
          xb = fma(q1.y * sc, tbk.y, xb);
          double2 ov = ((ADDV && LATE) || !implicit) ? make_double2(xa, xb) : make_double2(xa * scale, xb * scale);
          if constexpr (PLAIN && ACC >= 2) ov = make_double2(fabs(ov.x) * a.ring_s2, fabs(ov.y) * a.ring_s2);
          *reinterpret_cast<double2 *>(sb + ((kq << 4) ^ swz)) = ov;
        });
        if (PLAIN) {
          fence_async_smem();
          if constexpr (WSTORE) {
            __syncwarp();
            // every warp hands the boxes of its own chunks (lane 0's and, with 16-line tiles, lane 16's)
            // to the TMA unit
            const int q1 = warp_other_chunk<NLX>(a, tid, p);
            if ((tid & 31) == 0) {
              if constexpr (NLX >= 16) {
                if constexpr (ACC == 1) tma_reduce_add_3d(&tout, p * CT + g * G, (int)L0, 0, stage + (size_t)p * BOXB);
                else if constexpr (ACC == 2) tma_reduce_max_3d(&tout, p * CT + g * G, (int)L0, 0, stage + (size_t)p * BOXB);
                else tma_store_3d(&tout, p * CT + g * G, (int)L0, 0, stage + (size_t)p * BOXB);
                if (NLX == 16) {
                  if constexpr (ACC == 1) tma_reduce_add_3d(&tout, q1 * CT + g * G, (int)L0, 0, stage + (size_t)q1 * BOXB);
                  else if constexpr (ACC == 2) tma_reduce_max_3d(&tout, q1 * CT + g * G, (int)L0, 0, stage + (size_t)q1 * BOXB);
                  else tma_store_3d(&tout, q1 * CT + g * G, (int)L0, 0, stage + (size_t)q1 * BOXB);
                }
              } else {  // 8-line tiles: four chunks per warp
#pragma unroll
                for (int w = 0; w < 32 / NLX; ++w) {
                  const int q = a.perm[tid / NLX + w];
                  if constexpr (ACC == 1) tma_reduce_add_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
                  else if constexpr (ACC == 2) tma_reduce_max_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
                  else tma_store_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
                }
              }
              tma_store_commit();
            }
          } else {
            __syncthreads();
            if (tid == 0) {
              for (int q = 0; q < P; ++q) {
                if constexpr (ACC == 1) tma_reduce_add_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
                else if constexpr (ACC == 2) tma_reduce_max_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
                else tma_store_3d(&tout, q * CT + g * G, (int)L0, 0, stage + (size_t)q * BOXB);
              }
              tma_store_commit();
            }
          }
        } else {
          __syncthreads();
          constexpr int PIECES = NLX * PC * (G / 2);  // 16-byte pieces of a group
          constexpr int per_line = PC * (G / 2);
          const bool need_old = epi_needs_old(epi);
          double2 vadd[PIECES / kBlockThreads], oadd[PIECES / kBlockThreads];
#pragma unroll
          for (int it = 0; it < PIECES / kBlockThreads; ++it) {  // fetch add-back / previous output of the whole group first
            const int q = it * kBlockThreads + tid;
            const int line = q / per_line, rr = q - line * per_line;
            const long L = L0 + line;
            const long idx = (L < nlines ? L : nlines - 1) * (long)m + (rr >> 3) * CT + g * G + 2 * (rr & 7);
            if (ADDV && !LATE && implicit) vadd[it] = __ldg(reinterpret_cast<const double2 *>(v + idx));
            oadd[it] = need_old ? *reinterpret_cast<const double2 *>(out + idx) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int it = 0; it < PIECES / kBlockThreads; ++it) {
            const int q = it * kBlockThreads + tid;
            const int line = q / per_line, rr = q - line * per_line;
            const long L = L0 + line;
            const int chunk = rr >> 3, piece = rr & 7;
            double2 val = *reinterpret_cast<const double2 *>(stage + (size_t)chunk * BOXB + line * 128 + ((piece ^ (line & 7)) << 4));
            if (ADDV && !LATE && implicit) { val.x += vadd[it].x; val.y += vadd[it].y; }
            if (L < nlines) {
              const long idx = L * (long)m + chunk * CT + g * G + 2 * piece;
              *reinterpret_cast<double2 *>(out + idx) =
                  make_double2(epi_value(val.x, oadd[it].x, idx, epi), epi_value(val.y, oadd[it].y, idx + 1, epi));
            }
          }
        }
      });
    }
  }
  if (PLAIN && (tid & 31) == 0) tma_store_wait_read();
}

template <int FAM, int NLX, bool PLAIN, bool ADDV>
static cudaError_t launch_x_pipe(const SweepDev &a, const double *v, double *out, const EpiArgs &epi, cudaStream_t st) {
  const int m = a.m;
  if (a.C != 32 || NLX * a.P != kBlockThreads || m != 32 * a.P) return cudaErrorNotSupported;
  TileMap tin, tout;
  const uint64_t rowb = (uint64_t)m * 8;
  if (!encode_tile_map(&tin, v, (uint64_t)m, (uint64_t)a.nfast, 1, rowb, rowb * (uint64_t)a.nfast, 16, NLX, 1, true) ||
      !encode_tile_map(&tout, out, (uint64_t)m, (uint64_t)a.nfast, 1, rowb, rowb * (uint64_t)a.nfast, 16, NLX, 1, true, a.acc == 2))
    return cudaErrorNotSupported;
  const size_t smem = (size_t)(m / 16 + 2 + a.P) * NLX * 128 + 4 * (size_t)a.P * NLX * sizeof(double) + 16;
  static const bool late = getenv("PB_ADDV_LATE") ? atoi(getenv("PB_ADDV_LATE")) != 0 : true;
  void (*kfn)(SweepDev, TileMap, TileMap, const double *, double *, EpiArgs) = nullptr;
  int slot;
  if (!a.implicit) {
    if constexpr (FAM == F_R4 && ADDV) { kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 2, 0>; slot = 2; }
    else return cudaErrorNotSupported;
  } else if (ADDV && late) {
    kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 1, 0>; slot = 1;
  } else if (PLAIN && !ADDV && a.acc == 1) {
    if constexpr (PLAIN && !ADDV) kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 0, 1>;
    slot = 3;
  } else if (PLAIN && !ADDV && a.acc == 2) {
    if constexpr (PLAIN && !ADDV && FAM == F_R4) kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 0, 2>;
    slot = 4;
  } else if (PLAIN && !ADDV && a.ring) {
    if constexpr (PLAIN && !ADDV && FAM == F_R4) kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 0, 3>;
    slot = 5;
  } else {
    kfn = sweep_x_pipe_kernel<FAM, NLX, PLAIN, ADDV, 0, 0>; slot = 0;
  }
  if (kfn == nullptr) return cudaErrorNotSupported;
  static bool configured[6] = {false, false, false, false, false, false};
  if (!configured[slot]) {
    cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured[slot] = true;
  }
  const long ntiles = ((long)a.nfast + NLX - 1) / NLX;
  const long want = 2L * sm_count();
  PB_LAUNCH(kfn, dim3((unsigned)(ntiles < want ? ntiles : want)), dim3(kBlockThreads), smem, st, a, tin, tout, v, out, epi);
  ++g_launches;
  ++g_pipe_launches;
  return cudaGetLastError();
}

}  // namespace pb
#include "ring.cuh"
namespace pb {

// ---- launchers -----------------------------------------------------------------------------------
template <int FAM, int NL, bool PLAIN, bool ADDV>
static cudaError_t launch_yz_t(const SweepDev &a, const double *v, double *out, const double *hlo,
                               const double *hhi, double *iface, const EpiArgs &epi, cudaStream_t st) {
  const int tiles_i = (a.nfast + NL - 1) / NL;
  const dim3 grid((unsigned)(tiles_i * a.nouter)), block(NL * a.P);
  if (!a.implicit) {
    auto kfn = explicit_yz_kernel<FAM, NL, PLAIN, ADDV>;
    PB_LAUNCH(kfn, grid, block, 0, st, a, v, out, hlo, hhi, epi);
    ++g_launches;
    return cudaGetLastError();
  }
  if constexpr (NL == 8 || NL == 16 || NL == 32) {
    if (g_pipe_kernels && a.C == 32) {
      cudaError_t err;
      if (!PLAIN && epi.mode == EPI_ACC) {  // out += val: the plain-store kernel with TMA reduce-add stores
        SweepDev b = a;
        b.acc = 1;
        err = launch_yz_pipe<FAM, NL, true, ADDV>(b, v, out, hlo, hhi, iface, epi, st);
      } else if (!PLAIN && (epi.mode == EPI_RING_SET || epi.mode == EPI_RING_MAX) && epi.field == nullptr && FAM == F_R4) {
        SweepDev b = a;  // ring detector, constant length scale: |val| s^2 stored or max-reduced by the TMA unit
        b.ring = 1;
        b.ring_s2 = epi.s2;
        b.acc = epi.mode == EPI_RING_MAX ? 2 : 0;
        err = launch_yz_pipe<FAM, NL, true, ADDV>(b, v, out, hlo, hhi, iface, epi, st);
      } else {
        err = launch_yz_pipe<FAM, NL, PLAIN, ADDV>(a, v, out, hlo, hhi, iface, epi, st);
      }
      if (err != cudaErrorNotSupported) return err;
    }
  }
  if (g_reg_kernels && (a.C == 32 || (a.C == 16 && NL == 16))) {
    // register-resident kernels; the plain-store specialisation exists for the common 16-line tile only
    const size_t smem_r = 4 * (size_t)a.P * NL * sizeof(double);
    constexpr bool PL = PLAIN && NL == 16;
    if (a.C == 32) {
      auto kfn = sweep_yz_reg_kernel<FAM, 32, NL, PL, ADDV>;
      PB_LAUNCH(kfn, grid, block, smem_r, st, a, v, out, hlo, hhi, iface, epi);
    } else if constexpr (NL == 16) {
      auto kfn = sweep_yz_reg_kernel<FAM, 16, 16, false, ADDV>;
      PB_LAUNCH(kfn, grid, block, smem_r, st, a, v, out, hlo, hhi, iface, epi);
    }
    ++g_launches;
    return cudaGetLastError();
  }
  const size_t smem = ((size_t)a.m * NL + 4 * (size_t)a.P * NL + 4 * NL) * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(sweep_yz_kernel<FAM, NL, false, ADDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  auto kfn = sweep_yz_kernel<FAM, NL, false, ADDV>;
  PB_LAUNCH(kfn, grid, block, smem, st, a, v, out, hlo, hhi, iface, epi);
  ++g_launches;
  return cudaGetLastError();
}

template <int FAM, bool ADDV>
cudaError_t launch_yz_f(int lines, const SweepDev &a, const double *v, double *out, const double *hlo,
                               const double *hhi, double *iface, const EpiArgs &epi, cudaStream_t st) {
  const bool plain = epi.mode == EPI_STORE;
#define PB_YZ(NLV)                                                                                   \
  return plain ? launch_yz_t<FAM, NLV, true, ADDV>(a, v, out, hlo, hhi, iface, epi, st)              \
               : launch_yz_t<FAM, NLV, false, ADDV>(a, v, out, hlo, hhi, iface, epi, st)
  if (lines == 8) { PB_YZ(8); }
  if (lines == 32) { PB_YZ(32); }
  PB_YZ(16);
#undef PB_YZ
}

template <int FAM, int NLX, bool PLAIN, bool ADDV>
static cudaError_t launch_x_t(const SweepDev &a, const double *v, double *out, const EpiArgs &epi, cudaStream_t st) {
  const size_t tile = (((size_t)NLX * (a.m | 1) + 1) & ~(size_t)1);
  const size_t smem = (tile + 4 * (size_t)a.P * NLX + 4 * NLX) * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(sweep_x_kernel<FAM, NLX, false, ADDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  const long ntiles = ((long)a.nfast + NLX - 1) / NLX;
  int threads = NLX * a.P;
  threads = (threads + 31) / 32 * 32;
  if constexpr (NLX == 8 || NLX == 16 || NLX == 32) {
    if (g_pipe_kernels && a.C == 32) {
      cudaError_t err;
      if (!PLAIN && epi.mode == EPI_ACC) {  // out += val: the plain-store kernel with TMA reduce-add stores
        SweepDev b = a;
        b.acc = 1;
        err = launch_x_pipe<FAM, NLX, true, ADDV>(b, v, out, epi, st);
      } else if (!PLAIN && (epi.mode == EPI_RING_SET || epi.mode == EPI_RING_MAX) && epi.field == nullptr && FAM == F_R4) {
        SweepDev b = a;  // ring detector, constant length scale
        b.ring = 1;
        b.ring_s2 = epi.s2;
        b.acc = epi.mode == EPI_RING_MAX ? 2 : 0;
        err = launch_x_pipe<FAM, NLX, true, ADDV>(b, v, out, epi, st);
      } else {
        err = launch_x_pipe<FAM, NLX, PLAIN, ADDV>(a, v, out, epi, st);
      }
      if (err != cudaErrorNotSupported) return err;
    }
  }
  if (g_reg_kernels && a.implicit && (a.C == 32 || (a.C == 16 && NLX == 16))) {
    static size_t configured_r = 0;
    constexpr bool PL = PLAIN && NLX == 16;
    if (a.C == 32) {
      auto kfn = sweep_x_reg_kernel<FAM, 32, NLX, PL, ADDV>;
      if (smem > configured_r) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      PB_LAUNCH(kfn, dim3((unsigned)ntiles), dim3(threads), smem, st, a, v, out, epi);
    } else if constexpr (NLX == 16) {
      auto kfn = sweep_x_reg_kernel<FAM, 16, 16, false, ADDV>;
      if (smem > configured_r) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      PB_LAUNCH(kfn, dim3((unsigned)ntiles), dim3(threads), smem, st, a, v, out, epi);
    }
    if (smem > configured_r) configured_r = smem;
    ++g_launches;
    return cudaGetLastError();
  }
  auto kfn = sweep_x_kernel<FAM, NLX, false, ADDV>;
  PB_LAUNCH(kfn, dim3((unsigned)ntiles), dim3(threads), smem, st, a, v, out, epi);
  ++g_launches;
  return cudaGetLastError();
}

template <int FAM, bool ADDV>
cudaError_t launch_x_f(int lines, const SweepDev &a, const double *v, double *out, const EpiArgs &epi, cudaStream_t st) {
  const bool plain = epi.mode == EPI_STORE;
  if (!a.implicit && g_pipe_kernels && a.C == 32 && (lines == 8 || lines == 16 || lines == 32)) {  // the TMA-pipelined kernel without its solve phases
    cudaError_t err;
    if (lines == 8) err = plain ? launch_x_pipe<FAM, 8, true, ADDV>(a, v, out, epi, st) : launch_x_pipe<FAM, 8, false, ADDV>(a, v, out, epi, st);
    else if (lines == 16) err = plain ? launch_x_pipe<FAM, 16, true, ADDV>(a, v, out, epi, st) : launch_x_pipe<FAM, 16, false, ADDV>(a, v, out, epi, st);
    else err = plain ? launch_x_pipe<FAM, 32, true, ADDV>(a, v, out, epi, st) : launch_x_pipe<FAM, 32, false, ADDV>(a, v, out, epi, st);
    if (err != cudaErrorNotSupported) return err;
  }
  if (!a.implicit && a.m % 2 == 0 && a.m >= 16 && a.m <= 1024 && !((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    const long nblk = (a.nfast + kExplicitXLines - 1) / kExplicitXLines, cap = 3L * sm_count();
    const dim3 grid((unsigned)(nblk < cap ? nblk : cap));
    const size_t smem = 2 * kExplicitXLines * (size_t)(a.m + 8) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
      cudaError_t err = cudaFuncSetAttribute(explicit_x_kernel<FAM, true, ADDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err == cudaSuccess) err = cudaFuncSetAttribute(explicit_x_kernel<FAM, false, ADDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
      configured = smem;
    }
    if (plain) { auto kfn = explicit_x_kernel<FAM, true, ADDV>; PB_LAUNCH(kfn, grid, dim3(256), smem, st, a, v, out, epi); }
    else { auto kfn = explicit_x_kernel<FAM, false, ADDV>; PB_LAUNCH(kfn, grid, dim3(256), smem, st, a, v, out, epi); }
    ++g_launches;
    return cudaGetLastError();
  }
#define PB_X(NLV) \
  return plain ? launch_x_t<FAM, NLV, true, ADDV>(a, v, out, epi, st) : launch_x_t<FAM, NLV, false, ADDV>(a, v, out, epi, st)
  if (lines == 8) { PB_X(8); }
  if (lines == 32) { PB_X(32); }
  PB_X(16);
#undef PB_X
}


}  // namespace pb

// C ABI of libparcop_b200 (include/parcop_b200.h): plan construction and operator dispatch.
//
// Replaces, for the hot path only: parcop.f90 (the f2py surface), objects.f90 (global object
// tables -> one opaque plan), compact.f90:55-319 (operator suite setup), compact_operators.f90
// (null-op rule, metric scale) and the Cartesian / curvilinear branches of operators.f90.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/parcop_b200.h"
#include "kernels.cuh"
#include "tables.hpp"

using namespace pb;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

int g_chunk_len = 32;

#define PB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return fail(PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));     \
  } while (0)

struct SweepPlan {
  bool built = false, null_op = false, split = false;
  int kind = 0, dir = 0, np = 1, rank = 0, n = 0;
  Stencil st;
  SweepDev dev;
  double4 *RC = nullptr;  // rank-level spike columns [m]
  double *GR = nullptr;   // rank-level reduced-system rows [4][4np]
  int zone_lo = 0, zone_hi = 0;     // rows from either end of the slab the rank-level correction reaches
  unsigned long long rank_mask = 0; // ranks whose interface values enter this rank's correction
  // the same split line as consecutive chunks of the GLOBAL line (ring kernel): tables of this rank's
  // chunks from one factorisation of the whole line, and who exchanges which chunk states with whom
  bool xr_ok = false;
  SweepDev devx;
  XRing xr;
  std::vector<void *> owned;
};

}  // namespace

struct pb_plan {
  int n[3], p[3], c[3], a[3];
  int coordsys = 0, device = 0;
  double d[3], x1f[3], xnf[3];
  bool periodic[3], null_dir[3];
  size_t npts = 0;
  int bcode[3][2];                 // 0 NONE, 1 PERI, 2 SYMM
  int isym[3];                     // patch.f90:86-91: -1 when either end of the axis is a symmetry plane
  SweepPlan sw[K_COUNT][3];        // iop 1: even fields at symmetry planes (compact.f90:77-91)
  SweepPlan d1_odd[3];             // iop 2 of the first derivative: odd fields (the normal flux in divV)
  SweepPlan d8_odd[3];             // iop 2 of the 8th derivative (the normal component in ringV, operators.f90:661-671)
  SweepPlan custom_d1[3];
  std::vector<double *> scratch;   // device work fields, npts each
  double *red_partial = nullptr, *red_result = nullptr, *red_host = nullptr;
  std::map<std::string, double *> mesh;  // device mesh arrays
  bool mesh_set = false;
  // host-array entry points: copy / compute / copy-back pipeline over slabs
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};  // H2D, compute, D2H
  std::vector<cudaEvent_t> hev;
};

namespace {

template <typename T>
int upload(SweepPlan &sp, const std::vector<T> &h, const T **dptr) {
  T *d = nullptr;
  PB_CUDA(cudaMalloc(&d, sizeof(T) * (h.empty() ? 1 : h.size())));
  if (!h.empty()) PB_CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  sp.owned.push_back(d);
  *dptr = d;
  return PB_OK;
}

void free_sweep(SweepPlan &sp) {
  for (void *q : sp.owned) cudaFree(q);
  sp.owned.clear();
  sp.built = false;
}

// A split line as the global line's chunks (ring kernel).  One LU factorisation of the whole line of
// n rows (circulant limit when periodic): rank r owns chunks r P .. (r + 1) P - 1 of n / 32, every
// interior rank has constant coefficients from its first row, and the carried-state sums simply run
// across the slab faces -- the states they need from other ranks are what the kernels exchange.
int build_ring_tables(SweepPlan &sp, const std::vector<double> &bands, bool periodic) {
  sp.xr_ok = false;
  const int n = sp.n, np = sp.np, m = n / np, r = sp.rank;
  if (m % 32 != 0) return PB_OK;
  const int P = m / 32, Pg = n / 32;
  if (P != 4 && P != 8 && P != 16) return PB_OK;  // slabs one CTA's tile covers: 128, 256, 512 planes
  LineTables gt;
  try { gt = build_line_tables(n, bands, periodic, Pg); }
  catch (const std::exception &) { return PB_OK; }
  auto nF = [&](int rank, int q) { return gt.nF[rank * P + q]; };
  auto nB = [&](int rank, int q) { return gt.nB[rank * P + q]; };
  auto need_f = [&](int rank) { int v = 0; for (int q = 0; q < P; ++q) v = std::max(v, nF(rank, q) - q); return v; };
  auto need_b = [&](int rank) { int v = 0; for (int q = 0; q < P; ++q) v = std::max(v, nB(rank, q) - (P - 1 - q)); return v; };
  XRing xr;
  memset(&xr, 0, sizeof(xr));
  xr.on = 1;
  xr.need_f = need_f(r);
  xr.need_b = need_b(r);
  for (int k = 0; k < np - 1; ++k) {  // my top chunks as slots k P .. of rank r + 1 + k, my bottom chunks of rank r - 1 - k
    int up = r + 1 + k, dn = r - 1 - k;
    if (periodic) { up %= np; dn = (dn % np + np) % np; }
    const int cu = up < np ? std::min(P, std::max(0, need_f(up) - k * P)) : 0;
    const int cd = dn >= 0 ? std::min(P, std::max(0, need_b(dn) - k * P)) : 0;
    if (cu > 0) { if (k >= kXHops || xr.nup != k) return PB_OK; xr.cnt_up[k] = cu; xr.nup = k + 1; }
    if (cd > 0) { if (k >= kXHops || xr.ndn != k) return PB_OK; xr.cnt_dn[k] = cd; xr.ndn = k + 1; }
  }
  // every state must come from another rank, within the slots a CTA keeps, and a rank that sends to
  // a rank also hears from it (that is what orders consecutive sweeps on the record buffers)
  if (xr.need_f > kXExt || xr.need_b > kXExt) return PB_OK;
  if (xr.need_f > (np - 1) * P || xr.need_b > (np - 1) * P) return PB_OK;
  for (int rank = 0; rank < np; ++rank)
    if ((need_f(rank) + P - 1) / P != (need_b(rank) + P - 1) / P && periodic) return PB_OK;

  SweepDev &dv = sp.devx;
  dv = sp.dev;
  dv.P = P; dv.C = 32;
  dv.wrap = 0;
  dv.mconst = 0;  // the products of a slab's chunks come from the tables of the global line
  for (int q = 0; q < kMaxChunks; ++q) { dv.perm[q] = (unsigned char)q; dv.ctype[q] = 0; dv.nf[q] = dv.nb[q] = 0; }
  int jmax = 1;
  for (int q = 0; q < P; ++q) {
    dv.ctype[q] = gt.ctype[r * P + q];
    dv.nf[q] = (unsigned char)nF(r, q);
    dv.nb[q] = (unsigned char)nB(r, q);
    jmax = std::max(jmax, std::max(nF(r, q), nB(r, q)));
  }
  dv.has_const = gt.has_const ? 1 : 0;
  for (int q = 0; q < 5; ++q) dv.cst[q] = gt.cst[q];
  dv.cparam = gt.has_const ? 1 : 0;
  if (dv.cparam)
    for (int q = 0; q < 32; ++q) {
      dv.phi0[q] = make_double2(gt.phi[q * 2], gt.phi[q * 2 + 1]);
      dv.psi0[q] = make_double2(gt.psi[q * 2], gt.psi[q * 2 + 1]);
    }
  const size_t nt = (size_t)gt.ntypes * 32;
  std::vector<double2> luf(nt), phi(nt), psi(nt), chi(nt);
  std::vector<double4> lub(nt);
  for (size_t t = 0; t < nt; ++t) {
    luf[t] = make_double2(gt.luf[t * 2], gt.luf[t * 2 + 1]);
    phi[t] = make_double2(gt.phi[t * 2], gt.phi[t * 2 + 1]);
    psi[t] = make_double2(gt.psi[t * 2], gt.psi[t * 2 + 1]);
    chi[t] = make_double2(gt.chi[t * 2], gt.chi[t * 2 + 1]);
    lub[t] = make_double4(gt.lub[t * 4], gt.lub[t * 4] * gt.lub[t * 4 + 1], gt.lub[t * 4] * gt.lub[t * 4 + 2], 0.0);
  }
  if (dv.cparam)
    for (int q = 0; q < 32; ++q) dv.chi0[q] = chi[q];
  // Early form (nobody waits before solving): possible when every exchanged state comes from the
  // adjacent rank.  The backward states arrive as the rank above computed them from its own chunks;
  // what this rank's forward states add to them is  sum_j Mb[lp][j] X_e sum_j' Mf[e][j'] EN[c]  with
  // X_e the first two rows of chi of chunk e above: folded into one 2x2 block per (top chunk lp, top chunk c).
  static const bool want_early = getenv("PB_XR_EARLY") ? atoi(getenv("PB_XR_EARLY")) != 0 : true;
  // ... and worthwhile when few states cross a face: measured on 2 GPUs at 512^3 per GPU, ddz 0.52 -> 0.47 ms (two
  // states per face, the waits vanish) but the compact filter 0.80 -> 0.89 ms (six states: the correction and
  // closure sums cost more than the waits they remove) -- profiles/r2_xr_variants_2gpu.log
  static const int early_max = getenv("PB_XR_EARLY_MAX") ? atoi(getenv("PB_XR_EARLY_MAX")) : 3;
  // ... and, of those, for the first derivative only: the second / eighth derivative kernels (larger right-hand
  // sides, 56-64-byte spill frames) measured slower in the early form, dd8z 0.71 vs 0.62 ms and d2z 0.63 vs 0.55 ms
  // against ddz 0.48 vs 0.53 ms (2 GPUs, 512^3 per GPU, gpurun r2af in profiles/r2_xr_variants_2gpu.log)
  static const bool early_all = getenv("PB_XR_EARLY_ALL") ? atoi(getenv("PB_XR_EARLY_ALL")) != 0 : false;
  bool early = want_early && (early_all || sp.st.fam == F_D1);
  for (int rank = 0; rank < np; ++rank)
    if (need_f(rank) > std::min(P, early_max) || need_b(rank) > std::min(P, early_max)) early = false;
  std::vector<double4> Bc;
  int bc_n = 0;
  if (early && xr.need_b > 0) {
    const int up = periodic ? (r + 1) % np : r + 1;
    bc_n = std::max(1, need_f(up));
    Bc.assign((size_t)xr.need_b * bc_n, make_double4(0, 0, 0, 0));
    auto at4 = [](const std::vector<double> &v, size_t i) { return make_double4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]); };
    auto mul = [](const double4 &A, const double4 &B) {  // [[x, y], [z, w]]
      return make_double4(A.x * B.x + A.y * B.z, A.x * B.y + A.y * B.w, A.z * B.x + A.w * B.z, A.z * B.y + A.w * B.w);
    };
    for (int lp = P - xr.need_b; lp < P; ++lp)
      for (int j = P - lp; j <= nB(r, lp); ++j) {
        const int e = lp + j - P, g2 = up * P + e, t2 = gt.ctype[g2];
        const double4 X = make_double4(gt.chi[((size_t)t2 * 32 + 0) * 2], gt.chi[((size_t)t2 * 32 + 0) * 2 + 1],
                                       gt.chi[((size_t)t2 * 32 + 1) * 2], gt.chi[((size_t)t2 * 32 + 1) * 2 + 1]);
        const double4 M1 = at4(gt.Mb, (size_t)(r * P + lp) * (Pg + 1) + j);
        for (int j2 = e + 1; j2 <= nF(up, e); ++j2) {
          const int c = P + e - j2;  // this rank's chunk whose forward end state enters chunk e above
          if (c < P - bc_n || c < 0) { early = false; continue; }
          const double4 M2 = at4(gt.Mf, (size_t)g2 * (Pg + 1) + j2);
          const double4 T = mul(M1, mul(X, M2));
          double4 &dst = Bc[(size_t)(lp - (P - xr.need_b)) * bc_n + (c - (P - bc_n))];
          dst.x += T.x; dst.y += T.y; dst.z += T.z; dst.w += T.w;
        }
      }
  }
  dv.mstride = jmax + 1;
  std::vector<double4> Mf((size_t)P * dv.mstride), Mb((size_t)P * dv.mstride);
  for (int q = 0; q < P; ++q)
    for (int j = 0; j <= jmax; ++j) {
      const size_t src = ((size_t)(r * P + q) * (Pg + 1) + j) * 4, dst = (size_t)q * dv.mstride + j;
      const bool in = j <= Pg;
      Mf[dst] = in ? make_double4(gt.Mf[src], gt.Mf[src + 1], gt.Mf[src + 2], gt.Mf[src + 3]) : make_double4(0, 0, 0, 0);
      Mb[dst] = in ? make_double4(gt.Mb[src], gt.Mb[src + 1], gt.Mb[src + 2], gt.Mb[src + 3]) : make_double4(0, 0, 0, 0);
    }
  {  // chunks of a periodic global line share one row of products: a copy in the kernel parameters (SweepDev::Mf0 / Mb0)
    static const bool off = getenv("PB_NO_MCONST") != nullptr;
    int qf = 0, qb = 0;
    for (int q = 1; q < P; ++q) {
      if (dv.nf[q] > dv.nf[qf]) qf = q;
      if (dv.nb[q] > dv.nb[qb]) qb = q;
    }
    bool same = !off && dv.nf[qf] <= kMTerms && dv.nb[qb] <= kMTerms;
    for (int q = 0; q < P && same; ++q) {
      for (int j = 1; j <= dv.nf[q] && same; ++j)
        same = memcmp(&Mf[(size_t)q * dv.mstride + j], &Mf[(size_t)qf * dv.mstride + j], sizeof(double4)) == 0;
      for (int j = 1; j <= dv.nb[q] && same; ++j)
        same = memcmp(&Mb[(size_t)q * dv.mstride + j], &Mb[(size_t)qb * dv.mstride + j], sizeof(double4)) == 0;
    }
    dv.mconst = same ? 1 : 0;
    dv.nf0 = same ? dv.nf[qf] : 0;
    dv.nb0 = same ? dv.nb[qb] : 0;
    for (int j = 0; j < kMTerms; ++j) {
      dv.Mf0[j] = (same && j < dv.nf[qf]) ? Mf[(size_t)qf * dv.mstride + j + 1] : make_double4(0, 0, 0, 0);
      dv.Mb0[j] = (same && j < dv.nb[qb]) ? Mb[(size_t)qb * dv.mstride + j + 1] : make_double4(0, 0, 0, 0);
    }
  }
  int rcv;
  if ((rcv = upload(sp, luf, &dv.luf)) != PB_OK) return rcv;
  if ((rcv = upload(sp, lub, &dv.lub)) != PB_OK) return rcv;
  if ((rcv = upload(sp, phi, &dv.phi)) != PB_OK) return rcv;
  if ((rcv = upload(sp, psi, &dv.psi)) != PB_OK) return rcv;
  if ((rcv = upload(sp, Mf, &dv.Mf)) != PB_OK) return rcv;
  if ((rcv = upload(sp, Mb, &dv.Mb)) != PB_OK) return rcv;
  if ((rcv = upload(sp, chi, &dv.chi)) != PB_OK) return rcv;
  xr.early = early ? 1 : 0;
  xr.bc_n = bc_n;
  if (early && (rcv = upload(sp, Bc, &xr.Bc)) != PB_OK) return rcv;
  sp.xr = xr;
  sp.xr_ok = true;
  return PB_OK;
}

// One operator along one axis: compact_basetype.f90:65-209 re-expressed for the chunked kernels.
// bc_lo / bc_hi: 0 one-sided closure, +1 / -1 symmetry plane below an even / odd field
int build_sweep(pb_plan *pl, SweepPlan &sp, int kind, int dir, bool periodic, int bc_lo = 0, int bc_hi = 0) {
  sp.kind = kind; sp.dir = dir; sp.n = pl->n[dir]; sp.np = pl->p[dir]; sp.rank = pl->c[dir];
  try { sp.st = make_stencil((Kind)kind, periodic ? 0 : bc_lo, periodic ? 0 : bc_hi); }
  catch (const std::exception &ex) { return fail(PB_ERR_ARG, ex.what()); }
  sp.null_op = pl->n[dir] < 4;  // compact.f90:95-97, compact_basetype.f90:101-103
  sp.split = sp.np > 1;
  sp.built = true;
  if (sp.null_op) return PB_OK;
  const int n = sp.n, np = sp.np, m = n / np, r0 = sp.rank * m;
  if (m < 16)
    return fail(PB_ERR_UNSUPPORTED, "local extent " + std::to_string(m) + " along axis " + std::to_string(dir) +
                                        " is below 16 (the reference's own decomposition bound, pyrandaMPI.py:785-792)");
  const Stencil &st = sp.st;
  std::vector<double> bands = assemble_bands(st, n, periodic);
  std::vector<double> local(bands.begin() + (size_t)r0 * 5, bands.begin() + (size_t)(r0 + m) * 5);
  const bool cyclic_local = periodic && np == 1;
  const int P = choose_chunks(m, g_chunk_len);
  if (P > kMaxChunks) return fail(PB_ERR_UNSUPPORTED, "too many chunks");

  SweepDev &dv = sp.dev;
  memset(&dv, 0, sizeof(dv));
  dv.m = m; dv.P = P; dv.C = m / P;
  for (int q = 0; q < kMaxChunks; ++q) dv.perm[q] = (unsigned char)q;
  {
    static const int wstore = getenv("PB_WARP_STORE") ? atoi(getenv("PB_WARP_STORE")) : 1;
    dv.wstore = wstore;
    static const int skew = getenv("PB_SKEW_NS") ? atoi(getenv("PB_SKEW_NS")) : 0;
    dv.skew_ns = skew;
  }
  const int ax = pl->a[0], ay = pl->a[1], az = pl->a[2];
  if (dir == 0) { dv.nfast = ay * az; dv.nouter = 1; dv.rstride = 1; dv.ostride = 0; }
  else if (dir == 1) { dv.nfast = ax; dv.nouter = az; dv.rstride = ax; dv.ostride = (long)ax * ay; }
  else { dv.nfast = ax; dv.nouter = ay; dv.rstride = (long)ax * ay; dv.ostride = ax; }
  for (int l = 0; l < 9; ++l) dv.ari[l] = st.ari[l];
  // closure rows belong to the ranks that own the physical boundary
  for (int r = 0; r < 4; ++r)
    for (int l = 0; l < 9; ++l) { dv.arb_lo[r][l] = st.arb_lo[r][l]; dv.arb_hi[r][l] = st.arb_hi[r][l]; }
  dv.phys_lo = (!periodic && sp.rank == 0) ? 1 : 0;
  dv.phys_hi = (!periodic && sp.rank == np - 1) ? 1 : 0;
  dv.wrap = cyclic_local ? 1 : 0;
  dv.implicit = st.implicit ? 1 : 0;
  dv.add_v = st.add_back ? 1 : 0;
  const double dd = pl->d[dir];
  dv.scale = st.post == 1 ? 1.0 / dd : (st.post == 2 ? 1.0 / (dd * dd) : 1.0);  // compact_operators.f90:43,152

  if (st.fam == F_R3 && (dv.phys_lo || dv.phys_hi)) {
    // compact_r3.f90:89-104 multiplies zero ghost values by these weights; they must vanish
    for (int r = 0; r < 3; ++r)
      for (int l = 0; l < 3 - r; ++l)
        if (st.arb_lo[r][l] != 0.0 || st.arb_hi[3 - r][6 - l] != 0.0)
          return fail(PB_ERR_UNSUPPORTED, "r3 closure rows reach outside the domain");
  }
  if (st.implicit) {
    LineTables lt;
    // one-rank lines: the first / second / eighth derivative kernels measured faster with the third (1e-17)
    // term of their state sums kept, the compact filter with its sum cut at the rounding unit (six terms, not eight)
    try { lt = build_line_tables(m, local, cyclic_local, P, st.fam == F_R4 && st.add_back ? 0.0 : 1e-22); }
    catch (const std::exception &ex) { return fail(PB_ERR_ARG, ex.what()); }
    for (int q = 0; q < P; ++q) dv.ctype[q] = lt.ctype[q];
    dv.has_const = lt.has_const ? 1 : 0;
    for (int q = 0; q < 5; ++q) dv.cst[q] = lt.cst[q];
    dv.cparam = (lt.has_const && dv.C <= kParamRows) ? 1 : 0;
    if (dv.cparam)
      for (int q = 0; q < dv.C; ++q) {
        dv.phi0[q] = make_double2(lt.phi[q * 2], lt.phi[q * 2 + 1]);  // type 0 comes first in the tables
        dv.psi0[q] = make_double2(lt.psi[q * 2], lt.psi[q * 2 + 1]);
      }
    const size_t nt = (size_t)lt.ntypes * dv.C;
    std::vector<double2> luf(nt), phi(nt), psi(nt);
    std::vector<double4> lub(nt);
    for (size_t t = 0; t < nt; ++t) {
      luf[t] = make_double2(lt.luf[t * 2], lt.luf[t * 2 + 1]);
      phi[t] = make_double2(lt.phi[t * 2], lt.phi[t * 2 + 1]);
      psi[t] = make_double2(lt.psi[t * 2], lt.psi[t * 2 + 1]);
      // y / z kernels: {1/pivot, u1/pivot, u2/pivot} (one operation on the chain through x1); the x kernels sit at
      // the 128-register limit with the form (t - u1 x1 - u2 x2) / pivot and keep {1/pivot, u1, u2}
      const double sc = dir == 0 ? 1.0 : lt.lub[t * 4];
      lub[t] = make_double4(lt.lub[t * 4], sc * lt.lub[t * 4 + 1], sc * lt.lub[t * 4 + 2], 0.0);
    }
    std::vector<double4> Mf((size_t)P * (P + 1)), Mb((size_t)P * (P + 1));
    for (size_t t = 0; t < Mf.size(); ++t) {
      Mf[t] = make_double4(lt.Mf[t * 4], lt.Mf[t * 4 + 1], lt.Mf[t * 4 + 2], lt.Mf[t * 4 + 3]);
      Mb[t] = make_double4(lt.Mb[t * 4], lt.Mb[t * 4 + 1], lt.Mb[t * 4 + 2], lt.Mb[t * 4 + 3]);
    }
    for (int q = 0; q < P; ++q) { dv.nf[q] = (unsigned char)lt.nF[q]; dv.nb[q] = (unsigned char)lt.nB[q]; }
    dv.mstride = P + 1;
    {  // one row of products for every chunk (circulant lines): a copy in the kernel parameters
      static const bool off = getenv("PB_NO_MCONST") != nullptr;
      bool same = !off && lt.nF[0] >= 1 && lt.nB[0] >= 1 && lt.nF[0] <= kMTerms && lt.nB[0] <= kMTerms;
      for (int q = 1; q < P && same; ++q) {
        same = lt.nF[q] == lt.nF[0] && lt.nB[q] == lt.nB[0];
        for (int j = 1; j <= lt.nF[0] && same; ++j)
          same = memcmp(&Mf[(size_t)q * (P + 1) + j], &Mf[j], sizeof(double4)) == 0;
        for (int j = 1; j <= lt.nB[0] && same; ++j)
          same = memcmp(&Mb[(size_t)q * (P + 1) + j], &Mb[j], sizeof(double4)) == 0;
      }
      dv.mconst = same ? 1 : 0;
      dv.nf0 = same ? lt.nF[0] : 0;
      dv.nb0 = same ? lt.nB[0] : 0;
      for (int j = 0; j < kMTerms; ++j) {
        dv.Mf0[j] = (same && j < lt.nF[0]) ? Mf[j + 1] : make_double4(0, 0, 0, 0);
        dv.Mb0[j] = (same && j < lt.nB[0]) ? Mb[j + 1] : make_double4(0, 0, 0, 0);
      }
    }
    {
      int k = 0;
      for (int q = 0; q < P; ++q)
        if (!(dv.has_const && dv.ctype[q] == 0)) dv.perm[k++] = (unsigned char)q;
      for (int q = 0; q < P; ++q)
        if (dv.has_const && dv.ctype[q] == 0) dv.perm[k++] = (unsigned char)q;
    }
    int rcv;
    if ((rcv = upload(sp, luf, &dv.luf)) != PB_OK) return rcv;
    if ((rcv = upload(sp, lub, &dv.lub)) != PB_OK) return rcv;
    if ((rcv = upload(sp, phi, &dv.phi)) != PB_OK) return rcv;
    if ((rcv = upload(sp, psi, &dv.psi)) != PB_OK) return rcv;
    if ((rcv = upload(sp, Mf, &dv.Mf)) != PB_OK) return rcv;
    if ((rcv = upload(sp, Mb, &dv.Mb)) != PB_OK) return rcv;
    if (sp.split) {
      // rank level: the same partition algebra with ranks as chunks (compact_basetype.f90:150-198)
      Partition rp;
      try { rp = build_partition(n, bands, periodic, np); }
      catch (const std::exception &ex) { return fail(PB_ERR_ARG, ex.what()); }
      const int t = rp.ctype[sp.rank];
      std::vector<double4> RC(m);
      for (int i = 0; i < m; ++i) {
        const double *q = &rp.rc[((size_t)t * m + i) * 4];
        RC[i] = make_double4(q[0], q[1], q[2], q[3]);
      }
      std::vector<double> GR(rp.G.begin() + (size_t)sp.rank * 16 * np, rp.G.begin() + (size_t)(sp.rank + 1) * 16 * np);
      // The spike columns decay like rho^row away from the slab faces and the reduced-system rows
      // like rho^(az * distance in ranks): keep only what can change a double (1e-16 / 1e-19 relative).
      {
        double cmax = 0.0, gmax = 0.0;
        for (int i = 0; i < m; ++i)
          cmax = std::max({cmax, std::fabs(RC[i].x), std::fabs(RC[i].y), std::fabs(RC[i].z), std::fabs(RC[i].w)});
        sp.zone_lo = sp.zone_hi = 0;
        for (int i = 0; i < m; ++i) {
          if (std::max(std::fabs(RC[i].x), std::fabs(RC[i].y)) > 1e-16 * cmax) sp.zone_lo = i + 1;
          if (std::max(std::fabs(RC[m - 1 - i].z), std::fabs(RC[m - 1 - i].w)) > 1e-16 * cmax) sp.zone_hi = i + 1;
        }
        for (double gv : GR) gmax = std::max(gmax, std::fabs(gv));
        sp.rank_mask = 0;
        for (int r = 0; r < np; ++r)
          for (int c = 0; c < 4; ++c)
            for (int j = 0; j < 4; ++j)
              if (std::fabs(GR[(size_t)c * 4 * np + 4 * r + j]) > 1e-19 * gmax) sp.rank_mask |= 1ull << r;
      }
      const double4 *dRC; const double *dGR;
      if ((rcv = upload(sp, RC, &dRC)) != PB_OK) return rcv;
      if ((rcv = upload(sp, GR, &dGR)) != PB_OK) return rcv;
      sp.RC = const_cast<double4 *>(dRC);
      sp.GR = const_cast<double *>(dGR);
      if ((rcv = build_ring_tables(sp, bands, periodic)) != PB_OK) return rcv;
    }
  }
  return PB_OK;
}

int bc_code(const char *s, int *code) {
  if (!s) return PB_ERR_ARG;
  if (!strncmp(s, "NONE", 4)) { *code = 0; return PB_OK; }
  if (!strncmp(s, "PERI", 4)) { *code = 1; return PB_OK; }
  if (!strncmp(s, "SYMM", 4)) { *code = 2; return PB_OK; }
  return PB_ERR_ARG;
}

int get_scratch(pb_plan *pl, size_t k, double **out) {
  while (pl->scratch.size() <= k) {
    double *d = nullptr;
    PB_CUDA(cudaMalloc(&d, sizeof(double) * pl->npts));
    pl->scratch.push_back(d);
  }
  *out = pl->scratch[k];
  return PB_OK;
}

const EpiArgs kStore = {EPI_STORE, 0.0, nullptr};

// d1x..filterz of compact_operators.f90 for a non-split direction
int apply_dir(pb_plan *pl, SweepPlan &sp, const double *in, double *out, const EpiArgs &epi, cudaStream_t st) {
  const long N = (long)pl->npts;
  if (!sp.built) return fail(PB_ERR_STATE, "operator not built");
  if (in == out) return fail(PB_ERR_ARG, "operator output must not alias its input");
  if (sp.null_op) {
    // compact_operators.f90:24-29 (zero) / :397-402 (copy)
    if (sp.st.null_option == 1) {
      if (epi.mode != EPI_STORE) return fail(PB_ERR_ARG, "null filter with a composite epilogue");
      PB_CUDA(launch_copy(N, in, out, st));
    } else if (epi.mode == EPI_STORE || epi.mode == EPI_RING_SET) {
      PB_CUDA(launch_fill(N, 0.0, out, st));
    }
    return PB_OK;
  }
  if (sp.split)
    return fail(PB_ERR_STATE, "this axis is split across ranks: use pb_z_pack_halo / pb_z_local / pb_z_finish");
  if (sp.dir == 0) PB_CUDA(launch_sweep_x(sp.st.fam, 0, sp.dev, in, out, epi, st));
  else PB_CUDA(launch_sweep_yz(sp.st.fam, 0, sp.dev, in, out, nullptr, nullptr, nullptr, epi, st));
  return PB_OK;
}

// the same sweep restricted to `cnt` outer slices starting at `first`: z-planes for x / y sweeps,
// y-rows for z sweeps.  Used by the host-array pipeline, which works slab by slab.
int apply_dir_slab(pb_plan *pl, SweepPlan &sp, const double *in, double *out, int first, int cnt, cudaStream_t st) {
  if (!sp.built || sp.null_op || sp.split) return fail(PB_ERR_STATE, "slab sweeps need a local, non-null direction");
  SweepDev dv = sp.dev;
  long off;
  if (sp.dir == 0) { off = (long)first * pl->a[0] * pl->a[1]; dv.nfast = cnt * pl->a[1]; }
  else if (sp.dir == 1) { off = (long)first * pl->a[0] * pl->a[1]; dv.nouter = cnt; }
  else { off = (long)first * pl->a[0]; dv.nouter = cnt; }
  if (sp.dir == 0) PB_CUDA(launch_sweep_x(sp.st.fam, 0, dv, in + off, out + off, kStore, st));
  else PB_CUDA(launch_sweep_yz(sp.st.fam, 0, dv, in + off, out + off, nullptr, nullptr, nullptr, kStore, st));
  return PB_OK;
}

// An explicit z sweep (the Gaussian filter) restricted to planes [first, first + cnt): the slab is a
// short line whose neighbours' planes act as halo planes (compact_r4.f90:640-656 does the same with the
// planes received from the neighbouring ranks).  The first / last slab of a periodic axis wraps around
// the field; those of a bounded axis keep the one-sided closure rows.  Host-array pipeline only.
int apply_z_explicit_slab(pb_plan *pl, SweepPlan &sp, const double *in, double *out, int first, int cnt, cudaStream_t st) {
  if (!sp.built || sp.null_op || sp.split || sp.st.implicit || sp.dir != 2 || cnt % sp.dev.C != 0)
    return fail(PB_ERR_STATE, "slab z sweeps need a local explicit operator and whole chunks");
  const int az = pl->a[2], H = 4;
  const long plane = (long)pl->a[0] * pl->a[1];
  const bool lo_end = first == 0, hi_end = first + cnt == az;
  SweepDev dv = sp.dev;
  dv.m = cnt;
  dv.P = cnt / dv.C;
  dv.wrap = 0;
  dv.phys_lo = lo_end ? sp.dev.phys_lo : 0;
  dv.phys_hi = hi_end ? sp.dev.phys_hi : 0;
  const double *hlo = nullptr, *hhi = nullptr;
  if (!lo_end) hlo = in + (long)(first - H) * plane;
  else if (sp.dev.wrap) hlo = in + (long)(az - H) * plane;
  if (!hi_end) hhi = in + (long)(first + cnt) * plane;
  else if (sp.dev.wrap) hhi = in;
  if ((lo_end && !sp.dev.wrap && !sp.dev.phys_lo) || (hi_end && !sp.dev.wrap && !sp.dev.phys_hi))
    return fail(PB_ERR_STATE, "slab z sweeps need a periodic axis or closure rows at its ends");
  const long off = (long)first * plane;
  PB_CUDA(launch_sweep_yz(sp.st.fam, 0, dv, in + off, out + off, hlo, hhi, nullptr, kStore, st));
  return PB_OK;
}

double *mesh_arr(const pb_plan *pl, const char *name) {
  auto it = pl->mesh.find(name);
  return it == pl->mesh.end() ? nullptr : it->second;
}

int new_mesh_arr(pb_plan *pl, const char *name, double **out) {
  double *d = mesh_arr(pl, name);
  if (!d) {
    PB_CUDA(cudaMalloc(&d, sizeof(double) * pl->npts));
    pl->mesh[name] = d;
  }
  *out = d;
  return PB_OK;
}

int filter3(pb_plan *pl, int kind, const double *in, double *out, cudaStream_t st) {
  // operators.f90:849-851 / :873-875: x -> y -> z through one work array
  double *tmp;
  int rc;
  if ((rc = get_scratch(pl, 0, &tmp)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, pl->sw[kind][0], in, out, kStore, st)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, pl->sw[kind][1], out, tmp, kStore, st)) != PB_OK) return rc;
  return apply_dir(pl, pl->sw[kind][2], tmp, out, kStore, st);
}

// d1x(v, dv, bc) of compact_operators.f90:13-50: bc = -1 selects the operator built for odd fields
// when the axis has a symmetry plane, and the ordinary one otherwise (:35-38)
SweepPlan &d1_plan(pb_plan *pl, int dir, int bc) {
  return (bc == -1 && pl->d1_odd[dir].built) ? pl->d1_odd[dir] : pl->sw[K_D1][dir];
}

// the Cartesian divergence with a symmetry selector per direction (operators.f90:48-52, :106-120)
int div_cart(pb_plan *pl, const double *fx, const double *fy, const double *fz, double *out, int bx, int by, int bz,
             cudaStream_t st) {
  const EpiArgs acc = {EPI_ACC, 0.0, nullptr};
  int rc;
  if ((rc = apply_dir(pl, d1_plan(pl, 0, bx), fx, out, kStore, st)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, d1_plan(pl, 1, by), fy, out, acc, st)) != PB_OK) return rc;
  return apply_dir(pl, d1_plan(pl, 2, bz), fz, out, acc, st);
}

}  // namespace

extern "C" {

const char *pb_last_error(void) { return g_err.c_str(); }
const char *pb_version(void) { return "parcop_b200 0.1 (sm_100a)"; }
long pb_launch_count(void) { return launch_count(); }
long pb_pipe_launch_count(void) { return pipe_launch_count(); }
long pb_ring_launch_count(void) { return ring_launch_count(); }
int pb_set_ring(int mode, int lines) { set_ring_kernels(mode, lines); return PB_OK; }

int pb_set_tuning(int lines_yz, int lines_x, int chunk_len) {
  if (lines_yz > 0) set_yz_lines(lines_yz);
  if (lines_x > 0) set_x_lines(lines_x);
  if (chunk_len >= 16) g_chunk_len = chunk_len % 1000;
  if (chunk_len >= 1000) set_reg_kernels(chunk_len / 1000 - 1);  // 1xxx: shared-memory kernels, 2xxx: register kernels, 3xxx: + pipelined
  return PB_OK;
}

int pb_plan_create(pb_plan **plan, int nx, int ny, int nz, int px, int py, int pz, int cx, int cy, int cz,
                   int coordsys, double x1, double xn, double y1, double yn, double z1, double zn,
                   const char *bx1, const char *bxn, const char *by1, const char *byn, const char *bz1,
                   const char *bzn, int device) {
  if (!plan) return fail(PB_ERR_ARG, "plan is NULL");
  *plan = nullptr;
  if (nx < 1 || ny < 1 || nz < 1 || px < 1 || py < 1 || pz < 1) return fail(PB_ERR_ARG, "sizes must be positive");
  if (nx % px || ny % py || nz % pz) return fail(PB_ERR_ARG, "grid not divisible by the processor grid (comm.f90:189-206)");
  if (px != 1 || py != 1) return fail(PB_ERR_UNSUPPORTED, "only z-slab decompositions (px = py = 1) are supported");
  if (cx != 0 || cy != 0 || cz < 0 || cz >= pz) return fail(PB_ERR_ARG, "rank coordinates outside the processor grid");
  if (coordsys != 0 && coordsys != 3) return fail(PB_ERR_UNSUPPORTED, "coordsys must be 0 (Cartesian) or 3 (curvilinear)");
  const char *bs[3][2] = {{bx1, bxn}, {by1, byn}, {bz1, bzn}};
  int bcode[3][2];
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s) {
      if (bc_code(bs[d][s], &bcode[d][s]) != PB_OK) return fail(PB_ERR_ARG, "boundary strings must be NONE, PERI or SYMM");
    }
  if (device >= 0) PB_CUDA(cudaSetDevice(device));
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));

  pb_plan *pl = new pb_plan();
  pl->device = dev; pl->coordsys = coordsys;
  const int nn[3] = {nx, ny, nz}, pp[3] = {px, py, pz}, cc[3] = {cx, cy, cz};
  const double lo[3] = {x1, y1, z1}, hi[3] = {xn, yn, zn};
  for (int d = 0; d < 3; ++d) {
    pl->n[d] = nn[d]; pl->p[d] = pp[d]; pl->c[d] = cc[d]; pl->a[d] = nn[d] / pp[d];
    // parcop.f90:46-56 (nodes -> faces) then patch.f90:83-85
    const double dn = (hi[d] - lo[d]) / (double)(nn[d] - 1 > 1 ? nn[d] - 1 : 1);
    pl->x1f[d] = lo[d] - dn / 2.0; pl->xnf[d] = hi[d] + dn / 2.0;
    pl->d[d] = (pl->xnf[d] - pl->x1f[d]) / (double)nn[d];
    pl->periodic[d] = bcode[d][0] == 1;  // patch.f90:89-109: periodicity follows the lower string
    pl->bcode[d][0] = bcode[d][0]; pl->bcode[d][1] = bcode[d][1];
    pl->isym[d] = (bcode[d][0] == 2 || bcode[d][1] == 2) ? -1 : 1;
    pl->null_dir[d] = nn[d] < 4;
  }
  pl->npts = (size_t)pl->a[0] * pl->a[1] * pl->a[2];
  for (int k = 0; k < K_COUNT; ++k)
    for (int d = 0; d < 3; ++d) {
      // compact.f90:77-91: SYMM ends carry the pair (+1, -1); iop 1 takes the even member
      int rc = build_sweep(pl, pl->sw[k][d], k, d, pl->periodic[d], bcode[d][0] == 2 ? 1 : 0, bcode[d][1] == 2 ? 1 : 0);
      if (rc != PB_OK) { pb_plan_destroy(pl); return rc; }
    }
  for (int d = 0; d < 3; ++d)
    if (pl->isym[d] == -1) {
      int rc = build_sweep(pl, pl->d1_odd[d], K_D1, d, pl->periodic[d], bcode[d][0] == 2 ? -1 : 0, bcode[d][1] == 2 ? -1 : 0);
      if (rc == PB_OK)
        rc = build_sweep(pl, pl->d8_odd[d], K_D8, d, pl->periodic[d], bcode[d][0] == 2 ? -1 : 0, bcode[d][1] == 2 ? -1 : 0);
      if (rc != PB_OK) { pb_plan_destroy(pl); return rc; }
    }
  if (cudaMalloc(&pl->red_partial, sizeof(double) * 2048) != cudaSuccess ||
      cudaMalloc(&pl->red_result, sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&pl->red_host, sizeof(double)) != cudaSuccess) {
    pb_plan_destroy(pl);
    return fail(PB_ERR_CUDA, "allocation of reduction buffers failed");
  }
  *plan = pl;
  return PB_OK;
}

int pb_plan_destroy(pb_plan *pl) {
  if (!pl) return PB_OK;
  for (int k = 0; k < K_COUNT; ++k)
    for (int d = 0; d < 3; ++d) free_sweep(pl->sw[k][d]);
  for (int d = 0; d < 3; ++d) { free_sweep(pl->custom_d1[d]); free_sweep(pl->d1_odd[d]); free_sweep(pl->d8_odd[d]); }
  for (double *q : pl->scratch) cudaFree(q);
  for (auto &kv : pl->mesh) cudaFree(kv.second);
  if (pl->red_partial) cudaFree(pl->red_partial);
  if (pl->red_result) cudaFree(pl->red_result);
  if (pl->red_host) cudaFreeHost(pl->red_host);
  for (cudaEvent_t e : pl->hev) cudaEventDestroy(e);
  for (int k = 0; k < 3; ++k)
    if (pl->hs[k]) cudaStreamDestroy(pl->hs[k]);
  delete pl;
  return PB_OK;
}

int pb_plan_extents(const pb_plan *pl, int *ax, int *ay, int *az) {
  if (!pl) return fail(PB_ERR_ARG, "plan is NULL");
  if (ax) *ax = pl->a[0];
  if (ay) *ay = pl->a[1];
  if (az) *az = pl->a[2];
  return PB_OK;
}
int pb_plan_spacing(const pb_plan *pl, double *dx, double *dy, double *dz) {
  if (!pl) return fail(PB_ERR_ARG, "plan is NULL");
  if (dx) *dx = pl->d[0];
  if (dy) *dy = pl->d[1];
  if (dz) *dz = pl->d[2];
  return PB_OK;
}

int pb_plan_set_mesh(pb_plan *pl, const double *x, const double *y, const double *z, int periodic_grid) {
  if (!pl) return fail(PB_ERR_ARG, "plan is NULL");
  const size_t N = pl->npts;
  const int ax = pl->a[0], ay = pl->a[1], az = pl->a[2];
  const bool given = x && y && z;
  if (pl->coordsys == 3 && !given) return fail(PB_ERR_ARG, "curvilinear meshes need x, y, z (setup_mesh_x3)");
  double *dxg, *dyg, *dzg, *d1, *d2, *d3, *cv, *gl;
  int rc;
  if ((rc = new_mesh_arr(pl, "x", &dxg)) || (rc = new_mesh_arr(pl, "y", &dyg)) || (rc = new_mesh_arr(pl, "z", &dzg)) ||
      (rc = new_mesh_arr(pl, "d1", &d1)) || (rc = new_mesh_arr(pl, "d2", &d2)) || (rc = new_mesh_arr(pl, "d3", &d3)) ||
      (rc = new_mesh_arr(pl, "CellVol", &cv)) || (rc = new_mesh_arr(pl, "GridLen", &gl)))
    return rc;
  if (given) {
    PB_CUDA(cudaMemcpy(dxg, x, sizeof(double) * N, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(dyg, y, sizeof(double) * N, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(dzg, z, sizeof(double) * N, cudaMemcpyHostToDevice));
  } else {
    // mesh.f90:173-177: cell centres of the uniform grid, global index offset by the rank's slab
    std::vector<double> hx(N), hy(N), hz(N);
    for (int k = 0; k < az; ++k)
      for (int j = 0; j < ay; ++j)
        for (int i = 0; i < ax; ++i) {
          const size_t t = i + (size_t)ax * (j + (size_t)ay * k);
          hx[t] = pl->x1f[0] + (double)(2 * (pl->c[0] * ax + i + 1) - 1) * 0.5 * pl->d[0];
          hy[t] = pl->x1f[1] + (double)(2 * (pl->c[1] * ay + j + 1) - 1) * 0.5 * pl->d[1];
          hz[t] = pl->x1f[2] + (double)(2 * (pl->c[2] * az + k + 1) - 1) * 0.5 * pl->d[2];
        }
    PB_CUDA(cudaMemcpy(dxg, hx.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(dyg, hy.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(dzg, hz.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  }
  cudaStream_t st = 0;
  if (pl->coordsys == 0) {
    // mesh.f90:191-212
    const double dx = pl->d[0], dy = pl->d[1], dz = pl->d[2];
    const double v1 = pl->n[0] == 1 ? fmax(dy, dz) : dx, v2 = pl->n[1] == 1 ? fmax(dx, dz) : dy,
                 v3 = pl->n[2] == 1 ? fmax(dx, dy) : dz;
    PB_CUDA(launch_fill((long)N, v1, d1, st));
    PB_CUDA(launch_fill((long)N, v2, d2, st));
    PB_CUDA(launch_fill((long)N, v3, d3, st));
    PB_CUDA(launch_fill((long)N, dx * dy * dz, cv, st));
    PB_CUDA(launch_fill((long)N, fmin(v1, fmin(v2, v3)), gl, st));
    PB_CUDA(cudaStreamSynchronize(st));
    pl->mesh_set = true;
    return PB_OK;
  }
  // curvilinear, mesh.f90:244-358
  if (pl->p[2] > 1) return fail(PB_ERR_UNSUPPORTED, "curvilinear metrics on a split z axis are not implemented");
  const char *jn[9] = {"_dxdA", "_dxdB", "_dxdC", "_dydA", "_dydB", "_dydC", "_dzdA", "_dzdB", "_dzdC"};
  const char *in[9] = {"dAx", "dAy", "dAz", "dBx", "dBy", "dBz", "dCx", "dCy", "dCz"};
  double *J[9], *I[9], *det;
  for (int k = 0; k < 9; ++k)
    if ((rc = new_mesh_arr(pl, jn[k], &J[k])) || (rc = new_mesh_arr(pl, in[k], &I[k]))) return rc;
  if ((rc = new_mesh_arr(pl, "dtJ", &det))) return rc;
  const double *xyz[3] = {dxg, dyg, dzg};
  for (int a = 0; a < 3; ++a) {
    SweepPlan *sp = &pl->sw[K_D1][a];
    if (periodic_grid) {  // mesh.f90:251-304: non-periodic first derivative with one-sided ends
      if (!pl->custom_d1[a].built) {
        if ((rc = build_sweep(pl, pl->custom_d1[a], K_D1, a, false)) != PB_OK) return rc;
      }
      sp = &pl->custom_d1[a];
    }
    for (int c = 0; c < 3; ++c) {
      double *dst = J[c * 3 + a];
      if (sp->null_op) {
        // evalx on a null operator returns zeros (compact_d1.f90:51-63); the 2-D special cases
        // mesh.f90:332-334 then overwrite the diagonal entry with one
        PB_CUDA(launch_fill((long)N, (pl->n[a] == 1 && c == a) ? 1.0 : 0.0, dst, st));
      } else if ((rc = apply_dir(pl, *sp, xyz[c], dst, kStore, st)) != PB_OK) {
        return rc;  // note: the sweep already divides by dA (dv.scale = 1/d)
      }
    }
  }
  {
    const double *Jc[9]; double *Ic[9];
    for (int k = 0; k < 9; ++k) { Jc[k] = J[k]; Ic[k] = I[k]; }
    PB_CUDA(launch_metrics((long)N, Jc, pl->d[0], pl->d[1], pl->d[2], Ic, det, d1, d2, d3, cv, gl, st));
  }
  // filtered cell volumes, mesh.f90:374-383
  double *cvs, *cvg, *tmp;
  if ((rc = new_mesh_arr(pl, "CellVolS", &cvs)) || (rc = new_mesh_arr(pl, "CellVolG", &cvg)) || (rc = get_scratch(pl, 0, &tmp)))
    return rc;
  for (int which = 0; which < 2; ++which) {
    const int kind = which == 0 ? K_SF : K_GF;
    double *dst = which == 0 ? cvs : cvg;
    if ((rc = apply_dir(pl, pl->sw[kind][0], cv, dst, kStore, st)) != PB_OK) return rc;
    if ((rc = apply_dir(pl, pl->sw[kind][1], dst, tmp, kStore, st)) != PB_OK) return rc;
    if ((rc = apply_dir(pl, pl->sw[kind][2], tmp, dst, kStore, st)) != PB_OK) return rc;
  }
  PB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 9; ++k) {  // the Jacobian itself is not kept (mesh.f90:360-368)
    cudaFree(J[k]);
    pl->mesh.erase(jn[k]);
  }
  pl->mesh_set = true;
  return PB_OK;
}

int pb_getvar_device(const pb_plan *pl, const char *name, const double **dev_ptr) {
  if (!pl || !name || !dev_ptr) return fail(PB_ERR_ARG, "NULL argument");
  if (!pl->mesh_set) return fail(PB_ERR_STATE, "mesh arrays requested before pb_plan_set_mesh");
  double *d = mesh_arr(pl, name);
  if (!d || name[0] == '_') return fail(PB_ERR_ARG, std::string("unknown mesh variable: ") + name);  // parcop.f90:123-126
  *dev_ptr = d;
  return PB_OK;
}

int pb_getvar(const pb_plan *pl, const char *name, double *host_out) {
  const double *d;
  int rc = pb_getvar_device(pl, name, &d);
  if (rc != PB_OK) return rc;
  PB_CUDA(cudaMemcpy(host_out, d, sizeof(double) * pl->npts, cudaMemcpyDeviceToHost));
  return PB_OK;
}

int pb_apply(pb_plan *pl, int opcode, const double *in, double *out, void *stream) {
  if (!pl || !in || !out) return fail(PB_ERR_ARG, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long N = (long)pl->npts;
  int rc;
  switch (opcode) {
    case PB_OP_DDX: case PB_OP_DDY: case PB_OP_DDZ:
      return apply_dir(pl, pl->sw[K_D1][opcode - PB_OP_DDX], in, out, kStore, st);
    case PB_OP_DDX_ODD: case PB_OP_DDY_ODD: case PB_OP_DDZ_ODD:
      return apply_dir(pl, d1_plan(pl, opcode - PB_OP_DDX_ODD, -1), in, out, kStore, st);
    case PB_OP_DD8X: case PB_OP_DD8Y: case PB_OP_DD8Z:
      return apply_dir(pl, pl->sw[K_D8][opcode - PB_OP_DD8X], in, out, kStore, st);
    case PB_OP_D2X: case PB_OP_D2Y: case PB_OP_D2Z:
      return apply_dir(pl, pl->sw[K_D2][opcode - PB_OP_D2X], in, out, kStore, st);
    case PB_OP_DD4X: case PB_OP_DD4Y: case PB_OP_DD4Z:
      return apply_dir(pl, pl->sw[K_D4][opcode - PB_OP_DD4X], in, out, kStore, st);
    case PB_OP_DD8X_ODD: case PB_OP_DD8Y_ODD: case PB_OP_DD8Z_ODD: {
      const int d = opcode - PB_OP_DD8X_ODD;
      return apply_dir(pl, pl->d8_odd[d].built ? pl->d8_odd[d] : pl->sw[K_D8][d], in, out, kStore, st);
    }
    case PB_OP_GFILTERX: case PB_OP_GFILTERY: case PB_OP_GFILTERZ:
      return apply_dir(pl, pl->sw[K_GF][opcode - PB_OP_GFILTERX], in, out, kStore, st);
    case PB_OP_SFILTERX: case PB_OP_SFILTERY: case PB_OP_SFILTERZ:
      return apply_dir(pl, pl->sw[K_SF][opcode - PB_OP_SFILTERX], in, out, kStore, st);
    case PB_OP_LAPLACIAN: {  // operators.f90:523-526
      if (pl->coordsys != 0) return fail(PB_ERR_UNSUPPORTED, "laplacian has no curvilinear branch in the reference (operators.f90:521-560)");
      const EpiArgs acc = {EPI_ACC, 0.0, nullptr};
      if ((rc = apply_dir(pl, pl->sw[K_D2][0], in, out, kStore, st)) != PB_OK) return rc;
      if ((rc = apply_dir(pl, pl->sw[K_D2][1], in, out, acc, st)) != PB_OK) return rc;
      return apply_dir(pl, pl->sw[K_D2][2], in, out, acc, st);
    }
    case PB_OP_RING: {  // operators.f90:615-643 (L = 2), ringx/y/z :701-753
      if (!pl->mesh_set) return fail(PB_ERR_STATE, "ring needs the mesh length scales: call pb_plan_set_mesh first");
      const char *dn[3] = {"d1", "d2", "d3"};
      bool first = true;
      for (int d = 0; d < 3; ++d) {
        if (pl->n[d] == 1 || pl->sw[K_D8][d].null_op) continue;  // contributes max(.,0)
        EpiArgs e;
        e.mode = first ? EPI_RING_SET : EPI_RING_MAX;
        e.field = nullptr; e.s2 = 0.0;
        if (pl->coordsys == 0) {
          const double len = (d == 0) ? pl->d[0] : (d == 1 ? pl->d[1] : pl->d[2]);
          e.s2 = len * len;
        } else {
          e.field = mesh_arr(pl, dn[d]);
        }
        if ((rc = apply_dir(pl, pl->sw[K_D8][d], in, out, e, st)) != PB_OK) return rc;
        first = false;
      }
      if (first) PB_CUDA(launch_fill(N, 0.0, out, st));
      return PB_OK;
    }
    case PB_OP_GFILTER:  // 'smooth' operators.f90:848-853: no cell-volume weighting
      return filter3(pl, K_GF, in, out, st);
    case PB_OP_SFILTER: {  // 'spectral' operators.f90:860-862, 871-893
      if (pl->coordsys == 0) return filter3(pl, K_SF, in, out, st);
      if (!pl->mesh_set) return fail(PB_ERR_STATE, "curvilinear filter needs pb_plan_set_mesh");
      double *tmp, *tmp2;
      if ((rc = get_scratch(pl, 0, &tmp)) != PB_OK || (rc = get_scratch(pl, 1, &tmp2)) != PB_OK) return rc;
      PB_CUDA(launch_mul(N, in, mesh_arr(pl, "CellVol"), tmp2, st));
      if ((rc = apply_dir(pl, pl->sw[K_SF][0], tmp2, out, kStore, st)) != PB_OK) return rc;
      if ((rc = apply_dir(pl, pl->sw[K_SF][1], out, tmp, kStore, st)) != PB_OK) return rc;
      if ((rc = apply_dir(pl, pl->sw[K_SF][2], tmp, tmp2, kStore, st)) != PB_OK) return rc;
      PB_CUDA(launch_div(N, tmp2, mesh_arr(pl, "CellVolS"), out, st));
      return PB_OK;
    }
    default:
      return fail(PB_ERR_ARG, "unknown opcode");
  }
}

int pb_divergence(pb_plan *pl, const double *fx, const double *fy, const double *fz, double *out, void *stream) {
  if (!pl || !fx || !fy || !fz || !out) return fail(PB_ERR_ARG, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const EpiArgs acc = {EPI_ACC, 0.0, nullptr};
  int rc;
  if (pl->coordsys == 0)  // operators.f90:48-52: each flux component is odd across its own symmetry plane
    return div_cart(pl, fx, fy, fz, out, pl->isym[0], pl->isym[1], pl->isym[2], st);
  if (!pl->mesh_set) return fail(PB_ERR_STATE, "curvilinear divergence needs pb_plan_set_mesh");
  double *fA, *fB, *fC;
  if ((rc = get_scratch(pl, 1, &fA)) || (rc = get_scratch(pl, 2, &fB)) || (rc = get_scratch(pl, 3, &fC))) return rc;
  const char *in[9] = {"dAx", "dAy", "dAz", "dBx", "dBy", "dBz", "dCx", "dCy", "dCz"};
  const double *M[9];
  for (int k = 0; k < 9; ++k) M[k] = mesh_arr(pl, in[k]);
  const double *det = mesh_arr(pl, "dtJ");
  PB_CUDA(launch_contra((long)pl->npts, fx, fy, fz, M, det, fA, fB, fC, st));  // operators.f90:79-81
  if ((rc = apply_dir(pl, pl->sw[K_D1][0], fA, out, kStore, st)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, pl->sw[K_D1][1], fB, out, acc, st)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, pl->sw[K_D1][2], fC, out, acc, st)) != PB_OK) return rc;
  PB_CUDA(launch_div((long)pl->npts, out, det, out, st));  // operators.f90:90
  return PB_OK;
}

int pb_grads(pb_plan *pl, const double *in, double *gx, double *gy, double *gz, void *stream) {
  if (!pl || !in || !gx || !gy || !gz) return fail(PB_ERR_ARG, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = apply_dir(pl, pl->sw[K_D1][0], in, gx, kStore, st)) != PB_OK) return rc;  // operators.f90:191-193
  if ((rc = apply_dir(pl, pl->sw[K_D1][1], in, gy, kStore, st)) != PB_OK) return rc;
  if ((rc = apply_dir(pl, pl->sw[K_D1][2], in, gz, kStore, st)) != PB_OK) return rc;
  if (pl->coordsys == 3) {
    if (!pl->mesh_set) return fail(PB_ERR_STATE, "curvilinear gradient needs pb_plan_set_mesh");
    const char *names[9] = {"dAx", "dAy", "dAz", "dBx", "dBy", "dBz", "dCx", "dCy", "dCz"};
    const double *M[9];
    for (int k = 0; k < 9; ++k) M[k] = mesh_arr(pl, names[k]);
    PB_CUDA(launch_grad_contract((long)pl->npts, M, gx, gy, gz, st));  // operators.f90:204-209
  }
  return PB_OK;
}

// operators.f90:97-123 (Cartesian branch), parcop.f90:213-223: the divergence of each column of the tensor
int pb_divergence_tensor(pb_plan *pl, const double *fxx, const double *fxy, const double *fxz, const double *fyx,
                         const double *fyy, const double *fyz, const double *fzx, const double *fzy, const double *fzz,
                         double *dfx, double *dfy, double *dfz, void *stream) {
  if (!pl) return fail(PB_ERR_ARG, "plan is NULL");
  if (!fxx || !fxy || !fxz || !fyx || !fyy || !fyz || !fzx || !fzy || !fzz || !dfx || !dfy || !dfz)
    return fail(PB_ERR_ARG, "NULL argument");
  int rc;
  if (pl->coordsys == 3) {  // operators.f90:176-179: the divergence of each ROW of the tensor
    if ((rc = pb_divergence(pl, fxx, fxy, fxz, dfx, stream)) != PB_OK) return rc;
    if ((rc = pb_divergence(pl, fyx, fyy, fyz, dfy, stream)) != PB_OK) return rc;
    return pb_divergence(pl, fzx, fzy, fzz, dfz, stream);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int *is = pl->isym;  // operators.f90:106-120: the diagonal components are even (isym**2)
  if ((rc = div_cart(pl, fxx, fyx, fzx, dfx, 1, is[1], is[2], st)) != PB_OK) return rc;
  if ((rc = div_cart(pl, fxy, fyy, fzy, dfy, is[0], 1, is[2], st)) != PB_OK) return rc;
  return div_cart(pl, fxz, fyz, fzz, dfz, is[0], is[1], 1, st);
}

// operators.f90:645-699 with L = 1 (parcop.f90:324-333): max over the three directions of
// max(|d8(vx)|, |d8(vy)|, |d8(vz)|) * d -- nine 8th-derivative sweeps that keep a running maximum.
// The component normal to a symmetry plane is odd across it (ringx(f, ., isymX), :661-671).
// Cartesian: d is a constant, multiplied in the sweeps' epilogue (multiplying by the positive spacing
// before or after the maximum rounds identically).  Curvilinear: per direction the running maximum
// goes to a work field and one pointwise pass multiplies by mesh d1 / d2 / d3 and folds it in.
int pb_ring_vector(pb_plan *pl, const double *vx, const double *vy, const double *vz, double *out, void *stream) {
  if (!pl || !vx || !vy || !vz || !out) return fail(PB_ERR_ARG, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const double *comp[3] = {vx, vy, vz};
  const char *dn[3] = {"d1", "d2", "d3"};
  const bool curv = pl->coordsys != 0;
  double *tmp = nullptr;
  bool first = true;
  int rc;
  if (curv) {
    if (!pl->mesh_set) return fail(PB_ERR_STATE, "curvilinear ringV needs pb_plan_set_mesh");
    if ((rc = get_scratch(pl, 0, &tmp)) != PB_OK) return rc;
  }
  for (int d = 0; d < 3; ++d) {
    if (pl->n[d] == 1 || pl->sw[K_D8][d].null_op) continue;  // ringx/y/z give zero there (operators.f90:709-712)
    for (int c = 0; c < 3; ++c) {
      SweepPlan &sp = (c == d && pl->d8_odd[d].built) ? pl->d8_odd[d] : pl->sw[K_D8][d];
      EpiArgs e;
      e.field = nullptr;
      if (curv) { e.mode = c == 0 ? EPI_RING_SET : EPI_RING_MAX; e.s2 = 1.0; }
      else { e.mode = first ? EPI_RING_SET : EPI_RING_MAX; e.s2 = pl->d[d]; }
      if ((rc = apply_dir(pl, sp, comp[c], curv ? tmp : out, e, st)) != PB_OK) return rc;
      if (!curv) first = false;
    }
    if (curv) {
      PB_CUDA(launch_mul_max((long)pl->npts, tmp, mesh_arr(pl, dn[d]), out, first ? 1 : 0, st));
      first = false;
    }
  }
  if (first) PB_CUDA(launch_fill((long)pl->npts, 0.0, out, st));
  return PB_OK;
}

int pb_rk4_stage(pb_plan *pl, long n, double dt, double A, double B, const double *F, double *PHI, double *U, void *stream) {
  if (!pl || !F || !PHI || !U || n < 0) return fail(PB_ERR_ARG, "bad argument");
  PB_CUDA(launch_rk4_stage(n, dt, A, B, F, PHI, U, (cudaStream_t)stream));
  return PB_OK;
}

int pb_reduce(pb_plan *pl, int kind, long n, const double *d_val, double *host_out, void *stream) {
  if (!pl || !d_val || !host_out || n <= 0 || kind < 0 || kind > 2) return fail(PB_ERR_ARG, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  PB_CUDA(launch_reduce(kind, n, d_val, pl->red_partial, 2048, pl->red_result, st));
  PB_CUDA(cudaMemcpyAsync(pl->red_host, pl->red_result, sizeof(double), cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  *host_out = *pl->red_host;
  return PB_OK;
}

int pb_reduce_device(pb_plan *pl, int kind, long n, const double *d_val, double *d_out, void *stream) {
  if (!pl || !d_val || !d_out || n <= 0 || kind < 0 || kind > 2) return fail(PB_ERR_ARG, "bad argument");
  PB_CUDA(launch_reduce(kind, n, d_val, pl->red_partial, 2048, d_out, (cudaStream_t)stream));
  return PB_OK;
}

// ---- z-slab pieces -------------------------------------------------------------------------------
// the z sweep plan of a z-slab opcode: the odd variants exist only on axes with a symmetry plane
static SweepPlan &zplan(pb_plan *pl, int zop, int k) {
  if (zop == PB_OP_DDZ_ODD) return d1_plan(pl, 2, -1);
  if (zop == PB_OP_DD8Z_ODD && pl->d8_odd[2].built) return pl->d8_odd[2];
  return pl->sw[k][2];
}

static int zop_kind(int zop) {
  switch (zop) {
    case PB_OP_DDZ: case PB_OP_DDZ_ODD: return K_D1;
    case PB_OP_DD8Z: case PB_OP_DD8Z_ODD: return K_D8;
    case PB_OP_D2Z: return K_D2;
    case PB_OP_DD4Z: return K_D4;
    case PB_OP_SFILTERZ: return K_SF;
    case PB_OP_GFILTERZ: return K_GF;
    default: return -1;
  }
}

int pb_z_pack_halo(pb_plan *pl, int zop, const double *d_val, double *send_lo, double *send_hi, void *stream) {
  const int k = zop_kind(zop);
  if (!pl || k < 0 || !d_val || !send_lo || !send_hi) return fail(PB_ERR_ARG, "bad argument");
  const SweepPlan &sp = zplan(pl, zop, k);
  if (sp.null_op) return PB_OK;
  const long plane = (long)pl->a[0] * pl->a[1];
  PB_CUDA(launch_pack_planes(d_val, plane, pl->a[2], sp.st.nor, send_lo, send_hi, (cudaStream_t)stream));
  return PB_OK;
}

int pb_z_local(pb_plan *pl, int zop, const double *d_val, const double *recv_lo, const double *recv_hi, double *d_out,
               double *iface_local, void *stream) {
  const int k = zop_kind(zop);
  if (!pl || k < 0 || !d_val || !d_out) return fail(PB_ERR_ARG, "bad argument");
  SweepPlan &sp = zplan(pl, zop, k);
  cudaStream_t st = (cudaStream_t)stream;
  if (sp.null_op || !sp.split) {
    // not distributed: the whole operator in one call
    return apply_dir(pl, sp, d_val, d_out, kStore, st);
  }
  if ((!sp.dev.phys_lo && !recv_lo) || (!sp.dev.phys_hi && !recv_hi)) return fail(PB_ERR_ARG, "missing halo buffer");
  // the local pass already scales and adds back: pb_z_finish only adds the rank-level correction
  if (sp.st.implicit && !iface_local) return fail(PB_ERR_ARG, "missing interface buffer");
  PB_CUDA(launch_sweep_yz(sp.st.fam, 0, sp.dev, d_val, d_out, recv_lo, recv_hi, sp.st.implicit ? iface_local : nullptr, kStore, st));
  return PB_OK;
}

int pb_z_finish(pb_plan *pl, int zop, const double *d_val, const double *iface_all, double *d_out, void *stream) {
  const int k = zop_kind(zop);
  if (!pl || k < 0 || !d_out) return fail(PB_ERR_ARG, "bad argument");
  SweepPlan &sp = zplan(pl, zop, k);
  if (sp.null_op || !sp.split || !sp.st.implicit) return PB_OK;  // explicit operators are complete after pb_z_local
  if (!iface_all) return fail(PB_ERR_ARG, "missing interface buffer");
  (void)d_val;
  const long plane = (long)pl->a[0] * pl->a[1];
  PB_CUDA(launch_z_finish(d_out, plane, pl->a[2], sp.RC, sp.GR, sp.np, sp.rank_mask, sp.zone_lo, sp.zone_hi, iface_all,
                          sp.dev.scale, (cudaStream_t)stream));
  return PB_OK;
}

int pb_z_ring_info(pb_plan *pl, int zop, int *need_f, int *need_b, int *nup, int *ndn) {
  const int k = zop_kind(zop);
  if (!pl || k < 0) return fail(PB_ERR_ARG, "bad argument");
  const SweepPlan &sp = zplan(pl, zop, k);
  const bool ok = !sp.null_op && sp.split && sp.st.implicit && sp.xr_ok;
  if (need_f) *need_f = ok ? sp.xr.need_f : 0;
  if (need_b) *need_b = ok ? sp.xr.need_b : 0;
  if (nup) *nup = ok ? sp.xr.nup : 0;
  if (ndn) *ndn = ok ? sp.xr.ndn : 0;
  return PB_OK;
}

int pb_z_ring_mode(pb_plan *pl, int zop) {
  const int k = zop_kind(zop);
  if (!pl || k < 0) return 0;
  const SweepPlan &sp = zplan(pl, zop, k);
  if (sp.null_op || !sp.split || !sp.st.implicit || !sp.xr_ok) return 0;
  return sp.xr.early ? 2 : 1;
}

static int epi_from(int mode, double s2, EpiArgs *e) {
  if (mode < EPI_STORE || mode > EPI_RING_MAX) return PB_ERR_ARG;
  e->mode = mode; e->s2 = s2; e->field = nullptr;
  return PB_OK;
}

int pb_z_ring(pb_plan *pl, int zop, const double *d_val, const double *recv_lo, const double *recv_hi, double *d_out,
              const pb_xring *x, int epi_mode, double s2, void *stream) {
  const int k = zop_kind(zop);
  if (!pl || k < 0 || !d_val || !d_out || !x) return fail(PB_ERR_ARG, "bad argument");
  SweepPlan &sp = zplan(pl, zop, k);
  if (sp.null_op || !sp.split || !sp.st.implicit || !sp.xr_ok)
    return fail(PB_ERR_UNSUPPORTED, "this z operator has no fused ring form on this partition (pb_z_ring_info reports zeros): use pb_z_local / pb_z_finish");
  if ((!sp.dev.phys_lo && !recv_lo) || (!sp.dev.phys_hi && !recv_hi)) return fail(PB_ERR_ARG, "missing halo buffer");
  EpiArgs epi;
  if (epi_from(epi_mode, s2, &epi) != PB_OK) return fail(PB_ERR_ARG, "bad epilogue mode");
  XRing xr = sp.xr;
  xr.epoch = x->epoch;
  xr.plane = (long)pl->a[0] * pl->a[1];
  xr.en_in = (const unsigned long long *)x->en_in;
  xr.st_in = (const unsigned long long *)x->st_in;
  if ((xr.need_f > 0 && !xr.en_in) || (xr.need_b > 0 && !xr.st_in) || x->epoch == 0) return fail(PB_ERR_ARG, "missing record buffer / epoch 0");
#ifndef PB_EMULATE
  if (((uintptr_t)x->en_in | (uintptr_t)x->st_in) & 31) return fail(PB_ERR_ARG, "record buffers must be 32-byte aligned");
  for (int h = 0; h < kXHops; ++h)
    if (((uintptr_t)x->en_out[h] | (uintptr_t)x->st_out[h]) & 31) return fail(PB_ERR_ARG, "record buffers must be 32-byte aligned");
#endif
  for (int h = 0; h < kXHops; ++h) {
    xr.en_out[h] = (unsigned long long *)x->en_out[h];
    xr.st_out[h] = (unsigned long long *)x->st_out[h];
    if ((h < xr.nup && !xr.en_out[h]) || (h < xr.ndn && !xr.st_out[h])) return fail(PB_ERR_ARG, "missing peer record buffer");
  }
  static const int nopoll = getenv("PB_XR_NOPOLL") ? atoi(getenv("PB_XR_NOPOLL")) : 0;  // timing experiments only: results are wrong
  xr.nopoll = nopoll;
  xr.push = 0;
  if (x->push) {
    if (x->npeers < 0 || x->npeers > 2 || !x->counter || x->halo_epoch == 0) return fail(PB_ERR_ARG, "bad halo push arguments");
    const long plane = xr.plane;
    const int h = sp.st.nor;
    if ((plane * h) % 2) return fail(PB_ERR_UNSUPPORTED, "halo push needs 16-byte aligned planes");
    xr.push = 1;
    xr.npeers = x->npeers;
    xr.push_n = plane * h;
    xr.push_src[0] = d_val;                                  // my first planes -> the lower neighbour's upper halo
    xr.push_src[1] = d_val + (long)(pl->a[2] - h) * plane;   // my last planes -> the upper neighbour's lower halo
    xr.push_dst[0] = (double *)x->halo_dst[0];
    xr.push_dst[1] = (double *)x->halo_dst[1];
    for (int q = 0; q < x->npeers; ++q) {
      if (!x->flag_remote[q] || !x->flag_local[q]) return fail(PB_ERR_ARG, "missing flag address");
      xr.hflag_remote[q] = (unsigned long long *)x->flag_remote[q];
      xr.hflag_local[q] = (const volatile unsigned long long *)x->flag_local[q];
    }
    xr.hepoch = x->halo_epoch;
    xr.hcounter = (unsigned int *)x->counter;
  }
  // both halo planes or none reach the kernel: a rank at a physical end passes its own buffer for the unused side
  const double *lo = recv_lo ? recv_lo : recv_hi, *hi = recv_hi ? recv_hi : recv_lo;
  const cudaError_t err = launch_sweep_ring(sp.st.fam, 0, sp.devx, d_val, d_out, lo, hi, &xr, epi, (cudaStream_t)stream);
  if (err == cudaErrorNotSupported) return fail(PB_ERR_UNSUPPORTED, "ring kernel: geometry not supported");
  PB_CUDA(err);
  return PB_OK;
}

// one directional operator with a composite epilogue (what pb_apply's laplacian / ring do internally),
// for callers that assemble composites across ranks
int pb_apply_epi(pb_plan *pl, int opcode, const double *in, double *out, int epi_mode, double s2, void *stream) {
  if (!pl || !in || !out) return fail(PB_ERR_ARG, "NULL argument");
  EpiArgs epi;
  if (epi_from(epi_mode, s2, &epi) != PB_OK) return fail(PB_ERR_ARG, "bad epilogue mode");
  cudaStream_t st = (cudaStream_t)stream;
  int kind, dir;
  if (opcode >= PB_OP_DDX && opcode <= PB_OP_DDZ) { kind = K_D1; dir = opcode - PB_OP_DDX; }
  else if (opcode >= PB_OP_DD8X && opcode <= PB_OP_DD8Z) { kind = K_D8; dir = opcode - PB_OP_DD8X; }
  else if (opcode >= PB_OP_D2X && opcode <= PB_OP_D2Z) { kind = K_D2; dir = opcode - PB_OP_D2X; }
  else if (opcode >= PB_OP_DDX_ODD && opcode <= PB_OP_DDZ_ODD) return apply_dir(pl, d1_plan(pl, opcode - PB_OP_DDX_ODD, -1), in, out, epi, st);
  else if (opcode >= PB_OP_DD8X_ODD && opcode <= PB_OP_DD8Z_ODD) {
    const int d = opcode - PB_OP_DD8X_ODD;
    return apply_dir(pl, pl->d8_odd[d].built ? pl->d8_odd[d] : pl->sw[K_D8][d], in, out, epi, st);
  } else return fail(PB_ERR_ARG, "pb_apply_epi: directional derivative opcodes only");
  return apply_dir(pl, pl->sw[kind][dir], in, out, epi, st);
}

int pb_peer_exchange(int ncopies, void *const *dst, const void *const *src, const size_t *bytes, int npeers,
                     void *const *remote_flags, void *const *local_flags, unsigned long long epoch, void *counter,
                     void *stream) {
  if (ncopies < 0 || ncopies > 4 || npeers < 0 || npeers > 2 || !counter) return fail(PB_ERR_ARG, "bad peer exchange");
  PeerExchange x;
  memset(&x, 0, sizeof(x));
  x.ncopies = ncopies; x.npeers = npeers; x.epoch = epoch; x.counter = (unsigned int *)counter;
  for (int c = 0; c < ncopies; ++c) {
    if (!dst[c] || !src[c] || (bytes[c] & 15) || ((uintptr_t)dst[c] & 15) || ((uintptr_t)src[c] & 15))
      return fail(PB_ERR_ARG, "peer copies must be 16-byte aligned");
    x.dst[c] = dst[c]; x.src[c] = src[c]; x.bytes[c] = bytes[c];
  }
  for (int p = 0; p < npeers; ++p) {
    if (!remote_flags[p] || !local_flags[p]) return fail(PB_ERR_ARG, "missing flag address");
    x.remote_flag[p] = (unsigned long long *)remote_flags[p];
    x.local_flag[p] = (const volatile unsigned long long *)local_flags[p];
  }
  const cudaError_t err = launch_peer_exchange(x, (cudaStream_t)stream);
  if (err == cudaErrorNotSupported) return fail(PB_ERR_UNSUPPORTED, "peer exchange needs the CUDA build");
  PB_CUDA(err);
  return PB_OK;
}

int pb_z_exchange_ranks(pb_plan *pl, int zop, unsigned long long *mask) {
  const int k = zop_kind(zop);
  if (!pl || k < 0 || !mask) return fail(PB_ERR_ARG, "bad argument");
  const SweepPlan &sp = zplan(pl, zop, k);
  *mask = (sp.null_op || !sp.split || !sp.st.implicit) ? 0ull : sp.rank_mask;
  return PB_OK;
}

// ---- host-array wrappers -------------------------------------------------------------------------
// Host arrays in / out.  PCIe dominates (2 x 8 bytes per point against 16-48 bytes of HBM traffic
// at ~100x the bandwidth), so the field moves in slabs and the three engines overlap:
//   one-direction operators: slab s is copied in while slab s-1 is swept and slab s-2 copied out
//     (slabs of z-planes for x / y sweeps, slabs of y-rows for z sweeps);
//   compact filter (x -> y -> z): x and y sweeps run per z-slab as the slabs arrive, the z sweep
//     runs per y-slab and each finished slab leaves while the next is swept (every output plane depends
//     on every input plane, so the two copies cannot overlap);
//   Gaussian filter: its z sweep is explicit, so a z-slab of the result leaves as soon as the next slab
//     has passed x and y (apply_z_explicit_slab) -- copies in and out overlap as for one direction.
// Anything else (composites, curvilinear weighting, null or split directions) takes the plain path.
static int host_pipeline_ready(pb_plan *pl, size_t nev) {
  for (int k = 0; k < 3; ++k)
    if (!pl->hs[k]) PB_CUDA(cudaStreamCreateWithFlags(&pl->hs[k], cudaStreamNonBlocking));
  while (pl->hev.size() < nev) {
    cudaEvent_t e;
    PB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    pl->hev.push_back(e);
  }
  return PB_OK;
}

// copies `cnt` slices (z-planes when along_z, else y-rows of every plane) between host and device
static int copy_slab(pb_plan *pl, void *dst, const void *src, bool along_z, int first, int cnt, cudaMemcpyKind kind, cudaStream_t st) {
  const size_t ax = pl->a[0], ay = pl->a[1], az = pl->a[2];
  if (along_z) {
    const size_t off = (size_t)first * ax * ay * sizeof(double);
    PB_CUDA(cudaMemcpyAsync((char *)dst + off, (const char *)src + off, (size_t)cnt * ax * ay * sizeof(double), kind, st));
  } else {
    const size_t off = (size_t)first * ax * sizeof(double), pitch = ax * ay * sizeof(double);
    PB_CUDA(cudaMemcpy2DAsync((char *)dst + off, pitch, (const char *)src + off, pitch, (size_t)cnt * ax * sizeof(double), az, kind, st));
  }
  return PB_OK;
}

static long sz_chunk(const SweepPlan *sz) { return (sz && sz->built && sz->dev.C > 0) ? sz->dev.C : 1L << 30; }

static int host_apply_pipelined(pb_plan *pl, SweepPlan *sx, SweepPlan *sy, SweepPlan *sz, const double *h_val, double *h_out,
                                bool *done) {
  *done = false;
  SweepPlan *all[3] = {sx, sy, sz};
  int nsw = 0;
  for (SweepPlan *s : all)
    if (s) { if (!s->built || s->null_op || s->split) return PB_OK; ++nsw; }
  const int az = pl->a[2], ay = pl->a[1];
  static const int ns_env = getenv("PB_HOST_SLABS") ? atoi(getenv("PB_HOST_SLABS")) : 0;
  int NS = ns_env > 0 ? ns_env : 16;  // slabs per field: fill and drain of the pipeline cost 1 / NS of the copy time each
  while (NS > 1 && (az < 2 * NS || ay < 2 * NS)) NS >>= 1;
  if (nsw == 0 || NS < 4) return PB_OK;
  int rc;
  double *d0, *d1, *d2;
  if ((rc = get_scratch(pl, 4, &d0)) || (rc = get_scratch(pl, 5, &d1)) || (rc = get_scratch(pl, 6, &d2))) return rc;
  if ((rc = host_pipeline_ready(pl, 2 * NS + 2)) != PB_OK) return rc;
  cudaStream_t sIn = pl->hs[0], sC = pl->hs[1], sOut = pl->hs[2];
  auto slab = [&](int n, int s, int *first, int *cnt) { *first = (int)((long)n * s / NS); *cnt = (int)((long)n * (s + 1) / NS) - *first; };
  PB_CUDA(cudaStreamSynchronize(0));  // earlier work of the caller on the default stream
  const bool zin = sx || sy;           // the input arrives in z-slabs unless only z is swept
  // Explicit z sweep after x / y (the Gaussian filter): a z-slab of the result needs the x / y-swept planes of
  // its own slab and four planes of either neighbour, so slab s - 1 is swept along z and sent back as soon as
  // slab s has passed x and y -- input and output copies overlap instead of following each other.  On a
  // periodic axis slab 0 also needs the last planes and leaves last.
  static const bool no_stream = getenv("PB_HOST_NO_ZSTREAM") != nullptr;
  int NZ = NS;  // slabs of whole chunks
  while (NZ >= 4 && az % (NZ * sz_chunk(sz)) != 0) NZ >>= 1;
  if (sz && zin && !sz->st.implicit && !no_stream && NZ >= 4 && az / NZ >= 8 &&
      (sz->dev.wrap || (sz->dev.phys_lo && sz->dev.phys_hi))) {
    const bool wrap = sz->dev.wrap != 0;
    const int NS = NZ;
    auto slab = [&](int n, int s, int *first, int *cnt) { *first = (int)((long)n * s / NS); *cnt = (int)((long)n * (s + 1) / NS) - *first; };
    double *Y = (sx && sy) ? d2 : d1, *fin = (Y == d1) ? d2 : d1;  // x / y-swept field, result
    auto zsweep = [&](int s) -> int {
      int f, c, r2;
      slab(az, s, &f, &c);
      if ((r2 = apply_z_explicit_slab(pl, *sz, Y, fin, f, c, sC)) != PB_OK) return r2;
      PB_CUDA(cudaEventRecord(pl->hev[NS + s], sC));
      PB_CUDA(cudaStreamWaitEvent(sOut, pl->hev[NS + s], 0));
      return copy_slab(pl, h_out, fin, true, f, c, cudaMemcpyDeviceToHost, sOut);
    };
    for (int s = 0; s < NS; ++s) {
      int f, c;
      slab(az, s, &f, &c);
      if ((rc = copy_slab(pl, d0, h_val, true, f, c, cudaMemcpyHostToDevice, sIn)) != PB_OK) return rc;
      PB_CUDA(cudaEventRecord(pl->hev[s], sIn));
      PB_CUDA(cudaStreamWaitEvent(sC, pl->hev[s], 0));
      const double *cur = d0;
      if (sx) { if ((rc = apply_dir_slab(pl, *sx, cur, d1, f, c, sC)) != PB_OK) return rc; cur = d1; }
      if (sy) { if ((rc = apply_dir_slab(pl, *sy, cur, cur == d1 ? d2 : d1, f, c, sC)) != PB_OK) return rc; }
      if (s >= 1 && (s - 1 > 0 || !wrap)) { if ((rc = zsweep(s - 1)) != PB_OK) return rc; }
    }
    if ((rc = zsweep(NS - 1)) != PB_OK) return rc;
    if (wrap) { if ((rc = zsweep(0)) != PB_OK) return rc; }
    PB_CUDA(cudaStreamSynchronize(sOut));
    PB_CUDA(cudaStreamSynchronize(sC));
    *done = true;
    return PB_OK;
  }
  for (int s = 0; s < NS; ++s) {
    int f, c;
    slab(zin ? az : ay, s, &f, &c);
    if ((rc = copy_slab(pl, d0, h_val, zin, f, c, cudaMemcpyHostToDevice, sIn)) != PB_OK) return rc;
    PB_CUDA(cudaEventRecord(pl->hev[s], sIn));
    PB_CUDA(cudaStreamWaitEvent(sC, pl->hev[s], 0));
    const double *cur = d0;
    double *nxt = d1;
    if (sx) { if ((rc = apply_dir_slab(pl, *sx, cur, nxt, f, c, sC)) != PB_OK) return rc; cur = nxt; nxt = (nxt == d1) ? d2 : d1; }
    if (sy) { if ((rc = apply_dir_slab(pl, *sy, cur, nxt, f, c, sC)) != PB_OK) return rc; cur = nxt; nxt = (nxt == d1) ? d2 : d1; }
    if (sz && !zin) { if ((rc = apply_dir_slab(pl, *sz, cur, nxt, f, c, sC)) != PB_OK) return rc; cur = nxt; }
    if (!(sz && zin)) {  // this slab is final: send it back
      PB_CUDA(cudaEventRecord(pl->hev[NS + s], sC));
      PB_CUDA(cudaStreamWaitEvent(sOut, pl->hev[NS + s], 0));
      if ((rc = copy_slab(pl, h_out, cur, zin, f, c, cudaMemcpyDeviceToHost, sOut)) != PB_OK) return rc;
    }
  }
  if (sz && zin) {  // the z sweep needs whole lines: it runs per y-slab once every z-slab has passed x / y
    const double *cur = (sx && sy) ? d2 : d1;
    double *fin = (cur == d1) ? d2 : d1;
    for (int s = 0; s < NS; ++s) {
      int f, c;
      slab(ay, s, &f, &c);
      if ((rc = apply_dir_slab(pl, *sz, cur, fin, f, c, sC)) != PB_OK) return rc;
      PB_CUDA(cudaEventRecord(pl->hev[NS + s], sC));
      PB_CUDA(cudaStreamWaitEvent(sOut, pl->hev[NS + s], 0));
      if ((rc = copy_slab(pl, h_out, fin, false, f, c, cudaMemcpyDeviceToHost, sOut)) != PB_OK) return rc;
    }
  }
  PB_CUDA(cudaStreamSynchronize(sOut));
  PB_CUDA(cudaStreamSynchronize(sC));
  *done = true;
  return PB_OK;
}

int pb_host_apply(pb_plan *pl, int opcode, const double *h_val, double *h_out) {
  if (!pl || !h_val || !h_out) return fail(PB_ERR_ARG, "NULL argument");
  int rc;
  {
    SweepPlan *sx = nullptr, *sy = nullptr, *sz = nullptr;
    int kind = -1, dir = -1;
    if (opcode >= PB_OP_DDX && opcode <= PB_OP_DDZ) { kind = K_D1; dir = opcode - PB_OP_DDX; }
    else if (opcode >= PB_OP_DD8X && opcode <= PB_OP_DD8Z) { kind = K_D8; dir = opcode - PB_OP_DD8X; }
    else if (opcode >= PB_OP_D2X && opcode <= PB_OP_D2Z) { kind = K_D2; dir = opcode - PB_OP_D2X; }
    else if (opcode >= PB_OP_DD4X && opcode <= PB_OP_DD4Z) { kind = K_D4; dir = opcode - PB_OP_DD4X; }
    else if (opcode >= PB_OP_GFILTERX && opcode <= PB_OP_GFILTERZ) { kind = K_GF; dir = opcode - PB_OP_GFILTERX; }
    else if (opcode >= PB_OP_SFILTERX && opcode <= PB_OP_SFILTERZ) { kind = K_SF; dir = opcode - PB_OP_SFILTERX; }
    else if (opcode == PB_OP_GFILTER) kind = K_GF;
    else if (opcode == PB_OP_SFILTER && pl->coordsys == 0) kind = K_SF;
    if (kind >= 0) {
      if (dir == 0 || dir < 0) sx = &pl->sw[kind][0];
      if (dir == 1 || dir < 0) sy = &pl->sw[kind][1];
      if (dir == 2 || dir < 0) sz = &pl->sw[kind][2];
      bool done = false;
      if ((rc = host_apply_pipelined(pl, sx, sy, sz, h_val, h_out, &done)) != PB_OK) return rc;
      if (done) return PB_OK;
    }
  }
  double *din, *dout;
  if ((rc = get_scratch(pl, 4, &din)) || (rc = get_scratch(pl, 5, &dout))) return rc;
  const size_t bytes = sizeof(double) * pl->npts;
  PB_CUDA(cudaMemcpyAsync(din, h_val, bytes, cudaMemcpyHostToDevice, 0));
  if ((rc = pb_apply(pl, opcode, din, dout, nullptr)) != PB_OK) return rc;
  PB_CUDA(cudaMemcpyAsync(h_out, dout, bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaStreamSynchronize(0));
  return PB_OK;
}

// three host fields in, one out; sel = nullptr: the plan's own divergence (pb_divergence), otherwise
// the Cartesian sum with the given symmetry selectors
static int host_div(pb_plan *pl, const double *h_fx, const double *h_fy, const double *h_fz, double *h_out, const int *sel) {
  double *d[4];
  int rc;
  for (int k = 0; k < 4; ++k)
    if ((rc = get_scratch(pl, 4 + k, &d[k]))) return rc;
  const size_t bytes = sizeof(double) * pl->npts;
  PB_CUDA(cudaMemcpyAsync(d[0], h_fx, bytes, cudaMemcpyHostToDevice, 0));
  PB_CUDA(cudaMemcpyAsync(d[1], h_fy, bytes, cudaMemcpyHostToDevice, 0));
  PB_CUDA(cudaMemcpyAsync(d[2], h_fz, bytes, cudaMemcpyHostToDevice, 0));
  if (sel) rc = div_cart(pl, d[0], d[1], d[2], d[3], sel[0], sel[1], sel[2], nullptr);
  else rc = pb_divergence(pl, d[0], d[1], d[2], d[3], nullptr);
  if (rc != PB_OK) return rc;
  PB_CUDA(cudaMemcpyAsync(h_out, d[3], bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaStreamSynchronize(0));
  return PB_OK;
}

int pb_host_divergence(pb_plan *pl, const double *h_fx, const double *h_fy, const double *h_fz, double *h_out) {
  if (!pl || !h_fx || !h_fy || !h_fz || !h_out) return fail(PB_ERR_ARG, "NULL argument");
  return host_div(pl, h_fx, h_fy, h_fz, h_out, nullptr);
}

int pb_host_grads(pb_plan *pl, const double *h_val, double *h_gx, double *h_gy, double *h_gz) {
  if (!pl || !h_val || !h_gx || !h_gy || !h_gz) return fail(PB_ERR_ARG, "NULL argument");
  double *d[4];
  int rc;
  for (int k = 0; k < 4; ++k)
    if ((rc = get_scratch(pl, 4 + k, &d[k]))) return rc;
  const size_t bytes = sizeof(double) * pl->npts;
  PB_CUDA(cudaMemcpyAsync(d[0], h_val, bytes, cudaMemcpyHostToDevice, 0));
  if ((rc = pb_grads(pl, d[0], d[1], d[2], d[3], nullptr)) != PB_OK) return rc;
  PB_CUDA(cudaMemcpyAsync(h_gx, d[1], bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaMemcpyAsync(h_gy, d[2], bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaMemcpyAsync(h_gz, d[3], bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaStreamSynchronize(0));
  return PB_OK;
}

int pb_host_divergence_tensor(pb_plan *pl, const double *const *h_f9, double *const *h_out3) {
  // h_f9 = {fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz}; one divergence at a time through four staging fields
  if (!pl || !h_f9 || !h_out3) return fail(PB_ERR_ARG, "NULL argument");
  int rc;
  if (pl->coordsys == 3) {  // rows of the tensor through the curvilinear divergence (operators.f90:176-179)
    for (int c = 0; c < 3; ++c)
      if ((rc = host_div(pl, h_f9[3 * c], h_f9[3 * c + 1], h_f9[3 * c + 2], h_out3[c], nullptr)) != PB_OK) return rc;
    return PB_OK;
  }
  for (int c = 0; c < 3; ++c) {
    int sel[3] = {pl->isym[0], pl->isym[1], pl->isym[2]};
    sel[c] = 1;  // operators.f90:106,113,120: isym**2 on the diagonal
    if ((rc = host_div(pl, h_f9[c], h_f9[3 + c], h_f9[6 + c], h_out3[c], sel)) != PB_OK) return rc;
  }
  return PB_OK;
}

int pb_host_ring_vector(pb_plan *pl, const double *h_vx, const double *h_vy, const double *h_vz, double *h_out) {
  if (!pl || !h_vx || !h_vy || !h_vz || !h_out) return fail(PB_ERR_ARG, "NULL argument");
  double *d[4];
  int rc;
  for (int k = 0; k < 4; ++k)
    if ((rc = get_scratch(pl, 4 + k, &d[k]))) return rc;
  const size_t bytes = sizeof(double) * pl->npts;
  PB_CUDA(cudaMemcpyAsync(d[0], h_vx, bytes, cudaMemcpyHostToDevice, 0));
  PB_CUDA(cudaMemcpyAsync(d[1], h_vy, bytes, cudaMemcpyHostToDevice, 0));
  PB_CUDA(cudaMemcpyAsync(d[2], h_vz, bytes, cudaMemcpyHostToDevice, 0));
  if ((rc = pb_ring_vector(pl, d[0], d[1], d[2], d[3], nullptr)) != PB_OK) return rc;
  PB_CUDA(cudaMemcpyAsync(h_out, d[3], bytes, cudaMemcpyDeviceToHost, 0));
  PB_CUDA(cudaStreamSynchronize(0));
  return PB_OK;
}

}  // extern "C"

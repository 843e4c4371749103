// See tables.hpp.  Host only.
#include "tables.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace pb {

namespace {

// lower closure rows are given; the upper ones are their mirror image (sign -1 for the
// antisymmetric first derivative), as in stencils.f90:243-254 and the analogous lines of each set.
void mirror_closures(Stencil &s, double rhs_sign) {
  for (int r = 0; r < 4; ++r) {
    for (int l = 0; l < s.ncl; ++l) s.alb_hi[r][l] = s.alb_lo[3 - r][s.ncl - 1 - l];
    for (int l = 0; l < s.ncr; ++l) s.arb_hi[r][l] = rhs_sign * s.arb_lo[3 - r][s.ncr - 1 - l];
  }
}

void set_row(double *dst, std::initializer_list<double> v) {
  int i = 0;
  for (double x : v) dst[i++] = x;
}

// filters are evaluated as f + A^-1 (B - A) f: subtract the lhs row from the centre of the rhs row
// (stencils.f90:831-834 for the compact filter, :1472-1475 for the Gaussian)
void difference_form(Stencil &s) {
  const int off = s.nor - s.nol;
  for (int l = 0; l < s.ncl; ++l) s.ari[off + l] -= s.ali[l];
  for (int r = 0; r < 4; ++r)
    for (int l = 0; l < s.ncl; ++l) {
      s.arb_lo[r][off + l] -= s.alb_lo[r][l];
      s.arb_hi[r][off + l] -= s.alb_hi[r][l];
    }
}

// The R4 right-hand side is evaluated as sum_l w_l (v_l - v_ref) + sigma v_ref with v_ref the row's
// own point (interior) or the boundary point (closure rows, compact_r4.f90:123-126,174-177).  The
// slot of the reference column carries sigma, the row sum of the weights: zero for the 8th
// derivative and for both filters in the reference's difference form  v + A^-1 (B - A) v.
// (direct = true gives  A out = B v  without an add-back; it is NOT used for the compact filter:
// its matrix is nearly singular at the Nyquist wavenumber, cond ~ 2e3, and the direct form loses
// three digits against the reference -- measured 3e-13 instead of 2e-15 per application.)
// At an antisymmetric end (bc = -1) the reference evaluates the closure rows as plain dot products
// (compact_r4.f90:112-115,163-166): the folded weights do not sum to zero, and sigma carries their sum.
void set_reference_slots(Stencil &s, bool direct, bool direct_lo = false, bool direct_hi = false) {
  auto rowsum = [](const double *w) { double t = 0.0; for (int l = 0; l < 9; ++l) t += w[l]; return t; };
  s.ari[4] = direct ? rowsum(s.ari) : 0.0;
  for (int i = 0; i < 4; ++i) {
    s.arb_lo[i][4 - i] = (direct || direct_lo) ? rowsum(s.arb_lo[i]) : 0.0;
    s.arb_hi[i][7 - i] = (direct || direct_hi) ? rowsum(s.arb_hi[i]) : 0.0;
  }
}

// Symmetry planes (stencils.f90:2390-2453, lower/upper_symm_weights_int): the plane lies half a
// cell outside the first / last point, so the interior stencil's out-of-range weight at distance
// i beyond the plane folds onto the in-range point at distance i inside it, with the sign of the
// field's parity (syml for the lhs, symr for the rhs).  Every closure row of that end is the
// folded interior row.
void fold_lower(Stencil &s, int syml, int symr) {
  for (int r = 0; r < 4; ++r) {
    for (int l = 0; l < 5; ++l) s.alb_lo[r][l] = l < s.ncl ? s.ali[l] : 0.0;
    for (int l = 0; l < 9; ++l) s.arb_lo[r][l] = l < s.ncr ? s.ari[l] : 0.0;
    for (int k = s.nol - r, i = 1; i <= k; ++i) {  // columns k-i (outside) -> k+i-1 (inside), 0-based
      s.alb_lo[r][k + i - 1] += syml * s.alb_lo[r][k - i];
      s.alb_lo[r][k - i] = 0.0;
    }
    for (int k = s.nor - r, i = 1; i <= k; ++i) {
      s.arb_lo[r][k + i - 1] += symr * s.arb_lo[r][k - i];
      s.arb_lo[r][k - i] = 0.0;
    }
  }
}
void fold_upper(Stencil &s, int syml, int symr) {
  for (int r = 0; r < 4; ++r) {  // row r is point n-4+r: 3-r points remain above it
    for (int l = 0; l < 5; ++l) s.alb_hi[r][l] = l < s.ncl ? s.ali[l] : 0.0;
    for (int l = 0; l < 9; ++l) s.arb_hi[r][l] = l < s.ncr ? s.ari[l] : 0.0;
    for (int k = s.nol - (3 - r), i = 1; i <= k; ++i) {
      const int noff = s.ncl - k;  // first out-of-range column
      s.alb_hi[r][noff - i] += syml * s.alb_hi[r][noff + i - 1];
      s.alb_hi[r][noff + i - 1] = 0.0;
    }
    for (int k = s.nor - (3 - r), i = 1; i <= k; ++i) {
      const int noff = s.ncr - k;
      s.arb_hi[r][noff - i] += symr * s.arb_hi[r][noff + i - 1];
      s.arb_hi[r][noff + i - 1] = 0.0;
    }
  }
}

// bc = +1 (even field) / -1 (odd field) at either end; d1 flips the parity between lhs and rhs
// (stencils.f90:256-259), every other set keeps it (e.g. :407-410)
void apply_symmetry(Stencil &s, int bc_lo, int bc_hi) {
  const bool d1 = s.fam == F_D1;
  if (bc_lo) fold_lower(s, d1 ? -bc_lo : bc_lo, bc_lo);
  if (bc_hi) fold_upper(s, d1 ? -bc_hi : bc_hi, bc_hi);
}

// The first-derivative kernels difference the closure rows against the boundary point and have no
// sigma column of their own: it rides in the unused column 8 (compact_d1.f90:124-126,166-168 are
// plain dot products at bc = -1).
void set_sigma_column(Stencil &s, int bc_lo, int bc_hi) {
  auto rowsum = [&](const double *w) { double t = 0.0; for (int l = 0; l < s.ncr; ++l) t += w[l]; return t; };
  for (int r = 0; r < 4; ++r) {
    s.arb_lo[r][8] = bc_lo == -1 ? rowsum(s.arb_lo[r]) : 0.0;
    s.arb_hi[r][8] = bc_hi == -1 ? rowsum(s.arb_hi[r]) : 0.0;
  }
}

}  // namespace

Stencil make_stencil(Kind k, int bc_lo, int bc_hi) {
  Stencil s;
  switch (k) {
    case K_D1: {  // 10th-order compact first derivative, stencils.f90:207-254
      s.nol = 2; s.nor = 3; s.implicit = true; s.null_option = 0; s.post = 1; s.fam = F_D1;
      s.ncl = 5; s.ncr = 7;
      set_row(s.ali, {0.45, 4.5, 9.0, 4.5, 0.45});
      set_row(s.ari, {-0.015, -1.515, -6.375, 0.0, 6.375, 1.515, 0.015});
      set_row(s.alb_lo[0], {0.0, 0.0, 4.725, 9.45, 0.0});
      set_row(s.alb_lo[1], {0.0, 1.94578125, 7.783125, 1.94578125, 0.0});
      set_row(s.alb_lo[2], {0.2964375, 4.743, 10.67175, 4.743, 0.2964375});
      set_row(s.alb_lo[3], {0.451390625, 4.63271875, 9.38146875, 4.63271875, 0.451390625});
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, -11.8125, 9.45, 2.3625, 0.0});
      set_row(s.arb_lo[1], {0.0, 0.0, -5.83734375, 0.0, 5.83734375, 0.0, 0.0});
      set_row(s.arb_lo[2], {0.0, -1.23515625, -7.905, 0.0, 7.905, 1.23515625, 0.0});
      set_row(s.arb_lo[3], {-0.015, -1.53, -6.66984375, 0.0, 6.66984375, 1.53, 0.015});
      mirror_closures(s, -1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      set_sigma_column(s, bc_lo, bc_hi);
      break;
    }
    case K_D2: {  // 10th-order compact second derivative, stencils.f90:358-405
      s.nol = 2; s.nor = 3; s.implicit = true; s.null_option = 0; s.post = 2; s.fam = F_R3;
      s.ncl = 5; s.ncr = 7;
      set_row(s.ali, {387.0, 6012.0, 16182.0, 6012.0, 387.0});
      set_row(s.ari, {79.0, 4671.0, 9585.0, -28670.0, 9585.0, 4671.0, 79.0});
      set_row(s.alb_lo[0], {0.0, 0.0, 1.0, 11.0, 0.0});
      set_row(s.alb_lo[1], {0.0, 1.0, 10.0, 1.0, 0.0});
      set_row(s.alb_lo[2], {23.0, 688.0, 2358.0, 688.0, 23.0});
      set_row(s.alb_lo[3], {387.0, 6012.0, 16182.0, 6012.0, 387.0});
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, 13.0, -27.0, 15.0, -1.0});
      set_row(s.arb_lo[1], {0.0, 0.0, 12.0, -24.0, 12.0, 0.0, 0.0});
      set_row(s.arb_lo[2], {0.0, 465.0, 1920.0, -4770.0, 1920.0, 465.0, 0.0});
      set_row(s.arb_lo[3], {79.0, 4671.0, 9585.0, -28670.0, 9585.0, 4671.0, 79.0});
      mirror_closures(s, 1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      if (bc_lo == -1 || bc_hi == -1)  // no entry point of parcop.f90 reaches d2 with bc = -1 (operators.f90:513-526)
        throw std::invalid_argument("make_stencil: the second derivative has no antisymmetric variant on this path");
      break;
    }
    case K_D8: {  // compact 8th derivative (ringing detector), stencils.f90:515-591
      const double zeta = 29.0, alpha = 14.0, beta = 1.5;
      const double aa = 4200.0, bb = -3360.0, cc = 1680.0, dd = -480.0, ee = 60.0;
      s.nol = 2; s.nor = 4; s.implicit = true; s.null_option = 0; s.post = 0; s.fam = F_R4;
      s.ncl = 5; s.ncr = 9;
      set_row(s.ali, {beta, alpha, zeta, alpha, beta});
      set_row(s.ari, {ee, dd, cc, bb, aa, bb, cc, dd, ee});
      set_row(s.alb_lo[0], {0.0, 0.0, alpha + zeta, beta + alpha, beta});
      set_row(s.alb_lo[1], {0.0, beta + alpha, zeta, alpha, beta});
      set_row(s.alb_lo[2], {beta, alpha, zeta, alpha, beta});
      set_row(s.alb_lo[3], {beta, alpha, zeta, alpha, beta});
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, 0.0, bb + aa, cc + bb, dd + cc, ee + dd, ee});
      set_row(s.arb_lo[1], {0.0, 0.0, 0.0, cc + bb, dd + aa, ee + bb, cc, dd, ee});
      set_row(s.arb_lo[2], {0.0, 0.0, dd + cc, ee + bb, aa, bb, cc, dd, ee});
      set_row(s.arb_lo[3], {0.0, ee + dd, cc, bb, aa, bb, cc, dd, ee});
      mirror_closures(s, 1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      set_reference_slots(s, false, bc_lo == -1, bc_hi == -1);
      break;
    }
    case K_SF: {  // 8th-order compact "9/10" filter, telescoped closures, stencils.f90:713-835
      const double beta = 1.6688e-1, alpha = 6.6624e-1, zeta = 1.0;
      const double aa = 9.9965e-1, bb = 6.6652e-1, cc = 1.6674e-1, dd = 4.0e-5, ee = -5.0e-6;
      s.nol = 2; s.nor = 4; s.implicit = true; s.null_option = 1; s.post = 0; s.fam = F_R4;
      s.ncl = 5; s.ncr = 9;
      set_row(s.ali, {beta, alpha, zeta, alpha, beta});
      set_row(s.ari, {ee, dd, cc, bb, aa, bb, cc, dd, ee});
      set_row(s.alb_lo[0], {0.0, 0.0, 1.0, 0.0, 0.0});
      set_row(s.alb_lo[1], {0.0, 4.997e-1, 1.0, 4.997e-1, 0.0});
      set_row(s.alb_lo[2], {1.6688e-1, 6.6624e-1, 1.0, 6.6624e-1, 1.6688e-1});
      set_row(s.alb_lo[3], {1.6688e-1, 6.6624e-1, 1.0, 6.6624e-1, 1.6688e-1});
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0});
      set_row(s.arb_lo[1], {0.0, 0.0, 0.0, 4.9985e-1, 9.997e-1, 4.9985e-1, 0.0, 0.0, 0.0});
      set_row(s.arb_lo[2], {0.0, 0.0, 1.668e-1, 6.6656e-1, 9.9952e-1, 6.6656e-1, 1.668e-1, 0.0, 0.0});
      set_row(s.arb_lo[3], {0.0, 4.0e-5, 1.6672e-1, 6.6652e-1, 9.9968e-1, 6.6652e-1, 1.6672e-1, 4.0e-5, 0.0});
      mirror_closures(s, 1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      difference_form(s);
      set_reference_slots(s, false, bc_lo == -1, bc_hi == -1);
      s.add_back = true;
      break;
    }
    case K_GF: {  // explicit 9-point Gaussian, stencils.f90:1387-1476
      const double a = 3565.0 / 10368.0, b = 3091.0 / 12960.0, c = 1997.0 / 25920.0,
                   d = 149.0 / 12960.0, e = 107.0 / 103680.0;
      s.nol = 0; s.nor = 4; s.implicit = false; s.null_option = 1; s.post = 0; s.fam = F_R4;
      s.ncl = 1; s.ncr = 9;
      s.ali[0] = 1.0;
      set_row(s.ari, {e, d, c, b, a, b, c, d, e});
      for (int r = 0; r < 4; ++r) s.alb_lo[r][0] = 1.0;
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, 0.0, a + b, b + c, c + d, d + e, e});
      set_row(s.arb_lo[1], {0.0, 0.0, 0.0, b + c, a + d, b + e, c, d, e});
      set_row(s.arb_lo[2], {0.0, 0.0, c + d, b + e, a, b, c, d, e});
      set_row(s.arb_lo[3], {0.0, d + e, c, b, a, b, c, d, e});
      mirror_closures(s, 1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      difference_form(s);
      set_reference_slots(s, false, bc_lo == -1, bc_hi == -1);
      s.add_back = true;
      break;
    }
    case K_D4: {  // explicit 4th-order 4th derivative, stencils.f90:430-513; evaluated by the 7-point
      // difference-form sweep (compact.f90:41) without a metric scale (compact_operators.f90:229).
      // The reference assigns three of its four closure rows per end (:480-493) and places the upper
      // three one row early; as in every other set, row 4 is completed with the interior stencil and
      // the upper rows mirror the lower ones (periodic and SYMM axes never read them).
      const double a = 28.0 / 3.0, b = -6.5, c = 2.0, d = -1.0 / 6.0;
      s.nol = 0; s.nor = 3; s.implicit = false; s.null_option = 0; s.post = 0; s.fam = F_R3;
      s.ncl = 1; s.ncr = 7;
      s.ali[0] = 1.0;
      set_row(s.ari, {d, c, b, a, b, c, d});
      for (int r = 0; r < 4; ++r) s.alb_lo[r][0] = 1.0;
      set_row(s.arb_lo[0], {0.0, 0.0, 0.0, a + b, b + c, c + d, d});
      set_row(s.arb_lo[1], {0.0, 0.0, b + c, a + d, b, 0.0, 0.0});
      set_row(s.arb_lo[2], {0.0, c + d, b, a, b, c, 0.0});
      set_row(s.arb_lo[3], {d, c, b, a, b, c, d});
      mirror_closures(s, 1.0);
      apply_symmetry(s, bc_lo, bc_hi);
      if (bc_lo == -1 || bc_hi == -1)  // parcop.f90:255-277 calls d4x/y/z without a symmetry selector
        throw std::invalid_argument("make_stencil: the fourth derivative has no antisymmetric variant on this path");
      break;
    }
    default:
      throw std::invalid_argument("make_stencil: unknown operator kind");
  }
  return s;
}

std::vector<double> assemble_bands(const Stencil &st, int n, bool periodic) {
  // compact_basetype.f90:123-133: interior weights on every row, closures patched on the ends
  std::vector<double> b((size_t)n * 5, 0.0);
  if (!st.implicit) {
    for (int i = 0; i < n; ++i) b[(size_t)i * 5 + 2] = 1.0;
    return b;
  }
  for (int i = 0; i < n; ++i)
    for (int l = 0; l < 5; ++l) b[(size_t)i * 5 + l] = st.ali[l];
  if (!periodic) {
    for (int r = 0; r < 4 && r < n; ++r)
      for (int l = 0; l < 5; ++l) b[(size_t)r * 5 + l] = st.alb_lo[r][l];
    for (int r = 0; r < 4 && n - 4 + r >= 0; ++r)
      for (int l = 0; l < 5; ++l) b[(size_t)(n - 4 + r) * 5 + l] = st.alb_hi[r][l];
  }
  return b;
}

int choose_chunks(int m, int chunk_len) {
  if (chunk_len < 16) chunk_len = 16;
  int best = 1;
  // largest P (<= 32) dividing m whose chunk length stays >= chunk_len; otherwise the closest
  for (int P = 1; P <= 32; ++P) {
    if (m % P) continue;
    const int C = m / P;
    if (C < 16) break;
    if (C >= chunk_len) best = P;
  }
  return best;
}

namespace {

// LU of a bounded pentadiagonal block without pivoting; the reference's elimination order
// (pentadiagonal.f90:18-39).  c: C x 5 in place; pivots are stored inverted.
void lu_block(std::vector<double> &c, int C) {
  auto at = [&](int i, int l) -> double & { return c[(size_t)i * 5 + l]; };
  for (int i = 0; i + 2 < C; ++i) {
    at(i + 1, 1) /= at(i, 2);
    at(i + 1, 2) -= at(i, 3) * at(i + 1, 1);
    at(i + 1, 3) -= at(i, 4) * at(i + 1, 1);
    at(i + 2, 0) /= at(i, 2);
    at(i + 2, 1) -= at(i, 3) * at(i + 2, 0);
    at(i + 2, 2) -= at(i, 4) * at(i + 2, 0);
  }
  at(C - 1, 1) /= at(C - 2, 2);
  at(C - 1, 2) -= at(C - 2, 3) * at(C - 1, 1);
  for (int i = 0; i < C; ++i) at(i, 2) = 1.0 / at(i, 2);
  // entries that point outside the block are never used by the solve: zero them so the kernels
  // can run one uniform recurrence over every row
  at(0, 0) = at(0, 1) = at(1, 0) = 0.0;
  at(C - 1, 3) = at(C - 1, 4) = at(C - 2, 4) = 0.0;
}

// solve with the factors above (pentadiagonal.f90:42-61), pull form
void solve_block(const std::vector<double> &c, int C, double *r) {
  auto at = [&](int i, int l) { return c[(size_t)i * 5 + l]; };
  for (int i = 0; i < C; ++i) {
    double t = r[i];
    if (i >= 2) t -= at(i, 0) * r[i - 2];
    if (i >= 1) t -= at(i, 1) * r[i - 1];
    r[i] = t;
  }
  for (int i = C - 1; i >= 0; --i) {
    double t = r[i];
    if (i + 1 < C) t -= at(i, 3) * r[i + 1];
    if (i + 2 < C) t -= at(i, 4) * r[i + 2];
    r[i] = t * at(i, 2);
  }
}

// dense inverse by Gauss-Jordan with partial pivoting, long double
std::vector<long double> invert_dense(std::vector<long double> a, int n) {
  std::vector<long double> inv((size_t)n * n, 0.0L);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0L;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (fabsl(a[(size_t)r * n + col]) > fabsl(a[(size_t)piv * n + col])) piv = r;
    if (a[(size_t)piv * n + col] == 0.0L) throw std::runtime_error("reduced interface system is singular");
    if (piv != col)
      for (int c = 0; c < n; ++c) {
        std::swap(a[(size_t)piv * n + c], a[(size_t)col * n + c]);
        std::swap(inv[(size_t)piv * n + c], inv[(size_t)col * n + c]);
      }
    const long double d = 1.0L / a[(size_t)col * n + col];
    for (int c = 0; c < n; ++c) { a[(size_t)col * n + c] *= d; inv[(size_t)col * n + c] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const long double f = a[(size_t)r * n + col];
      if (f == 0.0L) continue;
      for (int c = 0; c < n; ++c) {
        a[(size_t)r * n + c] -= f * a[(size_t)col * n + c];
        inv[(size_t)r * n + c] -= f * inv[(size_t)col * n + c];
      }
    }
  }
  return inv;
}

}  // namespace

Partition build_partition(int m, const std::vector<double> &bands, bool cyclic, int P) {
  if (P < 1 || m % P != 0) throw std::invalid_argument("build_partition: P must divide m");
  const int C = m / P;
  if (C < 8) throw std::invalid_argument("build_partition: chunk shorter than 8 rows");
  Partition part;
  part.m = m; part.P = P; part.C = C; part.cyclic = cyclic;
  part.ctype.assign(P, 0);
  part.G.assign((size_t)P * 4 * 4 * P, 0.0);

  std::vector<std::vector<double>> lu_all(P), rc_all(P);
  auto band = [&](int i, int l) { return bands[(size_t)i * 5 + l]; };
  const int nred = 4 * P;
  std::vector<long double> R((size_t)nred * nred, 0.0L);
  for (int i = 0; i < nred; ++i) R[(size_t)i * nred + i] = 1.0L;
  std::vector<int> prev(P), next(P);

  for (int p = 0; p < P; ++p) {
    const int s = p * C, e = s + C;
    std::vector<double> c((size_t)C * 5);
    for (int i = 0; i < C; ++i)
      for (int l = 0; l < 5; ++l) c[(size_t)i * 5 + l] = band(s + i, l);
    lu_block(c, C);
    prev[p] = (p > 0) ? p - 1 : (cyclic ? P - 1 : -1);
    next[p] = (p + 1 < P) ? p + 1 : (cyclic ? 0 : -1);
    // spike columns: A_p^-1 applied to the dropped couplings (compact_basetype.f90:151-181)
    std::vector<double> rc((size_t)C * 4, 0.0), col(C);
    for (int q = 0; q < 4; ++q) {
      std::fill(col.begin(), col.end(), 0.0);
      bool any = false;
      if (q < 2 && prev[p] >= 0) {
        if (q == 0) { col[0] = band(s, 0); }
        else { col[0] = band(s, 1); col[1] = band(s + 1, 0); }
        any = true;
      } else if (q >= 2 && next[p] >= 0) {
        if (q == 2) { col[C - 2] = band(e - 2, 4); col[C - 1] = band(e - 1, 3); }
        else { col[C - 1] = band(e - 1, 4); }
        any = true;
      }
      if (any) {
        bool nz = false;
        for (double x : col) nz = nz || (x != 0.0);
        if (nz) solve_block(c, C, col.data());
      }
      for (int i = 0; i < C; ++i) rc[(size_t)i * 4 + q] = any ? col[i] : 0.0;
    }
    // reduced system rows of this chunk: y_p + V^ y_prev(3:4) + W^ y_next(1:2) = d_p
    const int rows[4] = {0, 1, C - 2, C - 1};
    for (int r = 0; r < 4; ++r)
      for (int q = 0; q < 4; ++q) {
        const int nb = (q < 2) ? prev[p] : next[p];
        if (nb < 0) continue;
        const int colidx = (q < 2) ? 4 * nb + 2 + q : 4 * nb + (q - 2);
        R[(size_t)(4 * p + r) * nred + colidx] += (long double)rc[(size_t)rows[r] * 4 + q];
      }
    lu_all[p] = std::move(c);
    rc_all[p] = std::move(rc);
  }

  std::vector<long double> Rinv = invert_dense(R, nred);
  for (int p = 0; p < P; ++p)
    for (int q = 0; q < 4; ++q) {
      const int nb = (q < 2) ? prev[p] : next[p];
      if (nb < 0) continue;
      const int row = (q < 2) ? 4 * nb + 2 + q : 4 * nb + (q - 2);
      for (int j = 0; j < nred; ++j)
        part.G[((size_t)p * 4 + q) * nred + j] = (double)Rinv[(size_t)row * nred + j];
    }

  // share identical chunk tables
  std::vector<int> rep;  // representative chunk of each type
  for (int p = 0; p < P; ++p) {
    int t = -1;
    for (size_t k = 0; k < rep.size(); ++k) {
      const int q = rep[k];
      if (!memcmp(lu_all[p].data(), lu_all[q].data(), sizeof(double) * C * 5) &&
          !memcmp(rc_all[p].data(), rc_all[q].data(), sizeof(double) * C * 4)) { t = (int)k; break; }
    }
    if (t < 0) { t = (int)rep.size(); rep.push_back(p); }
    part.ctype[p] = t;
  }
  part.ntypes = (int)rep.size();
  part.lu.resize((size_t)part.ntypes * C * 5);
  part.rc.resize((size_t)part.ntypes * C * 4);
  for (int t = 0; t < part.ntypes; ++t) {
    memcpy(&part.lu[(size_t)t * C * 5], lu_all[rep[t]].data(), sizeof(double) * C * 5);
    memcpy(&part.rc[(size_t)t * C * 4], rc_all[rep[t]].data(), sizeof(double) * C * 4);
  }
  return part;
}

LineTables build_line_tables(int m, const std::vector<double> &bands, bool cyclic, int P, double cut) {
  if (P < 1 || m % P != 0) throw std::invalid_argument("build_line_tables: P must divide m");
  const int C = m / P;
  if (C < 8) throw std::invalid_argument("build_line_tables: chunk shorter than 8 rows");
  LineTables T;
  T.m = m; T.P = P; T.C = C; T.cyclic = cyclic;
  T.ctype.assign(P, 0);
  auto band = [&](int i, int l) { return bands[(size_t)i * 5 + l]; };

  // one LU of the bounded line (corner couplings of a periodic line are handled below)
  std::vector<double> c(bands);
  lu_block(c, m);
  auto L2 = [&](int i) { return c[(size_t)i * 5 + 0]; };
  auto L1 = [&](int i) { return c[(size_t)i * 5 + 1]; };
  auto IP = [&](int i) { return c[(size_t)i * 5 + 2]; };
  auto U1 = [&](int i) { return c[(size_t)i * 5 + 3]; };
  auto U2 = [&](int i) { return c[(size_t)i * 5 + 4]; };

  // Periodic line: A is circulant, A = Lc * Uc with circulant band factors whose entries are the
  // Toeplitz limit of the LU recurrence (the spectral factorisation of the symbol).  Every chunk
  // then has the same constant coefficients and the carried state simply wraps around the ring of
  // chunks (closed geometric series below) -- no corner correction is needed.
  long double lim[5] = {0, 0, 0, 0, 0};
  {
    const long double a0 = band(m / 2, 0), a1 = band(m / 2, 1), a2 = band(m / 2, 2), a3 = band(m / 2, 3), a4 = band(m / 2, 4);
    long double pv1 = a2, pv2 = a2, u1m1 = a3, u1m2 = a3, u2 = a4;  // pivots / u1 of rows r-1, r-2
    long double l2 = 0, l1 = 0, pv = a2, u1 = a3;
    for (int it = 0; it < 20000; ++it) {
      const long double nl2 = a0 / pv2;
      const long double nl1 = (a1 - u1m2 * nl2) / pv1;
      const long double npv = a2 - u2 * nl2 - u1m1 * nl1;
      const long double nu1 = a3 - u2 * nl1;
      const long double d = fabsl(nl2 - l2) + fabsl(nl1 - l1) + fabsl(npv - pv) + fabsl(nu1 - u1);
      l2 = nl2; l1 = nl1; pv = npv; u1 = nu1;
      pv2 = pv1; pv1 = pv; u1m2 = u1m1; u1m1 = u1;
      if (it > 8 && d < 1e-19L * fabsl(pv)) break;
    }
    lim[0] = l2; lim[1] = l1; lim[2] = 1.0L / pv; lim[3] = u1; lim[4] = u2;
  }

  // Converged coefficients: the Toeplitz limit itself.  The double-precision LU recurrence of a bounded
  // line does not settle on one bit pattern: it ends in a limit cycle a few units in the last place
  // around the limit (measured 3e-15 relative for the compact filter from row ~100 on), so rows within
  // 64 ulp of the limit count as converged and take the limit's coefficients -- the same substitution,
  // of the same size, as the circulant factors of a periodic line.
  auto row_is_const = [&](int i) {
    for (int k = 0; k < 5; ++k) {
      const double a = c[(size_t)i * 5 + k], b = (double)lim[k];
      if (std::fabs(a - b) > 64.0 * 2.220446049250313e-16 * std::fabs(b)) return false;
    }
    return true;
  };
  std::vector<char> chunk_const(P, 0);
  int nconst = 0;
  for (int p = 0; p < P; ++p) {
    bool ok = true;
    for (int i = p * C; i < (p + 1) * C && ok; ++i) ok = row_is_const(i);
    chunk_const[p] = ok;
    nconst += ok;
  }
  T.has_const = nconst > 0;
  for (int k = 0; k < 5; ++k) T.cst[k] = (double)lim[k];
  if (cyclic) {
    T.has_const = true;
    for (int p = 0; p < P; ++p) chunk_const[p] = 1;
    for (int k = 0; k < 5; ++k) T.cst[k] = (double)lim[k];
  }

  struct ChunkTab { std::vector<double> luf, lub, phi, psi, chi; };
  auto make_chunk = [&](int p, bool use_const) {
    ChunkTab t;
    const int s = p * C;
    t.luf.resize((size_t)C * 2); t.lub.resize((size_t)C * 4); t.phi.resize((size_t)C * 2); t.psi.resize((size_t)C * 2);
    t.chi.resize((size_t)C * 2);
    std::vector<double> l2(C), l1(C), ip(C), u1(C), u2(C);
    for (int i = 0; i < C; ++i) {
      l2[i] = use_const ? T.cst[0] : L2(s + i); l1[i] = use_const ? T.cst[1] : L1(s + i);
      ip[i] = use_const ? T.cst[2] : IP(s + i); u1[i] = use_const ? T.cst[3] : U1(s + i); u2[i] = use_const ? T.cst[4] : U2(s + i);
      t.luf[i * 2 + 0] = l2[i]; t.luf[i * 2 + 1] = l1[i];
      t.lub[i * 4 + 0] = ip[i]; t.lub[i * 4 + 1] = u1[i]; t.lub[i * 4 + 2] = u2[i]; t.lub[i * 4 + 3] = 0.0;
    }
    // phi: forward recurrence with zero right-hand side and unit incoming state
    for (int col = 0; col < 2; ++col) {
      double rm1 = col == 0 ? 1.0 : 0.0, rm2 = col == 0 ? 0.0 : 1.0;  // r'[s-1], r'[s-2]
      for (int i = 0; i < C; ++i) {
        const double v = -l2[i] * rm2 - l1[i] * rm1;
        t.phi[i * 2 + col] = v;
        rm2 = rm1; rm1 = v;
      }
    }
    // psi: backward recurrence with zero r' and unit incoming (x[e], x[e+1])
    for (int col = 0; col < 2; ++col) {
      double x1 = col == 0 ? 1.0 : 0.0, x2 = col == 0 ? 0.0 : 1.0;  // x[i+1], x[i+2]
      for (int i = C - 1; i >= 0; --i) {
        const double v = (-u1[i] * x1 - u2[i] * x2) * ip[i];
        t.psi[i * 2 + col] = v;
        x2 = x1; x1 = v;
      }
    }
    // chi: the backward recurrence (zero incoming state) applied to phi -- what a forward state that
    // arrives late adds to the chunk's solution
    for (int col = 0; col < 2; ++col) {
      double x1 = 0.0, x2 = 0.0;
      for (int i = C - 1; i >= 0; --i) {
        const double v = (t.phi[i * 2 + col] - u1[i] * x1 - u2[i] * x2) * ip[i];
        t.chi[i * 2 + col] = v;
        x2 = x1; x1 = v;
      }
    }
    return t;
  };

  std::vector<ChunkTab> types;
  std::vector<int> rep;
  if (T.has_const) {
    int p0 = 0;
    while (!chunk_const[p0]) ++p0;
    types.push_back(make_chunk(p0, true));
    rep.push_back(-1);
  }
  for (int p = 0; p < P; ++p) {
    if (chunk_const[p]) { T.ctype[p] = 0; continue; }
    ChunkTab t = make_chunk(p, false);
    int found = -1;
    for (size_t k = T.has_const ? 1 : 0; k < types.size(); ++k)
      if (types[k].luf == t.luf && types[k].lub == t.lub && types[k].phi == t.phi && types[k].psi == t.psi) { found = (int)k; break; }
    if (found < 0) { found = (int)types.size(); types.push_back(std::move(t)); rep.push_back(p); }
    T.ctype[p] = found;
  }
  T.ntypes = (int)types.size();
  for (auto &t : types) {
    T.luf.insert(T.luf.end(), t.luf.begin(), t.luf.end());
    T.lub.insert(T.lub.end(), t.lub.begin(), t.lub.end());
    T.phi.insert(T.phi.end(), t.phi.begin(), t.phi.end());
    T.psi.insert(T.psi.end(), t.psi.begin(), t.psi.end());
    T.chi.insert(T.chi.end(), t.chi.begin(), t.chi.end());
  }

  // chunk transfer matrices and their truncated products
  {
    auto typ = [&](int q) -> const ChunkTab & { return types[T.ctype[q]]; };
    struct M2 { long double a, b, c, d; };
    auto mul = [](const M2 &x, const M2 &y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; };
    std::vector<M2> Tf(P), Tb(P);
    for (int q = 0; q < P; ++q) {
      const ChunkTab &t = typ(q);
      Tf[q] = M2{t.phi[(C - 1) * 2 + 0], t.phi[(C - 1) * 2 + 1], t.phi[(C - 2) * 2 + 0], t.phi[(C - 2) * 2 + 1]};
      Tb[q] = M2{t.psi[0], t.psi[1], t.psi[2], t.psi[3]};
    }
    T.Mf.assign((size_t)P * (P + 1) * 4, 0.0); T.Mb.assign((size_t)P * (P + 1) * 4, 0.0);
    T.nF.assign(P, 0); T.nB.assign(P, 0);
    // Terms are dropped once the transfer product falls below the rounding unit of the leading
    // (identity) term: what they would add is below the last bit of the state they are added to.
    static const long double tiny_env = getenv("PB_TINY") ? (long double)atof(getenv("PB_TINY")) : 0.0L;
    const long double tiny = tiny_env > 0.0L ? tiny_env : (cut > 0.0 ? (long double)cut : 2.2e-16L);
    auto mxabs = [](const M2 &x) { return std::max(std::max(fabsl(x.a), fabsl(x.b)), std::max(fabsl(x.c), fabsl(x.d))); };
    auto put4 = [](double *dst, const M2 &x) { dst[0] = (double)x.a; dst[1] = (double)x.b; dst[2] = (double)x.c; dst[3] = (double)x.d; };
    if (cyclic) {
      // s_in = (I - T^P)^-1 (end[p-1] + T end[p-2] + ... + T^(P-1) end[p-P]), indices modulo P
      for (int dir = 0; dir < 2; ++dir) {
        const M2 Tm = dir == 0 ? Tf[0] : Tb[0];
        M2 TP{1, 0, 0, 1};
        for (int q = 0; q < P; ++q) TP = mul(TP, Tm);
        const M2 ImT{1 - TP.a, -TP.b, -TP.c, 1 - TP.d};
        const long double det = ImT.a * ImT.d - ImT.b * ImT.c;
        const M2 G{ImT.d / det, -ImT.b / det, -ImT.c / det, ImT.a / det};
        M2 acc = G;
        int nterms = 0;
        std::vector<double> row((size_t)(P + 1) * 4, 0.0);
        for (int j = 1; j <= P; ++j) {
          if (j >= 2) acc = mul(acc, Tm);
          if (j >= 2 && mxabs(acc) < tiny) break;
          put4(&row[(size_t)j * 4], acc);
          nterms = j;
        }
        for (int p = 0; p < P; ++p) {
          std::copy(row.begin(), row.end(), (dir == 0 ? T.Mf : T.Mb).begin() + (size_t)p * (P + 1) * 4);
          (dir == 0 ? T.nF : T.nB)[p] = nterms;
        }
      }
    } else {
      for (int p = 0; p < P; ++p) {
        // forward: s_in[p] = end[p-1] + T[p-1] end[p-2] + T[p-1] T[p-2] end[p-3] + ...
        M2 acc{1, 0, 0, 1};
        for (int j = 1; j <= p; ++j) {
          if (j >= 2) acc = mul(acc, Tf[p - j + 1]);
          if (j >= 2 && mxabs(acc) < tiny) break;
          put4(&T.Mf[((size_t)p * (P + 1) + j) * 4], acc);
          T.nF[p] = j;
        }
        // backward: t_in[p] = start[p+1] + U[p+1] start[p+2] + U[p+1] U[p+2] start[p+3] + ...
        acc = M2{1, 0, 0, 1};
        for (int j = 1; p + j < P; ++j) {
          if (j >= 2) acc = mul(acc, Tb[p + j - 1]);
          if (j >= 2 && mxabs(acc) < tiny) break;
          put4(&T.Mb[((size_t)p * (P + 1) + j) * 4], acc);
          T.nB[p] = j;
        }
      }
    }
  }

  return T;
}

}  // namespace pb

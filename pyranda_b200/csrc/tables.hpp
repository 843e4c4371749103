// Host-side operator assembly for libparcop_b200: stencil coefficient sets, per-row bands,
// chunk-local LU factors, spike columns and the dense reduced-system rows the kernels consume.
//
// What it replaces in the reference (pyranda/parcop/): stencils.f90 (coefficient tables),
// compact_basetype.f90:65-209 (row assembly, LU, spikes, reduced system) and
// pentadiagonal.f90:18-61,160-224,351-454 (LU factorisations).  The design differs: one generic
// "partitioned line system" builder serves both levels of the solve -- chunks of a line inside a
// thread block, and z-slabs across GPUs -- and the reduced interface system is inverted once on
// the host (dense, long double) instead of being LU-solved per line on every rank.
#pragma once
#include <string>
#include <vector>

namespace pb {

enum Kind { K_D1 = 0, K_D2 = 1, K_D8 = 2, K_SF = 3, K_GF = 4, K_D4 = 5, K_COUNT = 6 };
enum Fam { F_D1 = 0, F_R3 = 1, F_R4 = 2 };

// One operator's coefficients with its boundary closures: one-sided ("NONE", bc = 0) or the
// interior stencil folded across a symmetry plane ("SYMM": bc = +1 for an even field, -1 for an
// odd one; compact.f90:77-91, stencils.f90:2390-2453).
struct Stencil {
  int nol = 0, nor = 0, ncl = 1, ncr = 1;
  bool implicit = false;
  int null_option = 0;  // 0: zero when the direction is a null-op, 1: copy   (stencils.f90:28)
  int post = 0;         // metric scale after the solve: 0 none, 1 /d, 2 /d^2  (compact_operators.f90)
  bool add_back = false;  // result = input + stencil (the explicit Gaussian in difference form)
  int fam = F_D1;
  double ali[5] = {0, 0, 0, 0, 0};
  double ari[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double alb_lo[4][5] = {}, alb_hi[4][5] = {};
  double arb_lo[4][9] = {}, arb_hi[4][9] = {};
};
Stencil make_stencil(Kind k, int bc_lo = 0, int bc_hi = 0);

// A line system A x = b of m unknowns, pentadiagonal (or identity for explicit operators), split
// into P equal chunks of C rows.  Per chunk: LU of the chunk's diagonal block (couplings to other
// chunks dropped), the four spike columns, and the rows of the inverse reduced system that give
// the four neighbour interface unknowns this chunk needs.  Identical chunks share one table
// ("type"): a periodic line has one type, a bounded line up to three.
struct Partition {
  int m = 0, P = 1, C = 0, ntypes = 0;
  bool cyclic = false;
  std::vector<int> ctype;   // [P]
  std::vector<double> lu;   // [ntypes][C][5]   l2, l1, 1/pivot, u1, u2 for row i (pull form)
  std::vector<double> rc;   // [ntypes][C][4]   spike columns (prev.last2, next.first2)
  std::vector<double> G;    // [P][4][4P]       g_p = G_p * d_all
};

// Tables for the in-kernel solve ("chunked Thomas with carried state"): ONE LU factorisation of
// the whole (bounded) line; a thread runs the forward / backward recurrences over its chunk with a
// zero incoming state, a short serial scan propagates the true 2-value state from chunk to chunk,
// and the homogeneous responses phi / psi add the carried state back.  Rows whose LU coefficients
// have converged to the Toeplitz limit share one "constant" chunk type whose coefficients live in
// registers.  A periodic line is circulant: it is factored into circulant band factors (the Toeplitz
// limit of the LU recurrence), every chunk is the constant type, and the carried state wraps around
// the ring of chunks through a closed geometric series of the 2x2 chunk transfer matrix.
struct LineTables {
  int m = 0, P = 1, C = 0, ntypes = 0;
  bool cyclic = false, has_const = false;
  double cst[5] = {0, 0, 0, 0, 0};  // l2, l1, 1/pivot, u1, u2 of the converged rows
  std::vector<int> ctype;           // [P]; type 0 is the constant type when has_const
  std::vector<double> luf;          // [ntypes][C][2]  l2, l1
  std::vector<double> lub;          // [ntypes][C][4]  1/pivot, u1, u2, 0
  std::vector<double> phi;          // [ntypes][C][2]  forward response to (r'[s-1], r'[s-2])
  std::vector<double> psi;          // [ntypes][C][2]  backward response to (x[e], x[e+1])
  std::vector<double> chi;          // [ntypes][C][2]  solution response (backward pass, zero incoming) to the forward state entering the chunk
  // carried state without a serial scan: the state entering chunk p is a short sum over the
  // local end values of the chunks before (after) it, weighted by products of 2x2 chunk transfer
  // matrices; the products decay like rho^(C*distance) and are truncated below 1e-22.
  std::vector<double> Mf;           // [P][P+1][4]  Mf[p][j], j >= 1, applies to the end values of chunk (p-j) mod P
  std::vector<double> Mb;           // [P][P+1][4]  Mb[p][j] applies to the start values of chunk (p+j) mod P
  std::vector<int> nF, nB;          // [P] number of terms kept (including the identity term)
};
// cut: carried-state terms whose transfer product is below this are dropped (0: the default, the rounding unit)
LineTables build_line_tables(int m, const std::vector<double> &bands, bool cyclic, int P, double cut = 0.0);

// bands: m rows x 5, row i multiplies x[i-2..i+2]; entries that fall outside [0,m) are couplings
// to the other end when `cyclic`, and are ignored otherwise.
Partition build_partition(int m, const std::vector<double> &bands, bool cyclic, int P);

// Global rows of one operator along one axis: n x 5 lhs bands (interior + closures).
std::vector<double> assemble_bands(const Stencil &st, int n, bool periodic);

// choose the number of chunks for a line of m rows (target chunk length ~chunk_len)
int choose_chunks(int m, int chunk_len);

}  // namespace pb

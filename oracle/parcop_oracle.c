/*
 * parcop_oracle.c -- CPU restatement of LLNL/pyranda's `parcop` compact-operator path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load it,
 * and only as the checker / CPU baseline.  The product (pyranda_b200/) never links or calls it.
 *
 * The reference's Fortran cannot be compiled in this image (no Fortran compiler, no MPI), so this
 * file restates it in plain C, function by function; every routine cites the reference file:line
 * (relative to /root/reference/pyranda/parcop/) it follows.  MPI ranks along a sweep axis are
 * emulated in-process: a "rank" is a contiguous segment of each grid line, MPI_Sendrecv becomes a
 * copy of neighbour rows and mpi_allgather a copy of every rank's 4 interface values.
 *
 * Parity pinning (tests/test_oracle_pins.py, tests/test_sim_oracle.py; table in DESIGN.md section 5):
 * the reference's own golden scalars (tests/cases/testUnit.py:3, testTaylorGreen.py:3,
 * test1DAdvection.py:3-5, testHeat1D.py) and golden curve files (tests/baselines/RT_2D.dat,
 * cylinder-2d-32/64.dat, cylinder_curved-2d-64.dat, cylinder_omesh-2d-64.dat, euler-2d-64/128.dat,
 * KelvinHelmholtzKH-2d-64.dat; packed into tests/golden/reference_baselines.npz), reproduced to 1e-12 or to the last
 * printed digit against the reference's own 1e-4 tolerance; the analytic transfer functions of the
 * stencils on periodic grids (1e-13); symmetry planes == periodic operators on the mirrored field;
 * np-rank emulation == 1-rank (1e-13).  A direct comparison against a Fortran/MPI binary is NOT
 * possible in this container.
 *
 * Lines are processed in bundles of PO_NB adjacent lines so the sequential recurrences vectorise
 * across lines (the same idea as the reference's transposed `bpp_lus_opt` CPU branch,
 * compact_d1.f90:229-237); the arithmetic per line is the reference's.
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_NB 8 /* lines per bundle */

enum { PO_D1 = 0, PO_D2 = 1, PO_D8 = 2, PO_SF = 3, PO_GF = 4, PO_D4 = 5, PO_NKIND = 6 };
enum { FAM_D1 = 0, FAM_R3 = 1, FAM_R4 = 2 };
enum { BC_NONE = 0, BC_PERI = 1, BC_SYMM = 2 };

/* ------------------------------------------------------------------------------------------ */
/* compact_weight  (stencils.f90:24-39)                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int nol, nor, ncl, ncr, nci, nbc1, nbc2, null_option, implicit_op;
  double ali[5], ari[9];
  /* [bc+1][row][l], bc in -1..2  (alb1(l,row,bc)) */
  double alb1[4][4][5], alb2[4][4][5];
  double arb1[4][4][9], arb2[4][4][9];
} po_weight;

static void cpy(double *d, const double *s, int n) { memcpy(d, s, sizeof(double) * n); }
static void rev(double *d, const double *s, int n, double sgn) {
  for (int i = 0; i < n; i++) d[i] = sgn * s[n - 1 - i];
}

/* stencils.f90:2390-2421 lower_symm_weights_int (shl = shr = 0) */
static void lower_symm(double alb[4][5], double arb[4][9], const double *ali, const double *ari,
                       int ncl, int ncr, int nol, int nor, int syml, int symr) {
  for (int r = 0; r < 4; r++) { cpy(alb[r], ali, ncl); cpy(arb[r], ari, ncr); }
  for (int j = 1; j <= nol; j++) {
    int k = nol - j + 1;
    for (int i = 1; i <= k; i++) {
      int i1 = k - i + 1, i2 = k + i;
      alb[j - 1][i2 - 1] += (double)syml * alb[j - 1][i1 - 1];
      alb[j - 1][i1 - 1] = 0.0;
    }
  }
  for (int j = 1; j <= nor; j++) {
    int k = nor - j + 1;
    for (int i = 1; i <= k; i++) {
      int i1 = k - i + 1, i2 = k + i;
      arb[j - 1][i2 - 1] += (double)symr * arb[j - 1][i1 - 1];
      arb[j - 1][i1 - 1] = 0.0;
    }
  }
}

/* stencils.f90:2423-2453 upper_symm_weights_int (shl = shr = 0), n = size(alb,2) = 4 */
static void upper_symm(double alb[4][5], double arb[4][9], const double *ali, const double *ari,
                       int ncl, int ncr, int nol, int nor, int syml, int symr) {
  const int n = 4;
  for (int r = 0; r < 4; r++) { cpy(alb[r], ali, ncl); cpy(arb[r], ari, ncr); }
  for (int j = 1; j <= nol; j++) {
    int jj = n + 1 - j, k = nol - j + 1;
    for (int i = 1; i <= k; i++) {
      int noff = ncl - k;
      alb[jj - 1][noff - i + 1 - 1] += (double)syml * alb[jj - 1][noff + i - 1];
      alb[jj - 1][noff + i - 1] = 0.0;
    }
  }
  for (int j = n - nor + 1; j <= n; j++) {
    int k = j - n + nor;
    for (int i = 1; i <= k; i++) {
      int noff = ncr - k;
      arb[j - 1][noff - i + 1 - 1] += (double)symr * arb[j - 1][noff + i - 1];
      arb[j - 1][noff + i - 1] = 0.0;
    }
  }
}

/* alb2(:,r,0) = alb1(ncl:1:-1,5-r,0); arb2(:,r,0) = sgn*arb1(ncr:1:-1,5-r,0) */
static void mirror_bc0(po_weight *w, double sgn) {
  for (int r = 0; r < 4; r++) {
    rev(w->alb2[1][r], w->alb1[1][3 - r], w->ncl, 1.0);
    rev(w->arb2[1][r], w->arb1[1][3 - r], w->ncr, sgn);
  }
}

static void weight_common(po_weight *w, int nol, int nor, int null_option, int implicit_op) {
  memset(w, 0, sizeof(*w));
  w->nol = nol; w->nor = nor; w->ncl = 2 * nol + 1; w->ncr = 2 * nor + 1; w->nci = 2 * nol;
  w->nbc1 = 4; w->nbc2 = 4; w->null_option = null_option;
  w->implicit_op = (nol == 0) ? 0 : implicit_op;
}

#define SET5(d, a, b, c, e, f) do { double t_[5] = {a, b, c, e, f}; cpy(d, t_, 5); } while (0)
#define SET7(d, a, b, c, e, f, g, h) do { double t_[7] = {a, b, c, e, f, g, h}; cpy(d, t_, 7); } while (0)
#define SET9(d, a, b, c, e, f, g, h, i, j) do { double t_[9] = {a, b, c, e, f, g, h, i, j}; cpy(d, t_, 9); } while (0)

/* stencils.f90:207-277 c10d1 (bc=2 "extended" rows are unreachable from pyranda and omitted) */
static void c10d1(po_weight *w) {
  weight_common(w, 2, 3, 0, 1);
  SET5(w->ali, 0.45, 4.5, 9.0, 4.5, 0.45);
  SET7(w->ari, -0.015, -1.515, -6.375, 0.0, 6.375, 1.515, 0.015);
  SET5(w->alb1[1][0], 0.0, 0.0, 4.725, 9.45, 0.0);
  SET5(w->alb1[1][1], 0.0, 1.94578125, 7.783125, 1.94578125, 0.0);
  SET5(w->alb1[1][2], 0.2964375, 4.743, 10.67175, 4.743, 0.2964375);
  SET5(w->alb1[1][3], 0.451390625, 4.63271875, 9.38146875, 4.63271875, 0.451390625);
  SET7(w->arb1[1][0], 0.0, 0.0, 0.0, -11.8125, 9.45, 2.3625, 0.0);
  SET7(w->arb1[1][1], 0.0, 0.0, -5.83734375, 0.0, 5.83734375, 0.0, 0.0);
  SET7(w->arb1[1][2], 0.0, -1.23515625, -7.905, 0.0, 7.905, 1.23515625, 0.0);
  SET7(w->arb1[1][3], -0.015, -1.53, -6.66984375, 0.0, 6.66984375, 1.53, 0.015);
  mirror_bc0(w, -1.0);
  /* :256-259 */
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 2, 3, -1, +1);
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 2, 3, +1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 2, 3, -1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 2, 3, +1, -1);
}

/* stencils.f90:358-428 c10d2 */
static void c10d2(po_weight *w) {
  weight_common(w, 2, 3, 0, 1);
  SET5(w->ali, 387.0, 6012.0, 16182.0, 6012.0, 387.0);
  SET7(w->ari, 79.0, 4671.0, 9585.0, -28670.0, 9585.0, 4671.0, 79.0);
  SET5(w->alb1[1][0], 0.0, 0.0, 1.0, 11.0, 0.0);
  SET5(w->alb1[1][1], 0.0, 1.0, 10.0, 1.0, 0.0);
  SET5(w->alb1[1][2], 23.0, 688.0, 2358.0, 688.0, 23.0);
  SET5(w->alb1[1][3], 387.0, 6012.0, 16182.0, 6012.0, 387.0);
  SET7(w->arb1[1][0], 0.0, 0.0, 0.0, 13.0, -27.0, 15.0, -1.0);
  SET7(w->arb1[1][1], 0.0, 0.0, 12.0, -24.0, 12.0, 0.0, 0.0);
  SET7(w->arb1[1][2], 0.0, 465.0, 1920.0, -4770.0, 1920.0, 465.0, 0.0);
  SET7(w->arb1[1][3], 79.0, 4671.0, 9585.0, -28670.0, 9585.0, 4671.0, 79.0);
  mirror_bc0(w, 1.0);
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 2, 3, +1, +1);
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 2, 3, -1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 2, 3, +1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 2, 3, -1, -1);
}

/* stencils.f90:515-621 c10d8 */
static void c10d8(po_weight *w) {
  const double zeta = 29.0, alpha = 14.0, beta = 1.5;
  const double aa = 4200.0, bb = -3360.0, cc = 1680.0, dd = -480.0, ee = 60.0;
  const double alpha2 = beta + alpha, zeta1 = alpha + zeta, alpha1 = beta + alpha;
  const double dd4 = ee + dd, cc3 = dd + cc, bb3 = ee + bb, bb2 = cc + bb, aa2 = dd + aa,
               b22 = ee + bb, aa1 = bb + aa, bb1 = cc + bb, cc1 = dd + cc, dd1 = ee + dd;
  weight_common(w, 2, 4, 0, 1);
  SET5(w->ali, beta, alpha, zeta, alpha, beta);
  SET9(w->ari, ee, dd, cc, bb, aa, bb, cc, dd, ee);
  SET5(w->alb1[1][0], 0.0, 0.0, zeta1, alpha1, beta);
  SET5(w->alb1[1][1], 0.0, alpha2, zeta, alpha, beta);
  SET5(w->alb1[1][2], beta, alpha, zeta, alpha, beta);
  SET5(w->alb1[1][3], beta, alpha, zeta, alpha, beta);
  SET9(w->arb1[1][0], 0.0, 0.0, 0.0, 0.0, aa1, bb1, cc1, dd1, ee);
  SET9(w->arb1[1][1], 0.0, 0.0, 0.0, bb2, aa2, b22, cc, dd, ee);
  SET9(w->arb1[1][2], 0.0, 0.0, cc3, bb3, aa, bb, cc, dd, ee);
  SET9(w->arb1[1][3], 0.0, dd4, cc, bb, aa, bb, cc, dd, ee);
  mirror_bc0(w, 1.0);
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 2, 4, +1, +1);
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 2, 4, -1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 2, 4, +1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 2, 4, -1, -1);
}

/* stencils.f90:713-835 c8ff8 (telescoping boundary rows; pre-differenced rhs :831-834) */
static void c8ff8(po_weight *w) {
  const double beta = 1.6688e-1, alpha = 6.6624e-1, zeta = 1.0;
  const double aa = 9.9965e-1, bb = 6.6652e-1, cc = 1.6674e-1, dd = 4.0e-5, ee = -5.0e-6;
  weight_common(w, 2, 4, 1, 1);
  SET5(w->ali, beta, alpha, zeta, alpha, beta);
  SET9(w->ari, ee, dd, cc, bb, aa, bb, cc, dd, ee);
  SET5(w->alb1[1][0], 0.0, 0.0, 1.0, 0.0, 0.0);
  SET5(w->alb1[1][1], 0.0, 4.997e-1, 1.0, 4.997e-1, 0.0);
  SET5(w->alb1[1][2], 1.6688e-1, 6.6624e-1, 1.0, 6.6624e-1, 1.6688e-1);
  SET5(w->alb1[1][3], 1.6688e-1, 6.6624e-1, 1.0, 6.6624e-1, 1.6688e-1);
  SET9(w->arb1[1][0], 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0);
  SET9(w->arb1[1][1], 0.0, 0.0, 0.0, 4.9985e-1, 9.997e-1, 4.9985e-1, 0.0, 0.0, 0.0);
  SET9(w->arb1[1][2], 0.0, 0.0, 1.668e-1, 6.6656e-1, 9.9952e-1, 6.6656e-1, 1.668e-1, 0.0, 0.0);
  SET9(w->arb1[1][3], 0.0, 4.0e-5, 1.6672e-1, 6.6652e-1, 9.9968e-1, 6.6652e-1, 1.6672e-1, 4.0e-5, 0.0);
  mirror_bc0(w, 1.0);
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 2, 4, +1, +1);
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 2, 4, -1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 2, 4, +1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 2, 4, -1, -1);
  /* difference :831-834 (ari(3:7) -= ali(1:5), all boundary rows / bcs likewise) */
  for (int l = 0; l < 5; l++) w->ari[l + 2] -= w->ali[l];
  for (int b = 0; b < 4; b++)
    for (int r = 0; r < 4; r++)
      for (int l = 0; l < 5; l++) {
        w->arb1[b][r][l + 2] -= w->alb1[b][r][l];
        w->arb2[b][r][l + 2] -= w->alb2[b][r][l];
      }
}

/* stencils.f90:1387-1476 cgfs4 (explicit Gaussian, nol = 0) */
static void cgfs4(po_weight *w) {
  const double agau = 3565.0 / 10368.0, bgau = 3091.0 / 12960.0, cgau = 1997.0 / 25920.0,
               dgau = 149.0 / 12960.0, egau = 107.0 / 103680.0;
  const double dg14 = dgau + egau, cg13 = cgau + dgau, bg23 = bgau + egau, bg12 = bgau + cgau,
               ag22 = agau + dgau, bg32 = bgau + egau, ag11 = agau + bgau, bg21 = bgau + cgau,
               cg31 = cgau + dgau, dg41 = dgau + egau;
  weight_common(w, 0, 4, 1, 0);
  w->ali[0] = 1.0;
  SET9(w->ari, egau, dgau, cgau, bgau, agau, bgau, cgau, dgau, egau);
  for (int r = 0; r < 4; r++) w->alb1[1][r][0] = 1.0;
  SET9(w->arb1[1][0], 0.0, 0.0, 0.0, 0.0, ag11, bg21, cg31, dg41, egau);
  SET9(w->arb1[1][1], 0.0, 0.0, 0.0, bg12, ag22, bg32, cgau, dgau, egau);
  SET9(w->arb1[1][2], 0.0, 0.0, cg13, bg23, agau, bgau, cgau, dgau, egau);
  SET9(w->arb1[1][3], 0.0, dg14, cgau, bgau, agau, bgau, cgau, dgau, egau);
  mirror_bc0(w, 1.0);
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 0, 4, +1, +1);
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 0, 4, -1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 0, 4, +1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 0, 4, -1, -1);
  /* difference :1472-1475 */
  w->ari[4] -= w->ali[0];
  for (int b = 0; b < 4; b++)
    for (int r = 0; r < 4; r++) {
      w->arb1[b][r][4] -= w->alb1[b][r][0];
      w->arb2[b][r][4] -= w->alb2[b][r][0];
    }
}

/* stencils.f90:430-513 e4d4 (explicit 4th derivative, d4spec = 1).  The reference declares four
 * closure rows per end (nbc1 = nbc2 = 4, :436) but assigns three (:480-493): lower row 4 and the
 * last upper row are never written (unallocated-initialised memory), and the three upper rows it does
 * assign land one row early (alb2/arb2(:,1:3) are rows n-3..n-1, compact_basetype.f90:128-133).
 * NONE ends are therefore not reproducible from the source; this restatement completes them the
 * way every other set of the file is built (c10d1 :243-254): row 4 is the interior stencil and the
 * upper rows mirror the lower ones.  Periodic and SYMM axes never read those rows. */
static void e4d4(po_weight *w) {
  const double agau = 28.0 / 3.0, bgau = -6.5, cgau = 2.0, dgau = -1.0 / 6.0;
  weight_common(w, 0, 3, 0, 0);
  w->ali[0] = 1.0;
  SET7(w->ari, dgau, cgau, bgau, agau, bgau, cgau, dgau);
  for (int r = 0; r < 4; r++) w->alb1[1][r][0] = 1.0;
  SET7(w->arb1[1][0], 0.0, 0.0, 0.0, agau + bgau, bgau + cgau, cgau + dgau, dgau);
  SET7(w->arb1[1][1], 0.0, 0.0, bgau + cgau, agau + dgau, bgau, 0.0, 0.0);
  SET7(w->arb1[1][2], 0.0, cgau + dgau, bgau, agau, bgau, cgau, 0.0);
  cpy(w->arb1[1][3], w->ari, 7);
  mirror_bc0(w, 1.0);
  lower_symm(w->alb1[2], w->arb1[2], w->ali, w->ari, w->ncl, w->ncr, 0, 3, +1, +1); /* :496-499 */
  lower_symm(w->alb1[0], w->arb1[0], w->ali, w->ari, w->ncl, w->ncr, 0, 3, -1, -1);
  upper_symm(w->alb2[2], w->arb2[2], w->ali, w->ari, w->ncl, w->ncr, 0, 3, +1, +1);
  upper_symm(w->alb2[0], w->arb2[0], w->ali, w->ari, w->ncl, w->ncr, 0, 3, -1, -1);
}

static void make_weight(int kind, po_weight *w) {
  switch (kind) {
    case PO_D1: c10d1(w); break;
    case PO_D2: c10d2(w); break;
    case PO_D8: c10d8(w); break;
    case PO_SF: c8ff8(w); break;   /* sfspec = 2, compact.f90:28 */
    case PO_GF: cgfs4(w); break;   /* gfspec = 6 */
    case PO_D4: e4d4(w); break;    /* d4spec = 1 */
    default: memset(w, 0, sizeof(*w));
  }
}
static int kind_family(int kind) {
  return kind == PO_D1 ? FAM_D1 : ((kind == PO_D2 || kind == PO_D4) ? FAM_R3 : FAM_R4); /* compact.f90:39-44 */
}

/* export for tests: ali(5) ari(9) alb1 alb2 (4*4*5) arb1 arb2 (4*4*9), plus ints */
void po_get_weight(int kind, int *ints, double *ali, double *ari, double *alb1, double *alb2,
                   double *arb1, double *arb2) {
  po_weight w;
  make_weight(kind, &w);
  ints[0] = w.nol; ints[1] = w.nor; ints[2] = w.ncl; ints[3] = w.ncr;
  ints[4] = w.null_option; ints[5] = w.implicit_op;
  cpy(ali, w.ali, 5); cpy(ari, w.ari, 9);
  cpy(alb1, &w.alb1[0][0][0], 80); cpy(alb2, &w.alb2[0][0][0], 80);
  cpy(arb1, &w.arb1[0][0][0], 144); cpy(arb2, &w.arb2[0][0][0], 144);
}

/* ------------------------------------------------------------------------------------------ */
/* pentadiagonal.f90                                                                           */
/* ------------------------------------------------------------------------------------------ */
#define C_(i, c) cm[((c)-1) * (size_t)n + ((i)-1)] /* c(i,c), column-major (n,ncol) */

/* pentadiagonal.f90:18-39 */
static void bpentLUD1(double *cm, int n) {
  for (int i = 1; i <= n - 2; i++) {
    C_(i + 1, 2) = C_(i + 1, 2) / C_(i, 3);
    C_(i + 1, 3) = C_(i + 1, 3) - C_(i, 4) * C_(i + 1, 2);
    C_(i + 1, 4) = C_(i + 1, 4) - C_(i, 5) * C_(i + 1, 2);
    C_(i + 2, 1) = C_(i + 2, 1) / C_(i, 3);
    C_(i + 2, 2) = C_(i + 2, 2) - C_(i, 4) * C_(i + 2, 1);
    C_(i + 2, 3) = C_(i + 2, 3) - C_(i, 5) * C_(i + 2, 1);
  }
  C_(n, 2) = C_(n, 2) / C_(n - 1, 3);
  C_(n, 3) = C_(n, 3) - C_(n - 1, 4) * C_(n, 2);
  for (int i = 1; i <= n; i++) C_(i, 3) = 1.0 / C_(i, 3);
}

/* pentadiagonal.f90:42-61 */
static void bpentLUS1(const double *cm, double *r, int n) {
#define R_(i) r[(i)-1]
  for (int i = 1; i <= n - 2; i++) {
    R_(i + 1) = R_(i + 1) - R_(i) * C_(i + 1, 2);
    R_(i + 2) = R_(i + 2) - R_(i) * C_(i + 2, 1);
  }
  R_(n) = R_(n) - R_(n - 1) * C_(n, 2);
  R_(n) = R_(n) * C_(n, 3);
  R_(n - 1) = (R_(n - 1) - C_(n - 1, 4) * R_(n)) * C_(n - 1, 3);
  for (int i = n - 2; i >= 1; i--) R_(i) = (R_(i) - C_(i, 4) * R_(i + 1) - C_(i, 5) * R_(i + 2)) * C_(i, 3);
#undef R_
}

/* pentadiagonal.f90:64-131 */
static void ppentLUD1(double *cm, int n) {
  const int N = n;
  const double one = 1.0;
  int K;
  C_(1, 8) = C_(1, 1);
  C_(1, 9) = C_(1, 2);
  C_(1, 3) = one / C_(1, 3);
  C_(1, 6) = C_(N - 1, 5) * C_(1, 3);
  C_(1, 7) = C_(N, 4) * C_(1, 3);

  C_(2, 2) = C_(2, 2) * C_(1, 3);
  C_(2, 3) = C_(2, 3) - C_(2, 2) * C_(1, 4);
  C_(2, 4) = C_(2, 4) - C_(2, 2) * C_(1, 5);
  C_(2, 8) = -C_(2, 2) * C_(1, 8);
  C_(2, 9) = C_(2, 1) - C_(2, 2) * C_(1, 9);
  C_(2, 3) = one / C_(2, 3);
  C_(2, 6) = -C_(1, 6) * C_(1, 4) * C_(2, 3);
  C_(2, 7) = (C_(N, 5) - C_(1, 7) * C_(1, 4)) * C_(2, 3);

  for (K = 3; K <= N - 4; K++) {
    C_(K, 1) = C_(K, 1) * C_(K - 2, 3);
    C_(K, 2) = (C_(K, 2) - C_(K, 1) * C_(K - 2, 4)) * C_(K - 1, 3);
    C_(K, 3) = C_(K, 3) - (C_(K, 2) * C_(K - 1, 4) + C_(K, 1) * C_(K - 2, 5));
    C_(K, 4) = C_(K, 4) - C_(K, 2) * C_(K - 1, 5);
    C_(K, 8) = -(C_(K, 2) * C_(K - 1, 8) + C_(K, 1) * C_(K - 2, 8));
    C_(K, 9) = -(C_(K, 2) * C_(K - 1, 9) + C_(K, 1) * C_(K - 2, 9));
    C_(K, 3) = one / C_(K, 3);
    C_(K, 6) = -(C_(K - 1, 6) * C_(K - 1, 4) + C_(K - 2, 6) * C_(K - 2, 5)) * C_(K, 3);
    C_(K, 7) = -(C_(K - 1, 7) * C_(K - 1, 4) + C_(K - 2, 7) * C_(K - 2, 5)) * C_(K, 3);
  }

  C_(N - 3, 1) = C_(N - 3, 1) * C_(N - 5, 3);
  C_(N - 3, 2) = (C_(N - 3, 2) - C_(N - 3, 1) * C_(N - 5, 4)) * C_(N - 4, 3);
  C_(N - 3, 3) = C_(N - 3, 3) - (C_(N - 3, 2) * C_(N - 4, 4) + C_(N - 3, 1) * C_(N - 5, 5));
  C_(N - 3, 4) = C_(N - 3, 4) - C_(N - 3, 2) * C_(N - 4, 5);
  C_(N - 3, 8) = C_(N - 3, 5) - (C_(N - 3, 2) * C_(N - 4, 8) + C_(N - 3, 1) * C_(N - 5, 8));
  C_(N - 3, 9) = -(C_(N - 3, 2) * C_(N - 4, 9) + C_(N - 3, 1) * C_(N - 5, 9));
  C_(N - 3, 3) = one / C_(N - 3, 3);
  C_(N - 3, 6) = (C_(N - 1, 1) - (C_(N - 4, 6) * C_(N - 4, 4) + C_(N - 5, 6) * C_(N - 5, 5))) * C_(N - 3, 3);
  C_(N - 3, 7) = -(C_(N - 4, 7) * C_(N - 4, 4) + C_(N - 5, 7) * C_(N - 5, 5)) * C_(N - 3, 3);

  C_(N - 2, 1) = C_(N - 2, 1) * C_(N - 4, 3);
  C_(N - 2, 2) = (C_(N - 2, 2) - C_(N - 2, 1) * C_(N - 4, 4)) * C_(N - 3, 3);
  C_(N - 2, 3) = C_(N - 2, 3) - (C_(N - 2, 2) * C_(N - 3, 4) + C_(N - 2, 1) * C_(N - 4, 5));
  C_(N - 2, 8) = C_(N - 2, 4) - (C_(N - 2, 2) * C_(N - 3, 8) + C_(N - 2, 1) * C_(N - 4, 8));
  C_(N - 2, 9) = C_(N - 2, 5) - (C_(N - 2, 2) * C_(N - 3, 9) + C_(N - 2, 1) * C_(N - 4, 9));
  C_(N - 2, 3) = one / C_(N - 2, 3);
  C_(N - 2, 6) = (C_(N - 1, 2) - C_(N - 3, 6) * C_(N - 3, 4) - C_(N - 4, 6) * C_(N - 4, 5)) * C_(N - 2, 3);
  C_(N - 2, 7) = (C_(N, 1) - C_(N - 3, 7) * C_(N - 3, 4) - C_(N - 4, 7) * C_(N - 4, 5)) * C_(N - 2, 3);

  for (K = 1; K <= N - 2; K++) {
    C_(N - 1, 3) = C_(N - 1, 3) - C_(K, 6) * C_(K, 8);
    C_(N - 1, 4) = C_(N - 1, 4) - C_(K, 6) * C_(K, 9);
  }
  C_(N - 1, 9) = C_(N - 1, 4);
  for (K = 1; K <= N - 2; K++) C_(N, 2) = C_(N, 2) - C_(K, 7) * C_(K, 8);
  C_(N - 1, 3) = one / C_(N - 1, 3);
  C_(N - 1, 7) = C_(N, 2) * C_(N - 1, 3);
  for (K = 1; K <= N - 1; K++) C_(N, 3) = C_(N, 3) - C_(K, 7) * C_(K, 9);
  C_(N, 3) = one / C_(N, 3);
}

/* Bundle solves: r is [n][PO_NB].  pentadiagonal.f90:629-651 (bpentLUS3x/y/z) */
#define RB(i) (r + ((size_t)(i)-1) * PO_NB)
static void bpentLUS_bundle(const double *cm, double *r, int n) {
  for (int i = 1; i <= n - 2; i++) {
    const double c2 = C_(i + 1, 2), c1 = C_(i + 2, 1);
    double *r0 = RB(i), *r1 = RB(i + 1), *r2 = RB(i + 2);
    for (int b = 0; b < PO_NB; b++) {
      r1[b] = r1[b] - r0[b] * c2;
      r2[b] = r2[b] - r0[b] * c1;
    }
  }
  {
    double *rn = RB(n), *rn1 = RB(n - 1);
    for (int b = 0; b < PO_NB; b++) {
      rn[b] = (rn[b] - rn1[b] * C_(n, 2)) * C_(n, 3);
      rn1[b] = (rn1[b] - C_(n - 1, 4) * rn[b]) * C_(n - 1, 3);
    }
  }
  for (int i = n - 2; i >= 1; i--) {
    const double c3 = C_(i, 3), c4 = C_(i, 4), c5 = C_(i, 5);
    double *r0 = RB(i), *r1 = RB(i + 1), *r2 = RB(i + 2);
    for (int b = 0; b < PO_NB; b++) r0[b] = (r0[b] - c4 * r1[b] - c5 * r2[b]) * c3;
  }
}

/* pentadiagonal.f90:654-684 (ppentLUS3x/y/z) */
static void ppentLUS_bundle(const double *cm, double *r, int n) {
  const int N = n;
  double tmp1[PO_NB], tmp2[PO_NB];
  {
    double *r1 = RB(1), *r2 = RB(2);
    for (int b = 0; b < PO_NB; b++) {
      r2[b] = r2[b] - C_(2, 2) * r1[b];
      tmp1[b] = C_(1, 6) * r1[b] + C_(2, 6) * r2[b];
      tmp2[b] = C_(1, 7) * r1[b] + C_(2, 7) * r2[b];
    }
  }
  for (int i = 3; i <= N - 2; i++) {
    const double c1 = C_(i, 1), c2 = C_(i, 2), c6 = C_(i, 6), c7 = C_(i, 7);
    double *r0 = RB(i), *rm1 = RB(i - 1), *rm2 = RB(i - 2);
    for (int b = 0; b < PO_NB; b++) {
      r0[b] = r0[b] - (c2 * rm1[b] + c1 * rm2[b]);
      tmp1[b] = tmp1[b] + c6 * r0[b];
      tmp2[b] = tmp2[b] + c7 * r0[b];
    }
  }
  {
    double *rN = RB(N), *rN1 = RB(N - 1), *rN2 = RB(N - 2), *rN3 = RB(N - 3);
    for (int b = 0; b < PO_NB; b++) {
      rN1[b] = rN1[b] - tmp1[b];
      rN[b] = (rN[b] - tmp2[b] - C_(N - 1, 7) * rN1[b]) * C_(N, 3);
      rN1[b] = (rN1[b] - C_(N - 1, 9) * rN[b]) * C_(N - 1, 3);
      rN2[b] = (rN2[b] - C_(N - 2, 8) * rN1[b] - C_(N - 2, 9) * rN[b]) * C_(N - 2, 3);
      rN3[b] = (rN3[b] - (C_(N - 3, 4) * rN2[b] + C_(N - 3, 8) * rN1[b] + C_(N - 3, 9) * rN[b])) * C_(N - 3, 3);
    }
  }
  {
    const double *rN = RB(N), *rN1 = RB(N - 1);
    for (int i = N - 4; i >= 1; i--) {
      const double c3 = C_(i, 3), c4 = C_(i, 4), c5 = C_(i, 5), c8 = C_(i, 8), c9 = C_(i, 9);
      double *r0 = RB(i), *r1 = RB(i + 1), *r2 = RB(i + 2);
      for (int b = 0; b < PO_NB; b++)
        r0[b] = (r0[b] - (c4 * r1[b] + c5 * r2[b] + c8 * rN1[b] + c9 * rN[b])) * c3;
    }
  }
}
#undef RB
#undef C_

/* ------------------------------------------------------------------------------------------ */
/* blockmath.f90:553-616 (2x2 / 4x4 inverse) and the block-tridiagonal LU, column-major 4x4    */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double a[4][4]; } m4; /* a[r][c] */
typedef struct { double a[2][2]; } m2;

static m2 m2mul(m2 x, m2 y) {
  m2 z;
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 2; c++) z.a[r][c] = x.a[r][0] * y.a[0][c] + x.a[r][1] * y.a[1][c];
  return z;
}
static m2 m2add(m2 x, m2 y) { m2 z; for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) z.a[r][c] = x.a[r][c] + y.a[r][c]; return z; }
static m2 m2neg(m2 x) { m2 z; for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) z.a[r][c] = -x.a[r][c]; return z; }
/* blockmath.f90:553-562 */
static m2 m2inv(m2 a) {
  m2 b;
  b.a[1][1] = 1. / (a.a[0][0] * a.a[1][1] - a.a[1][0] * a.a[0][1]);
  b.a[0][0] = a.a[1][1] * b.a[1][1];
  b.a[1][0] = -a.a[1][0] * b.a[1][1];
  b.a[0][1] = -a.a[0][1] * b.a[1][1];
  b.a[1][1] = a.a[0][0] * b.a[1][1];
  return b;
}
/* blockmath.f90:596-616 */
static m4 m4inv(m4 aa) {
  m2 a, b, c, d;
  for (int r = 0; r < 2; r++)
    for (int q = 0; q < 2; q++) {
      a.a[r][q] = aa.a[r][q]; b.a[r][q] = aa.a[r][q + 2];
      c.a[r][q] = aa.a[r + 2][q]; d.a[r][q] = aa.a[r + 2][q + 2];
    }
  d = m2inv(d);
  c = m2neg(m2mul(d, c));
  a = m2inv(m2add(a, m2mul(b, c)));
  b = m2mul(m2neg(m2mul(a, b)), d);
  d = m2add(d, m2mul(c, b));
  c = m2mul(c, a);
  m4 bb;
  for (int r = 0; r < 2; r++)
    for (int q = 0; q < 2; q++) {
      bb.a[r][q] = a.a[r][q]; bb.a[r][q + 2] = b.a[r][q];
      bb.a[r + 2][q] = c.a[r][q]; bb.a[r + 2][q + 2] = d.a[r][q];
    }
  return bb;
}
static m4 m4mul(m4 x, m4 y) {
  m4 z;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += x.a[r][k] * y.a[k][c];
      z.a[r][c] = s;
    }
  return z;
}
static m4 m4sub(m4 x, m4 y) { m4 z; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) z.a[r][c] = x.a[r][c] - y.a[r][c]; return z; }
static m4 m4add(m4 x, m4 y) { m4 z; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) z.a[r][c] = x.a[r][c] + y.a[r][c]; return z; }
static m4 m4neg(m4 x) { m4 z; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) z.a[r][c] = -x.a[r][c]; return z; }
static m4 m4zero(void) { m4 z; memset(&z, 0, sizeof(z)); return z; }

/* aa(r,c,s,j): Fortran (4,4,naa,0:np-1) */
#define AA(r, c, s, j) aa[((((size_t)(j)) * naa + ((s)-1)) * 4 + ((c)-1)) * 4 + ((r)-1)]
static m4 aa_get(const double *aa, int naa, int s, int j) {
  m4 z;
  for (int r = 1; r <= 4; r++) for (int c = 1; c <= 4; c++) z.a[r - 1][c - 1] = AA(r, c, s, j);
  return z;
}
static void aa_put(double *aa, int naa, int s, int j, m4 z) {
  for (int r = 1; r <= 4; r++) for (int c = 1; c <= 4; c++) AA(r, c, s, j) = z.a[r - 1][c - 1];
}

/* pentadiagonal.f90:160-180 (j is 0-based here; Fortran 1..n) */
static void btrid_block4_lud(double *aa, int n) {
  const int naa = 3;
  m4 *a = malloc(sizeof(m4) * n), *b = malloc(sizeof(m4) * n), *c = malloc(sizeof(m4) * n);
  for (int j = 0; j < n; j++) { a[j] = aa_get(aa, naa, 1, j); b[j] = aa_get(aa, naa, 2, j); c[j] = aa_get(aa, naa, 3, j); }
  b[0] = m4inv(b[0]);
  aa_put(aa, naa, 2, 0, b[0]);
  for (int j = 1; j < n; j++) {
    a[j] = m4mul(a[j], b[j - 1]);
    b[j] = m4inv(m4sub(b[j], m4mul(a[j], c[j - 1])));
    aa_put(aa, naa, 1, j, a[j]);
    aa_put(aa, naa, 2, j, b[j]);
  }
  free(a); free(b); free(c);
}

/* pentadiagonal.f90:351-388 */
static void ptrid_block4_lud(double *aa, int n) {
  const int naa = 4;
  m4 *a = malloc(sizeof(m4) * n), *b = malloc(sizeof(m4) * n), *c = malloc(sizeof(m4) * n),
     *ax = malloc(sizeof(m4) * n), *cx = malloc(sizeof(m4) * n);
  for (int j = 0; j < n; j++) {
    a[j] = aa_get(aa, naa, 1, j); b[j] = aa_get(aa, naa, 2, j); c[j] = aa_get(aa, naa, 3, j);
    ax[j] = m4zero(); cx[j] = m4zero();
  }
  const int L = n - 1; /* Fortran index n */
  b[0] = m4inv(b[0]);
  ax[0] = a[0];
  cx[0] = m4mul(c[L], b[0]);
  for (int j = 1; j <= n - 2; j++) {
    b[L] = m4sub(b[L], m4mul(cx[j - 1], ax[j - 1]));
    c[L] = m4neg(m4mul(cx[j - 1], c[j - 1]));
    a[j] = m4mul(a[j], b[j - 1]);
    ax[j] = m4neg(m4mul(a[j], ax[j - 1]));
    b[j] = m4inv(m4sub(b[j], m4mul(a[j], c[j - 1])));
    cx[j] = m4mul(c[L], b[j]);
  }
  {
    int j = L;
    a[j] = m4add(m4mul(a[j], b[j - 1]), cx[j - 1]);
    b[j] = m4inv(m4sub(b[j], m4mul(a[j], m4add(c[j - 1], ax[j - 1]))));
  }
  for (int j = 0; j < n; j++) {
    aa_put(aa, naa, 1, j, a[j]);
    aa_put(aa, naa, 2, j, b[j]);
    for (int r = 1; r <= 4; r++) {
      AA(r, 1, 4, j) = cx[j].a[r - 1][0]; AA(r, 2, 4, j) = cx[j].a[r - 1][1];
      AA(r, 3, 4, j) = ax[j].a[r - 1][2]; AA(r, 4, 4, j) = ax[j].a[r - 1][3];
    }
  }
  free(a); free(b); free(c); free(ax); free(cx);
}

/* Reduced-system solves on a bundle: r(l,k) -> r[(k*4 + l)*PO_NB + b], l=0..3, k=0..np-1.
 * pentadiagonal.f90:184-224 (bounded), :393-454 (periodic).  Fortran k=1..n -> C k-1. */
#define RR(l, k) (r + (((size_t)(k)-1) * 4 + ((l)-1)) * PO_NB)
static void btrid_block4_lus(const double *aa, double *r, int n) {
  const int naa = 3;
  for (int k = 2; k <= n; k++)
    for (int b = 0; b < PO_NB; b++) {
      double p1 = RR(1, k - 1)[b], p2 = RR(2, k - 1)[b], p3 = RR(3, k - 1)[b], p4 = RR(4, k - 1)[b];
      for (int l = 1; l <= 4; l++)
        RR(l, k)[b] = RR(l, k)[b] - AA(l, 1, 1, k - 1) * p1 - AA(l, 2, 1, k - 1) * p2 -
                      AA(l, 3, 1, k - 1) * p3 - AA(l, 4, 1, k - 1) * p4;
    }
  for (int b = 0; b < PO_NB; b++) {
    double t[4];
    for (int l = 1; l <= 4; l++) t[l - 1] = RR(l, n)[b];
    for (int l = 1; l <= 4; l++)
      RR(l, n)[b] = AA(l, 1, 2, n - 1) * t[0] + AA(l, 2, 2, n - 1) * t[1] + AA(l, 3, 2, n - 1) * t[2] + AA(l, 4, 2, n - 1) * t[3];
  }
  for (int k = n - 1; k >= 1; k--)
    for (int b = 0; b < PO_NB; b++) {
      double t[4], q1 = RR(1, k + 1)[b], q2 = RR(2, k + 1)[b];
      for (int l = 1; l <= 4; l++) t[l - 1] = RR(l, k)[b] - AA(l, 1, 3, k - 1) * q1 - AA(l, 2, 3, k - 1) * q2;
      for (int l = 1; l <= 4; l++)
        RR(l, k)[b] = AA(l, 1, 2, k - 1) * t[0] + AA(l, 2, 2, k - 1) * t[1] + AA(l, 3, 2, k - 1) * t[2] + AA(l, 4, 2, k - 1) * t[3];
    }
}

static void ptrid_block4_lus(const double *aa, double *r, int n) {
  const int naa = 4;
  if (n > 2)
    for (int b = 0; b < PO_NB; b++) {
      double p1 = RR(1, 1)[b], p2 = RR(2, 1)[b];
      for (int l = 1; l <= 4; l++) RR(l, n)[b] = RR(l, n)[b] - AA(l, 1, 4, 0) * p1 - AA(l, 2, 4, 0) * p2;
    }
  for (int k = 2; k <= n - 2; k++)
    for (int b = 0; b < PO_NB; b++) {
      double p1 = RR(1, k - 1)[b], p2 = RR(2, k - 1)[b], p3 = RR(3, k - 1)[b], p4 = RR(4, k - 1)[b];
      for (int l = 1; l <= 4; l++)
        RR(l, k)[b] = RR(l, k)[b] - AA(l, 1, 1, k - 1) * p1 - AA(l, 2, 1, k - 1) * p2 -
                      AA(l, 3, 1, k - 1) * p3 - AA(l, 4, 1, k - 1) * p4;
      double q1 = RR(1, k)[b], q2 = RR(2, k)[b];
      for (int l = 1; l <= 4; l++) RR(l, n)[b] = RR(l, n)[b] - AA(l, 1, 4, k - 1) * q1 - AA(l, 2, 4, k - 1) * q2;
    }
  if (n > 2)
    for (int b = 0; b < PO_NB; b++) {
      double p1 = RR(1, n - 2)[b], p2 = RR(2, n - 2)[b], p3 = RR(3, n - 2)[b], p4 = RR(4, n - 2)[b];
      for (int l = 1; l <= 4; l++)
        RR(l, n - 1)[b] = RR(l, n - 1)[b] - AA(l, 1, 1, n - 2) * p1 - AA(l, 2, 1, n - 2) * p2 -
                          AA(l, 3, 1, n - 2) * p3 - AA(l, 4, 1, n - 2) * p4;
    }
  for (int b = 0; b < PO_NB; b++) {
    double t[4], p1 = RR(1, n - 1)[b], p2 = RR(2, n - 1)[b], p3 = RR(3, n - 1)[b], p4 = RR(4, n - 1)[b];
    for (int l = 1; l <= 4; l++)
      t[l - 1] = RR(l, n)[b] - AA(l, 1, 1, n - 1) * p1 - AA(l, 2, 1, n - 1) * p2 - AA(l, 3, 1, n - 1) * p3 - AA(l, 4, 1, n - 1) * p4;
    for (int l = 1; l <= 4; l++)
      RR(l, n)[b] = AA(l, 1, 2, n - 1) * t[0] + AA(l, 2, 2, n - 1) * t[1] + AA(l, 3, 2, n - 1) * t[2] + AA(l, 4, 2, n - 1) * t[3];
  }
  for (int k = n - 1; k >= 1; k--)
    for (int b = 0; b < PO_NB; b++) {
      double t[4], q1 = RR(1, k + 1)[b], q2 = RR(2, k + 1)[b], s3 = RR(3, n)[b], s4 = RR(4, n)[b];
      for (int l = 1; l <= 4; l++)
        t[l - 1] = RR(l, k)[b] - AA(l, 1, 3, k - 1) * q1 - AA(l, 2, 3, k - 1) * q2 -
                   AA(l, 3, 4, k - 1) * s3 - AA(l, 4, 4, k - 1) * s4;
      for (int l = 1; l <= 4; l++)
        RR(l, k)[b] = AA(l, 1, 2, k - 1) * t[0] + AA(l, 2, 2, k - 1) * t[1] + AA(l, 3, 2, k - 1) * t[2] + AA(l, 4, 2, k - 1) * t[3];
    }
}
#undef RR

/* ------------------------------------------------------------------------------------------ */
/* compact_op1 + setup_compact_op1 (compact_basetype.f90:17-42, 65-209)                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int m, n, np, id, lo, hi; /* lo/hi: neighbour rank id or -1 == MPI_PROC_NULL */
  int nol, nor, ncl, ncr, nci, nal, naa;
  int null_op, null_option, implicit_op, periodic;
  int bc[2], range1;
  double *ar; /* ar(l,i)  -> ar[(i-1)*ncr + (l-1)] */
  double *al; /* al(i,c)  -> al[(c-1)*m + (i-1)]   */
  double *rc; /* rc(i,c)  -> rc[(c-1)*m + (i-1)]   */
  double *aa; /* AA() */
} po_op;

typedef struct {
  int kind, fam, n, np, m, periodic, bc[2], null_op;
  po_weight w;
  po_op *ops;
} po_opset;

static void opset_free(po_opset *s) {
  if (!s) return;
  if (s->ops)
    for (int r = 0; r < s->np; r++) { free(s->ops[r].ar); free(s->ops[r].al); free(s->ops[r].rc); free(s->ops[r].aa); }
  free(s->ops);
  free(s);
}

/* bc[] are the reference's integer codes (-1,0,1,2); bci = bc+1 indexes the weight tables */
static po_opset *opset_create(int kind, int n, int np, int periodic, int bc1, int bcn, int null_req) {
  po_opset *s = calloc(1, sizeof(*s));
  s->kind = kind; s->fam = kind_family(kind); s->n = n; s->np = np; s->m = n / np;
  s->periodic = periodic; s->bc[0] = bc1; s->bc[1] = bcn;
  make_weight(kind, &s->w);
  const po_weight *w = &s->w;
  /* compact_basetype.f90:101-106 */
  s->null_op = (n < 4) ? 1 : null_req;
  s->ops = calloc(np, sizeof(po_op));
  const int nol = w->nol, nl = w->ncl, ni = w->nci, nr = w->ncr, m = s->m;
  const int nal = periodic ? nl + 4 : nl, naa = periodic ? 4 : 3;
  for (int id = 0; id < np; id++) {
    po_op *op = &s->ops[id];
    op->m = m; op->n = n; op->np = np; op->id = id;
    op->nol = nol; op->nor = w->nor; op->ncl = nl; op->ncr = nr; op->nci = ni; op->nal = nal; op->naa = naa;
    op->null_op = s->null_op; op->null_option = w->null_option; op->implicit_op = w->implicit_op;
    op->periodic = periodic; op->bc[0] = bc1; op->bc[1] = bcn; op->range1 = id * m + 1;
    /* MPI_CART_SHIFT (comm.f90:174-186) */
    op->lo = id - 1; op->hi = id + 1;
    if (periodic) { op->lo = (id - 1 + np) % np; op->hi = (id + 1) % np; }
    else { if (id == 0) op->lo = -1; if (id == np - 1) op->hi = -1; }
  }
  if (s->null_op) return s;
  /* global rows (:118-133) */
  double *gal = calloc((size_t)nal * n, sizeof(double)); /* gal(c,i) -> gal[(i-1)*nal + c-1] */
  double *gar = calloc((size_t)nr * n, sizeof(double));
  for (int i = 0; i < n; i++) { cpy(gal + (size_t)i * nal, w->ali, nl); cpy(gar + (size_t)i * nr, w->ari, nr); }
  if (!periodic) {
    for (int r = 0; r < w->nbc1; r++) cpy(gal + (size_t)r * nal, w->alb1[bc1 + 1][r], nl);
    for (int r = 0; r < w->nbc2; r++) cpy(gal + (size_t)(n - w->nbc2 + r) * nal, w->alb2[bcn + 1][r], nl);
    for (int r = 0; r < w->nbc1; r++) cpy(gar + (size_t)r * nr, w->arb1[bc1 + 1][r], nr);
    for (int r = 0; r < w->nbc2; r++) cpy(gar + (size_t)(n - w->nbc2 + r) * nr, w->arb2[bcn + 1][r], nr);
  }
  double *ro = calloc((size_t)16 * np, sizeof(double)); /* ro(r,c,j) -> ro[(j*4 + c-1)*4 + r-1] */
  for (int id = 0; id < np; id++) {
    po_op *op = &s->ops[id];
    op->ar = malloc(sizeof(double) * nr * m);
    op->al = calloc((size_t)nal * m, sizeof(double));
    /* :145-146 */
    for (int i = 1; i <= m; i++) {
      int gi = op->range1 + i - 1;
      for (int c = 1; c <= nal; c++) op->al[(size_t)(c - 1) * m + (i - 1)] = gal[(size_t)(gi - 1) * nal + (c - 1)];
      cpy(op->ar + (size_t)(i - 1) * nr, gar + (size_t)(gi - 1) * nr, nr);
    }
    if (!op->implicit_op) continue;
#define AL(i, c) op->al[((c)-1) * (size_t)m + ((i)-1)]
#define RC(i, c) op->rc[((c)-1) * (size_t)m + ((i)-1)]
    double al1[2][2] = {{0, 0}, {0, 0}}, al2[2][2] = {{0, 0}, {0, 0}};
    if (op->lo != -1) /* :153-158 */
      for (int i = 1; i <= nol; i++)
        for (int q = i; q <= nol; q++) al1[i - 1][q - 1] = AL(i, q - i + 1);
    if (op->hi != -1) /* :159-164 */
      for (int i = 1; i <= nol; i++)
        for (int q = 1; q <= nol - i + 1; q++) al2[nol - i][q - 1] = AL(m - i + 1, nl - nol + i + q - 1);
    if (np == 1) { /* :165-170 */
      if (periodic) ppentLUD1(op->al, m); else bpentLUD1(op->al, m);
    } else { /* :171-183 */
      op->rc = calloc((size_t)ni * m, sizeof(double));
      bpentLUD1(op->al, m);
      for (int i = 1; i <= nol; i++) for (int q = 1; q <= nol; q++) RC(i, q) = al1[i - 1][q - 1];
      for (int i = 1; i <= nol; i++) for (int q = 1; q <= nol; q++) RC(m - nol + i, nol + q) = al2[i - 1][q - 1];
      double rop[4][4]; /* rop(r,c) */
      for (int c = 1; c <= ni; c++) {
        int any = 0;
        for (int i = 1; i <= m; i++) if (RC(i, c) != 0.0) { any = 1; break; }
        if (any) bpentLUS1(op->al, &RC(1, c), m);
        for (int r = 1; r <= nol; r++) rop[r - 1][c - 1] = RC(r, c);
        for (int r = 1; r <= nol; r++) rop[nol + r - 1][c - 1] = RC(m - nol + r, c);
      }
      if (op->lo == -1) for (int r = 0; r < nol; r++) for (int c = 0; c < ni; c++) rop[r][c] = 0.0;
      if (op->hi == -1) for (int r = nol; r < ni; r++) for (int c = 0; c < ni; c++) rop[r][c] = 0.0;
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) ro[((size_t)id * 4 + c) * 4 + r] = rop[r][c];
    }
#undef AL
#undef RC
  }
  if (w->implicit_op && np > 1) { /* :184-197, identical on every rank */
    double *aa = calloc((size_t)16 * naa * np, sizeof(double));
    for (int j = 0; j < np; j++) {
      for (int r = 1; r <= 4; r++) {
        for (int c = 1; c <= nol; c++) AA(r, nol + c, 1, j) = ro[((size_t)j * 4 + (c - 1)) * 4 + (r - 1)];
        AA(r, r, 2, j) = 1.0;
        for (int c = 1; c <= nol; c++) AA(r, c, 3, j) = ro[((size_t)j * 4 + (nol + c - 1)) * 4 + (r - 1)];
      }
    }
    if (periodic) ptrid_block4_lud(aa, np); else btrid_block4_lud(aa, np);
    for (int id = 0; id < np; id++) {
      s->ops[id].aa = malloc(sizeof(double) * 16 * naa * np);
      cpy(s->ops[id].aa, aa, 16 * naa * np);
    }
    free(aa);
  }
  free(gal); free(gar); free(ro);
  return s;
}

/* ------------------------------------------------------------------------------------------ */
/* Directional evaluators on one bundle of PO_NB lines.                                        */
/* V, DV: [n][PO_NB] (global line, all ranks).  Halo "messages" are row copies.                */
/* ------------------------------------------------------------------------------------------ */
#define AR(l, i) op->ar[((size_t)(i)-1) * ncr + ((l)-1)]
#define VS(i) (vs + ((ptrdiff_t)(i)-1) * PO_NB)   /* local row i (1-based) of this rank */
#define DS(i) (ds + ((ptrdiff_t)(i)-1) * PO_NB)
#define H1(q) (vbr1 + ((q)-1) * PO_NB)            /* vbr1(q), q=1..nor */
#define H2(q) (vbr2 + ((q)-1) * PO_NB)
#define FORB for (int b = 0; b < PO_NB; b++)

/* compact_d1.f90:119-193 (x), :415-..., :752-... (y, z): rhs of the first derivative */
static void rhs_d1(const po_op *op, const double *vs, double *ds, const double *vbr1, const double *vbr2) {
  const int m = op->m, ncr = op->ncr;
  if (op->lo == -1) {
    if (op->bc[0] == -1) { /* :124-126 */
      FORB {
        double s;
        s = 0; for (int l = 4; l <= 7; l++) s += AR(l, 1) * VS(l - 3)[b]; DS(1)[b] = s;
        s = 0; for (int l = 3; l <= 7; l++) s += AR(l, 2) * VS(l - 2)[b]; DS(2)[b] = s;
        s = 0; for (int l = 2; l <= 7; l++) s += AR(l, 3) * VS(l - 1)[b]; DS(3)[b] = s;
      }
    } else { /* :134-136 */
      FORB {
        double s, v1 = VS(1)[b];
        s = 0; for (int l = 5; l <= 7; l++) s += AR(l, 1) * (VS(l - 3)[b] - v1); DS(1)[b] = s;
        s = 0; for (int l = 4; l <= 7; l++) s += AR(l, 2) * (VS(l - 2)[b] - v1); DS(2)[b] = s;
        DS(3)[b] = AR(5, 3) * (VS(4)[b] - VS(2)[b]) + AR(6, 3) * (VS(5)[b] - v1) + AR(7, 3) * (VS(6)[b] - v1);
      }
    }
  } else { /* :145-147 */
    FORB {
      DS(1)[b] = AR(5, 1) * (VS(2)[b] - H1(3)[b]) + AR(6, 1) * (VS(3)[b] - H1(2)[b]) + AR(7, 1) * (VS(4)[b] - H1(1)[b]);
      DS(2)[b] = AR(5, 2) * (VS(3)[b] - VS(1)[b]) + AR(6, 2) * (VS(4)[b] - H1(3)[b]) + AR(7, 2) * (VS(5)[b] - H1(2)[b]);
      DS(3)[b] = AR(5, 3) * (VS(4)[b] - VS(2)[b]) + AR(6, 3) * (VS(5)[b] - VS(1)[b]) + AR(7, 3) * (VS(6)[b] - H1(3)[b]);
    }
  }
  for (int i = 4; i <= m - 3; i++) { /* :153-159 */
    const double a5 = AR(5, i), a6 = AR(6, i), a7 = AR(7, i);
    FORB DS(i)[b] = a5 * (VS(i + 1)[b] - VS(i - 1)[b]) + a6 * (VS(i + 2)[b] - VS(i - 2)[b]) + a7 * (VS(i + 3)[b] - VS(i - 3)[b]);
  }
  if (op->hi == -1) {
    if (op->bc[1] == -1) { /* :166-168 */
      FORB {
        double s;
        s = 0; for (int l = 1; l <= 6; l++) s += AR(l, m - 2) * VS(m - 5 + l - 1)[b]; DS(m - 2)[b] = s;
        s = 0; for (int l = 1; l <= 5; l++) s += AR(l, m - 1) * VS(m - 4 + l - 1)[b]; DS(m - 1)[b] = s;
        s = 0; for (int l = 1; l <= 4; l++) s += AR(l, m) * VS(m - 3 + l - 1)[b]; DS(m)[b] = s;
      }
    } else { /* :176-178 */
      FORB {
        double s, vm = VS(m)[b];
        DS(m - 2)[b] = AR(1, m - 2) * (VS(m - 5)[b] - vm) + AR(2, m - 2) * (VS(m - 4)[b] - vm) + AR(3, m - 2) * (VS(m - 3)[b] - VS(m - 1)[b]);
        s = 0; for (int l = 1; l <= 4; l++) s += AR(l, m - 1) * (VS(m - 4 + l - 1)[b] - vm); DS(m - 1)[b] = s;
        s = 0; for (int l = 1; l <= 3; l++) s += AR(l, m) * (VS(m - 3 + l - 1)[b] - vm); DS(m)[b] = s;
      }
    }
  } else { /* :187-189 */
    FORB {
      DS(m - 2)[b] = AR(5, m - 2) * (VS(m - 1)[b] - VS(m - 3)[b]) + AR(6, m - 2) * (VS(m)[b] - VS(m - 4)[b]) + AR(7, m - 2) * (H2(1)[b] - VS(m - 5)[b]);
      DS(m - 1)[b] = AR(5, m - 1) * (VS(m)[b] - VS(m - 2)[b]) + AR(6, m - 1) * (H2(1)[b] - VS(m - 3)[b]) + AR(7, m - 1) * (H2(2)[b] - VS(m - 4)[b]);
      DS(m)[b] = AR(5, m) * (H2(1)[b] - VS(m - 1)[b]) + AR(6, m) * (H2(2)[b] - VS(m - 2)[b]) + AR(7, m) * (H2(3)[b] - VS(m - 3)[b]);
    }
  }
}

/* compact_r4.f90:107-193: 9-point difference-form rhs (d8 and every filter) */
static void rhs_r4(const po_op *op, const double *vs, double *ds, const double *vbr1, const double *vbr2) {
  const int m = op->m, ncr = op->ncr;
  if (op->lo == -1) {
    if (op->bc[0] == -1) { /* :112-115 */
      FORB for (int i = 1; i <= 4; i++) {
        double s = 0; for (int l = 6 - i; l <= 9; l++) s += AR(l, i) * VS(l - (5 - i))[b]; DS(i)[b] = s;
      }
    } else { /* :123-126 */
      FORB {
        double v1 = VS(1)[b];
        for (int i = 1; i <= 4; i++) {
          double s = 0; for (int l = 7 - i; l <= 9; l++) s += AR(l, i) * (VS(l - (5 - i))[b] - v1); DS(i)[b] = s;
        }
      }
    }
  } else { /* :135-138 */
    FORB for (int i = 1; i <= 4; i++) {
      double vc = VS(i)[b], s1 = 0, s2 = 0;
      for (int l = 1; l <= 5 - i; l++) s1 += AR(l, i) * (H1(l + i - 1)[b] - vc);
      for (int l = 6 - i; l <= 9; l++) s2 += AR(l, i) * (VS(l - (5 - i))[b] - vc);
      DS(i)[b] = s1 + s2;
    }
  }
  for (int i = 5; i <= m - 4; i++) { /* :147-153 */
    FORB {
      double s = 0.0, vc = VS(i)[b];
      for (int l = 1; l <= 9; l++) s = s + (VS(l + i - 5)[b] - vc) * AR(l, i);
      DS(i)[b] = s;
    }
  }
  if (op->hi == -1) {
    if (op->bc[1] == -1) { /* :163-166: dv(m-3+q) = sum(ar(1:8-q) * v(m-7+q : m)), q=0..3 */
      FORB for (int q = 0; q <= 3; q++) {
        int i = m - 3 + q; double s = 0;
        for (int l = 1; l <= 8 - q; l++) s += AR(l, i) * VS(m - 7 + q + l - 1)[b];
        DS(i)[b] = s;
      }
    } else { /* :174-177: dv(m-3+q) = sum(ar(1:7-q)*(v(m-7+q : m-1) - v(m))) */
      FORB {
        double vm = VS(m)[b];
        for (int q = 0; q <= 3; q++) {
          int i = m - 3 + q; double s = 0;
          for (int l = 1; l <= 7 - q; l++) s += AR(l, i) * (VS(m - 7 + q + l - 1)[b] - vm);
          DS(i)[b] = s;
        }
      }
    }
  } else { /* :186-189 */
    FORB for (int q = 0; q <= 3; q++) {
      int i = m - 3 + q; double vc = VS(i)[b], s1 = 0, s2 = 0;
      for (int l = 1; l <= 8 - q; l++) s1 += AR(l, i) * (VS(m - 7 + q + l - 1)[b] - vc);
      for (int l = 9 - q; l <= 9; l++) s2 += AR(l, i) * (H2(l - (9 - q) + 1)[b] - vc);
      DS(i)[b] = s1 + s2;
    }
  }
}

/* compact_r3.f90:89-149: 7-point difference-form rhs (d2); halos are zero at physical ends */
static void rhs_r3(const po_op *op, const double *vs, double *ds, const double *vbr1in, const double *vbr2in) {
  const int m = op->m, ncr = op->ncr;
  double z1[3 * PO_NB], z2[3 * PO_NB];
  const double *vbr1 = vbr1in, *vbr2 = vbr2in;
  if (op->lo == -1) { memset(z1, 0, sizeof(z1)); vbr1 = z1; } /* :89-96 */
  if (op->hi == -1) { memset(z2, 0, sizeof(z2)); vbr2 = z2; } /* :97-104 */
  const int anti1 = (op->lo == -1 && op->bc[0] == -1), anti2 = (op->hi == -1 && op->bc[1] == -1);
  FORB for (int i = 1; i <= 3; i++) { /* :110-112 / :118-120 */
    double vc = VS(i)[b], s1 = 0, s2 = 0;
    for (int l = 1; l <= 4 - i; l++) s1 += AR(l, i) * (H1(l + i - 1)[b] - vc);
    for (int l = 5 - i; l <= 7; l++) s2 += AR(l, i) * (VS(l - (4 - i))[b] - vc);
    double d = s1 + s2;
    if (anti1) { double sumr = 0; for (int l = 1; l <= 7; l++) sumr += AR(l, i); d += sumr * vc; }
    DS(i)[b] = d;
  }
  for (int i = 4; i <= m - 3; i++) { /* :126-129 */
    FORB {
      double s = 0.0, vc = VS(i)[b];
      for (int l = 1; l <= 7; l++) s += AR(l, i) * (VS(i - 4 + l)[b] - vc);
      DS(i)[b] = s;
    }
  }
  FORB for (int q = 0; q <= 2; q++) { /* :136-138 / :144-146 */
    int i = m - 2 + q; double vc = VS(i)[b], s1 = 0, s2 = 0;
    for (int l = 1; l <= 6 - q; l++) s1 += AR(l, i) * (VS(m - 5 + q + l - 1)[b] - vc);
    for (int l = 7 - q; l <= 7; l++) s2 += AR(l, i) * (H2(l - (7 - q) + 1)[b] - vc);
    double d = s1 + s2;
    if (anti2) { double sumr = 0; for (int l = 1; l <= 7; l++) sumr += AR(l, i); d += sumr * vc; }
    DS(i)[b] = d;
  }
}

/* One bundle through eval_compact_op1{x,y,z}_{d1,r3,r4}:
 * compact_d1.f90:38-339, compact_r4.f90:37-318, compact_r3.f90:37-214 */
static void eval_bundle(const po_opset *s, const double *V, double *DV, double *dvo /* [np][4][NB] */) {
  const int n = s->n, np = s->np, m = s->m, nor = s->w.nor;
  if (s->null_op) { /* compact_d1.f90:51-63, compact_r4.f90:48-66 */
    if (s->w.null_option == 1) memcpy(DV, V, sizeof(double) * (size_t)n * PO_NB);
    else memset(DV, 0, sizeof(double) * (size_t)n * PO_NB);
    return;
  }
  double vbr1[4 * PO_NB], vbr2[4 * PO_NB];
  for (int id = 0; id < np; id++) {
    const po_op *op = &s->ops[id];
    const int r0 = id * m; /* 0-based first global row */
    const double *vs = V + (size_t)r0 * PO_NB;
    double *ds = DV + (size_t)r0 * PO_NB;
    /* ghost data: MPI_Sendrecv from lo/hi, or periodic self copy (compact_d1.f90:82-110) */
    if (op->lo != -1)
      for (int q = 1; q <= nor; q++) {
        int g = ((r0 - nor + q - 1) % n + n) % n;
        memcpy(vbr1 + (q - 1) * PO_NB, V + (size_t)g * PO_NB, sizeof(double) * PO_NB);
      }
    if (op->hi != -1)
      for (int q = 1; q <= nor; q++) {
        int g = (r0 + m + q - 1) % n;
        memcpy(vbr2 + (q - 1) * PO_NB, V + (size_t)g * PO_NB, sizeof(double) * PO_NB);
      }
    switch (s->fam) {
      case FAM_D1: rhs_d1(op, vs, ds, vbr1, vbr2); break;
      case FAM_R3: rhs_r3(op, vs, ds, vbr1, vbr2); break;
      default: rhs_r4(op, vs, ds, vbr1, vbr2); break;
    }
    if (!op->implicit_op) { /* compact_r4.f90:209-218 */
      if (op->null_option == 1) for (size_t t = 0; t < (size_t)m * PO_NB; t++) ds[t] = ds[t] + vs[t];
      continue;
    }
    /* implicit part (compact_d1.f90:218-241 with gpu_kernel=1) */
    if (np == 1 && op->periodic) ppentLUS_bundle(op->al, ds, m);
    else bpentLUS_bundle(op->al, ds, m);
    if (np == 1) {
      if (op->null_option == 1) for (size_t t = 0; t < (size_t)m * PO_NB; t++) ds[t] = ds[t] + vs[t];
      continue;
    }
    /* pack the 4 interface values (compact_d1.f90:243-268) */
    double *o = dvo + (size_t)id * 4 * PO_NB;
    memcpy(o + 0 * PO_NB, ds + 0 * PO_NB, sizeof(double) * PO_NB);
    memcpy(o + 1 * PO_NB, ds + 1 * PO_NB, sizeof(double) * PO_NB);
    memcpy(o + 2 * PO_NB, ds + (size_t)(m - 2) * PO_NB, sizeof(double) * PO_NB);
    memcpy(o + 3 * PO_NB, ds + (size_t)(m - 1) * PO_NB, sizeof(double) * PO_NB);
    if (op->lo == -1) memset(o, 0, sizeof(double) * 2 * PO_NB);
    if (op->hi == -1) memset(o + 2 * PO_NB, 0, sizeof(double) * 2 * PO_NB);
  }
  if (np == 1 || !s->w.implicit_op) return;
  /* mpi_allgather == dvo is already complete; redundant reduced solve (compact_d1.f90:273-279) */
  if (s->periodic) ptrid_block4_lus(s->ops[0].aa, dvo, np);
  else btrid_block4_lus(s->ops[0].aa, dvo, np);
  for (int id = 0; id < np; id++) {
    const po_op *op = &s->ops[id];
    const int r0 = id * m;
    const double *vs = V + (size_t)r0 * PO_NB;
    double *ds = DV + (size_t)r0 * PO_NB;
#define RCF(i, c) op->rc[((c)-1) * (size_t)m + ((i)-1)]
#define DVO(l, k) (dvo + (((size_t)(k)) * 4 + ((l)-1)) * PO_NB)
    if (s->fam == FAM_D1) { /* compact_d1.f90:280-311 */
      if (op->lo != -1 && op->hi != -1) {
        for (int i = 1; i <= m; i++) FORB
          DS(i)[b] = DS(i)[b] - RCF(i, 1) * DVO(3, op->lo)[b] - RCF(i, 2) * DVO(4, op->lo)[b]
                              - RCF(i, 3) * DVO(1, op->hi)[b] - RCF(i, 4) * DVO(2, op->hi)[b];
      } else if (op->lo != -1) {
        for (int i = 1; i <= m; i++) FORB DS(i)[b] = DS(i)[b] - RCF(i, 1) * DVO(3, op->lo)[b] - RCF(i, 2) * DVO(4, op->lo)[b];
      } else if (op->hi != -1) {
        for (int i = 1; i <= m; i++) FORB DS(i)[b] = DS(i)[b] - RCF(i, 3) * DVO(1, op->hi)[b] - RCF(i, 4) * DVO(2, op->hi)[b];
      }
    } else { /* compact_r4.f90:270-283, compact_r3.f90:185-186 */
      if (op->lo != -1)
        for (int i = 1; i <= m; i++) FORB DS(i)[b] = DS(i)[b] - (RCF(i, 1) * DVO(3, op->lo)[b] + RCF(i, 2) * DVO(4, op->lo)[b]);
      if (op->hi != -1)
        for (int i = 1; i <= m; i++) FORB DS(i)[b] = DS(i)[b] - (RCF(i, 3) * DVO(1, op->hi)[b] + RCF(i, 4) * DVO(2, op->hi)[b]);
    }
    if (op->null_option == 1) for (size_t t = 0; t < (size_t)m * PO_NB; t++) ds[t] = ds[t] + vs[t]; /* compact_r4.f90:310-316 */
#undef RCF
#undef DVO
  }
}
#undef AR
#undef VS
#undef DS
#undef H1
#undef H2

/* Sweep a whole Fortran-ordered field (nx,ny,nz) along axis dir (0,1,2).  post: 0 none, 1 divide
 * by `scale` (compact_operators.f90:41-46), 2 divide by scale**2 (:152). */
static void sweep(const po_opset *s, int dir, int nx, int ny, int nz, const double *v, double *dv,
                  int post, double scale) {
  const int dims[3] = {nx, ny, nz};
  const size_t strides[3] = {1, (size_t)nx, (size_t)nx * ny};
  const int n = dims[dir];
  const int o1 = (dir == 0) ? 1 : 0, o2 = (dir == 2) ? 1 : 2; /* the two non-sweep axes, o1 faster */
  const long nlines = (long)dims[o1] * dims[o2];
  const long nbund = (nlines + PO_NB - 1) / PO_NB;
  const size_t st = strides[dir];
  const double div = (post == 2) ? scale * scale : scale;
#pragma omp parallel
  {
    double *V = malloc(sizeof(double) * (size_t)n * PO_NB);
    double *DV = malloc(sizeof(double) * (size_t)n * PO_NB);
    double *dvo = malloc(sizeof(double) * (size_t)(s->np > 0 ? s->np : 1) * 4 * PO_NB);
#pragma omp for schedule(static)
    for (long bu = 0; bu < nbund; bu++) {
      size_t base[PO_NB];
      int valid[PO_NB];
      for (int b = 0; b < PO_NB; b++) {
        long L = bu * PO_NB + b;
        valid[b] = L < nlines;
        if (!valid[b]) L = nlines - 1;
        base[b] = (size_t)(L % dims[o1]) * strides[o1] + (size_t)(L / dims[o1]) * strides[o2];
      }
      for (int r = 0; r < n; r++)
        for (int b = 0; b < PO_NB; b++) V[(size_t)r * PO_NB + b] = v[base[b] + (size_t)r * st];
      eval_bundle(s, V, DV, dvo);
      if (post && !s->null_op)
        for (size_t t = 0; t < (size_t)n * PO_NB; t++) DV[t] = DV[t] / div;
      for (int b = 0; b < PO_NB; b++) {
        if (!valid[b]) continue;
        for (int r = 0; r < n; r++) dv[base[b] + (size_t)r * st] = DV[(size_t)r * PO_NB + b];
      }
    }
    free(V); free(DV); free(dvo);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Plan == patch + comm + compact_ops + mesh of one (patch, level)  (objects.f90:35-49)         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int nx, ny, nz, px, py, pz, coordsys;
  int periodic[3], bcs[3][2]; /* BC_NONE / BC_PERI / BC_SYMM per side */
  int nop[3], mbc[3][2][2];   /* mbc[dir][side][iop]  (compact.f90:77-91) */
  int null_dir[3];
  double d[3], x1f[3], xnf[3];
  po_opset *ops[PO_NKIND][3][2];
  po_opset *custom_d1[3];
  size_t npts;
  double *xg, *yg, *zg, *d1, *d2, *d3, *gridlen, *cellvol, *cellvolS, *cellvolG;
  double *detxyz, *dA[3], *dB[3], *dC[3]; /* dAdx,dAdy,dAdz ... */
} po_plan;

void po_free(po_plan *p) {
  if (!p) return;
  for (int k = 0; k < PO_NKIND; k++) for (int d = 0; d < 3; d++) for (int o = 0; o < 2; o++) opset_free(p->ops[k][d][o]);
  for (int d = 0; d < 3; d++) opset_free(p->custom_d1[d]);
  free(p->xg); free(p->yg); free(p->zg); free(p->d1); free(p->d2); free(p->d3); free(p->gridlen);
  free(p->cellvol); free(p->cellvolS); free(p->cellvolG); free(p->detxyz);
  for (int d = 0; d < 3; d++) { free(p->dA[d]); free(p->dB[d]); free(p->dC[d]); }
  free(p);
}

/* parcop.f90:23-61 setup -> patch.f90:48-112, compact.f90:55-319.  bc codes: BC_NONE/PERI/SYMM.
 * x1..zn are the node extents Python passes. */
po_plan *po_setup(int nx, int ny, int nz, int px, int py, int pz, int coordsys, double x1, double xn,
                  double y1, double yn, double z1, double zn, int bx1, int bxn, int by1, int byn,
                  int bz1, int bzn) {
  po_plan *p = calloc(1, sizeof(*p));
  p->nx = nx; p->ny = ny; p->nz = nz; p->px = px; p->py = py; p->pz = pz; p->coordsys = coordsys;
  p->npts = (size_t)nx * ny * nz;
  const int nn[3] = {nx, ny, nz}, pp[3] = {px, py, pz};
  const double lo[3] = {x1, y1, z1}, hi[3] = {xn, yn, zn};
  const int b1[3] = {bx1, by1, bz1}, bn[3] = {bxn, byn, bzn};
  for (int d = 0; d < 3; d++) {
    if (nn[d] % pp[d] != 0) { free(p); return NULL; } /* comm.f90:189-206 */
    /* parcop.f90:46-56: nodes -> faces; patch.f90:83-85: dx = (xnf-x1f)/nx */
    double dn = (hi[d] - lo[d]) / (double)(nn[d] - 1 > 1 ? nn[d] - 1 : 1);
    p->x1f[d] = lo[d] - dn / 2.0; p->xnf[d] = hi[d] + dn / 2.0;
    p->d[d] = (p->xnf[d] - p->x1f[d]) / (double)nn[d];
    p->bcs[d][0] = b1[d]; p->bcs[d][1] = bn[d];
    p->periodic[d] = (b1[d] == BC_PERI); /* patch.f90:89-109 */
    p->nop[d] = 1;
    if (b1[d] == BC_SYMM) { p->mbc[d][0][0] = 1; p->mbc[d][0][1] = -1; p->nop[d] = 2; }
    if (bn[d] == BC_SYMM) { p->mbc[d][1][0] = 1; p->mbc[d][1][1] = -1; p->nop[d] = 2; }
    p->null_dir[d] = nn[d] < 4; /* compact.f90:95-97 */
    for (int k = 0; k < PO_NKIND; k++)
      for (int o = 0; o < p->nop[d]; o++)
        p->ops[k][d][o] = opset_create(k, nn[d], pp[d], p->periodic[d], p->mbc[d][0][o], p->mbc[d][1][o], p->null_dir[d]);
  }
  return p;
}

static const int post_of_kind[PO_NKIND] = {1, 2, 0, 0, 0, 0}; /* d4: "No metric... unity assumed" (compact_operators.f90:229) */

/* d1x/d2x/d8x/filterx dispatch (compact_operators.f90:13-50,130-154,280-313,385-421).
 * bc: 0 => iop 1; -1 => iop 2 if it exists; >0 => iop = bc. */
static void dir_op(const po_plan *p, int kind, int dir, int bc, const double *v, double *dv) {
  const size_t N = p->npts;
  if (p->null_dir[dir]) {
    if (kind == PO_SF || kind == PO_GF) memcpy(dv, v, sizeof(double) * N);
    else memset(dv, 0, sizeof(double) * N);
    return;
  }
  int iop = 1;
  if (bc > 0) iop = bc;
  if (bc == -1 && p->nop[dir] > 1) iop = 2;
  sweep(p->ops[kind][dir][iop - 1], dir, p->nx, p->ny, p->nz, v, dv, post_of_kind[kind], p->d[dir]);
}

void po_dir_op(const po_plan *p, int kind, int dir, int bc, const double *v, double *dv) { dir_op(p, kind, dir, bc, v, dv); }
/* raw evaluator without the metric scale (op%evalx etc.) */
void po_eval_raw(const po_plan *p, int kind, int dir, int iop, const double *v, double *dv) {
  sweep(p->ops[kind][dir][iop - 1], dir, p->nx, p->ny, p->nz, v, dv, 0, 1.0);
}

/* parcop.f90:225-253 */
void po_ddx(const po_plan *p, const double *v, double *dv) { dir_op(p, PO_D1, 0, 0, v, dv); }
void po_ddy(const po_plan *p, const double *v, double *dv) { dir_op(p, PO_D1, 1, 0, v, dv); }
void po_ddz(const po_plan *p, const double *v, double *dv) { dir_op(p, PO_D1, 2, 0, v, dv); }
/* parcop.f90:279-301 */
void po_dd8(const po_plan *p, int dir, const double *v, double *dv) { dir_op(p, PO_D8, dir, 0, v, dv); }
void po_d2(const po_plan *p, int dir, const double *v, double *dv) { dir_op(p, PO_D2, dir, 0, v, dv); }
/* parcop.f90:255-277 dd4x/dd4y/dd4z -> d4x/d4y/d4z (compact_operators.f90:208-278) */
void po_dd4(const po_plan *p, int dir, const double *v, double *dv) { dir_op(p, PO_D4, dir, 0, v, dv); }

static int isym(const po_plan *p, int d) { return (p->bcs[d][0] == BC_SYMM || p->bcs[d][1] == BC_SYMM) ? -1 : 1; } /* patch.f90:86-88 */

/* operators.f90:38-53 (Cartesian), :77-91 (curvilinear) */
void po_div(const po_plan *p, const double *fx, const double *fy, const double *fz, double *df) {
  const size_t N = p->npts;
  double *fA = malloc(sizeof(double) * N), *fB = malloc(sizeof(double) * N), *fC = malloc(sizeof(double) * N);
  if (p->coordsys == 0) {
    dir_op(p, PO_D1, 0, isym(p, 0), fx, fA);
    dir_op(p, PO_D1, 1, isym(p, 1), fy, fB);
    dir_op(p, PO_D1, 2, isym(p, 2), fz, fC);
    for (size_t t = 0; t < N; t++) df[t] = fA[t] + fB[t] + fC[t];
  } else {
    double *tmp = malloc(sizeof(double) * N);
    for (size_t t = 0; t < N; t++) {
      fA[t] = (fx[t] * p->dA[0][t] + fy[t] * p->dA[1][t] + fz[t] * p->dA[2][t]) * p->detxyz[t];
      fB[t] = (fx[t] * p->dB[0][t] + fy[t] * p->dB[1][t] + fz[t] * p->dB[2][t]) * p->detxyz[t];
      fC[t] = (fx[t] * p->dC[0][t] + fy[t] * p->dC[1][t] + fz[t] * p->dC[2][t]) * p->detxyz[t];
    }
    dir_op(p, PO_D1, 0, 0, fA, df);
    dir_op(p, PO_D1, 1, 0, fB, tmp);
    for (size_t t = 0; t < N; t++) df[t] = df[t] + tmp[t];
    dir_op(p, PO_D1, 2, 0, fC, tmp);
    for (size_t t = 0; t < N; t++) df[t] = (df[t] + tmp[t]) / p->detxyz[t];
    free(tmp);
  }
  free(fA); free(fB); free(fC);
}

/* operators.f90:184-212 */
void po_grad(const po_plan *p, const double *f, double *gx, double *gy, double *gz) {
  dir_op(p, PO_D1, 0, 0, f, gx);
  dir_op(p, PO_D1, 1, 0, f, gy);
  dir_op(p, PO_D1, 2, 0, f, gz);
  if (p->coordsys == 3) {
    for (size_t t = 0; t < p->npts; t++) {
      double a = gx[t], b = gy[t], c = gz[t];
      gx[t] = a * p->dA[0][t] + b * p->dB[0][t] + c * p->dC[0][t];
      gy[t] = a * p->dA[1][t] + b * p->dB[1][t] + c * p->dC[1][t];
      gz[t] = a * p->dA[2][t] + b * p->dB[2][t] + c * p->dC[2][t];
    }
  }
}

/* operators.f90:513-526 (Cartesian only; no coordsys=3 branch exists in the reference) */
void po_lap(const po_plan *p, const double *f, double *lap) {
  const size_t N = p->npts;
  double *tmp = malloc(sizeof(double) * N), *dum = malloc(sizeof(double) * N);
  dir_op(p, PO_D2, 0, 0, f, lap);
  dir_op(p, PO_D2, 1, 0, f, tmp);
  dir_op(p, PO_D2, 2, 0, f, dum);
  for (size_t t = 0; t < N; t++) lap[t] = lap[t] + tmp[t] + dum[t];
  free(tmp); free(dum);
}

/* operators.f90:615-643 ringS with L (parcop.f90:320 uses L=2), :701-753 ringx/y/z */
void po_ring(const po_plan *p, const double *f, double *out) {
  const size_t N = p->npts;
  const int nn[3] = {p->nx, p->ny, p->nz};
  double *r[3];
  for (int d = 0; d < 3; d++) {
    r[d] = malloc(sizeof(double) * N);
    if (nn[d] == 1) memset(r[d], 0, sizeof(double) * N);
    else dir_op(p, PO_D8, d, 0, f, r[d]);
  }
  for (size_t t = 0; t < N; t++) {
    double a = fabs(r[0][t]) * (p->d1[t] * p->d1[t]), b = fabs(r[1][t]) * (p->d2[t] * p->d2[t]),
           c = fabs(r[2][t]) * (p->d3[t] * p->d3[t]);
    double mx = a > b ? a : b;
    out[t] = mx > c ? mx : c;
  }
  for (int d = 0; d < 3; d++) free(r[d]);
}

/* operators.f90:97-123 divT, Cartesian branch (parcop.f90:213-223 divergenceTensor): three
 * divergences of the tensor's columns, each direction with its symmetry selector. */
int po_divT(const po_plan *p, const double *fxx, const double *fxy, const double *fxz, const double *fyx, const double *fyy,
            const double *fyz, const double *fzx, const double *fzy, const double *fzz, double *dfx, double *dfy, double *dfz) {
  if (p->coordsys == 3) { /* :176-179: divV of each row */
    po_div(p, fxx, fxy, fxz, dfx);
    po_div(p, fyx, fyy, fyz, dfy);
    po_div(p, fzx, fzy, fzz, dfz);
    return 0;
  }
  if (p->coordsys != 0) return -1;
  const size_t N = p->npts;
  double *fA = malloc(sizeof(double) * N), *fB = malloc(sizeof(double) * N), *fC = malloc(sizeof(double) * N);
  const double *in[3][3] = {{fxx, fyx, fzx}, {fxy, fyy, fzy}, {fxz, fyz, fzz}};
  double *out[3] = {dfx, dfy, dfz};
  for (int c = 0; c < 3; c++) {
    /* :106-108,112-114,118-120: selector isym(d), squared (= +1) on the diagonal */
    dir_op(p, PO_D1, 0, c == 0 ? 1 : isym(p, 0), in[c][0], fA);
    dir_op(p, PO_D1, 1, c == 1 ? 1 : isym(p, 1), in[c][1], fB);
    dir_op(p, PO_D1, 2, c == 2 ? 1 : isym(p, 2), in[c][2], fC);
    for (size_t t = 0; t < N; t++) out[c][t] = fA[t] + fB[t] + fC[t]; /* :109,114,119 */
  }
  free(fA); free(fB); free(fC);
  return 0;
}

/* operators.f90:645-699 ringV with L = 1 (parcop.f90:324-333 pRingV), ringx/y/z :701-753 */
void po_ringV(const po_plan *p, const double *f, const double *g, const double *h, double *out) {
  const size_t N = p->npts;
  const int nn[3] = {p->nx, p->ny, p->nz};
  const double *comp[3] = {f, g, h};
  double *r[3][3];
  for (int c = 0; c < 3; c++)
    for (int d = 0; d < 3; d++) {
      r[c][d] = malloc(sizeof(double) * N);
      if (nn[d] == 1) memset(r[c][d], 0, sizeof(double) * N);
      else dir_op(p, PO_D8, d, c == d ? isym(p, d) : 0, comp[c], r[c][d]); /* :661-671 */
    }
  for (size_t t = 0; t < N; t++) {
    double L[3];
    const double dd[3] = {p->d1[t], p->d2[t], p->d3[t]};
    for (int d = 0; d < 3; d++) { /* :680-682 */
      double a = fabs(r[0][d][t]), b = fabs(r[1][d][t]), c = fabs(r[2][d][t]);
      double mx = a > b ? a : b;
      mx = mx > c ? mx : c;
      L[d] = mx * dd[d];
    }
    double mx = L[0] > L[1] ? L[0] : L[1];
    out[t] = mx > L[2] ? mx : L[2]; /* :683 */
  }
  for (int c = 0; c < 3; c++)
    for (int d = 0; d < 3; d++) free(r[c][d]);
}

/* operators.f90:781-896: filtype 'spectral' (which=0, parcop.f90:337-346) or 'smooth' (which=1,
 * parcop.f90:348-357); scalar => x/y/zasym = 1 */
void po_filter(const po_plan *p, int which, const double *fun, double *bar) {
  const size_t N = p->npts;
  double *tmp = calloc(N, sizeof(double));
  if (which == 1) { /* 'smooth' :848-853 */
    dir_op(p, PO_GF, 0, 1, fun, bar);
    dir_op(p, PO_GF, 1, 1, bar, tmp);
    dir_op(p, PO_GF, 2, 1, tmp, bar);
  } else if (p->coordsys == 0) { /* :872-875 */
    dir_op(p, PO_SF, 0, 1, fun, bar);
    dir_op(p, PO_SF, 1, 1, bar, tmp);
    dir_op(p, PO_SF, 2, 1, tmp, bar);
  } else { /* :887-892 */
    for (size_t t = 0; t < N; t++) tmp[t] = fun[t] * p->cellvol[t];
    dir_op(p, PO_SF, 0, 1, tmp, bar);
    dir_op(p, PO_SF, 1, 1, bar, tmp);
    dir_op(p, PO_SF, 2, 1, tmp, bar);
    for (size_t t = 0; t < N; t++) bar[t] = bar[t] / p->cellvolS[t];
  }
  free(tmp);
}

/* operators.f90:755-778 (direction 1..3) */
void po_gfilter_dir(const po_plan *p, int direction, const double *fun, double *bar) {
  dir_op(p, PO_GF, direction - 1, 1, fun, bar);
}

/* ------------------------------------------------------------------------------------------ */
/* mesh.f90:59-390                                                                             */
/* ------------------------------------------------------------------------------------------ */
static double *dalloc(size_t n) { return calloc(n, sizeof(double)); }

static int setup_mesh_common(po_plan *p, const double *xpy, const double *ypy, const double *zpy, int custom_per) {
  const int nx = p->nx, ny = p->ny, nz = p->nz;
  const size_t N = p->npts;
  const double dx = p->d[0], dy = p->d[1], dz = p->d[2];
  p->xg = dalloc(N); p->yg = dalloc(N); p->zg = dalloc(N);
  p->d1 = dalloc(N); p->d2 = dalloc(N); p->d3 = dalloc(N);
  p->gridlen = dalloc(N); p->cellvol = dalloc(N); p->cellvolS = dalloc(N); p->cellvolG = dalloc(N);
  if (xpy) { cpy(p->xg, xpy, N); cpy(p->yg, ypy, N); cpy(p->zg, zpy, N); } /* :168-171 */
  else /* :173-177 */
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
      size_t t = i + (size_t)nx * (j + (size_t)ny * k);
      p->xg[t] = p->x1f[0] + (double)(2 * (i + 1) - 1) * 0.5 * dx;
      p->yg[t] = p->x1f[1] + (double)(2 * (j + 1) - 1) * 0.5 * dy;
      p->zg[t] = p->x1f[2] + (double)(2 * (k + 1) - 1) * 0.5 * dz;
    }
  if (p->coordsys == 0) { /* :191-212 */
    double d1 = nx == 1 ? fmax(dy, dz) : dx, d2 = ny == 1 ? fmax(dx, dz) : dy, d3 = nz == 1 ? fmax(dx, dy) : dz;
    for (size_t t = 0; t < N; t++) {
      p->cellvol[t] = dx * dy * dz; p->d1[t] = d1; p->d2[t] = d2; p->d3[t] = d3;
      p->gridlen[t] = fmin(d1, fmin(d2, d3));
      p->cellvolS[t] = p->cellvol[t]; p->cellvolG[t] = p->cellvol[t]; /* unused by Cartesian filter */
    }
    return 0;
  }
  if (p->coordsys != 3) return -1; /* cylindrical / spherical: out of scope */
  /* :244-358 */
  double *J[3][3]; /* J[c][a] = d(x_c)/d(A_a) */
  for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) J[c][a] = dalloc(N);
  const double *xyz[3] = {p->xg, p->yg, p->zg};
  const double dABC[3] = {dx, dy, dz};
  const int nn[3] = {nx, ny, nz}, pp[3] = {p->px, p->py, p->pz};
  for (int a = 0; a < 3; a++) {
    const po_opset *s;
    if (custom_per) { /* :251-304: non-periodic d1 with one-sided ends, bc = [0,0] */
      if (!p->custom_d1[a]) p->custom_d1[a] = opset_create(PO_D1, nn[a], pp[a], 0, 0, 0, 0);
      s = p->custom_d1[a];
    } else s = p->ops[PO_D1][a][0]; /* :308-327 */
    for (int c = 0; c < 3; c++) {
      sweep(s, a, nx, ny, nz, xyz[c], J[c][a], 0, 1.0);
      for (size_t t = 0; t < N; t++) J[c][a][t] = J[c][a][t] / dABC[a];
    }
  }
  if (nx == 1) for (size_t t = 0; t < N; t++) J[0][0][t] = 1.0; /* :332-334 */
  if (ny == 1) for (size_t t = 0; t < N; t++) J[1][1][t] = 1.0;
  if (nz == 1) for (size_t t = 0; t < N; t++) J[2][2][t] = 1.0;
  p->detxyz = dalloc(N);
  for (int c = 0; c < 3; c++) { p->dA[c] = dalloc(N); p->dB[c] = dalloc(N); p->dC[c] = dalloc(N); }
  for (size_t t = 0; t < N; t++) {
    const double dxdA = J[0][0][t], dxdB = J[0][1][t], dxdC = J[0][2][t];
    const double dydA = J[1][0][t], dydB = J[1][1][t], dydC = J[1][2][t];
    const double dzdA = J[2][0][t], dzdB = J[2][1][t], dzdC = J[2][2][t];
    const double det = -dxdC * dydB * dzdA + dxdB * dydC * dzdA + dxdC * dydA * dzdB
                       - dxdA * dydC * dzdB - dxdB * dydA * dzdC + dxdA * dydB * dzdC; /* :337-338 */
    p->detxyz[t] = det;
    p->dA[0][t] = (-dydC * dzdB + dydB * dzdC) / det; /* :341-351 */
    p->dA[1][t] = (dxdC * dzdB - dxdB * dzdC) / det;
    p->dA[2][t] = (-dxdC * dydB + dxdB * dydC) / det;
    p->dB[0][t] = (dydC * dzdA - dydA * dzdC) / det;
    p->dB[1][t] = (-dxdC * dzdA + dxdA * dzdC) / det;
    p->dB[2][t] = (dxdC * dydA - dxdA * dydC) / det;
    p->dC[0][t] = (-dydB * dzdA + dydA * dzdB) / det;
    p->dC[1][t] = (dxdB * dzdA - dxdA * dzdB) / det;
    p->dC[2][t] = (-dxdB * dydA + dxdA * dydB) / det;
    const double dA = dx, dB = dy, dC = dz;
    p->d1[t] = sqrt((dxdA * dA) * (dxdA * dA) + (dydA * dA) * (dydA * dA) + (dzdA * dA) * (dzdA * dA)); /* :354-356 */
    p->d2[t] = sqrt((dxdB * dB) * (dxdB * dB) + (dydB * dB) * (dydB * dB) + (dzdB * dB) * (dzdB * dB));
    p->d3[t] = sqrt((dxdC * dC) * (dxdC * dC) + (dydC * dC) * (dydC * dC) + (dzdC * dC) * (dzdC * dC));
    p->cellvol[t] = p->d1[t] * p->d2[t] * p->d3[t];
    p->gridlen[t] = fmin(p->d1[t], fmin(p->d2[t], p->d3[t]));
  }
  for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) free(J[c][a]);
  /* :374-383 filtered cell volumes (evalx/y/z called directly: no null-direction short cut
   * other than the operator's own null_op) */
  double *tmp = dalloc(N);
  sweep(p->ops[PO_SF][0][0], 0, nx, ny, nz, p->cellvol, p->cellvolS, 0, 1.0);
  sweep(p->ops[PO_SF][1][0], 1, nx, ny, nz, p->cellvolS, tmp, 0, 1.0);
  sweep(p->ops[PO_SF][2][0], 2, nx, ny, nz, tmp, p->cellvolS, 0, 1.0);
  sweep(p->ops[PO_GF][0][0], 0, nx, ny, nz, p->cellvol, p->cellvolG, 0, 1.0);
  sweep(p->ops[PO_GF][1][0], 1, nx, ny, nz, p->cellvolG, tmp, 0, 1.0);
  sweep(p->ops[PO_GF][2][0], 2, nx, ny, nz, tmp, p->cellvolG, 0, 1.0);
  free(tmp);
  return 0;
}

int po_setup_mesh(po_plan *p) { return setup_mesh_common(p, NULL, NULL, NULL, 0); }                 /* parcop.f90:64-70 */
int po_setup_mesh_x3(po_plan *p, const double *x, const double *y, const double *z, int mesh_per) { /* parcop.f90:73-81 */
  return setup_mesh_common(p, x, y, z, mesh_per);
}

/* parcop.f90:84-128 getVar */
int po_getvar(const po_plan *p, const char *name, double *out) {
  const double *src = NULL;
  if (!strcmp(name, "x")) src = p->xg; else if (!strcmp(name, "y")) src = p->yg; else if (!strcmp(name, "z")) src = p->zg;
  else if (!strcmp(name, "d1")) src = p->d1; else if (!strcmp(name, "d2")) src = p->d2; else if (!strcmp(name, "d3")) src = p->d3;
  else if (!strcmp(name, "dAx")) src = p->dA[0]; else if (!strcmp(name, "dAy")) src = p->dA[1]; else if (!strcmp(name, "dAz")) src = p->dA[2];
  else if (!strcmp(name, "dBx")) src = p->dB[0]; else if (!strcmp(name, "dBy")) src = p->dB[1]; else if (!strcmp(name, "dBz")) src = p->dB[2];
  else if (!strcmp(name, "dCx")) src = p->dC[0]; else if (!strcmp(name, "dCy")) src = p->dC[1]; else if (!strcmp(name, "dCz")) src = p->dC[2];
  else if (!strcmp(name, "dtJ")) src = p->detxyz;
  else if (!strcmp(name, "CellVol")) src = p->cellvol; else if (!strcmp(name, "GridLen")) src = p->gridlen;
  else if (!strcmp(name, "CellVolS")) src = p->cellvolS; else if (!strcmp(name, "CellVolG")) src = p->cellvolG;
  if (!src) return -1;
  cpy(out, src, p->npts);
  return 0;
}

double po_spacing(const po_plan *p, int dir) { return p->d[dir]; }

/* LU / spike tables of one emulated rank, for table-level tests of the product's plan builder */
int po_get_tables(const po_plan *p, int kind, int dir, int iop, int rank, double *al, double *rc, double *aa) {
  const po_opset *s = p->ops[kind][dir][iop - 1];
  if (!s || s->null_op) return -1;
  const po_op *op = &s->ops[rank];
  if (al && op->al) cpy(al, op->al, (size_t)op->nal * op->m);
  if (rc && op->rc) cpy(rc, op->rc, (size_t)4 * op->m);
  if (aa && op->aa) cpy(aa, op->aa, (size_t)16 * op->naa * op->np);
  return op->nal;
}

int po_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

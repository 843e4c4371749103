// Device-side declarations shared by kernels.cu and api.cu.
#pragma once
#ifdef PB_EMULATE
#include "cuda_emul.h"  // tests/emul: host emulation of the CUDA sources (test infrastructure)
#else
#include <cuda_runtime.h>
#endif

namespace pb {

constexpr int kMaxChunks = 32;
constexpr int kParamRows = 32;   // chunk length up to which the constant-chunk tables ride in the kernel parameters
constexpr int kMTerms = 8;       // transfer products per direction that fit the parameter copy (SweepDev::Mf0 / Mb0)
constexpr int kBlockThreads = 256; // every sweep block has at most this many threads

// epilogue applied when a sweep writes its result
enum Epi {
  EPI_STORE = 0,     // out = val
  EPI_ACC = 1,       // out += val                         (div / laplacian accumulation)
  EPI_RING_SET = 2,  // out = |val| * s^2                  (ring, first direction)
  EPI_RING_MAX = 3   // out = max(out, |val| * s^2)        (ring, later directions)
};

// Everything a directional sweep needs; passed by value (kernel parameter / constant bank).
struct SweepDev {
  // line geometry --------------------------------------------------------------------------
  int m;         // rows along the sweep axis on this rank
  int nfast;     // y/z sweeps: lines along x (ax);            x sweep: number of lines (ay*az)
  int nouter;    // y sweep: az, z sweep: ay;                   x sweep: unused
  long rstride;  // stride between consecutive rows (y: ax, z: ax*ay, x: 1)
  long ostride;  // stride between outer slices (y: ax*ay, z: ax)
  // partition of a line into chunks --------------------------------------------------------
  int P, C;
  int ctype[kMaxChunks];
  int has_const;        // chunk type 0 has row-independent LU coefficients, kept in cst[]
  double cst[5];        // l2, l1, 1/pivot, u1, u2 of the converged rows
  const double2 *luf;   // [ntypes][C]  {l2, l1}   forward multipliers (rows i-2, i-1)
  const double4 *lub;   // [ntypes][C]  y / z sweeps: {1/pivot, u1/pivot, u2/pivot, 0}; x sweeps: {1/pivot, u1, u2, 0}
  const double2 *phi;   // [ntypes][C]  forward response to the state entering the chunk
  const double2 *psi;   // [ntypes][C]  backward response to the state entering the chunk
  const double4 *Mf;    // [P][P+1]     2x2 products giving the forward state entering a chunk
  const double4 *Mb;    // [P][P+1]     same for the backward state
  unsigned char nf[kMaxChunks], nb[kMaxChunks];  // terms kept per chunk
  unsigned char perm[kMaxChunks];  // thread slot -> chunk: chunks that read coefficient tables come first, so the
                                   // two chunks sharing a warp (16-line tiles) take the same code path
  int cparam;           // phi0 / psi0 below are valid (has_const and C <= kParamRows)
  double2 phi0[kParamRows], psi0[kParamRows];  // phi / psi of the constant chunk type: constant-bank operands
  // right-hand side ------------------------------------------------------------------------
  double ari[9];
  double arb_lo[4][9], arb_hi[4][9];
  int phys_lo, phys_hi;  // one-sided closure rows at the local ends
  int wrap;              // periodic and not split: halo rows come from the field itself
  int implicit;          // solve after the stencil
  int add_v;             // filters: result += input (null_option == 1)
  double scale;          // 1/d, 1/d^2 or 1
  int acc;               // pipelined kernels, plain-store path: 0 store, 1 out += val (TMA reduce-add), 2 out = max(out, val)
  int ring;              // pipelined kernels, plain-store path: val = |val| * ring_s2 first (ring detector, Cartesian)
  double ring_s2;
  int wstore;            // pipelined kernels: every warp issues the TMA stores of its own chunks (else one thread per block)
  int skew_ns;           // persistent kernels: the second wave of CTAs (the co-residents of the first) starts this much later
  int mstride;           // entries per chunk row of Mf / Mb (P + 1, or the compact stride of a z-slab line)
  // lines whose chunks all share ONE row of transfer products (periodic lines): the products ride in the kernel
  // parameters, so the carried-state sums take them as constant-bank operands and load nothing
  int mconst, nf0, nb0;
  double4 Mf0[kMTerms], Mb0[kMTerms];
  const double2 *chi;    // [ntypes][C]  z-slab ring: solution response to a forward state that arrives after the chunk's solve
  double2 chi0[kParamRows];  // the same for the constant chunk type
};

// Cross-rank exchange of chunk states inside a z sweep (ring kernel): a z-slab line is the global
// line cut into chunks; the states of the chunks next to a slab face travel to the neighbouring
// ranks through peer memory as self-validating records (4 words of {32 data bits, 32-bit epoch}).
constexpr int kXHops = 3;  // ranks above / below that one rank's chunk states can reach
constexpr int kXExt = 8;   // chunk states a rank takes in from either side
struct XRing {
  int on;
  unsigned int epoch;
  int need_f, need_b;                                    // chunk states polled: forward ends from below, backward starts from above
  int nup, ndn;                                          // ranks written to above / below
  int cnt_up[kXHops], cnt_dn[kXHops];                    // my top (bottom) chunks rank r+1+k (r-1-k) takes; its slot = k P + e
  unsigned long long *en_out[kXHops], *st_out[kXHops];   // those ranks' incoming buffers (peer-mapped)
  const unsigned long long *en_in, *st_in;               // this rank's incoming buffers [slot][line][4]
  long plane;                                            // lines per xy-plane
  // halo planes pushed by the sweep kernel itself (push = 0: the caller has exchanged them): every CTA
  // copies its share of this rank's first / last planes into the neighbours' halo buffers, the last
  // one publishes hepoch in the neighbours' flag words; halo boxes are requested once theirs is seen
  // early = 1: nobody waits for a neighbour before solving.  A chunk solves with the states of this
  // rank's chunks, sends its backward start state at once, and adds what the forward states from
  // below contribute (chi) when they have arrived; the top chunks add, to the backward states from
  // above, what this rank's own forward states change in them (Bc: [need_b][bc_n] 2x2 blocks applied
  // to the forward end states of this rank's top bc_n chunks)
  int early, bc_n;
  const double4 *Bc;
  int push, npeers, nopoll;
  const double *push_src[2];
  double *push_dst[2];
  long push_n;                                           // doubles per copy
  unsigned long long *hflag_remote[2];
  const volatile unsigned long long *hflag_local[2];
  unsigned long long hepoch;
  unsigned int *hcounter;
};

struct EpiArgs {
  int mode;
  double s2;           // constant length-scale squared for EPI_RING_* when field == nullptr
  const double *field; // per-point length scale (curvilinear d1/d2/d3); squared in the kernel
};

// launches (return cudaError_t of the launch)
cudaError_t launch_sweep_yz(int fam, int lines, const SweepDev &a, const double *v, double *out,
                            const double *halo_lo, const double *halo_hi, double *iface,
                            const EpiArgs &epi, cudaStream_t st);
cudaError_t launch_sweep_x(int fam, int lines, const SweepDev &a, const double *v, double *out,
                           const EpiArgs &epi, cudaStream_t st);
// the ring kernel directly: clusters for long lines, cross-rank chunk states for z-slabs (xr may be null)
cudaError_t launch_sweep_ring(int fam, int lines, const SweepDev &a, const double *v, double *out,
                              const double *halo_lo, const double *halo_hi, const XRing *xr,
                              const EpiArgs &epi, cudaStream_t st);

// z-slab helpers
cudaError_t launch_pack_planes(const double *v, long plane, int m, int h, double *send_lo,
                               double *send_hi, cudaStream_t st);
cudaError_t launch_z_finish(double *out, long plane, int m, const double4 *RC, const double *GR, int np,
                            unsigned long long rank_mask, int zone_lo, int zone_hi, const double *iface_all,
                            double scale, cudaStream_t st);

// peer exchange: copies into the neighbours' memory + flag handshake in one launch
struct PeerExchange {
  int ncopies, npeers;
  void *dst[4];
  const void *src[4];
  size_t bytes[4];                              // multiples of 16
  unsigned long long *remote_flag[2];           // in each neighbour's memory: the word this rank sets
  const volatile unsigned long long *local_flag[2];  // in this rank's memory: the word each neighbour sets
  unsigned long long epoch;
  unsigned int *counter;                        // zero-initialised scratch word in local memory
};
cudaError_t launch_peer_exchange(const PeerExchange &x, cudaStream_t st);

// pointwise / reductions
cudaError_t launch_rk4_stage(long n, double dt, double A, double B, const double *F, double *PHI,
                             double *U, cudaStream_t st);
cudaError_t launch_reduce(int kind, long n, const double *v, double *partial, int nblocks,
                          double *result, cudaStream_t st);
cudaError_t launch_copy(long n, const double *a, double *out, cudaStream_t st);
cudaError_t launch_fill(long n, double val, double *out, cudaStream_t st);
cudaError_t launch_mul(long n, const double *a, const double *b, double *out, cudaStream_t st);
cudaError_t launch_div(long n, const double *a, const double *b, double *out, cudaStream_t st);
cudaError_t launch_mul_max(long n, const double *a, const double *b, double *out, int first, cudaStream_t st);
// curvilinear divergence pre-contraction: fA/fB/fC = (fx*dAdx + fy*dAdy + fz*dAdz)*det ...
cudaError_t launch_contra(long n, const double *fx, const double *fy, const double *fz,
                          const double *const *metric9, const double *det, double *fA, double *fB,
                          double *fC, cudaStream_t st);
// curvilinear gradient contraction in place on (gx, gy, gz)
cudaError_t launch_grad_contract(long n, const double *const *metric9, double *gx, double *gy,
                                 double *gz, cudaStream_t st);
// metrics from the nine Jacobian entries (mesh.f90:337-358)
cudaError_t launch_metrics(long n, const double *const *J9, double dA, double dB, double dC,
                           double *const *inv9, double *det, double *d1, double *d2, double *d3,
                           double *cellvol, double *gridlen, cudaStream_t st);

long launch_count();
long pipe_launch_count();
long ring_launch_count();
void set_ring_kernels(int mode, int lines);
void set_yz_lines(int nl);
void set_reg_kernels(int on);
void set_x_lines(int nl);

}  // namespace pb

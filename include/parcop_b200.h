/*
 * parcop_b200.h -- C ABI of libparcop_b200.so, the B200 (sm_100a) replacement for the hot path of
 * LLNL/pyranda's f2py extension module `parcop` (reference: pyranda/parcop/parcop.f90, built by
 * pyranda/parcop/makefile:96-102 and bound in pyranda/pyrandaMPI.py:151-174,664-740).
 *
 * Conventions
 *   - Every entry point returns PB_OK (0) or a negative PB_ERR_* code; pb_last_error() gives the
 *     text.  (The reference prints and STOPs: compact_d1.f90:65-72.)
 *   - Fields are double precision, Fortran order (i fastest), local extents ax*ay*az of this rank,
 *     exactly what f2py hands to `real(8), dimension(nx,ny,nz)` arguments (parcop.f90:228).
 *   - `pb_*` operator entry points take DEVICE pointers and a cudaStream_t (passed as void*);
 *     output must not alias input.  `pb_host_*` take HOST pointers (pageable or pinned), do the
 *     H2D copy, the device operator and the D2H copy, and are what the f2py call shapes
 *     (`dval = parcop.ddx(val)`) map onto one-to-one.
 *   - One pb_plan replaces one (patch, level) slot of the reference's module-global object tables
 *     (objects.f90:22-30, set_patch parcop.f90:196-200).
 *   - No torch / C++ types cross this boundary.
 */
#ifndef PARCOP_B200_H
#define PARCOP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_plan pb_plan;

enum {
  PB_OK = 0,
  PB_ERR_ARG = -1,         /* bad argument / size mismatch */
  PB_ERR_UNSUPPORTED = -2, /* valid in the reference but outside this library's scope */
  PB_ERR_CUDA = -3,        /* CUDA runtime failure (no device, launch error, ...) */
  PB_ERR_STATE = -4        /* call out of order (e.g. mesh arrays before pb_plan_set_mesh) */
};

/* operator codes for pb_apply / pb_host_apply (one per f2py subroutine of parcop.f90) */
enum {
  PB_OP_DDX = 0, PB_OP_DDY = 1, PB_OP_DDZ = 2,          /* parcop.f90:225-253  ddx/ddy/ddz        */
  PB_OP_DD8X = 3, PB_OP_DD8Y = 4, PB_OP_DD8Z = 5,       /* parcop.f90:279-301  dd8x/dd8y/dd8z     */
  PB_OP_D2X = 6, PB_OP_D2Y = 7, PB_OP_D2Z = 8,          /* compact_operators.f90:130-206 d2x..z   */
  PB_OP_LAPLACIAN = 9,                                  /* parcop.f90:303-311  plaplacian         */
  PB_OP_RING = 10,                                      /* parcop.f90:313-322  pRing              */
  PB_OP_SFILTER = 11,                                   /* parcop.f90:337-346  sFilter            */
  PB_OP_GFILTER = 12,                                   /* parcop.f90:348-357  gFilter            */
  PB_OP_GFILTERX = 13, PB_OP_GFILTERY = 14, PB_OP_GFILTERZ = 15, /* parcop.f90:359-368 gFilterDir */
  PB_OP_SFILTERX = 16, PB_OP_SFILTERY = 17, PB_OP_SFILTERZ = 18, /* compact_operators.f90:385-497 */
  /* d1x/d1y/d1z(v, dv, bc = -1) (compact_operators.f90:13-50,52-89,91-128): the first derivative
   * of a field that is ODD across the axis' symmetry planes ("SYMM"); without a symmetry plane they
   * are ddx/ddy/ddz.  The reference reaches them through divV / divT (operators.f90:48-50,106-120). */
  PB_OP_DDX_ODD = 19, PB_OP_DDY_ODD = 20, PB_OP_DDZ_ODD = 21,
  /* parcop.f90:255-277 dd4x/dd4y/dd4z: explicit 4th derivative, no metric scale
   * (stencils.f90:430-513 e4d4, compact_operators.f90:208-278) */
  PB_OP_DD4X = 22, PB_OP_DD4Y = 23, PB_OP_DD4Z = 24,
  /* d8x/d8y/d8z(v, dv, bc = -1) (compact_operators.f90:280-383): the 8th derivative of a field that is
   * odd across the axis' symmetry planes; the reference reaches it inside ringV (operators.f90:661-671) */
  PB_OP_DD8X_ODD = 25, PB_OP_DD8Y_ODD = 26, PB_OP_DD8Z_ODD = 27,
  PB_OP_COUNT = 28
};

enum { PB_REDUCE_SUM = 0, PB_REDUCE_MAX = 1, PB_REDUCE_MIN = 2 };

const char *pb_last_error(void);
/* library / build info, e.g. "parcop_b200 0.1 sm_100a" */
const char *pb_version(void);

/* ---- setup: replaces parcop.setup (parcop.f90:23-61) -------------------------------------------
 * nx,ny,nz: GLOBAL sizes; px,py,pz: processor grid; cx,cy,cz: this rank's coordinates in it
 * (the reference derives them from MPI_CART_COORDS, comm.f90:150); x1..zn: node extents exactly as
 * Python passes them; b??: 4-character boundary strings "NONE" / "PERI" / "SYMM".
 * "SYMM" ends build the reference's symmetric / antisymmetric operator pair (compact.f90:77-91):
 * every one-field operator uses the even member, pb_divergence / pb_divergence_tensor pick the odd
 * first derivative for flux components normal to a symmetry plane (patch.f90:86-91 isymX/Y/Z).
 * Round-1 scope: px = py = 1 (z-slab), coordsys 0 (Cartesian) or 3 (curvilinear).
 * `device` is the CUDA ordinal (-1: current device). */
int pb_plan_create(pb_plan **plan, int nx, int ny, int nz, int px, int py, int pz, int cx, int cy,
                   int cz, int coordsys, double x1, double xn, double y1, double yn, double z1,
                   double zn, const char *bx1, const char *bxn, const char *by1, const char *byn,
                   const char *bz1, const char *bzn, int device);
int pb_plan_destroy(pb_plan *plan);

/* local extents of this rank (patch%ax/ay/az, patch.f90:80-82) and nominal spacings (:83-85) */
int pb_plan_extents(const pb_plan *plan, int *ax, int *ay, int *az);
int pb_plan_spacing(const pb_plan *plan, double *dx, double *dy, double *dz);

/* ---- mesh: replaces setup_mesh / setup_mesh_x3 (parcop.f90:64-81, mesh.f90:59-390) -------------
 * x,y,z: HOST arrays of the local coordinates (ax*ay*az) or all NULL for the uniform Cartesian
 * grid.  For coordsys 3 the Jacobian, inverse metrics, d1/d2/d3, CellVol and the filtered cell
 * volumes are computed on the device with this library's own operators. */
int pb_plan_set_mesh(pb_plan *plan, const double *x, const double *y, const double *z,
                     int periodic_grid);
/* replaces getVar / xGrid / dxGrid / mesh_getCellVol / mesh_getGridLen (parcop.f90:84-191).
 * name in: x y z d1 d2 d3 dAx dAy dAz dBx dBy dBz dCx dCy dCz dtJ CellVol GridLen.
 * pb_getvar_device returns a borrowed device pointer owned by the plan. */
int pb_getvar(const pb_plan *plan, const char *name, double *host_out);
int pb_getvar_device(const pb_plan *plan, const char *name, const double **dev_ptr);

/* ---- operators on device-resident fields ------------------------------------------------------ */
/* one-in / one-out operators, opcode PB_OP_* */
int pb_apply(pb_plan *plan, int opcode, const double *d_val, double *d_out, void *stream);
/* divergence (parcop.f90:202-211, operators.f90:38-53,77-91) */
int pb_divergence(pb_plan *plan, const double *d_fx, const double *d_fy, const double *d_fz,
                  double *d_out, void *stream);
/* gradS (parcop.f90:371-379, operators.f90:184-212) */
int pb_grads(pb_plan *plan, const double *d_val, double *d_gx, double *d_gy, double *d_gz,
             void *stream);
/* divergenceTensor (parcop.f90:213-223, operators.f90:97-123,176-179).  Cartesian: the divergence of
 * each COLUMN, dfx = ddx fxx + ddy fyx + ddz fzx ...; curvilinear (coordsys 3): pb_divergence of
 * each ROW, dfx = div(fxx, fxy, fxz) ... -- the reference's own (different) groupings. */
int pb_divergence_tensor(pb_plan *plan, const double *d_fxx, const double *d_fxy, const double *d_fxz,
                         const double *d_fyx, const double *d_fyy, const double *d_fyz,
                         const double *d_fzx, const double *d_fzy, const double *d_fzz,
                         double *d_dfx, double *d_dfy, double *d_dfz, void *stream);
/* pRingV (parcop.f90:324-333, operators.f90:645-699 with L = 1): max over directions of
 * max(|d8 vx|, |d8 vy|, |d8 vz|) * d, d = the spacing (Cartesian) or the mesh's d1 / d2 / d3 fields
 * (curvilinear, needs pb_plan_set_mesh) */
int pb_ring_vector(pb_plan *plan, const double *d_vx, const double *d_vy, const double *d_vz,
                   double *d_out, void *stream);

/* ---- RK4 stage update and reductions (pyranda.py:797-807, pyrandaMPI.py:307-326) --------------
 * PHI = dt*F + A*PHI ; U += B*PHI  for n points, fused (40 B/point). */
int pb_rk4_stage(pb_plan *plan, long n, double dt, double A, double B, const double *d_F,
                 double *d_PHI, double *d_U, void *stream);
/* local (this rank) reduction of n doubles; result written to *host_out after a stream sync */
int pb_reduce(pb_plan *plan, int kind, long n, const double *d_val, double *host_out, void *stream);
/* the same reduction with the result left in device memory (*d_out, one double): no stream sync, so
 * a time-step controller can stay on the device until the host really needs the number */
int pb_reduce_device(pb_plan *plan, int kind, long n, const double *d_val, double *d_out, void *stream);

/* ---- z-slab multi-GPU pieces (compact_d1.f90:719-746,858-928; compact_r4.f90:640-656,...) ------
 * A distributed z operator is: halo exchange (ncclSend/Recv by the caller, planes sent straight
 * from the field or packed with pb_z_pack_halo) -> pb_z_local (rhs with halos + local bounded
 * solve + scale / add-back; also writes the 4 interface unknowns per line) -> exchange of the
 * interface buffers by the caller (mpi_allgather in the reference; pb_z_exchange_ranks tells which
 * ranks' values actually matter, normally the two neighbours) -> pb_z_finish (reduced
 * block-tridiagonal solve + spike correction, applied only to the rows near the slab faces that
 * it can change).  Buffers are device pointers:
 *   send_lo/send_hi, recv_lo/recv_hi : 4*ax*ay doubles each (first / last planes, 3 or 4 used)
 *   iface_local : 4*ax*ay doubles,  iface_all : pz*4*ax*ay doubles (rank-major, as mpi_allgather)
 * zop in {PB_OP_DDZ, PB_OP_DD8Z, PB_OP_D2Z, PB_OP_SFILTERZ, PB_OP_GFILTERZ}. */
int pb_z_pack_halo(pb_plan *plan, int zop, const double *d_val, double *d_send_lo,
                   double *d_send_hi, void *stream);
int pb_z_local(pb_plan *plan, int zop, const double *d_val, const double *d_recv_lo,
               const double *d_recv_hi, double *d_out, double *d_iface_local, void *stream);
int pb_z_finish(pb_plan *plan, int zop, const double *d_val, const double *d_iface_all,
                double *d_out, void *stream);
/* The exchange itself, over NVLink peer memory (replaces MPI_Sendrecv / mpi_allgather of
 * compact_d1.f90:719-735,890): ONE launch that copies up to four local buffers into peer-mapped
 * destinations, publishes `epoch` in the flag word of each of up to two neighbours and waits until
 * the neighbours have published the same epoch in this rank's flag words.  All addresses are
 * device addresses valid on this GPU (peer memory mapped through CUDA IPC / symmetric memory);
 * `counter` is a zero-initialised 4-byte scratch word in local memory. */
int pb_peer_exchange(int ncopies, void *const *d_dst, const void *const *d_src, const size_t *bytes,
                     int npeers, void *const *d_remote_flags, void *const *d_local_flags,
                     unsigned long long epoch, void *d_counter, void *stream);
/* ---- the fused z-slab sweep (ring kernel) -------------------------------------------------------
 * The same distributed operator as ONE kernel per sweep and without a correction pass: the slabs
 * are consecutive chunks of the global line, factored once as a whole (compact_basetype.f90:150-198
 * builds a per-rank factorisation + reduced system instead), and the two-value states of the chunks
 * next to a slab face are written straight into the neighbouring ranks' memory (NVLink peer
 * mapping) as self-validating records which the consuming CTA polls, tile by tile.  The caller
 * provides, per rank, two record buffers of  slots * ax * ay * 32 bytes  (slots >= need_f / need_b,
 * zero-initialised, peer-visible) and the peer-mapped addresses of the buffers of the ranks it
 * writes to: en_out[k] = en_in of rank r+1+k, st_out[k] = st_in of rank r-1-k (periodic wrap).
 * `epoch` must be non-zero, identical on all ranks for one sweep and different from the previous
 * sweep's.  Halo planes are exchanged as for pb_z_local.  epi_mode: 0 store, 1 out += val,
 * 2 out = |val| s2, 3 out = max(out, |val| s2).
 * pb_z_ring_info reports the slots / hops this operator needs on this rank; all zeros means the
 * fused form is unavailable for it (explicit operator, slab not a multiple of 32 planes, states that
 * would wrap onto the rank itself) and pb_z_local / pb_z_finish must be used. */
typedef struct pb_xring {
  unsigned int epoch;
  void *en_in, *st_in;
  void *en_out[3], *st_out[3];
  /* push = 1: the kernel also moves the halo planes (no separate exchange): halo_dst[0] = the lower
   * neighbour's upper halo buffer, halo_dst[1] = the upper neighbour's lower halo buffer (NULL at a
   * physical end), flags / counter / epoch as for pb_peer_exchange */
  int push, npeers;
  void *halo_dst[2];
  void *flag_remote[2], *flag_local[2];
  unsigned long long halo_epoch;
  void *counter;
} pb_xring;
int pb_z_ring_info(pb_plan *plan, int zop, int *need_f, int *need_b, int *nup, int *ndn);
/* 0: no fused form; 1: fused, a CTA waits for the neighbours' states of its tile before it uses them
 * (states from more than one rank away); 2: fused, nobody waits before solving -- a chunk solves with
 * this rank's states, sends its own at once and adds the neighbours' contribution when it has landed */
int pb_z_ring_mode(pb_plan *plan, int zop);
int pb_z_ring(pb_plan *plan, int zop, const double *d_val, const double *d_recv_lo, const double *d_recv_hi,
              double *d_out, const pb_xring *x, int epi_mode, double s2, void *stream);
/* one directional derivative (PB_OP_DDX.., DD8X.., D2X.., the _ODD variants) with a composite epilogue
 * (epi_mode as above): the pieces pb_apply's laplacian / ring / divergence are made of */
int pb_apply_epi(pb_plan *plan, int opcode, const double *d_val, double *d_out, int epi_mode, double s2,
                 void *stream);
/* bit r of *mask is set when rank r's interface values enter this rank's correction (0: the
 * operator needs no exchange); slots of d_iface_all belonging to other ranks are never read */
int pb_z_exchange_ranks(pb_plan *plan, int zop, unsigned long long *mask);

/* ---- host-array convenience wrappers: the exact f2py call shapes ------------------------------
 * `dval = parcop.parcop.ddx(val)`  ==  pb_host_apply(plan, PB_OP_DDX, val, dval)
 * (H2D, operator, D2H inside the call; synchronous on return). */
int pb_host_apply(pb_plan *plan, int opcode, const double *h_val, double *h_out);
int pb_host_divergence(pb_plan *plan, const double *h_fx, const double *h_fy, const double *h_fz,
                       double *h_out);
int pb_host_grads(pb_plan *plan, const double *h_val, double *h_gx, double *h_gy, double *h_gz);
/* h_f9 = {fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz}, h_out3 = {dfx, dfy, dfz} */
int pb_host_divergence_tensor(pb_plan *plan, const double *const *h_f9, double *const *h_out3);
int pb_host_ring_vector(pb_plan *plan, const double *h_vx, const double *h_vy, const double *h_vz,
                        double *h_out);

/* ---- introspection for tests / benchmarks ----------------------------------------------------- */
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long pb_launch_count(void);
/* of those, launches of the TMA-pipelined persistent sweep kernels (tests assert the fast path ran) */
long pb_pipe_launch_count(void);
/* of those, launches of the ring kernel (thread-block clusters for long lines / cross-rank chunk states) */
long pb_ring_launch_count(void);
/* ring kernel policy: mode 0 never, 1 where the one-CTA pipelined kernel does not fit (default), 2 wherever
 * it fits; lines = 16 / 32 / 64 lines per tile (0: automatic); negative values keep the current setting */
int pb_set_ring(int mode, int lines);
/* tuning knobs (lines per tile, chunk length); 0 keeps the default.  Affects plans created later. */
int pb_set_tuning(int lines_yz, int lines_x, int chunk_len);

#ifdef __cplusplus
}
#endif
#endif /* PARCOP_B200_H */

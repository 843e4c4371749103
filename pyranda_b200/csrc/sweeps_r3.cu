// Instantiates the sweep kernels of one stencil family (see sweeps.cuh).
#include "sweeps.cuh"

namespace pb {
template cudaError_t launch_yz_f<F_R3, false>(int, const SweepDev &, const double *, double *, const double *, const double *, double *, const EpiArgs &, cudaStream_t);
template cudaError_t launch_x_f<F_R3, false>(int, const SweepDev &, const double *, double *, const EpiArgs &, cudaStream_t);
template cudaError_t launch_ring_f<F_R3, false>(int, const SweepDev &, const double *, double *, const double *, const double *, const XRing *, cudaStream_t);
}  // namespace pb

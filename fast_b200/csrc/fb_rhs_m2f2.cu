// explicit instantiation of the fused rhs kernels for MESH=2, FLOW=2 (see fb_rhs.cuh)
#include "fb_rhs.cuh"
template void launch_rhs_k<2, 2>(const ZoneDev &, RhsArgs &, bool, cudaStream_t);

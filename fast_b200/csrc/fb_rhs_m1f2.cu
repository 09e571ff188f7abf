// explicit instantiation of the fused rhs kernels for MESH=1, FLOW=2 (see fb_rhs.cuh)
#include "fb_rhs.cuh"
template void launch_rhs_k<1, 2>(const ZoneDev &, RhsArgs &, bool, cudaStream_t);

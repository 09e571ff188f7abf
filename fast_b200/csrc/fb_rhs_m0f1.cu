// explicit instantiation of the fused rhs kernels for MESH=0, FLOW=1 (see fb_rhs.cuh)
#include "fb_rhs.cuh"
template void launch_rhs_k<0, 1>(const ZoneDev &, RhsArgs &, bool, cudaStream_t);

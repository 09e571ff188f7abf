// explicit instantiation of the fused rhs kernels for MESH=0, FLOW=3 (see fb_rhs.cuh)
#include "fb_rhs.cuh"
template void launch_rhs_k<0, 3>(const ZoneDev &, RhsArgs &, bool, cudaStream_t);

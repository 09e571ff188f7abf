// fb_implicit.cu -- placeholder until the implicit chain lands (next commit): explicit zones never call these.
#include <cstdio>
#include "fb_implicit.cuh"
int  lu_plan_build(LuPlan &p, const ZoneDev &, const int box[6]) { for (int i = 0; i < 6; i++) p.box[i] = box[i]; return 0; }
void lu_plan_free(LuPlan &p) { if (p.scratch) cudaFree(p.scratch); p.scratch = nullptr; }
void launch_coe(const ZoneDev &, const double *, int, int, double *, cudaStream_t) {}
void launch_sa_ring(const ZoneDev &, const double *, cudaStream_t) {}
void launch_resid(const ZoneDev &, double *, int, cudaStream_t) {}
int  launch_lu(const ZoneDev &, const LuPlan &, const double *, double *, cudaStream_t) { return 0; }

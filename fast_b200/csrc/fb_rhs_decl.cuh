// fb_rhs_decl.cuh -- argument block of the fused rhs kernel + per-(MESH,FLOW) launcher declaration
#pragma once
#include "fb_common.cuh"
struct RhsArgs {
    const double *rop;      // rop_ssiter: flux input
    double *rop_out;        // explicit: stage output (ro or ro_P1)
    const double *ro_n;     // implicit: ro    (time level n)
    const double *ro_n1;    // implicit: ro_M1 (time level n-1)
    double coefH, coefW;    // low-storage RK (init_rhs.for:31-33, core3ark3.for:22-25)
    double c1, c2, c3;      // BDF weights (core3as2.for:37-46)
    int nitcfg;
    // halo-exchange overlap (fb_api.cu: gsdr3, explicit stages with cross-rank joins): with done != nullptr the CTAs take the
    // tiles that touch a zone face first (fb_rhs.cuh: rhs_face_first; their cells are the donors of the exchange) and the first
    // nbnd CTAs count their threads into *done when their results are in memory; the exchange stream waits for that count
    // (nsig = nbnd x threads per CTA, set by the launcher) while the interior tiles run
    int nbnd, nsig; int *done;
    int wfuse;              // SENSOR switch evaluated inside the first-sub-iteration kernel (default) instead of k_sensor
};

template <int MESH, int FLOW>
void launch_rhs_k(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st);

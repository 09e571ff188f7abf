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
    // halo-exchange overlap (fb_api.cu: gsdr3, explicit stages with cross-rank joins): part 1 / 2 = only the tiles that touch
    // a zone face / only the interior tiles, as a 1-D grid whose CTA c works on tile rhs_face_first(tile0 + c) of the
    // gx x gy x gz tile grid (fb_rhs.cuh); part 0 = the whole zone as a 3-D grid (gx = 0)
    int part, tile0, gx, gy, gz;
    int faces;              // in: faces of the zone that carry cross-rank joins, bit 0..5 = imin, imax, jmin, jmax, kmin, kmax;
                            // in the kernel: tile layers of part 1 behind each face, two bits per face (fb_tiles.cuh: rhs_face_flags)
    int thin;               // bit d: the last tile of direction d is less than two cells thick
    int wfuse;              // SENSOR switch evaluated inside the first-sub-iteration kernel (default) instead of k_sensor
    int pdl;                // 1: this launch may start while the previous kernel of the stream (the rhs of ANOTHER zone) drains its
                            // last wave (programmatic dependent launch); it consumes nothing that kernel writes
};

template <int MESH, int FLOW>
void launch_rhs_k(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st);

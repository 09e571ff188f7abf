// fb_tiles.cuh -- face-first order of the rhs tile grid (halo-exchange overlap, see fb_rhs.cuh); shared by the kernels, their
// launchers and the C ABI's fb_debug_tile_order (CPU test of the bijection).
#pragma once
#include <cuda_runtime.h>

// fl: how many tile layers behind each face of the zone belong to part 1 (the tiles whose cells are donors of cross-rank
// joins), two bits per face in the order imin, imax, jmin, jmax, kmin, kmax (rhs_face_flags has resolved degenerate tile
// grids, so that lo + hi <= g in every direction).  A layer count of 2 occurs on an upper face whose last tile is a single
// cell thick: the donors are the TWO cell layers behind a face, so the second one lives in the penultimate tile.
__host__ __device__ __forceinline__ void rhs_face_first(unsigned c, unsigned gx, unsigned gy, unsigned gz, unsigned fl,
                                                        unsigned &x, unsigned &y, unsigned &z)
{
    const unsigned xl = fl & 3u, xh = (fl >> 2) & 3u, yl = (fl >> 4) & 3u, yh = (fl >> 6) & 3u, zl = (fl >> 8) & 3u, zh = (fl >> 10) & 3u;
    const unsigned sxy = gx * gy;
    unsigned n = zl * sxy;
    if (c < n) { z = c / sxy; c -= z * sxy; y = c / gx; x = c - y * gx; return; }                            // slabs z < zl
    c -= n; n = zh * sxy;
    if (c < n) { const unsigned q = c / sxy; z = gz - zh + q; c -= q * sxy; y = c / gx; x = c - y * gx; return; }   // slabs z >= gz-zh
    c -= n;
    const unsigned nzm = gz - zl - zh, sxz = gx * nzm;                                       // layers z = zl .. zl+nzm-1
    n = yl * sxz;
    if (c < n) { y = c / sxz; c -= y * sxz; z = zl + c / gx; x = c % gx; return; }                             // rows y < yl
    c -= n; n = yh * sxz;
    if (c < n) { const unsigned q = c / sxz; y = gy - yh + q; c -= q * sxz; z = zl + c / gx; x = c % gx; return; }  // rows y >= gy-yh
    c -= n;
    const unsigned nym = gy - yl - yh, syz = nym * nzm;                                      // y = yl .. yl+nym-1
    n = xl * syz;
    if (c < n) { x = c / syz; c -= x * syz; z = zl + c / nym; y = yl + c % nym; return; }                      // columns x < xl
    c -= n; n = xh * syz;
    if (c < n) { const unsigned q = c / syz; x = gx - xh + q; c -= q * syz; z = zl + c / nym; y = yl + c % nym; return; }   // columns x >= gx-xh
    c -= n;
    const unsigned nxm = gx - xl - xh;                                                       // interior
    x = xl + c % nxm; c /= nxm; y = yl + c % nym; z = zl + c / nym;
}
// layer counts usable on a gx x gy x gz tile grid.  faces: bit 0..5 = imin .. kmax faces that carry cross-rank joins; thin: bit d
// = the last tile of direction d is less than two cells thick (its face then needs the penultimate tile layer as well)
static inline unsigned rhs_face_flags(unsigned faces, unsigned thin, unsigned gx, unsigned gy, unsigned gz)
{
    const unsigned g[3] = {gx, gy, gz};
    unsigned fl = 0;
    for (int d = 0; d < 3; d++) {
        unsigned lo = (faces >> (2 * d)) & 1u, hi = (faces >> (2 * d + 1)) & 1u;
        if (hi && ((thin >> d) & 1u)) hi = 2;
        if (lo + hi > g[d]) hi = g[d] - lo;
        fl |= (lo | (hi << 2)) << (4 * d);
    }
    return fl;
}
// number of face tiles (host)
static inline unsigned rhs_face_tiles(unsigned gx, unsigned gy, unsigned gz, unsigned fl)
{
    const unsigned ix = gx - (fl & 3u) - ((fl >> 2) & 3u), iy = gy - ((fl >> 4) & 3u) - ((fl >> 6) & 3u), iz = gz - ((fl >> 8) & 3u) - ((fl >> 10) & 3u);
    return gx * gy * gz - ix * iy * iz;
}

// fb_tiles.cuh -- face-first order of the rhs tile grid (halo-exchange overlap, see fb_rhs.cuh); shared by the kernels, their
// launchers and the C ABI's fb_debug_tile_order (CPU test of the bijection).
#pragma once
#include <cuda_runtime.h>

// fl: which faces of the zone carry cross-rank joins, bit 0..5 = imin, imax, jmin, jmax, kmin, kmax (rhs_face_flags has
// resolved degenerate tile grids, so that lo + hi <= g in every direction)
__host__ __device__ __forceinline__ void rhs_face_first(unsigned c, unsigned gx, unsigned gy, unsigned gz, unsigned fl,
                                                        unsigned &x, unsigned &y, unsigned &z)
{
    const unsigned xl = fl & 1u, xh = (fl >> 1) & 1u, yl = (fl >> 2) & 1u, yh = (fl >> 3) & 1u, zl = (fl >> 4) & 1u, zh = (fl >> 5) & 1u;
    const unsigned sxy = gx * gy;
    unsigned n = zl * sxy;
    if (c < n) { z = 0; y = c / gx; x = c - y * gx; return; }                                // slab z = 0
    c -= n; n = zh * sxy;
    if (c < n) { z = gz - 1; y = c / gx; x = c - y * gx; return; }                           // slab z = gz-1
    c -= n;
    const unsigned nzm = gz - zl - zh, sxz = gx * nzm;                                       // layers z = zl .. zl+nzm-1
    n = yl * sxz;
    if (c < n) { y = 0; z = zl + c / gx; x = c % gx; return; }                               // rows y = 0
    c -= n; n = yh * sxz;
    if (c < n) { y = gy - 1; z = zl + c / gx; x = c % gx; return; }                          // rows y = gy-1
    c -= n;
    const unsigned nym = gy - yl - yh, syz = nym * nzm;                                      // y = yl .. yl+nym-1
    n = xl * syz;
    if (c < n) { x = 0; z = zl + c / nym; y = yl + c % nym; return; }                        // columns x = 0
    c -= n; n = xh * syz;
    if (c < n) { x = gx - 1; z = zl + c / nym; y = yl + c % nym; return; }                   // columns x = gx-1
    c -= n;
    const unsigned nxm = gx - xl - xh;                                                       // interior
    x = xl + c % nxm; c /= nxm; y = yl + c % nym; z = zl + c / nym;
}
// face flags usable on a gx x gy x gz tile grid: a direction with a single tile keeps one flag at most
static inline unsigned rhs_face_flags(unsigned fl, unsigned gx, unsigned gy, unsigned gz)
{
    const unsigned g[3] = {gx, gy, gz};
    for (int d = 0; d < 3; d++) if (g[d] < 2 && ((fl >> (2 * d)) & 3u) == 3u) fl &= ~(2u << (2 * d));
    return fl;
}
// number of face tiles (host)
static inline unsigned rhs_face_tiles(unsigned gx, unsigned gy, unsigned gz, unsigned fl)
{
    const unsigned ix = gx - (fl & 1u) - ((fl >> 1) & 1u), iy = gy - ((fl >> 2) & 1u) - ((fl >> 3) & 1u), iz = gz - ((fl >> 4) & 1u) - ((fl >> 5) & 1u);
    return gx * gy * gz - ix * iy * iz;
}

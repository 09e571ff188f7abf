// fb_bc.cuh -- ghost-cell fill kernels (included by fb_kernels.cu).
//
// Reference: S/BC/BCzone.cpp:60-865 dispatches one bvbs_* routine per BC window; every routine walks the window's
// (j,k) lines, the ghost row next to the interior first ("first rank"), deeper rows after.  Here ONE launch per zone
// covers all windows: blockIdx.y = window, one thread per tangential cell, the thread walks its ghost line.
// In scope: 0 extrapolate, 1 farfield, 3 wall inviscid, 4/6 wall viscous adiabatic, 11 autoperiodic (22 overlap = no-op).
#pragma once

struct BcGeo { int d, low, ir, iref, gface, g0, g1, first, t1, t2; };

__device__ __forceinline__ int addr3(const ZoneDev &z, const BcGeo &g, int n, int a, int b)
{
    int ijk[3]; ijk[g.d] = n; ijk[g.t1] = a; ijk[g.t2] = b;
    return inddm(z, ijk[0], ijk[1], ijk[2]);
}
// face normal as the bvbs routines read it (C/Compute/shape_tab_mtr.for)
__device__ __forceinline__ void bc_normal(const ZoneDev &z, const BcGeo &g, int a, int b, double tc[3])
{
    int ijk[3]; ijk[g.d] = g.gface; ijk[g.t1] = a; ijk[g.t2] = b;
    const int lmtr = z.typ == 2 ? 0 : indmtr(z, ijk[0], ijk[1], ijk[2]);
    tc[0] = tc[1] = tc[2] = 0.;
    const double *T = g.d == 0 ? z.ti : (g.d == 1 ? z.tj : z.tk);
    if (z.typ == 0) { tc[0] = T[lmtr]; tc[1] = T[lmtr + z.nm]; tc[2] = T[lmtr + 2 * z.nm]; }
    else if (z.typ == 1 && g.d < 2) { tc[0] = T[lmtr]; tc[1] = T[lmtr + z.nm]; }
    else tc[g.d] = T[lmtr];
}

// S/BC/bvbs_farfield.for, non_reflection.for, BCFarfield_firstrank[_SA].for, BC_nextrank[SA].for
__device__ void bc_farfield_line(const ZoneDev &z, double *__restrict__ rop, const BcGeo &g, int a, int b)
{
    const int nd = z.nd, neq = z.neq;
    const double snorm = g.low ? -1. : 1.;
    const double gam1 = z.gam - 1., gam1_1 = 1. / gam1, cvinv = 1. / z.cv;
    const double roe = z.roinf, rue = z.vinf[0] * z.roinf, rve = z.vinf[1] * z.roinf, rwe = z.vinf[2] * z.roinf;
    const double ete = z.pinf / (z.gam - 1.) + 0.5 * (z.vinf[0] * rue + z.vinf[1] * rve + z.vinf[2] * rwe);
    const double nue = neq == 6 ? z.ronutildeinf / roe : 0.;
    double tc[3]; bc_normal(z, g, a, b, tc);
    const double tcx = tc[0] * snorm, tcy = tc[1] * snorm, tcz = tc[2] * snorm;
    {
        const int l = addr3(z, g, g.first, a, b), l1 = addr3(z, g, g.ir, a, b);
        const double sn = sqrt(tcx * tcx + tcy * tcy + tcz * tcz);
        const double roi = rop[l1], rui = rop[l1 + nd] * roi, rvi = rop[l1 + 2 * nd] * roi, rwi = rop[l1 + 3 * nd] * roi;
        const double eti = rop[l1 + 4 * nd] * z.cv * roi + .5 * (rui * rui + rvi * rvi + rwi * rwi) / roi;
        const double qref1 = .5 * (roe + roi), qref2 = .5 * (rue + rui), qref3 = .5 * (rve + rvi),
                     qref4 = .5 * (rwe + rwi), qref5 = .5 * (ete + eti);
        const double r = qref1, ri = 1. / r;
        double u = qref2 * ri, v = qref3 * ri, w = qref4 * ri;
        const double p = gam1 * (qref5 - .5 * r * (u * u + v * v + w * w));
        const double c = sqrt(z.gam * p * ri), ci = 1. / c, s_1 = 1. / sn;
        const double qn = tcx * u + tcy * v + tcz * w - 0.;
        double tnx, tny;
        if (sn == 0.) { tnx = 1.; tny = 0.; } else { tnx = tcx / sn; tny = tcy / sn; }
        const double eigv1 = qn, eigv4 = qn + sn * c, eigv5 = qn - sn * c;
        double ex[5], in[5];
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const double q1 = side ? roi : roe, q2 = side ? rui : rue, q3 = side ? rvi : rve, q4 = side ? rwi : rwe,
                         q5 = side ? eti : ete;
            const double s1 = q1, s2 = (q2 - u * q1) * ri, s3 = (q3 - v * q1) * ri, s4 = (q4 - w * q1) * ri;
            const double s5 = gam1 * (.5 * (u * u + v * v + w * w) * q1 - u * q2 - v * q3 - w * q4 + q5);
            double *o = side ? in : ex;
            o[0] = s1 - ci * ci * s5;
            o[1] = tcz * s_1 * (tnx * s2 + tny * s3) - s4;
            o[2] = -tny * s2 + tnx * s3;
            o[3] = .5 * s_1 * (tcx * s2 + tcy * s3 + tcz * s4) + .5 * ri * ci * s5;
            o[4] = -.5 * s_1 * (tcx * s2 + tcy * s3 + tcz * s4) + .5 * ri * ci * s5;
        }
        const double rvar1 = eigv1 < 0. ? ex[0] : in[0], rvar2 = eigv1 < 0. ? ex[1] : in[1], rvar3 = eigv1 < 0. ? ex[2] : in[2],
                     rvar4 = eigv4 < 0. ? ex[3] : in[3], rvar5 = eigv5 < 0. ? ex[4] : in[4];
        const double svar1 = rvar1 + r * ci * (rvar4 + rvar5);
        const double svar2 = tcz * s_1 * tnx * rvar2 - tny * rvar3 + tcx * s_1 * (rvar4 - rvar5);
        const double svar3 = tcz * s_1 * tny * rvar2 + tnx * rvar3 + tcy * s_1 * (rvar4 - rvar5);
        const double svar4 = -rvar2 + tcz * s_1 * (rvar4 - rvar5);
        const double svar5 = r * c * (rvar4 + rvar5);
        const double qvar5 = .5 * (u * u + v * v + w * w) * svar1 + r * (u * svar2 + v * svar3 + w * svar4) + gam1_1 * svar5;
        const double roinv = 1. / svar1;
        u = u + r * svar2 * roinv; v = v + r * svar3 * roinv; w = w + r * svar4 * roinv;
        rop[l] = svar1; rop[l + nd] = u; rop[l + 2 * nd] = v; rop[l + 3 * nd] = w;
        rop[l + 4 * nd] = (qvar5 * roinv - .5 * (u * u + v * v + w * w)) * cvinv;
        if (neq == 6) rop[l + 5 * nd] = eigv1 < 0. ? nue : rop[l1 + 5 * nd];
    }
    const double c4 = 5. / 6., c5 = 2. / 6., c6 = -1. / 6.;
    const double c0 = 1. / c6, c1 = -(c4 + c5) * c0, c2 = -c6 * c0, c3 = (2. - c5 - c4) * c0;
    const int incv = g.d == 0 ? 1 : (g.d == 1 ? z.incj : z.inck);
    const int inci = g.low ? incv : -incv;
    const int step = g.low ? -1 : 1;
    for (int n = g.first + step; n >= g.g0 && n <= g.g1; n += step) {
        const int l = addr3(z, g, n, a, b);
        const int l0 = l + inci, l1 = l + 2 * inci, l2 = l + 3 * inci;
        const double r0 = rop[l0];
        const double rl = r0 * c3 + rop[l1] * c1 + rop[l2] * c2;
        const bool bad = neq == 6 ? (rl <= 0.3 * r0) : (fabs(rl / r0 - 1.) >= 0.7);
        rop[l] = bad ? r0 : rl;
        for (int e = 1; e < neq; e++)
            rop[l + e * nd] = bad ? rop[l0 + e * nd] : rop[l0 + e * nd] * c3 + rop[l1 + e * nd] * c1 + rop[l2 + e * nd] * c2;
    }
}

__global__ void k_bc(ZoneDev z, const BcWin *__restrict__ wins, double *__restrict__ rop, int lrhs)
{
    const BcWin W = wins[blockIdx.y];
    BcGeo g;
    g.d = (W.idir - 1) / 2; g.low = W.idir % 2;
    g.g0 = W.w[2 * g.d]; g.g1 = W.w[2 * g.d + 1];
    if (g.low) { g.first = g.g1; g.ir = g.g1 + 1; g.iref = 2 * g.g1 + 1; g.gface = g.g1 + 1; }
    else { g.first = g.g0; g.ir = g.g0 - 1; g.iref = 2 * g.g0 - 1; g.gface = g.g0; }
    g.t1 = g.d == 0 ? 1 : 0; g.t2 = g.d == 2 ? 1 : 2;
    const int na = W.w[2 * g.t1 + 1] - W.w[2 * g.t1] + 1, nb = W.w[2 * g.t2 + 1] - W.w[2 * g.t2] + 1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (na <= 0 || nb <= 0 || t >= na * nb) return;
    const int a = W.w[2 * g.t1] + t % na, b = W.w[2 * g.t2] + t / na;
    const int nd = z.nd, neq = z.neq;
    switch (W.type) {
    case 0: {   // bvbs_extrapolate.for + BCExtrapolate[SA].for (lrhs: zero)
        const double c1 = lrhs == 1 ? 0. : 1.;
        const int l1 = addr3(z, g, g.ir, a, b);
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr3(z, g, n, a, b);
            for (int e = 0; e < neq; e++) rop[l + e * nd] = fmax(-1.e30, rop[l1 + e * nd]) * c1;
        }
    } break;
    case 11: {  // bvbs_periodique.for + BCPeriodic.for
        const double c1 = lrhs == 1 ? 0. : 1.;
        const int n_d = g.d == 0 ? z.iv : (g.d == 1 ? z.jv : z.kv);
        const int shift = g.low ? n_d : -n_d;
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr3(z, g, n, a, b), l1 = addr3(z, g, n + shift, a, b);
            for (int e = 0; e < neq; e++) rop[l + e * nd] = rop[l1 + e * nd] * c1;
        }
    } break;
    case 3: {   // bvbs_wall_inviscid.for + BCWallInviscid.for
        double tc[3]; bc_normal(z, g, a, b, tc);
        double surf = sqrt(tc[0] * tc[0] + tc[1] * tc[1] + tc[2] * tc[2]);
        surf = fmax(surf, 1e-30);
        const double s_1 = 1. / surf;
        const double tcx = tc[0] * s_1, tcy = tc[1] * s_1, tcz = tc[2] * s_1;
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr3(z, g, n, a, b), ldjr = addr3(z, g, g.iref - n, a, b);
            const double u = rop[ldjr + nd], v = rop[ldjr + 2 * nd], w = rop[ldjr + 3 * nd];
            const double qn = (u * tcx + v * tcy + w * tcz);
            const double ut = u - qn * tcx, vt = v - qn * tcy, wt = w - qn * tcz;
            rop[l] = rop[ldjr];
            rop[l + nd] = 2. * ut - u; rop[l + 2 * nd] = 2. * vt - v; rop[l + 3 * nd] = 2. * wt - w;
            rop[l + 4 * nd] = rop[ldjr + 4 * nd];
            if (neq == 6) rop[l + 5 * nd] = rop[ldjr + 5 * nd];
        }
    } break;
    case 4: case 6: {   // bvbs_wall_viscous_adia.for + BCWallViscous.for (+SA: nutilde' = -nutilde)
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr3(z, g, n, a, b), ldjr = addr3(z, g, g.iref - n, a, b);
            rop[l] = rop[ldjr];
            rop[l + nd] = 0. - rop[ldjr + nd]; rop[l + 2 * nd] = 0. - rop[ldjr + 2 * nd]; rop[l + 3 * nd] = 0. - rop[ldjr + 3 * nd];
            rop[l + 4 * nd] = rop[ldjr + 4 * nd];
            if (neq == 6) rop[l + 5 * nd] = -rop[ldjr + 5 * nd];
        }
    } break;
    case 1: bc_farfield_line(z, rop, g, a, b); break;
    default: break;   // 22 (overlap): ghost cells come from the transfers
    }
}

void launch_bc(const ZoneDev &z, const BcWin *d_wins, int nwin, int maxcells, double *rop, int lrhs, cudaStream_t st)
{
    if (nwin <= 0 || maxcells <= 0) return;
    dim3 b(128, 1, 1), g((maxcells + 127) / 128, nwin, 1);
    k_bc<<<g, b, 0, st>>>(z, d_wins, rop, lrhs);
}

// S/BC/correct_coins.for: edge ghost cells = mean of the two face-ghost neighbours, corner ghost cells = mean of three.
// One thread per edge/corner ghost cell; the sources are face ghosts only, so every cell is independent.
__global__ void k_corners(ZoneDev z, double *__restrict__ rop, int total)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int nf = z.nf, nfk = z.nfk, iv = z.iv, jv = z.jv, kv = z.kv;
    // region sizes: 4 i-j edges (nf*nf*kv), 4 i-k edges (nf*nfk*jv), 4 j-k edges (nf*nfk*iv), 8 corners (nf*nf*nfk)
    const int s_ij = nf * nf * kv, s_ik = nf * nfk * jv, s_jk = nf * nfk * iv, s_c = nf * nf * nfk;
    int r = t, gi = 0, gj = 0, gk = 0, si = 0, sj = 0, sk = 0, i = 0, j = 0, k = 0;
    if (r < 4 * s_ij) {
        const int q = r / s_ij; r %= s_ij; si = (q & 1) ? 1 : -1; sj = (q & 2) ? 1 : -1;
        gi = 1 + r % nf; gj = 1 + (r / nf) % nf; k = 1 + r / (nf * nf);
    } else if ((r -= 4 * s_ij) < 4 * s_ik) {
        const int q = r / s_ik; r %= s_ik; si = (q & 1) ? 1 : -1; sk = (q & 2) ? 1 : -1;
        gi = 1 + r % nf; gk = 1 + (r / nf) % nfk; j = 1 + r / (nf * nfk);
    } else if ((r -= 4 * s_ik) < 4 * s_jk) {
        const int q = r / s_jk; r %= s_jk; sj = (q & 1) ? 1 : -1; sk = (q & 2) ? 1 : -1;
        i = 1 + r % iv; r /= iv; gj = 1 + r % nf; gk = 1 + r / nf;
    } else {
        r -= 4 * s_jk;
        const int q = r / s_c; r %= s_c; si = (q & 1) ? 1 : -1; sj = (q & 2) ? 1 : -1; sk = (q & 4) ? 1 : -1;
        gi = 1 + r % nf; gj = 1 + (r / nf) % nf; gk = 1 + r / (nf * nf);
    }
    const int I0 = si > 0 ? iv : 1, J0 = sj > 0 ? jv : 1, K0 = sk > 0 ? kv : 1;
    const int i1 = min(nf, gi + 1), j1 = min(nf, gj + 1), k1 = min(nfk, gk + 1);
    const int nd = z.nd;
    int l, la, lb, lc = -1;
    if (si && sj && !sk) {
        l = inddm(z, I0 + gi * si, J0 + gj * sj, k); la = inddm(z, I0 + i1 * si, J0, k); lb = inddm(z, I0, J0 + j1 * sj, k);
    } else if (si && sk && !sj) {
        l = inddm(z, I0 + gi * si, j, K0 + gk * sk); la = inddm(z, I0, j, K0 + k1 * sk); lb = inddm(z, I0 + i1 * si, j, K0);
    } else if (sj && sk && !si) {
        l = inddm(z, i, J0 + gj * sj, K0 + gk * sk); la = inddm(z, i, J0, K0 + k1 * sk); lb = inddm(z, i, J0 + j1 * sj, K0);
    } else {
        l = inddm(z, I0 + gi * si, J0 + gj * sj, K0 + gk * sk);
        la = inddm(z, I0, J0, K0 + k1 * sk); lb = inddm(z, I0 + i1 * si, J0, K0); lc = inddm(z, I0, J0 + j1 * sj, K0);
    }
    const double c3 = 1. / 3.;
    for (int e = 0; e < z.neq; e++) {
        if (lc < 0) rop[l + e * nd] = (rop[la + e * nd] + rop[lb + e * nd]) * 0.5;
        else rop[l + e * nd] = (rop[la + e * nd] + rop[lb + e * nd] + rop[lc + e * nd]) * c3;
    }
}

void launch_corners(const ZoneDev &z, double *rop, const int box[6], cudaStream_t st)
{
    (void)box;   // correct_coins clamps its box to the interior (correct_coins.for:17-22)
    const int nf = z.nf, nfk = z.nfk;
    const int total = 4 * nf * nf * z.kv + 4 * nf * nfk * z.jv + 4 * nf * nfk * z.iv + 8 * nf * nf * nfk;
    if (total <= 0) return;
    k_corners<<<(total + 127) / 128, 128, 0, st>>>(z, rop, total);
}

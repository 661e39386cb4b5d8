// fp64 small linear algebra for the registration kernels: SE3 action, SO3 exp/update,
// 5x4 homogeneous plane fit, 6x6 Gauss-Newton solve, 3x3 symmetric eigen-decomposition.
// Replaces the Eigen/Sophus calls the reference makes inside its per-point loop
// (icp_registration.cpp:68,84,185-198,287-289,364-366; math_utils.h:112-136;
//  ndt_registration.cpp:118-130,403,423-427,445-449).
#pragma once
#include "common.cuh"

namespace locreg {

// fp64 reciprocal / square roots for the per-point plane fit.  The IEEE-rounded device routines carry slow-path
// subroutine calls (denormals, exceptional inputs) that force every live register across them to be spilled; here
// the inputs are ordinary positive numbers, so the hardware seed (rcp/rsqrt.approx.ftz.f64, ~20 bits) plus two
// Newton steps - accurate to an ulp or two, no calls - is used instead.  Host builds use the plain operators.
LR_HD double lr_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}
LR_HD double lr_rsqrt(double x) {  // x > 0
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = fma(r, fma(-hx * r, r, 0.5), r);
    r = fma(r, fma(-hx * r, r, 0.5), r);
    return r;
#else
    return 1.0 / sqrt(x);
#endif
}
LR_HD double lr_sqrt(double x) {  // x >= 0
#if defined(__CUDA_ARCH__)
    if (!(x > 0.0)) return 0.0;
    const double r = lr_rsqrt(x);
    const double s = x * r;
    return fma(fma(-s, s, x), 0.5 * r, s);  // one correction step on the root itself
#else
    return sqrt(x);
#endif
}

// Pose as the kernels use it: unit quaternion + translation (Sophus::SE3d::data() order) and the
// rotation matrix (row-major) derived from it once per Gauss-Newton iteration.
struct Pose {
    double qx, qy, qz, qw, tx, ty, tz;
    double R[9];
};

// Eigen::Quaterniond::toRotationMatrix
LR_HD void pose_refresh_R(Pose& T) {
    const double tx = 2 * T.qx, ty = 2 * T.qy, tz = 2 * T.qz;
    const double twx = tx * T.qw, twy = ty * T.qw, twz = tz * T.qw;
    const double txx = tx * T.qx, txy = ty * T.qx, txz = tz * T.qx;
    const double tyy = ty * T.qy, tyz = tz * T.qy, tzz = tz * T.qz;
    T.R[0] = 1 - (tyy + tzz); T.R[1] = txy - twz;       T.R[2] = txz + twy;
    T.R[3] = txy + twz;       T.R[4] = 1 - (txx + tzz); T.R[5] = tyz - twx;
    T.R[6] = txz - twy;       T.R[7] = tyz + twx;       T.R[8] = 1 - (txx + tyy);
}
LR_HD void pose_load(Pose& T, const double* p7) {
    T.qx = p7[0]; T.qy = p7[1]; T.qz = p7[2]; T.qw = p7[3];
    T.tx = p7[4]; T.ty = p7[5]; T.tz = p7[6];
    pose_refresh_R(T);
}
LR_HD void pose_store(const Pose& T, double* p7) {
    p7[0] = T.qx; p7[1] = T.qy; p7[2] = T.qz; p7[3] = T.qw;
    p7[4] = T.tx; p7[5] = T.ty; p7[6] = T.tz;
}

// Sophus SE3d * Vec3d: p + w*uv + v x uv with uv = 2 (v x p), then + t.  Written without FMA
// contraction so the float32 cast of the result (the k-NN query, icp_registration.cpp:70,170)
// is the same on every platform.
LR_HD void pose_apply(const Pose& T, double px, double py, double pz, double& ox, double& oy, double& oz) {
    double ux = LR_DSUB(LR_DMUL(T.qy, pz), LR_DMUL(T.qz, py));
    double uy = LR_DSUB(LR_DMUL(T.qz, px), LR_DMUL(T.qx, pz));
    double uz = LR_DSUB(LR_DMUL(T.qx, py), LR_DMUL(T.qy, px));
    ux = LR_DADD(ux, ux); uy = LR_DADD(uy, uy); uz = LR_DADD(uz, uz);
    const double cx = LR_DSUB(LR_DMUL(T.qy, uz), LR_DMUL(T.qz, uy));
    const double cy = LR_DSUB(LR_DMUL(T.qz, ux), LR_DMUL(T.qx, uz));
    const double cz = LR_DSUB(LR_DMUL(T.qx, uy), LR_DMUL(T.qy, ux));
    ox = LR_DADD(LR_DADD(LR_DADD(px, LR_DMUL(T.qw, ux)), cx), T.tx);
    oy = LR_DADD(LR_DADD(LR_DADD(py, LR_DMUL(T.qw, uy)), cy), T.ty);
    oz = LR_DADD(LR_DADD(LR_DADD(pz, LR_DMUL(T.qw, uz)), cz), T.tz);
}

// pose.so3() = pose.so3() * SO3::exp(w); pose.translation() += dt  (quirk Q10: split update)
LR_HD void pose_update(Pose& T, const double* dx) {
    const double wx = dx[0], wy = dx[1], wz = dx[2];
    const double theta_sq = wx * wx + wy * wy + wz * wz;
    double imag, real;
    if (theta_sq < 1e-20) {
        const double theta_po4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        const double theta = sqrt(theta_sq);
        const double half = 0.5 * theta;
        imag = sin(half) / theta;
        real = cos(half);
    }
    const double bx = imag * wx, by = imag * wy, bz = imag * wz, bw = real;
    const double ax = T.qx, ay = T.qy, az = T.qz, aw = T.qw;
    const double nw = aw * bw - ax * bx - ay * by - az * bz;
    const double nx = aw * bx + ax * bw + ay * bz - az * by;
    const double ny = aw * by + ay * bw + az * bx - ax * bz;
    const double nz = aw * bz + az * bw + ax * by - ay * bx;
    const double n = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
    T.qx = nx / n; T.qy = ny / n; T.qz = nz / n; T.qw = nw / n;
    T.tx += dx[3]; T.ty += dx[4]; T.tz += dx[5];
    pose_refresh_R(T);
}

// math::FitPlane's SVD (math_utils.h:118-125): right singular vector of the smallest singular value
// of A = [x y z 1] (5 x 4), by one-sided (Hestenes) Jacobi kept entirely in registers.
// One-sided Jacobi works on A itself, so the 1e4-1e5 condition number of a far-from-origin
// neighbourhood is not squared.  Returns the unit 4-vector (n, d) (quirk Q5: ||n|| != 1).
// (Not inlined on the device: it is the rare fallback of plane_fit5_fast and must not set the kernels' register budget.)
static LR_HD_NOINLINE void plane_svd5(const double (&P)[5][3], double (&coef)[4]) {
    double A[5][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 5; ++i) { A[i][0] = P[i][0]; A[i][1] = P[i][1]; A[i][2] = P[i][2]; A[i][3] = 1.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q];
                }
                if (gamma != 0.0 && fabs(gamma) > 1e-15 * sqrt(alpha * beta)) {
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const double ap = A[i][p], aq = A[i][q];
                        A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    double best = 0;
    int bj = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) s += A[i][j] * A[i][j];
        if (j == 0 || s < best) { best = s; bj = j; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) coef[i] = bj == 0 ? V[i][0] : (bj == 1 ? V[i][1] : (bj == 2 ? V[i][2] : V[i][3]));
}

// Plane fit without an SVD.  The smallest right singular vector x = (n, d) of A = [p 1] minimises
// ||A x||^2 = n^T S n + 5 (n.c + d)^2 subject to |n|^2 + d^2 = 1, with c the centroid of the five points and
// S = sum (p-c)(p-c)^T.  Eliminating d through its stationarity condition leaves the 3x3 problem
//     (S - lambda I - lambda kappa c c^T) n = 0,   kappa = 5 / (5 - lambda),   d = -kappa (c.n),
// and lambda = sigma_min^2 is the smallest root of the quartic (matrix determinant lemma, adj(S - lambda I) =
// adj(S) - lambda (tr(S) I - S) + lambda^2 I)
//     f(lambda) = lambda^4 - (5 + t + 5 a2) lambda^3 + (5 t + e + 5 a1) lambda^2 - (5 e + D + 5 a0) lambda + 5 D
// with t = tr S, e = sum of the principal 2x2 minors, D = det S, a0 = c^T adj(S) c, a1 = t |c|^2 - c^T S c,
// a2 = |c|^2: every coefficient is a sum of non-negative terms, so the centroid shift keeps the 1e4-1e5 condition
// number of the unshifted [p 1] out of the arithmetic.  f has four real non-negative roots (the squared singular
// values); Laguerre's iteration started at 0 increases monotonically to the SMALLEST one with cubic convergence,
// so - unlike a Rayleigh-quotient iteration - it cannot settle on a neighbouring singular value when two of them
// are close (common far from the origin, where quirk Q5 makes an in-plane direction compete with the normal).
// n is then the null vector of K = S - lambda I - lambda kappa c c^T, read off the adjugate column with the
// largest diagonal.  Returns false - the caller then runs plane_svd5 - only when the points are collinear /
// coincident (adj(K) vanishes) or a NaN got in.
// The five points enter one at a time (PlaneAcc::add) and are not kept: sums are taken relative to the first point
// (differences of float32 coordinates are exact in double and of the size of the neighbourhood), so one pass gives
// the centroid and S = sum d d^T - 5 dbar dbar^T without the cancellation a sum of raw p p^T would suffer, and the
// kernel does not have to hold 15 coordinates in registers across the solve.
struct PlaneAcc {
    double ox, oy, oz;                       // first point
    double sx, sy, sz;                       // sum of d = p - o
    double sxx, sxy, sxz, syy, syz, szz;     // sum of d d^T
    LR_HD void start(double x, double y, double z) {
        ox = x; oy = y; oz = z;
        sx = sy = sz = 0.0;
        sxx = sxy = sxz = syy = syz = szz = 0.0;
    }
    LR_HD void add(double x, double y, double z) {
        const double dx = x - ox, dy = y - oy, dz = z - oz;
        sx += dx; sy += dy; sz += dz;
        sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
};
LR_HD bool plane_fit5_solve(const PlaneAcc& a, double (&coef)[4]) {
    const double mx = a.sx * 0.2, my = a.sy * 0.2, mz = a.sz * 0.2;  // centroid relative to the first point
    const double cx = a.ox + mx, cy = a.oy + my, cz = a.oz + mz;
    const double sxx = a.sxx - a.sx * mx, sxy = a.sxy - a.sx * my, sxz = a.sxz - a.sx * mz;
    const double syy = a.syy - a.sy * my, syz = a.syz - a.sy * mz, szz = a.szz - a.sz * mz;
    // adj(S), the invariants of S and the three quadratic forms in c
    const double axx = syy * szz - syz * syz, axy = sxz * syz - sxy * szz, axz = sxy * syz - sxz * syy;
    const double ayy = sxx * szz - sxz * sxz, ayz = sxy * sxz - sxx * syz, azz = sxx * syy - sxy * sxy;
    const double t = sxx + syy + szz;
    const double e = axx + ayy + azz;
    double D = sxx * axx + sxy * axy + sxz * axz;
    if (D < 0.0) D = 0.0;  // rounding on an exactly planar set
    const double cc = cx * cx + cy * cy + cz * cz;
    const double cSc = cx * (sxx * cx + sxy * cy + sxz * cz) + cy * (sxy * cx + syy * cy + syz * cz) +
                       cz * (sxz * cx + syz * cy + szz * cz);
    double a0 = cx * (axx * cx + axy * cy + axz * cz) + cy * (axy * cx + ayy * cy + ayz * cz) +
                cz * (axz * cx + ayz * cy + azz * cz);
    if (a0 < 0.0) a0 = 0.0;
    const double a1 = t * cc - cSc;
    const double k3 = 5.0 + t + 5.0 * cc, k2 = 5.0 * t + e + 5.0 * a1, k1 = 5.0 * e + D + 5.0 * a0, k0 = 5.0 * D;
    // Laguerre from 0 (degree 4): monotone from below towards the smallest root
    double lam = 0.0;
    for (int it = 0; it < 12; ++it) {
        const double f = (((lam - k3) * lam + k2) * lam - k1) * lam + k0;
        if (!(f > 0.0)) break;  // on the root (or past it by rounding)
        const double f1 = ((4.0 * lam - 3.0 * k3) * lam + 2.0 * k2) * lam - k1;
        const double f2 = (12.0 * lam - 6.0 * k3) * lam + 2.0 * k2;
        // Laguerre step for degree n = 4, written without dividing by f:  -n f / (f' - sqrt((n-1)^2 f'^2 - n (n-1) f f''))
        double disc = 9.0 * f1 * f1 - 12.0 * f * f2;
        if (disc < 0.0) disc = 0.0;
        const double den = f1 - lr_sqrt(disc);  // f' < 0 left of the smallest root: the larger magnitude denominator
        const double step = -4.0 * f * lr_rcp(den);
        if (!(step > 0.0)) break;
        const double nl = lam + step;
        if (!(nl > lam)) break;              // step below one ulp
        const bool done = step <= 1e-15 * nl;
        lam = nl;
        if (done) break;
    }
    const double kappa = 5.0 * lr_rcp(5.0 - lam);
    const double mk = lam * kappa;
    const double kxx = sxx - lam - mk * cx * cx, kxy = sxy - mk * cx * cy, kxz = sxz - mk * cx * cz;
    const double kyy = syy - lam - mk * cy * cy, kyz = syz - mk * cy * cz, kzz = szz - lam - mk * cz * cz;
    const double bxx = kyy * kzz - kyz * kyz, bxy = kxz * kyz - kxy * kzz, bxz = kxy * kyz - kxz * kyy;
    const double byy = kxx * kzz - kxz * kxz, byz = kxy * kxz - kxx * kyz, bzz = kxx * kyy - kxy * kxy;
    double nx, ny, nz;
    if (bxx >= byy && bxx >= bzz) { nx = bxx; ny = bxy; nz = bxz; }
    else if (byy >= bzz) { nx = bxy; ny = byy; nz = byz; }
    else { nx = bxz; ny = byz; nz = bzz; }
    const double nn = nx * nx + ny * ny + nz * nz;
    // adj(K) = (product of K's two non-zero eigenvalues) n n^T: it vanishes when the null space is not
    // one-dimensional (collinear / coincident points, or a double smallest singular value)
    if (!(nn > 1e-30 * (t * t * t * t + 1e-300))) return false;
    const double d = -kappa * (cx * nx + cy * ny + cz * nz);
    const double inv = lr_rsqrt(nn + d * d);
    coef[0] = nx * inv; coef[1] = ny * inv; coef[2] = nz * inv; coef[3] = d * inv;
    return true;
}

// array flavour (tests/hostsim, readability)
LR_HD bool plane_fit5_fast(const double (&P)[5][3], double (&coef)[4]) {
    PlaneAcc a;
    a.start(P[0][0], P[0][1], P[0][2]);
#pragma unroll
    for (int i = 1; i < 5; ++i) a.add(P[i][0], P[i][1], P[i][2]);
    return plane_fit5_solve(a, coef);
}

// math::FitLine's direction (math_utils.h:138-152): right singular vector of the LARGEST singular value of
// Y = p - mean (5 x 3), i.e. the principal eigenvector of S = Y^T Y.  Unlike the plane fit's smallest singular vector
// this one is well conditioned in S (the squared spectrum only widens the gap that separates it), so it is taken
// from the symmetric 3x3 eigen-decomposition.  Its sign is arbitrary and cancels in J^T J and J^T e.
// S = {xx, xy, xz, yy, yz, zz} of the centred points.
LR_HD void sym3_eigen(const double* S, double* lam, double* Q);
LR_HD void line_dir_from_scatter(const double (&S)[6], double (&dir)[3]) {
    double lam[3], Q[9];
    sym3_eigen(S, lam, Q);  // eigenvalues sorted descending, eigenvectors in the columns of Q
    dir[0] = Q[0]; dir[1] = Q[3]; dir[2] = Q[6];
}

// Gauss-Newton step: solves H dx = b by partial-pivot LU (what Matrix6d::inverse()/determinant()
// do in Eigen for size 6).  Hu = packed upper triangle.  Returns false if det(H) == 0
// (icp_registration.cpp:100,210; ndt_registration.cpp:435).
LR_HD bool gn_solve6(const double* Hu, const double* b, double* dx) {
    double lu[6][6], y[6];
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) { lu[r][c] = Hu[hidx(r, c)]; lu[c][r] = lu[r][c]; }
    for (int i = 0; i < 6; ++i) y[i] = b[i];
    double det = 1.0;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(lu[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabs(lu[i][k]) > best) { best = fabs(lu[i][k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) { const double t = lu[k][j]; lu[k][j] = lu[piv][j]; lu[piv][j] = t; }
            const double t = y[k]; y[k] = y[piv]; y[piv] = t;
        }
        det *= lu[k][k];
        if (lu[k][k] != 0.0) {
            for (int i = k + 1; i < 6; ++i) {
                const double f = lu[i][k] / lu[k][k];
                for (int j = k + 1; j < 6; ++j) lu[i][j] -= f * lu[k][j];
                y[i] -= f * y[k];
            }
        }
    }
    if (det == 0.0 || !(det == det)) return false;
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int j = i + 1; j < 6; ++j) s -= lu[i][j] * dx[j];
        dx[i] = s / lu[i][i];
    }
    return true;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi: S = Q diag(lam) Q^T, lam sorted descending.
// For the symmetric PSD covariance this is what JacobiSVD(sigma) yields with U = V = Q
// (ndt_registration.cpp:118-119).  S = {xx, xy, xz, yy, yz, zz}; Q row-major.
LR_HD void sym3_eigen(const double* S, double* lam, double* Q) {
    double a[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-18 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    double d[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[order[j]] > d[order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    for (int j = 0; j < 3; ++j) {
        lam[j] = d[order[j]];
        for (int k = 0; k < 3; ++k) Q[k * 3 + j] = v[k][order[j]];
    }
}

}  // namespace locreg

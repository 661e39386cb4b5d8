// fp64 small linear algebra for the registration kernels: SE3 action, SO3 exp/update,
// 5x4 homogeneous plane fit, 6x6 Gauss-Newton solve, 3x3 symmetric eigen-decomposition.
// Replaces the Eigen/Sophus calls the reference makes inside its per-point loop
// (icp_registration.cpp:68,84,185-198,287-289,364-366; math_utils.h:112-136;
//  ndt_registration.cpp:118-130,403,423-427,445-449).
#pragma once
#include "common.cuh"

namespace locreg {

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

// Fast path of the plane fit.  The smallest right singular vector x = (n, d) of A = [p 1] minimises
// ||A x||^2 = n^T S n + 5 (n.c + d)^2 subject to |n|^2 + d^2 = 1, with c the centroid of the five points and
// S = sum (p-c)(p-c)^T.  Eliminating d through its stationarity condition leaves a 3x3 problem
//     S n = lambda (I + kappa c c^T) n,   kappa = 5 / (5 - lambda),   d = -kappa (c.n),
// whose smallest eigenpair is found by Rayleigh-quotient iteration (cubic convergence, 3-4 steps) started from
// the smallest eigenvector of S.  Each step solves (S - mu B) z = B n through the adjugate, which stays accurate
// when the matrix is (by design) nearly singular, and the centroid shift keeps every quantity at the scale of
// the neighbourhood, so the 1e4-1e5 condition number of the unshifted [p 1] is never squared.
// Returns false — the caller then runs plane_svd5 — if the points are collinear, the iteration has not
// settled, or it settled on an eigenpair that is not the smallest one (inertia check on S - mu B).
constexpr int kPlaneFitWarmup = 3;
LR_HD bool plane_fit5_fast(const double (&P)[5][3], double (&coef)[4]) {
    const double cx = (P[0][0] + P[1][0] + P[2][0] + P[3][0] + P[4][0]) * 0.2;
    const double cy = (P[0][1] + P[1][1] + P[2][1] + P[3][1] + P[4][1]) * 0.2;
    const double cz = (P[0][2] + P[1][2] + P[2][2] + P[3][2] + P[4][2]) * 0.2;
    double sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double x = P[i][0] - cx, y = P[i][1] - cy, z = P[i][2] - cz;
        sxx += x * x; sxy += x * y; sxz += x * z; syy += y * y; syz += y * z; szz += z * z;
    }
    // start: column of adj(S) with the largest diagonal = one inverse-iteration step on S
    double nx, ny, nz;
    {
        const double axx = syy * szz - syz * syz, axy = sxz * syz - sxy * szz, axz = sxy * syz - sxz * syy;
        const double ayy = sxx * szz - sxz * sxz, ayz = sxy * sxz - sxx * syz, azz = sxx * syy - sxy * sxy;
        if (axx >= ayy && axx >= azz) { nx = axx; ny = axy; nz = axz; }
        else if (ayy >= azz) { nx = axy; ny = ayy; nz = ayz; }
        else { nx = axz; ny = ayz; nz = azz; }
    }
    double nn = nx * nx + ny * ny + nz * nz;
    if (!(nn > 0.0)) return false;  // collinear / coincident points (or NaN)
    double inv = 1.0 / sqrt(nn);
    nx *= inv; ny *= inv; nz *= inv;
    // The weight c c^T (|c|^2 ~ 1e4 far from the origin, quirk Q5) can make the pencil's smallest eigenvector very
    // different from S's, so pull the start into the right basin with a few plain inverse iterations
    // n <- S^-1 B n (adjugate form, kappa = 1) before switching to Rayleigh-quotient shifts.
    {
        const double axx = syy * szz - syz * syz, axy = sxz * syz - sxy * szz, axz = sxy * syz - sxz * syy;
        const double ayy = sxx * szz - sxz * sxz, ayz = sxy * sxz - sxx * syz, azz = sxx * syy - sxy * sxy;
#pragma unroll
        for (int it = 0; it < kPlaneFitWarmup; ++it) {
            const double cn = cx * nx + cy * ny + cz * nz;
            const double bx = nx + cn * cx, by = ny + cn * cy, bz = nz + cn * cz;
            const double zx = axx * bx + axy * by + axz * bz, zy = axy * bx + ayy * by + ayz * bz,
                         zz = axz * bx + ayz * by + azz * bz;
            nn = zx * zx + zy * zy + zz * zz;
            if (!(nn > 0.0)) return false;
            inv = 1.0 / sqrt(nn);
            nx = zx * inv; ny = zy * inv; nz = zz * inv;
        }
    }
    double kappa = 1.0, mu = 0.0;
    bool settled = false;
    double e2 = 0.0, trk = 0.0;
    for (int it = 0; it < 8; ++it) {
        const double cn = cx * nx + cy * ny + cz * nz;
        const double bx = nx + kappa * cn * cx, by = ny + kappa * cn * cy, bz = nz + kappa * cn * cz;  // B n
        const double snx = sxx * nx + sxy * ny + sxz * nz, sny = sxy * nx + syy * ny + syz * nz,
                     snz = sxz * nx + syz * ny + szz * nz;
        mu = (nx * snx + ny * sny + nz * snz) / (nx * bx + ny * by + nz * bz);  // Rayleigh quotient
        kappa = 5.0 / (5.0 - mu);
        const double mk = mu * kappa;
        const double kxx = sxx - mu - mk * cx * cx, kxy = sxy - mk * cx * cy, kxz = sxz - mk * cx * cz;
        const double kyy = syy - mu - mk * cy * cy, kyz = syz - mk * cy * cz, kzz = szz - mu - mk * cz * cz;
        const double axx = kyy * kzz - kyz * kyz, axy = kxz * kyz - kxy * kzz, axz = kxy * kyz - kxz * kyy;
        const double ayy = kxx * kzz - kxz * kxz, ayz = kxy * kxz - kxx * kyz, azz = kxx * kyy - kxy * kxy;
        e2 = axx + ayy + azz;   // second elementary symmetric polynomial of eig(K)
        trk = kxx + kyy + kzz;
        double zx = axx * bx + axy * by + axz * bz;
        double zy = axy * bx + ayy * by + ayz * bz;
        double zz = axz * bx + ayz * by + azz * bz;
        nn = zx * zx + zy * zy + zz * zz;
        if (!(nn > 0.0)) return false;
        inv = 1.0 / sqrt(nn);
        if (zx * nx + zy * ny + zz * nz < 0) inv = -inv;
        zx *= inv; zy *= inv; zz *= inv;
        const double ch = fabs(zx - nx) + fabs(zy - ny) + fabs(zz - nz);
        nx = zx; ny = zy; nz = zz;
        if (ch < 1e-9) { settled = true; break; }  // cubic convergence: the step just taken is exact to rounding
    }
    // K = S - mu B must be positive semi-definite with a one-dimensional null space for mu to be the SMALLEST
    // eigenvalue (Sylvester): two positive eigenvalues <=> e2 > 0 and trace > 0.
    if (!settled || !(e2 > 0.0) || !(trk > 0.0)) return false;
    const double d = -kappa * (cx * nx + cy * ny + cz * nz);
    inv = 1.0 / sqrt(1.0 + d * d);
    coef[0] = nx * inv; coef[1] = ny * inv; coef[2] = nz * inv; coef[3] = d * inv;
    return true;
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

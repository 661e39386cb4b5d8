// ORACLE (test infrastructure, NOT product code).
//
// Small fixed-size fp64 linear algebra + SE3/SO3 restating the third-party
// arithmetic the reference's registration loop calls (Eigen 3.3.x, Sophus 1.x;
// neither is vendored under /root/reference):
//   * Sophus::SE3d * Vec3d            (icp_registration.cpp:68,113,169; ndt_registration.cpp:293,403)
//   * Sophus::SO3d::matrix / hat / exp (icp_registration.cpp:84,194,288,365; ndt_registration.cpp:423,448)
//   * Matrix<double,6,6>::inverse / determinant (icp_registration.cpp:100,210,287,364; ndt_registration.cpp:435,445)
//   * Eigen::JacobiSVD (math_utils.h:124; ndt_registration.cpp:118)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg may use anything under oracle/.  Parity status: UNPINNED by the reference
// (it has no golden vectors); pinned instead by analytic KATs and
// an independent numpy cross-check (tests/test_oracle_*.py).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Row-major 3x3.
struct Mat3 {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
};
inline Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline Vec3 mul(const Mat3& a, const Vec3& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
// Sophus::SO3d::hat
inline Mat3 hat(const Vec3& v) {
    Mat3 r;
    r.m[0][1] = -v.z; r.m[0][2] = v.y;
    r.m[1][0] = v.z;  r.m[1][2] = -v.x;
    r.m[2][0] = -v.y; r.m[2][1] = v.x;
    return r;
}

// Unit quaternion (x,y,z,w) + translation: the memory layout of Sophus::SE3d::data().
struct SE3 {
    double qx = 0, qy = 0, qz = 0, qw = 1;
    Vec3 t;
    static SE3 from7(const double* p) {
        SE3 T;
        T.qx = p[0]; T.qy = p[1]; T.qz = p[2]; T.qw = p[3];
        T.t = {p[4], p[5], p[6]};
        return T;
    }
    void to7(double* p) const {
        p[0] = qx; p[1] = qy; p[2] = qz; p[3] = qw;
        p[4] = t.x; p[5] = t.y; p[6] = t.z;
    }
    // Sophus SO3::operator*(Point): p + w*uv + vec x uv, uv = 2 * (vec x p).
    Vec3 rotate(const Vec3& p) const {
        Vec3 v{qx, qy, qz};
        Vec3 uv = cross(v, p);
        uv = uv + uv;
        return p + qw * uv + cross(v, uv);
    }
    Vec3 operator*(const Vec3& p) const { return rotate(p) + t; }
    // Eigen::Quaterniond::toRotationMatrix
    Mat3 matrix() const {
        Mat3 R;
        const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
        const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
        const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz;       R.m[0][2] = txz + twy;
        R.m[1][0] = txy + twz;       R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
        R.m[2][0] = txz - twy;       R.m[2][1] = tyz + twx;       R.m[2][2] = 1 - (txx + tyy);
        return R;
    }
    // pose.so3() = pose.so3() * SO3::exp(w)  (icp_registration.cpp:288,365; ndt_registration.cpp:448)
    void right_mul_exp(const Vec3& w) {
        const double theta_sq = dot(w, w);
        double imag, real;
        if (theta_sq < 1e-10 * 1e-10) {  // Sophus::Constants<double>::epsilon()^2
            const double theta_po4 = theta_sq * theta_sq;
            imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
            real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
        } else {
            const double theta = std::sqrt(theta_sq);
            const double half = 0.5 * theta;
            imag = std::sin(half) / theta;
            real = std::cos(half);
        }
        const double bx = imag * w.x, by = imag * w.y, bz = imag * w.z, bw = real;
        const double ax = qx, ay = qy, az = qz, aw = qw;
        double nw = aw * bw - ax * bx - ay * by - az * bz;
        double nx = aw * bx + ax * bw + ay * bz - az * by;
        double ny = aw * by + ay * bw + az * bx - ax * bz;
        double nz = aw * bz + az * bw + ax * by - ay * bx;
        const double n = std::sqrt(nx * nx + ny * ny + nz * nz + nw * nw);  // SO3 ctor normalises
        qx = nx / n; qy = ny / n; qz = nz / n; qw = nw / n;
    }
};

// Column-major 6x6 (Eigen default storage), H(r,c) = a[c*6+r].
struct Mat6 {
    double a[36];
    Mat6() { for (double& v : a) v = 0; }
    double& operator()(int r, int c) { return a[c * 6 + r]; }
    double operator()(int r, int c) const { return a[c * 6 + r]; }
};
struct Vec6 {
    double v[6] = {0, 0, 0, 0, 0, 0};
    double norm() const {
        double s = 0;
        for (double e : v) s += e * e;
        return std::sqrt(s);
    }
};

// Partial-pivot LU (Eigen::PartialPivLU semantics: row of max |.| in the column, first wins).
// Returns determinant; if inv != nullptr also writes the inverse.
inline double lu6(const Mat6& H, Mat6* inv) {
    double lu[6][6];
    int perm[6];
    for (int i = 0; i < 6; ++i) {
        perm[i] = i;
        for (int j = 0; j < 6; ++j) lu[i][j] = H(i, j);
    }
    double det = 1.0;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = std::fabs(lu[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (std::fabs(lu[i][k]) > best) { best = std::fabs(lu[i][k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) std::swap(lu[k][j], lu[piv][j]);
            std::swap(perm[k], perm[piv]);
            det = -det;
        }
        det *= lu[k][k];
        if (lu[k][k] != 0.0) {
            for (int i = k + 1; i < 6; ++i) {
                lu[i][k] /= lu[k][k];
                for (int j = k + 1; j < 6; ++j) lu[i][j] -= lu[i][k] * lu[k][j];
            }
        }
    }
    if (inv) {
        for (int c = 0; c < 6; ++c) {
            double y[6];
            for (int i = 0; i < 6; ++i) {  // forward: L y = P e_c
                double s = (perm[i] == c) ? 1.0 : 0.0;
                for (int j = 0; j < i; ++j) s -= lu[i][j] * y[j];
                y[i] = s;
            }
            for (int i = 5; i >= 0; --i) {  // backward: U x = y
                double s = y[i];
                for (int j = i + 1; j < 6; ++j) s -= lu[i][j] * (*inv)(j, c);
                (*inv)(i, c) = s / lu[i][i];
            }
        }
    }
    return det;
}
inline Vec6 mul(const Mat6& A, const Vec6& b) {
    Vec6 r;
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += A(i, j) * b.v[j];
        r.v[i] = s;
    }
    return r;
}

// One-sided (Hestenes) Jacobi SVD of a dense m x n matrix (m >= n), heap-allocated like the
// reference's Eigen::MatrixXd call sites.  Outputs singular values (descending) and V (n x n,
// column j = right singular vector of sigma[j]), and optionally U (m x n thin).
// A is row-major m x n.  Accuracy: columns orthogonal to ~1e-15 relative.
inline void jacobi_svd(std::vector<double> A, int m, int n, std::vector<double>& sigma,
                       std::vector<double>& V, std::vector<double>* U = nullptr) {
    V.assign(static_cast<size_t>(n) * n, 0.0);
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    const double tol = 1e-15;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < m; ++i) {
                    const double ap = A[i * n + p], aq = A[i * n + q];
                    alpha += ap * ap; beta += aq * aq; gamma += ap * aq;
                }
                if (gamma == 0.0 || std::fabs(gamma) <= tol * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; ++i) {
                    const double ap = A[i * n + p], aq = A[i * n + q];
                    A[i * n + p] = c * ap - s * aq;
                    A[i * n + q] = s * ap + c * aq;
                }
                for (int i = 0; i < n; ++i) {
                    const double vp = V[i * n + p], vq = V[i * n + q];
                    V[i * n + p] = c * vp - s * vq;
                    V[i * n + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    sigma.assign(n, 0.0);
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += A[i * n + j] * A[i * n + j];
        sigma[j] = std::sqrt(s);
    }
    // sort descending (selection sort, swapping columns of A and V)
    for (int j = 0; j < n - 1; ++j) {
        int big = j;
        for (int k = j + 1; k < n; ++k)
            if (sigma[k] > sigma[big]) big = k;
        if (big != j) {
            std::swap(sigma[j], sigma[big]);
            for (int i = 0; i < m; ++i) std::swap(A[i * n + j], A[i * n + big]);
            for (int i = 0; i < n; ++i) std::swap(V[i * n + j], V[i * n + big]);
        }
    }
    if (U) {
        U->assign(static_cast<size_t>(m) * n, 0.0);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) (*U)[i * n + j] = sigma[j] > 0 ? A[i * n + j] / sigma[j] : (m == n ? V[i * n + j] : 0.0);
    }
}

}  // namespace oracle

// Direct NDT on the device (kernels K5/K6 of SURVEY.md §2.3): voxel grid build bodies and the
// per-point body of AlignNdt.  Restates NdtRegistration::SetDirectNdtTargetCloud
// (ndt_registration.cpp:87-148) and the loop body of ::AlignNdt (ndt_registration.cpp:399-433).
//
// Layout in HBM: an open-addressing table of 16 B slots {packed voxel key, voxel id, point count}
// (one 16 B load per probe) and an array of 96 B voxel records {mu[3], info[9]} (three 32 B sectors),
// matching the algorithmic-bytes accounting of SURVEY.md §8d (7 x 16 B probes + h x 96 B per point).
#pragma once
#include "la.cuh"

namespace locreg {

struct __attribute__((aligned(16))) NdtSlot {
    unsigned long long key;  // packed (kx,ky,kz), kNdtEmpty if unused
    int vid;                 // index into voxels[], -1 while/if the voxel has too few points
    unsigned int count;      // points that fell into the voxel
};
struct __attribute__((aligned(32))) NdtVoxel {
    double mu[3];
    double info[9];  // row-major
};
static_assert(sizeof(NdtSlot) == 16 && sizeof(NdtVoxel) == 96, "NDT record sizes");

constexpr unsigned long long kNdtEmpty = ~0ull;
constexpr int kNdtBias = 1 << 20;
constexpr int kNdtKeyLimit = (1 << 20) - 2;

struct NdtMapView {
    const NdtSlot* slots;
    const NdtVoxel* voxels;
    unsigned int slot_mask;
    unsigned int n_voxels;
    double inv_voxel;  // 1.0 / voxel_size_, recomputed from voxel_size_ (ndt_registration.cpp:25)
};

struct NdtParams {
    double res_outlier_th;
    double eps;
    int max_iteration;
    int min_effective_pts;
    int min_pts_in_voxel;
    int n_nearby;  // 1 (CENTER) or 7 (NEARBY6)
};

// (pt * inv_voxel_size_).cast<int>(): C++ double -> int conversion truncates toward zero (quirk Q9).
LR_HD int ndt_trunc(double v) {
    if (!(v > -2147483000.0)) return -2147483000;
    if (!(v < 2147483000.0)) return 2147483000;
    return static_cast<int>(v);
}
LR_HD bool ndt_key_ok(int kx, int ky, int kz) {
    return kx >= -kNdtKeyLimit && kx <= kNdtKeyLimit && ky >= -kNdtKeyLimit && ky <= kNdtKeyLimit &&
           kz >= -kNdtKeyLimit && kz <= kNdtKeyLimit;
}
LR_HD unsigned long long ndt_pack(int kx, int ky, int kz) {
    return (static_cast<unsigned long long>(static_cast<unsigned int>(kx + kNdtBias)) << 42) |
           (static_cast<unsigned long long>(static_cast<unsigned int>(ky + kNdtBias)) << 21) |
           static_cast<unsigned long long>(static_cast<unsigned int>(kz + kNdtBias));
}
LR_HD void ndt_unpack(unsigned long long k, int& kx, int& ky, int& kz) {
    kx = static_cast<int>((k >> 42) & 0x1FFFFFu) - kNdtBias;
    ky = static_cast<int>((k >> 21) & 0x1FFFFFu) - kNdtBias;
    kz = static_cast<int>(k & 0x1FFFFFu) - kNdtBias;
}
LR_HD unsigned int ndt_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return static_cast<unsigned int>(k);
}

// NEARBY6 offsets in the reference's order (GenerateNearbyGrids, ndt_registration.cpp:57-58)
LR_HD void ndt_offset(int i, int& dx, int& dy, int& dz) {
    dx = (i == 1) ? -1 : (i == 2 ? 1 : 0);
    dy = (i == 3) ? 1 : (i == 4 ? -1 : 0);
    dz = (i == 5) ? -1 : (i == 6 ? 1 : 0);
}

// ---- build bodies -------------------------------------------------------------------------------
// Step 1: voxel key of map point i -> slot (inserted on first sight); slot.count++.
// counters: [0] voxels inserted, [1] overflow flag
template <class A>
LR_HD void ndt_insert_body(size_t i, const void* xyz, size_t stride, double inv_voxel, NdtSlot* slots,
                           unsigned int slot_mask, unsigned int* pt_slot, unsigned int* counters) {
    const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + i * stride);
    pt_slot[i] = 0xFFFFFFFFu;
    if (!finite3(p[0], p[1], p[2])) return;  // deviation D1
    const int kx = ndt_trunc(LR_DMUL(static_cast<double>(p[0]), inv_voxel));
    const int ky = ndt_trunc(LR_DMUL(static_cast<double>(p[1]), inv_voxel));
    const int kz = ndt_trunc(LR_DMUL(static_cast<double>(p[2]), inv_voxel));
    if (!ndt_key_ok(kx, ky, kz)) return;
    const unsigned long long key = ndt_pack(kx, ky, kz);
    unsigned int h = ndt_hash(key) & slot_mask;
    unsigned int probes = 0;
    while (true) {
        const unsigned long long k = A::cas64(&slots[h].key, kNdtEmpty, key);
        if (k == kNdtEmpty) { A::add32(&counters[0], 1u); break; }
        if (k == key) break;
        h = (h + 1) & slot_mask;
        if (++probes > slot_mask) { counters[1] = 1u; return; }
    }
    A::add32(&slots[h].count, 1u);
    pt_slot[i] = h;
}

// Step 4: statistics of one voxel from its (index-sorted) member list:
// math::ComputeMeanAndCov (math_utils.h:55-72) in index order, then the clamped inverse
// (ndt_registration.cpp:118-130): info = V diag(1/lambda) U^T with lambda_1,2 >= 1e-3 lambda_0.
// For the symmetric PSD covariance U = V = eigenvectors; a numerically zero lambda takes u = v.
LR_HD void ndt_voxel_stats(const unsigned int* idx, unsigned int cnt, const void* xyz, size_t stride, NdtVoxel& out) {
    double sx = 0, sy = 0, sz = 0;
    for (unsigned int j = 0; j < cnt; ++j) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + static_cast<size_t>(idx[j]) * stride);
        sx = LR_DADD(sx, static_cast<double>(p[0]));
        sy = LR_DADD(sy, static_cast<double>(p[1]));
        sz = LR_DADD(sz, static_cast<double>(p[2]));
    }
    const double len = static_cast<double>(cnt);
    const double mx = sx / len, my = sy / len, mz = sz / len;
    double c[6] = {0, 0, 0, 0, 0, 0};  // xx xy xz yy yz zz
    for (unsigned int j = 0; j < cnt; ++j) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + static_cast<size_t>(idx[j]) * stride);
        const double dx = LR_DSUB(static_cast<double>(p[0]), mx), dy = LR_DSUB(static_cast<double>(p[1]), my),
                     dz = LR_DSUB(static_cast<double>(p[2]), mz);
        c[0] = LR_DADD(c[0], LR_DMUL(dx, dx)); c[1] = LR_DADD(c[1], LR_DMUL(dx, dy)); c[2] = LR_DADD(c[2], LR_DMUL(dx, dz));
        c[3] = LR_DADD(c[3], LR_DMUL(dy, dy)); c[4] = LR_DADD(c[4], LR_DMUL(dy, dz)); c[5] = LR_DADD(c[5], LR_DMUL(dz, dz));
    }
    const double len1 = static_cast<double>(cnt - 1);
    for (int k = 0; k < 6; ++k) c[k] = c[k] / len1;
    double lam[3], Q[9];
    sym3_eigen(c, lam, Q);
    double sgn[3] = {1.0, 1.0, 1.0};
    for (int k = 0; k < 3; ++k) {  // singular value = |eigenvalue|; u = sign * v unless numerically zero
        if (lam[k] < 0) { lam[k] = -lam[k]; sgn[k] = (lam[k] > 1e-12 * fabs(lam[0])) ? -1.0 : 1.0; }
    }
    // sort by singular value descending (only matters if a negative eigenvalue flipped the order)
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (lam[b] > lam[a]) {
                double t = lam[a]; lam[a] = lam[b]; lam[b] = t;
                t = sgn[a]; sgn[a] = sgn[b]; sgn[b] = t;
                for (int r = 0; r < 3; ++r) { t = Q[r * 3 + a]; Q[r * 3 + a] = Q[r * 3 + b]; Q[r * 3 + b] = t; }
            }
    if (lam[1] < lam[0] * 1e-3) lam[1] = lam[0] * 1e-3;
    if (lam[2] < lam[0] * 1e-3) lam[2] = lam[0] * 1e-3;
    const double il[3] = {1.0 / lam[0], 1.0 / lam[1], 1.0 / lam[2]};
    out.mu[0] = mx; out.mu[1] = my; out.mu[2] = mz;
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Q[r * 3 + k] * il[k] * (sgn[k] * Q[cc * 3 + k]);
            out.info[r * 3 + cc] = s;
        }
}

// ---- per-point body of AlignNdt (ndt_registration.cpp:399-433) ---------------------------------
// Returns the number of voxels that passed the chi-square gate.  J = [ -R hat(q) , I ] is the same for
// every hit of the point, so H += hits * J^T J and err += -J^T (sum of e) — the information matrix
// only gates (quirk Q8).
template <class Acc>
LR_HD unsigned char ndt_point(const NdtMapView& map, const NdtParams& prm, const Pose& T, float sx, float sy, float sz,
                              Acc& acc) {
    if (!finite3(sx, sy, sz)) return 0;  // deviation D1
    const double qx = sx, qy = sy, qz = sz;
    double wx, wy, wz;
    pose_apply(T, qx, qy, qz, wx, wy, wz);
    const int kx = ndt_trunc(LR_DMUL(wx, map.inv_voxel)), ky = ndt_trunc(LR_DMUL(wy, map.inv_voxel)),
              kz = ndt_trunc(LR_DMUL(wz, map.inv_voxel));
    acc.inc_eff();  // effective_num++ per point, unconditionally (:432)
    // The probes of the (up to seven) voxels are independent: issue all first-slot loads, then all voxel loads, before
    // anything is consumed - the kernel is bound by the latency of these dependent accesses, not by arithmetic.
    int vid[7];
    unsigned int slot_h[7];
    unsigned long long keys[7];
    NdtSlot first[7];
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        int dx, dy, dz;
        ndt_offset(o, dx, dy, dz);
        const int cx = kx + dx, cy = ky + dy, cz = kz + dz;
        const bool use = o < prm.n_nearby && ndt_key_ok(cx, cy, cz);
        keys[o] = use ? ndt_pack(cx, cy, cz) : kNdtEmpty;
        slot_h[o] = ndt_hash(keys[o]) & map.slot_mask;
        if (use) first[o] = map.slots[slot_h[o]];
        else { first[o].key = kNdtEmpty; first[o].vid = -1; first[o].count = 0; }
    }
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        vid[o] = -1;
        if (keys[o] == kNdtEmpty) continue;
        NdtSlot s = first[o];
        unsigned int h = slot_h[o];
        while (true) {  // linear probing past the first slot is rare (load factor <= 0.5)
            if (s.key == keys[o]) { vid[o] = s.vid; break; }
            if (s.key == kNdtEmpty) break;
            h = (h + 1) & map.slot_mask;
            s = map.slots[h];
        }
    }
    int hits = 0;
    double ex = 0, ey = 0, ez = 0, ss = 0;
#pragma unroll
    for (int o = 0; o < 7; ++o) {  // the reference's order (GenerateNearbyGrids)
        if (vid[o] < 0) continue;
        const NdtVoxel& v = map.voxels[vid[o]];
        const double e0 = wx - v.mu[0], e1 = wy - v.mu[1], e2 = wz - v.mu[2];
        const double i0 = v.info[0] * e0 + v.info[1] * e1 + v.info[2] * e2;
        const double i1 = v.info[3] * e0 + v.info[4] * e1 + v.info[5] * e2;
        const double i2 = v.info[6] * e0 + v.info[7] * e1 + v.info[8] * e2;
        const double res = e0 * i0 + e1 * i1 + e2 * i2;
        if (!(res == res) || res > prm.res_outlier_th) continue;
        ++hits;
        ex += e0; ey += e1; ez += e2;
        ss += e0 * e0 + e1 * e1 + e2 * e2;
    }
    if (hits == 0) return 0;
    // A = -R hat(q); hat(q) = [0 -qz qy; qz 0 -qx; -qy qx 0]
    double A[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double r0 = T.R[r * 3], r1 = T.R[r * 3 + 1], r2 = T.R[r * 3 + 2];
        A[r][0] = -(r1 * qz - r2 * qy);
        A[r][1] = -(-r0 * qz + r2 * qx);
        A[r][2] = -(r0 * qy - r1 * qx);
    }
    const double w = static_cast<double>(hits);
    // H = [A^T A, A^T; A, I] * hits
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = i; j < 3; ++j)
            acc.add(hidx(i, j), w * (A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j]));
#pragma unroll
        for (int j = 0; j < 3; ++j) acc.add(hidx(i, 3 + j), w * A[j][i]);
    }
    acc.add(hidx(3, 3), w); acc.add(hidx(4, 4), w); acc.add(hidx(5, 5), w);
    // err = -J^T sum(e)
#pragma unroll
    for (int i = 0; i < 3; ++i) acc.add(21 + i, -(A[0][i] * ex + A[1][i] * ey + A[2][i] * ez));
    acc.add(24, -ex); acc.add(25, -ey); acc.add(26, -ez);
    acc.add(27, ss);
    acc.inc_inl(static_cast<unsigned int>(hits));
    return static_cast<unsigned char>(hits);
}

// ---- incremental NDT (ndt_registration.cpp:150-236, 262-372) ------------------------------------------------------
// UpdateVoxel as it actually runs: flag_first_scan_ is set back to true after every SetIncNdtTargetCloud
// (:181), so only its first branch (:186-198) is ever taken.  The statistics of a voxel are those of the points the
// CURRENT cloud put into it (pts_ is cleared after every update): mean and covariance (/(n-1), math_utils.h:55-72) in
// arrival order and info = (sigma + 1e-3 I)^-1 for two or more points, mu = the point and info = 100 I for one.
// The member points of a voxel come either as an index list (tests/hostsim) or as RUNS of consecutive point indices
// (the device path: a scan line enters and leaves a voxel a few times, so a voxel's points are a handful of runs).
struct IncMembersList {
    const unsigned int* idx;
    unsigned int cnt;
    template <class F> LR_HD void for_each(F f) const {
        for (unsigned int j = 0; j < cnt; ++j) f(idx[j]);
    }
    LR_HD unsigned int first() const { return idx[0]; }
};
struct IncRun { unsigned int first, len; };
struct IncMembersRuns {
    const IncRun* runs;
    unsigned int n_runs;
    template <class F> LR_HD void for_each(F f) const {
        for (unsigned int r = 0; r < n_runs; ++r)
            for (unsigned int j = 0; j < runs[r].len; ++j) f(runs[r].first + j);
    }
    LR_HD unsigned int first() const { return runs[0].first; }
};
template <class Members>
LR_HD void inc_ndt_voxel_stats_of(const Members& mem, unsigned int cnt, const void* xyz, size_t stride, NdtVoxel& out) {
    if (cnt == 1) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + static_cast<size_t>(mem.first()) * stride);
        out.mu[0] = p[0]; out.mu[1] = p[1]; out.mu[2] = p[2];
        for (int k = 0; k < 9; ++k) out.info[k] = (k % 4 == 0) ? 1e2 : 0.0;
        return;
    }
    double sx = 0, sy = 0, sz = 0;
    mem.for_each([&](unsigned int i) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + static_cast<size_t>(i) * stride);
        sx = LR_DADD(sx, static_cast<double>(p[0])); sy = LR_DADD(sy, static_cast<double>(p[1])); sz = LR_DADD(sz, static_cast<double>(p[2]));
    });
    const double len = static_cast<double>(cnt);
    const double mx = sx / len, my = sy / len, mz = sz / len;
    double c[6] = {0, 0, 0, 0, 0, 0};  // xx xy xz yy yz zz
    mem.for_each([&](unsigned int i) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + static_cast<size_t>(i) * stride);
        const double dx = LR_DSUB(static_cast<double>(p[0]), mx), dy = LR_DSUB(static_cast<double>(p[1]), my),
                     dz = LR_DSUB(static_cast<double>(p[2]), mz);
        c[0] = LR_DADD(c[0], LR_DMUL(dx, dx)); c[1] = LR_DADD(c[1], LR_DMUL(dx, dy)); c[2] = LR_DADD(c[2], LR_DMUL(dx, dz));
        c[3] = LR_DADD(c[3], LR_DMUL(dy, dy)); c[4] = LR_DADD(c[4], LR_DMUL(dy, dz)); c[5] = LR_DADD(c[5], LR_DMUL(dz, dz));
    });
    const double len1 = static_cast<double>(cnt - 1);
    // A = sigma + 1e-3 I (symmetric), info = adj(A) / det(A): Eigen's fixed-size 3x3 inverse is the cofactor formula
    const double a = c[0] / len1 + 1e-3, b = c[1] / len1, cc = c[2] / len1, d = c[3] / len1 + 1e-3, e = c[4] / len1,
                 f = c[5] / len1 + 1e-3;
    const double A00 = d * f - e * e, A01 = cc * e - b * f, A02 = b * e - cc * d;
    const double A11 = a * f - cc * cc, A12 = b * cc - a * e, A22 = a * d - b * b;
    const double det = a * A00 + b * A01 + cc * A02;
    const double inv = 1.0 / det;
    out.mu[0] = mx; out.mu[1] = my; out.mu[2] = mz;
    out.info[0] = A00 * inv; out.info[1] = A01 * inv; out.info[2] = A02 * inv;
    out.info[3] = A01 * inv; out.info[4] = A11 * inv; out.info[5] = A12 * inv;
    out.info[6] = A02 * inv; out.info[7] = A12 * inv; out.info[8] = A22 * inv;
}

LR_HD void inc_ndt_voxel_stats(const unsigned int* idx, unsigned int cnt, const void* xyz, size_t stride, NdtVoxel& out) {
    inc_ndt_voxel_stats_of(IncMembersList{idx, cnt}, cnt, xyz, stride, out);
}

// Per-point body of AlignIncNdt (ndt_registration.cpp:289-324, 334-347): every gated-in voxel is one residual,
// weighted by its information matrix: H += J^T info J, err += -J^T info e, total_res += e^T info e, with
// J = [ -R hat(q) , I ] the same for all voxels of the point, so the sums W = sum info_k and g = sum info_k e_k are
// formed first.  effective_num counts RESIDUALS (:341), not points.
template <class Acc>
LR_HD unsigned char inc_ndt_point(const NdtMapView& map, const NdtParams& prm, const Pose& T, float sx, float sy, float sz,
                                  Acc& acc) {
    if (!finite3(sx, sy, sz)) return 0;  // deviation D1
    const double qx = sx, qy = sy, qz = sz;
    double wx, wy, wz;
    pose_apply(T, qx, qy, qz, wx, wy, wz);
    const int kx = ndt_trunc(LR_DMUL(wx, map.inv_voxel)), ky = ndt_trunc(LR_DMUL(wy, map.inv_voxel)),
              kz = ndt_trunc(LR_DMUL(wz, map.inv_voxel));
    int hits = 0;
    double W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, rr = 0;
    for (int o = 0; o < prm.n_nearby; ++o) {
        int dx, dy, dz;
        ndt_offset(o, dx, dy, dz);
        const int cx = kx + dx, cy = ky + dy, cz = kz + dz;
        if (!ndt_key_ok(cx, cy, cz)) continue;
        const unsigned long long key = ndt_pack(cx, cy, cz);
        unsigned int h = ndt_hash(key) & map.slot_mask;
        int vid = -1;
        while (true) {
            const NdtSlot s = map.slots[h];
            if (s.key == key) { vid = s.vid; break; }
            if (s.key == kNdtEmpty) break;
            h = (h + 1) & map.slot_mask;
        }
        if (vid < 0) continue;
        const NdtVoxel& v = map.voxels[vid];
        const double e0 = wx - v.mu[0], e1 = wy - v.mu[1], e2 = wz - v.mu[2];
        const double i0 = v.info[0] * e0 + v.info[1] * e1 + v.info[2] * e2;
        const double i1 = v.info[3] * e0 + v.info[4] * e1 + v.info[5] * e2;
        const double i2 = v.info[6] * e0 + v.info[7] * e1 + v.info[8] * e2;
        const double res = e0 * i0 + e1 * i1 + e2 * i2;
        if (!(res == res) || res > prm.res_outlier_th) continue;
        ++hits;
#pragma unroll
        for (int k = 0; k < 9; ++k) W[k] += v.info[k];
        g[0] += i0; g[1] += i1; g[2] += i2;
        rr += res;
    }
    if (hits == 0) return 0;
    double A[3][3];  // A = -R hat(q)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double r0 = T.R[r * 3], r1 = T.R[r * 3 + 1], r2 = T.R[r * 3 + 2];
        A[r][0] = -(r1 * qz - r2 * qy);
        A[r][1] = -(-r0 * qz + r2 * qx);
        A[r][2] = -(r0 * qy - r1 * qx);
    }
    // WA = W A (3 x 3); H = [A^T W A, A^T W; W A, W] (upper triangle kept), err = -[A^T g; g]
    double WA[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) WA[r][c] = W[r * 3] * A[0][c] + W[r * 3 + 1] * A[1][c] + W[r * 3 + 2] * A[2][c];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = i; j < 3; ++j) acc.add(hidx(i, j), A[0][i] * WA[0][j] + A[1][i] * WA[1][j] + A[2][i] * WA[2][j]);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc.add(hidx(i, 3 + j), A[0][i] * W[j] + A[1][i] * W[3 + j] + A[2][i] * W[6 + j]);  // (A^T W)(i, j)
    }
    acc.add(hidx(3, 3), W[0]); acc.add(hidx(3, 4), W[1]); acc.add(hidx(3, 5), W[2]);
    acc.add(hidx(4, 4), W[4]); acc.add(hidx(4, 5), W[5]); acc.add(hidx(5, 5), W[8]);
#pragma unroll
    for (int i = 0; i < 3; ++i) acc.add(21 + i, -(A[0][i] * g[0] + A[1][i] * g[1] + A[2][i] * g[2]));
    acc.add(24, -g[0]); acc.add(25, -g[1]); acc.add(26, -g[2]);
    acc.add(27, rr);
    acc.inc_eff(static_cast<unsigned int>(hits));
    acc.inc_inl(static_cast<unsigned int>(hits));
    return static_cast<unsigned char>(hits);
}
// Tail of one AlignIncNdt iteration (:349-367): too few residuals -> result_pose = pose, return false (4: stop, the
// pose IS written); otherwise solve, update, converged?  The reference never looks at det(H) here; a singular H
// (not reachable with >= min_effective_pts_ residuals on real data) stops the loop with the pose unchanged.
LR_HD int inc_ndt_gn_update(const double* acc28, unsigned int n_eff, const NdtParams& prm, Pose& T) {
    if (static_cast<long long>(n_eff) < static_cast<long long>(prm.min_effective_pts)) return 4;
    double dx[6];
    if (!gn_solve6(acc28, acc28 + 21, dx)) return 4;
    pose_update(T, dx);
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm += dx[i] * dx[i];
    return sqrt(nrm) < prm.eps ? 2 : 1;
}

// Tail of one AlignNdt iteration (ndt_registration.cpp:435-459).
// 3 = det(H)==0: return false WITHOUT writing result_pose (quirk Q11); 0 = too few points (`continue`);
// 1 = updated; 2 = updated and converged.
LR_HD int ndt_gn_update(const double* acc28, unsigned int n_eff, const NdtParams& prm, Pose& T) {
    double dx[6];
    if (!gn_solve6(acc28, acc28 + 21, dx)) return 3;
    if (static_cast<long long>(n_eff) < static_cast<long long>(prm.min_effective_pts)) return 0;
    pose_update(T, dx);
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm += dx[i] * dx[i];
    return sqrt(nrm) < prm.eps ? 2 : 1;
}

}  // namespace locreg

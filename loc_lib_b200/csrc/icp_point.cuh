// Per-source-point body of the ICP Gauss-Newton accumulation (kernel K3 of SURVEY.md §2.3):
// transform -> exact k-NN in the voxel-hash map -> (plane fit) -> gates -> J, r -> H += J^T J, B += -J^T r.
// Restates the loop bodies of IcpRegistration::CaculateMatrixHAndBP2P (icp_registration.cpp:62-92)
// and ::CaculateMatrixHAndBP2Plane (icp_registration.cpp:166-202) for one point.
#pragma once
#include "la.cuh"
#include "voxel_map.cuh"

namespace locreg {

enum Method : int { kIcpP2P = 0, kIcpP2Line = 1, kIcpP2Plane = 2, kNdtDirect = 3 };

struct IcpParams {
    double max_nn_distance;     // IcpOptions::max_nn_distance_ (compared against a SQUARED distance, quirk Q6)
    double max_plane_distance;  // IcpOptions::max_plane_distance_
    double max_line_distance;   // IcpOptions::max_line_distance_ (also FitLine's eps, icp_registration.cpp:123)
    double plane_fit_eps;       // math::FitPlane's eps (1e-2, math_utils.h:113)
    double eps;                 // IcpOptions::eps_
    int max_iteration;
    int min_effective_pts;
};

// gate codes written by the debug probe (same meaning as oracle_icp_compute_hb's gate[])
enum Gate : unsigned char { kGateSkipped = 0, kGateFitFailed = 1, kGateResidual = 2, kGateInlier = 3 };

// The accumulator policy `Acc` of the per-point bodies provides row(J, r) - one residual row, meaning
// H += J^T J, B += -J^T r, sum_sq += r^2 - and the two counters inc_eff() / inc_inl().  On the host (oracle-style
// serial loop, tests/hostsim) it is `Accum`; in k_icp_post it stages the row in shared memory (RowSink) and the
// block forms the products cooperatively.
// math::FitPlane on the five neighbours (icp_registration.cpp:171-181): kPlaneOk + coefficients, kPlaneFailed (the eps
// check rejected the fit) or kPlaneNone (fewer than five neighbours).
enum PlaneStatus : unsigned char { kPlaneNone = 0, kPlaneOk = 1, kPlaneFailed = 2 };
LR_HD unsigned char icp_fit_plane(const VoxelMapView& map, const IcpParams& prm, const KnnResult<5>& nn, double (&n)[4]) {
    // a map with fewer than 5 leaves makes KdTree::GetClosestPoint refuse (kdtree.cpp:149): no neighbours
    if (knn_count(nn) < 5) return kPlaneNone;
    PlaneAcc pa;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = map.pts[nn.pos[j]];
        if (j == 0) pa.start(p.x, p.y, p.z); else pa.add(p.x, p.y, p.z);
    }
    if (!plane_fit5_solve(pa, n)) {  // collinear / coincident neighbours only: the SVD of math::FitPlane (:179)
        // only arrays local to this branch have their address taken (the SVD is not inlined)
        double Pd[5][3], ns[4];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float4 p = map.pts[nn.pos[j]];
            Pd[j][0] = p.x; Pd[j][1] = p.y; Pd[j][2] = p.z;
        }
        plane_svd5(Pd, ns);
        n[0] = ns[0]; n[1] = ns[1]; n[2] = ns[2]; n[3] = ns[3];
    }
    // FitPlane's own check (math_utils.h:128-133); the neighbours are read again (L1) rather than kept in registers
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = map.pts[nn.pos[j]];
        const double err = n[0] * static_cast<double>(p.x) + n[1] * static_cast<double>(p.y) + n[2] * static_cast<double>(p.z) + n[3];
        if (err * err > prm.plane_fit_eps) return kPlaneFailed;
    }
    return kPlaneOk;
}
// Point-to-plane residual of one source point against a fitted plane (icp_registration.cpp:183-201).
template <class Acc>
LR_HD unsigned char icp_p2plane_residual(const IcpParams& prm, const Pose& T, double qx, double qy, double qz, double wx, double wy,
                                         double wz, unsigned char plane_status, const double (&n)[4], Acc& acc) {
    if (plane_status == kPlaneNone) return kGateSkipped;
    if (plane_status == kPlaneFailed) return kGateFitFailed;
    acc.inc_eff();  // quirk Q4: counted before the distance gate (:184)
    const double dis = n[0] * wx + n[1] * wy + n[2] * wz + n[3];
    if (fabs(dis) > prm.max_plane_distance) return kGateResidual;
    // J = [ -n^T R hat(q) , n^T ]  (:193-195): with m = R^T n, -m^T hat(q) = (q x m)^T
    const double mx = T.R[0] * n[0] + T.R[3] * n[1] + T.R[6] * n[2];
    const double my = T.R[1] * n[0] + T.R[4] * n[1] + T.R[7] * n[2];
    const double mz = T.R[2] * n[0] + T.R[5] * n[1] + T.R[8] * n[2];
    const double J[6] = {qy * mz - qz * my, qz * mx - qx * mz, qx * my - qy * mx, n[0], n[1], n[2]};
    acc.row(J, dis);
    acc.inc_inl();
    return kGateInlier;
}
// Point-to-plane, everything after the neighbour search (icp_registration.cpp:171-201): plane fit, gates,
// Jacobian row, accumulation.  q = source point, w = predict_pose * q, nn = its 5 nearest map points.
template <class Acc>
LR_HD unsigned char icp_p2plane_post(const VoxelMapView& map, const IcpParams& prm, const Pose& T, double qx, double qy,
                                     double qz, double wx, double wy, double wz, const KnnResult<5>& nn, Acc& acc) {
    double n[4] = {0, 0, 0, 0};
    const unsigned char st = icp_fit_plane(map, prm, nn, n);
    return icp_p2plane_residual(prm, T, qx, qy, qz, wx, wy, wz, st, n, acc);
}

// Point-to-line after the neighbour search (icp_registration.cpp:115-147): line fit through the five neighbours
// (math::FitLine, math_utils.h:138-163), e = d x (qs - p0), J = [ -hat(d) R hat(q) , hat(d) ] (3 x 6).
template <class Acc>
LR_HD unsigned char icp_p2line_post(const VoxelMapView& map, const IcpParams& prm, const Pose& T, double qx, double qy,
                                    double qz, double wx, double wy, double wz, const KnnResult<5>& nn, Acc& acc) {
    if (knn_count(nn) != 5) return kGateSkipped;  // nn.size() == 5 (:115)
    // origin = mean of the five points, summed in neighbour order like std::accumulate (math_utils.h:144)
    double ox = 0, oy = 0, oz = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = map.pts[nn.pos[j]];
        ox += p.x; oy += p.y; oz += p.z;
    }
    ox /= 5; oy /= 5; oz /= 5;
    double S[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = map.pts[nn.pos[j]];
        const double x = p.x - ox, y = p.y - oy, z = p.z - oz;
        S[0] += x * x; S[1] += x * y; S[2] += x * z; S[3] += y * y; S[4] += y * z; S[5] += z * z;
    }
    double d[3];
    line_dir_from_scatter(S, d);
#pragma unroll
    for (int j = 0; j < 5; ++j) {  // FitLine's eps check: |d x (p - origin)|^2 > eps  (math_utils.h:155-159)
        const float4 p = map.pts[nn.pos[j]];
        const double x = p.x - ox, y = p.y - oy, z = p.z - oz;
        const double cx = d[1] * z - d[2] * y, cy = d[2] * x - d[0] * z, cz = d[0] * y - d[1] * x;
        if (cx * cx + cy * cy + cz * cz > prm.max_line_distance) return kGateFitFailed;
    }
    acc.inc_eff();
    const double vx = wx - ox, vy = wy - oy, vz = wz - oz;
    const double e[3] = {d[1] * vz - d[2] * vy, d[2] * vx - d[0] * vz, d[0] * vy - d[1] * vx};
    if (sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) > prm.max_line_distance) return kGateResidual;
    acc.inc_inl();
    // M = R hat(q) (3 x 3); row r of -hat(d) M is -(d x M_col) taken per column, i.e. row r of hat(d) applied to M
    const double hq[3][3] = {{0.0, -qz, qy}, {qz, 0.0, -qx}, {-qy, qx, 0.0}};
    double M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[r][c] = T.R[r * 3 + 0] * hq[0][c] + T.R[r * 3 + 1] * hq[1][c] + T.R[r * 3 + 2] * hq[2][c];
    const double hd[3][3] = {{0.0, -d[2], d[1]}, {d[2], 0.0, -d[0]}, {-d[1], d[0], 0.0}};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double J[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) J[c] = -(hd[r][0] * M[0][c] + hd[r][1] * M[1][c] + hd[r][2] * M[2][c]);
        J[3] = hd[r][0]; J[4] = hd[r][1]; J[5] = hd[r][2];
        acc.row(J, e[r]);
    }
    return kGateInlier;
}

// Point-to-point after the neighbour search (icp_registration.cpp:71-91).  J = [ R hat(q) / 16 , -I ]  (quirk Q6).
template <class Acc>
LR_HD unsigned char icp_p2p_post(const VoxelMapView& map, const IcpParams& prm, const Pose& T, double qx, double qy,
                                 double qz, double wx, double wy, double wz, const KnnResult<1>& nn, Acc& acc) {
    if (nn.pos[0] == kNoPos) return kGateSkipped;
    const float4 p = map.pts[nn.pos[0]];
    const double e[3] = {static_cast<double>(p.x) - wx, static_cast<double>(p.y) - wy, static_cast<double>(p.z) - wz};
    const double dis2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    if (dis2 > prm.max_nn_distance) return kGateResidual;  // squared vs unsquared threshold (:75)
    acc.inc_eff();
    acc.inc_inl();
    const double hq[3][3] = {{0.0, -qz, qy}, {qz, 0.0, -qx}, {-qy, qx, 0.0}};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double J[6];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            J[c] = (T.R[r * 3 + 0] * hq[0][c] + T.R[r * 3 + 1] * hq[1][c] + T.R[r * 3 + 2] * hq[2][c]) / 16;
        J[3] = r == 0 ? -1.0 : 0.0;
        J[4] = r == 1 ? -1.0 : 0.0;
        J[5] = r == 2 ? -1.0 : 0.0;
        acc.row(J, e[r]);  // sum_sq accumulates e[0]^2 + e[1]^2 + e[2]^2 = dis2
    }
    return kGateInlier;
}

// Whole per-point bodies with the serial (one query at a time) search: used by tests/hostsim and kept as the
// readable statement of the algorithm; the kernels split the same steps over k_icp_nn / k_icp_post.
// nn_pos (optional, K entries, in/out): seeds from the previous iteration on entry, this iteration's neighbour
// positions on exit.
LR_HD unsigned char icp_point_p2plane(const VoxelMapView& map, const CoarseLevels& coarse, const IcpParams& prm, const Pose& T, float sx, float sy,
                                      float sz, Accum& acc, int* nn_out, unsigned int* nn_pos = nullptr) {
    if (nn_out) {
#pragma unroll
        for (int j = 0; j < 5; ++j) nn_out[j] = -1;
    }
    if (!finite3(sx, sy, sz)) return kGateSkipped;  // deviation D1: the reference would poison H with NaN
    const double qx = sx, qy = sy, qz = sz;
    double wx, wy, wz;
    pose_apply(T, qx, qy, qz, wx, wy, wz);  // qs = predict_pose * q  (:169)
    KnnResult<5> nn;
    knn_query<5>(map, coarse, true, static_cast<float>(wx), static_cast<float>(wy), static_cast<float>(wz), nn, nn_pos);  // (:170)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (nn_out) nn_out[j] = nn.pos[j] != kNoPos ? knn_index_of(map.pts, nn.pos[j]) : -1;
        if (nn_pos) nn_pos[j] = nn.pos[j];
    }
    return icp_p2plane_post(map, prm, T, qx, qy, qz, wx, wy, wz, nn, acc);
}

LR_HD unsigned char icp_point_p2line(const VoxelMapView& map, const CoarseLevels& coarse, const IcpParams& prm, const Pose& T,
                                     float sx, float sy, float sz, Accum& acc, int* nn_out, unsigned int* nn_pos = nullptr) {
    if (nn_out) {
#pragma unroll
        for (int j = 0; j < 5; ++j) nn_out[j] = -1;
    }
    if (!finite3(sx, sy, sz)) return kGateSkipped;  // deviation D1
    const double qx = sx, qy = sy, qz = sz;
    double wx, wy, wz;
    pose_apply(T, qx, qy, qz, wx, wy, wz);
    KnnResult<5> nn;
    knn_query<5>(map, coarse, true, static_cast<float>(wx), static_cast<float>(wy), static_cast<float>(wz), nn, nn_pos);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (nn_out) nn_out[j] = nn.pos[j] != kNoPos ? knn_index_of(map.pts, nn.pos[j]) : -1;
        if (nn_pos) nn_pos[j] = nn.pos[j];
    }
    return icp_p2line_post(map, prm, T, qx, qy, qz, wx, wy, wz, nn, acc);
}

LR_HD unsigned char icp_point_p2p(const VoxelMapView& map, const CoarseLevels& coarse, const IcpParams& prm, const Pose& T, float sx, float sy,
                                  float sz, Accum& acc, int* nn_out, unsigned int* nn_pos = nullptr) {
    if (nn_out) nn_out[0] = -1;
    if (!finite3(sx, sy, sz)) return kGateSkipped;  // pcl::isFinite (:64)
    const double qx = sx, qy = sy, qz = sz;
    double wx, wy, wz;
    pose_apply(T, qx, qy, qz, wx, wy, wz);
    KnnResult<1> nn;
    knn_query<1>(map, coarse, true, static_cast<float>(wx), static_cast<float>(wy), static_cast<float>(wz), nn, nn_pos);
    if (nn_out) nn_out[0] = nn.pos[0] != kNoPos ? knn_index_of(map.pts, nn.pos[0]) : -1;
    if (nn_pos) nn_pos[0] = nn.pos[0];
    return icp_p2p_post(map, prm, T, qx, qy, qz, wx, wy, wz, nn, acc);
}

template <int METHOD>
LR_HD unsigned char icp_point(const VoxelMapView& map, const CoarseLevels& coarse, const IcpParams& prm, const Pose& T, float sx, float sy, float sz,
                              Accum& acc, int* nn_out, unsigned int* nn_pos = nullptr) {
    if (METHOD == kIcpP2P) return icp_point_p2p(map, coarse, prm, T, sx, sy, sz, acc, nn_out, nn_pos);
    if (METHOD == kIcpP2Line) return icp_point_p2line(map, coarse, prm, T, sx, sy, sz, acc, nn_out, nn_pos);
    return icp_point_p2plane(map, coarse, prm, T, sx, sy, sz, acc, nn_out, nn_pos);
}

// One Gauss-Newton update from the reduced accumulator: the tail of AlignP2P / AlignP2Plane
// (icp_registration.cpp:284-299, 362-375).  Returns 0 = evaluation failed (pose unchanged, quirk Q11),
// 1 = pose updated, 2 = pose updated and ||dx|| < eps (converged).
template <int METHOD>
LR_HD int icp_gn_update(const double* acc28, unsigned int n_eff, const IcpParams& prm, Pose& T) {
    if (static_cast<long long>(n_eff) < static_cast<long long>(prm.min_effective_pts)) return 0;
    double dx[6];
    if (!gn_solve6(acc28, acc28 + 21, dx)) return 0;
    if (METHOD == kIcpP2P) {
#pragma unroll
        for (int i = 0; i < 6; ++i) dx[i] = dx[i] / 16;  // H.inverse()/16 * err (:287)
    }
    pose_update(T, dx);
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm += dx[i] * dx[i];
    return sqrt(nrm) < prm.eps ? 2 : 1;
}

}  // namespace locreg

// ORACLE (test infrastructure, NOT product code) — see oracle.h / la.hpp headers.
//
// std-only C++17 restatement of the reference's registration hot path:
//   KdTree                      LocUtils/src/model/search_point/kdtree/kdtree.cpp (whole file)
//   BfnnRegistration            LocUtils/src/model/search_point/bfnn/bfnn.cpp:24-50
//   math::FitPlane              LocUtils/include/LocUtils/common/math_utils.h:112-136
//   math::ComputeMeanAndCov*    LocUtils/include/LocUtils/common/math_utils.h:35-72
//   IcpRegistration             LocUtils/src/model/matching/3d/icp/icp_registration.cpp:31-103,161-303,345-381
//   NdtRegistration (direct)    LocUtils/src/model/matching/3d/ndt/ndt_registration.cpp:51-63,87-148,374-464
//   hash_vec<3>                 LocUtils/include/LocUtils/common/eigen_types.h:104-107
// including quirks Q1-Q12 of SURVEY.md §8.  The reference itself cannot be compiled here
// (needs Eigen, Sophus, PCL, glog, ROS: none installed), so there is no oracle/_ref.
// Build flags mirror the reference's (-std=c++17 -O3, no -march, no OpenMP; LocUtils/CMakeLists.txt:4-6).
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <list>
#include <memory>
#include <set>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

#include "la.hpp"

namespace oracle {

struct Vec3f {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

static inline const float* pt_at(const float* base, size_t i, size_t stride) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + i * stride);
}
static inline float* pt_at(float* base, size_t i, size_t stride) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + i * stride);
}

// KdTree::Dis2 (kdtree.h:94): (p1-p2).squaredNorm() in float32.  Eigen 3.3's fixed-size
// unrolled reduction associates as x^2 + (y^2 + z^2) (redux_novec_unroller, HalfLength = 1).
static inline float dis2f(const Vec3f& a, const Vec3f& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    const float yy_zz = dy * dy + dz * dz;
    return dx * dx + yy_zz;
}

// ---------------------------------------------------------------------------------------------
// KdTree (kdtree.cpp)
// ---------------------------------------------------------------------------------------------
struct KdTreeNode {  // kdtree.h:20-29
    int id_ = -1;
    int point_idx_ = 0;
    int axis_index_ = 0;
    float split_thresh_ = 0.0f;
    KdTreeNode* left_ = nullptr;
    KdTreeNode* right_ = nullptr;
    bool IsLeaf() const { return left_ == nullptr && right_ == nullptr; }
};
struct NodeAndDistance {  // kdtree.h:32-38
    NodeAndDistance(KdTreeNode* n, float d) : node_(n), distance2_(d) {}
    KdTreeNode* node_;
    float distance2_;
    bool operator<(const NodeAndDistance& o) const { return distance2_ < o.distance2_; }
};
struct IdxDist {  // total order (dis2, index) used by the parity contract
    float d2;
    int idx;
};
static inline bool better(const IdxDist& a, const IdxDist& b) {
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

class KdTree {
   public:
    ~KdTree() { Clear(); }

    // kdtree.cpp:10-31
    bool BuildTree(const std::vector<Vec3f>& cloud) {
        if (cloud.empty()) return false;
        cloud_ = cloud;
        Clear();
        tree_node_id_ = 0;
        root_ = new KdTreeNode();
        root_->id_ = tree_node_id_++;
        size_ = 0;
        std::vector<int> idx(cloud_.size());
        for (size_t i = 0; i < cloud_.size(); ++i) idx[i] = static_cast<int>(i);
        Insert(idx, root_);
        return true;
    }
    void Clear() {  // kdtree.cpp:34-48
        for (auto& np : nodes_) delete np.second;
        nodes_.clear();
        root_ = nullptr;
        size_ = 0;
        tree_node_id_ = 0;
    }
    size_t size() const { return size_; }
    const std::vector<Vec3f>& cloud() const { return cloud_; }

    // kdtree.cpp:147-167 (+ SetEnableANN kdtree.h:57-60).  approximate/alpha are arguments here so
    // the search is re-entrant (the reference keeps k_ / approximate_ as members).
    bool GetClosestPoint(const Vec3f& pt, std::vector<int>& closest_idx, int k, bool approximate, float alpha) const {
        if (static_cast<size_t>(k) > size_) { closest_idx.clear(); return false; }
        std::priority_queue<NodeAndDistance> knn_result;
        Knn(pt, root_, knn_result, k, approximate, alpha);
        closest_idx.resize(knn_result.size());
        for (int i = static_cast<int>(closest_idx.size()) - 1; i >= 0; --i) {
            closest_idx[i] = knn_result.top().node_->point_idx_;
            knn_result.pop();
        }
        return true;
    }

    // Exact k-NN over the tree's leaves under the total order (dis2_f32, index): the parity target
    // (SURVEY.md §8 Q1).  Pruning uses d*d > worst (not >=) so equal-distance, lower-index points on
    // the far side are still reached; see DESIGN.md "NN contract" for the monotonicity argument.
    bool GetClosestPointExact(const Vec3f& pt, std::vector<int>& closest_idx, int k) const {
        if (static_cast<size_t>(k) > size_) { closest_idx.clear(); return false; }
        std::vector<IdxDist> best;
        best.reserve(k + 1);
        KnnExact(pt, root_, best, k);
        closest_idx.resize(best.size());
        for (size_t i = 0; i < best.size(); ++i) closest_idx[i] = best[i].idx;
        return true;
    }

   private:
    // kdtree.cpp:58-94
    void Insert(const std::vector<int>& points, KdTreeNode* node) {
        nodes_.insert({node->id_, node});
        if (points.empty()) return;
        if (points.size() == 1) {
            size_++;
            node->point_idx_ = points[0];
            return;
        }
        std::vector<int> left, right;
        if (!FindSplitAxisAndThresh(points, node->axis_index_, node->split_thresh_, left, right)) {
            size_++;
            node->point_idx_ = points[0];  // quirk Q3: the rest of an all-equal set is dropped
            return;
        }
        if (!left.empty()) {
            node->left_ = new KdTreeNode;
            node->left_->id_ = tree_node_id_++;
            Insert(left, node->left_);
        }
        if (!right.empty()) {
            node->right_ = new KdTreeNode;
            node->right_->id_ = tree_node_id_++;
            Insert(right, node->right_);
        }
    }
    // kdtree.cpp:96-123 with math::ComputeMeanAndCovDiag (math_utils.h:35-47) in float32
    bool FindSplitAxisAndThresh(const std::vector<int>& point_idx, int& axis, float& th, std::vector<int>& left,
                                std::vector<int>& right) const {
        const size_t len = point_idx.size();
        float sx = 0, sy = 0, sz = 0;
        for (int i : point_idx) { sx = sx + cloud_[i].x; sy = sy + cloud_[i].y; sz = sz + cloud_[i].z; }
        const float flen = static_cast<float>(len);
        const float mean[3] = {sx / flen, sy / flen, sz / flen};
        float vx = 0, vy = 0, vz = 0;
        for (int i : point_idx) {
            const float dx = cloud_[i].x - mean[0], dy = cloud_[i].y - mean[1], dz = cloud_[i].z - mean[2];
            vx = vx + dx * dx; vy = vy + dy * dy; vz = vz + dz * dz;
        }
        const float flen1 = static_cast<float>(len - 1);
        const float var[3] = {vx / flen1, vy / flen1, vz / flen1};
        int max_i = 0;  // DenseBase::maxCoeff: strict '>' keeps the first maximum
        for (int i = 1; i < 3; ++i)
            if (var[i] > var[max_i]) max_i = i;
        axis = max_i;
        th = mean[axis];
        for (int idx : point_idx) {
            if (cloud_[idx][axis] < th) left.emplace_back(idx);
            else right.emplace_back(idx);
        }
        if (point_idx.size() > 1 && (left.empty() || right.empty())) return false;
        return true;
    }
    // kdtree.cpp:169-195
    void Knn(const Vec3f& pt, KdTreeNode* node, std::priority_queue<NodeAndDistance>& knn_result, int k,
             bool approximate, float alpha) const {
        if (node->IsLeaf()) {
            ComputeDisForLeaf(pt, node, knn_result, k);
            return;
        }
        KdTreeNode *this_side, *that_side;
        if (pt[node->axis_index_] < node->split_thresh_) { this_side = node->left_; that_side = node->right_; }
        else { this_side = node->right_; that_side = node->left_; }
        Knn(pt, this_side, knn_result, k, approximate, alpha);
        if (NeedExpand(pt, node, knn_result, k, approximate, alpha)) Knn(pt, that_side, knn_result, k, approximate, alpha);
    }
    // kdtree.cpp:197-212
    void ComputeDisForLeaf(const Vec3f& pt, KdTreeNode* node, std::priority_queue<NodeAndDistance>& knn_result,
                           int k) const {
        const float dis2 = dis2f(pt, cloud_[node->point_idx_]);
        if (static_cast<int>(knn_result.size()) < k) {
            knn_result.emplace(node, dis2);
        } else if (dis2 < knn_result.top().distance2_) {
            knn_result.emplace(node, dis2);
            knn_result.pop();
        }
    }
    // kdtree.cpp:214-236
    bool NeedExpand(const Vec3f& pt, KdTreeNode* node, std::priority_queue<NodeAndDistance>& knn_result, int k,
                    bool approximate, float alpha) const {
        if (static_cast<int>(knn_result.size()) < k) return true;
        const float d = pt[node->axis_index_] - node->split_thresh_;
        if (approximate) return (d * d) < knn_result.top().distance2_ * alpha;
        return (d * d) < knn_result.top().distance2_;
    }

    void KnnExact(const Vec3f& pt, const KdTreeNode* node, std::vector<IdxDist>& best, int k) const {
        if (node->IsLeaf()) {
            IdxDist c{dis2f(pt, cloud_[node->point_idx_]), node->point_idx_};
            if (static_cast<int>(best.size()) == k && !better(c, best.back())) return;
            auto it = std::upper_bound(best.begin(), best.end(), c, better);
            best.insert(it, c);
            if (static_cast<int>(best.size()) > k) best.pop_back();
            return;
        }
        const KdTreeNode *this_side, *that_side;
        if (pt[node->axis_index_] < node->split_thresh_) { this_side = node->left_; that_side = node->right_; }
        else { this_side = node->right_; that_side = node->left_; }
        KnnExact(pt, this_side, best, k);
        const float d = pt[node->axis_index_] - node->split_thresh_;
        if (static_cast<int>(best.size()) < k || !((d * d) > best.back().d2)) KnnExact(pt, that_side, best, k);
    }

    KdTreeNode* root_ = nullptr;
    std::vector<Vec3f> cloud_;
    std::unordered_map<int, KdTreeNode*> nodes_;  // bookkeeping, as in the reference (kdtree.h:122)
    size_t size_ = 0;
    int tree_node_id_ = 0;
};

// BfnnRegistration::FindNearstPoints (bfnn.cpp:24-50) with the (dis2, index) total order in place of the
// reference's unstable std::sort on dis2 alone.
static void bfnn_one(const float* map, size_t n, size_t stride, const Vec3f& q, int k, int32_t* out) {
    std::vector<IdxDist> best;
    best.reserve(k + 1);
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(map, i, stride);
        IdxDist c{dis2f(Vec3f{p[0], p[1], p[2]}, q), static_cast<int>(i)};
        if (static_cast<int>(best.size()) == k && !better(c, best.back())) continue;
        auto it = std::upper_bound(best.begin(), best.end(), c, better);
        best.insert(it, c);
        if (static_cast<int>(best.size()) > k) best.pop_back();
    }
    for (int j = 0; j < k; ++j) out[j] = j < static_cast<int>(best.size()) ? best[j].idx : -1;
}

// math::FitPlane (math_utils.h:112-136): right singular vector of the smallest singular value of
// [x y z 1] (n x 4), then every point must satisfy (n.p + d)^2 <= eps.
static bool FitPlane(const std::vector<Vec3>& data, double coeffs[4], double eps = 1e-2) {
    if (data.size() < 3) return false;
    const int n = static_cast<int>(data.size());
    std::vector<double> A(static_cast<size_t>(n) * 4);
    for (int i = 0; i < n; ++i) {
        A[i * 4 + 0] = data[i].x; A[i * 4 + 1] = data[i].y; A[i * 4 + 2] = data[i].z; A[i * 4 + 3] = 1.0;
    }
    std::vector<double> sigma, V;
    if (n >= 4) {
        jacobi_svd(A, n, 4, sigma, V);
    } else {
        // 3 x 4: thin V has 3 columns in Eigen; col(3) would be out of range. Pad with a zero row (n=3 is
        // unreachable from the P2Plane path, which requires nn.size() > 3).
        A.resize(16, 0.0);
        jacobi_svd(A, 4, 4, sigma, V);
    }
    for (int i = 0; i < 4; ++i) coeffs[i] = V[i * 4 + 3];
    for (int i = 0; i < n; ++i) {
        const double err = coeffs[0] * data[i].x + coeffs[1] * data[i].y + coeffs[2] * data[i].z + coeffs[3];
        if (err * err > eps) return false;
    }
    return true;
}

static inline bool finite3(const float* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }

// math::FitLine (math_utils.h:138-163): origin = mean of the points, dir = right singular vector of the LARGEST
// singular value of Y = data - origin (JacobiSVD(Y, ComputeFullV), V.col(0)); fails if any point is farther from
// the line than sqrt(eps) (|dir x (p - origin)|^2 > eps).
static bool FitLine(const std::vector<Vec3>& data, Vec3& origin, Vec3& dir, double eps) {
    if (data.size() < 2) return false;
    const int n = static_cast<int>(data.size());
    Vec3 sum{0, 0, 0};
    for (const Vec3& d : data) sum = sum + d;  // std::accumulate, in order
    origin = Vec3{sum.x / n, sum.y / n, sum.z / n};
    std::vector<double> Y(static_cast<size_t>(n) * 3);
    for (int i = 0; i < n; ++i) {
        Y[i * 3 + 0] = data[i].x - origin.x; Y[i * 3 + 1] = data[i].y - origin.y; Y[i * 3 + 2] = data[i].z - origin.z;
    }
    std::vector<double> sigma, V;
    if (n >= 3) {
        jacobi_svd(Y, n, 3, sigma, V);
    } else {
        Y.resize(9, 0.0);
        jacobi_svd(Y, 3, 3, sigma, V);
    }
    dir = Vec3{V[0 * 3 + 0], V[1 * 3 + 0], V[2 * 3 + 0]};  // singular values are sorted descending
    for (const Vec3& d : data) {
        const Vec3 c = cross(dir, d - origin);
        if (dot(c, c) > eps) return false;
    }
    return true;
}

// pcl::transformPointCloud(in, out, Matrix4f) as PCL 1.8 writes it: per coordinate
// m(r,0)*x + m(r,1)*y + m(r,2)*z + m(r,3), float32, left to right; non-finite points pass through.
static void transform_cloud(const float* src, size_t n, size_t stride, const SE3& T, float* out) {
    const Mat3 R = T.matrix();
    float m[3][4];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) m[r][c] = static_cast<float>(R.m[r][c]);
        m[r][3] = static_cast<float>(T.t[r]);
    }
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        float* o = pt_at(out, i, stride);
        if (o != p) std::memcpy(o, p, stride < 16 ? 12 : stride);
        if (!finite3(p)) continue;
        const float x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; ++r) o[r] = m[r][0] * x + m[r][1] * y + m[r][2] * z + m[r][3];
    }
}

// pcl::transformPointCloud<PointT, double> with a Matrix4d (lio.cpp:243,278: pose.matrix() is passed WITHOUT the
// .cast<float>() of ScanMatch): PCL 1.8 transforms.hpp evaluates m00*x + m01*y + m02*z + m03 in double, left to right,
// and casts once to float.
static void transform_cloud_d(const float* src, size_t n, size_t stride, const SE3& T, float* out) {
    const Mat3 R = T.matrix();
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        float* o = pt_at(out, i, stride);
        if (o != p) std::memcpy(o, p, stride < 16 ? 12 : stride);
        if (!finite3(p)) continue;
        const double x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; ++r) o[r] = static_cast<float>(R.m[r][0] * x + R.m[r][1] * y + R.m[r][2] * z + T.t[r]);
    }
}

// ---------------------------------------------------------------------------------------------
// IcpRegistration
// ---------------------------------------------------------------------------------------------
struct Icp {
    oracle_icp_options opt;
    std::vector<Vec3f> target;  // target_ (icp_registration.cpp:16)
    KdTree tree;                // kdtree_ptr_ (icp_registration.cpp:18)

    // KdtreeRegistration::FindNearstPoints (kdtree.cpp:272-283)
    void FindNearstPoints(const Vec3f& q, int k, int mode, std::vector<int>& out) const {
        switch (mode) {
            case ORACLE_NN_LITERAL_ANN: tree.GetClosestPoint(q, out, k, true, 0.1f); break;
            case ORACLE_NN_LITERAL_EXACT: tree.GetClosestPoint(q, out, k, false, 0.1f); break;
            case ORACLE_NN_EXACT_TIEBREAK: tree.GetClosestPointExact(q, out, k); break;
            default: {
                out.assign(k, -1);
                bfnn_one(&target[0].x, target.size(), sizeof(Vec3f), q, k, out.data());
                while (!out.empty() && out.back() < 0) out.pop_back();
            }
        }
    }

    // CaculateMatrixHAndBP2P (icp_registration.cpp:57-103)
    bool HB_P2P(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& B, oracle_result& res,
                uint8_t* gate, int32_t* nn_out) const {
        size_t effective_num = 0;
        double total_res = 0;
        const Mat3 R = pose.matrix();
        std::vector<int> nn;
        for (size_t i = 0; i < n; ++i) {
            const float* sp = pt_at(src, i, stride);
            if (gate) gate[i] = 0;
            if (nn_out) nn_out[i] = -1;
            if (!finite3(sp)) continue;  // pcl::isFinite (icp_registration.cpp:64)
            const Vec3 q{sp[0], sp[1], sp[2]};
            const Vec3 qs = pose * q;
            FindNearstPoints(Vec3f{static_cast<float>(qs.x), static_cast<float>(qs.y), static_cast<float>(qs.z)}, 1,
                             opt.nn_mode, nn);
            if (nn.empty()) continue;
            if (nn_out) nn_out[i] = nn[0];
            const Vec3f& pf = target[nn[0]];
            const Vec3 p{pf.x, pf.y, pf.z};
            const Vec3 e = p - qs;
            const double dis2 = dot(e, e);
            if (dis2 > opt.max_nn_distance) { if (gate) gate[i] = 2; continue; }  // quirk Q6: squared vs unsquared
            effective_num++;
            if (gate) gate[i] = 3;
            // J = [ R*hat(q)/16 , -I ]  (icp_registration.cpp:84-85)
            const Mat3 Rh = mul(R, hat(q));
            double J[3][6];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) { J[r][c] = Rh.m[r][c] / 16; J[r][3 + c] = (r == c) ? -1.0 : 0.0; }
            }
            const double ev[3] = {e.x, e.y, e.z};
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) H(a, b) += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
                B.v[a] += -(J[0][a] * ev[0] + J[1][a] * ev[1] + J[2][a] * ev[2]);
            }
            total_res += dot(e, e);
        }
        res.n_effective = static_cast<int64_t>(effective_num);
        res.n_inlier = static_cast<int64_t>(effective_num);
        res.sum_sq_res = total_res;
        if (effective_num < static_cast<size_t>(opt.min_effective_pts)) return false;
        if (lu6(H, nullptr) == 0) return false;
        return true;
    }

    // CaculateMatrixHAndBP2Plane (icp_registration.cpp:161-213)
    bool HB_P2Plane(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& B, oracle_result& res,
                    uint8_t* gate, int32_t* nn_out) const {
        size_t effective_num = 0, inliers = 0;
        double sum_sq = 0;
        const Mat3 R = pose.matrix();
        std::vector<int> nn;
        for (size_t i = 0; i < n; ++i) {
            const float* sp = pt_at(src, i, stride);
            if (gate) gate[i] = 0;
            if (nn_out) for (int j = 0; j < 5; ++j) nn_out[i * 5 + j] = -1;
            if (opt.skip_nonfinite && !finite3(sp)) continue;  // deviation D1 (the reference would poison H with NaN)
            const Vec3 q{sp[0], sp[1], sp[2]};
            const Vec3 qs = pose * q;
            FindNearstPoints(Vec3f{static_cast<float>(qs.x), static_cast<float>(qs.y), static_cast<float>(qs.z)}, 5,
                             opt.nn_mode, nn);
            if (nn_out) for (size_t j = 0; j < nn.size() && j < 5; ++j) nn_out[i * 5 + j] = nn[j];
            if (nn.size() > 3) {
                std::vector<Vec3> nn_eigen;
                for (size_t j = 0; j < nn.size(); ++j) nn_eigen.emplace_back(target[nn[j]].x, target[nn[j]].y, target[nn[j]].z);
                double nrm[4];
                if (!FitPlane(nn_eigen, nrm)) { if (gate) gate[i] = 1; continue; }
                effective_num++;  // quirk Q4: counted before the distance gate
                const Vec3 n3{nrm[0], nrm[1], nrm[2]};
                const double dis = dot(n3, qs) + nrm[3];
                if (std::fabs(dis) > opt.max_plane_distance) { if (gate) gate[i] = 2; continue; }
                if (gate) gate[i] = 3;
                // J = [ -n^T * R * hat(q) , n^T ]  (icp_registration.cpp:193-195), evaluated left to right
                const double nR[3] = {-(n3.x * R.m[0][0] + n3.y * R.m[1][0] + n3.z * R.m[2][0]),
                                      -(n3.x * R.m[0][1] + n3.y * R.m[1][1] + n3.z * R.m[2][1]),
                                      -(n3.x * R.m[0][2] + n3.y * R.m[1][2] + n3.z * R.m[2][2])};
                const Mat3 hq = hat(q);
                double J[6];
                for (int c = 0; c < 3; ++c) J[c] = nR[0] * hq.m[0][c] + nR[1] * hq.m[1][c] + nR[2] * hq.m[2][c];
                J[3] = n3.x; J[4] = n3.y; J[5] = n3.z;
                for (int a = 0; a < 6; ++a) {
                    for (int b = 0; b < 6; ++b) H(a, b) += J[a] * J[b];
                    B.v[a] += -J[a] * dis;
                }
                inliers++;
                sum_sq += dis * dis;
            }
        }
        res.n_effective = static_cast<int64_t>(effective_num);
        res.n_inlier = static_cast<int64_t>(inliers);
        res.sum_sq_res = sum_sq;
        if (effective_num < static_cast<size_t>(opt.min_effective_pts)) return false;
        if (lu6(H, nullptr) == 0) return false;
        return true;
    }

    // CaculateMatrixHAndBP2Line (icp_registration.cpp:105-159)
    bool HB_P2Line(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& B, oracle_result& res,
                   uint8_t* gate, int32_t* nn_out) const {
        size_t effective_num = 0, inliers = 0;
        double total_res = 0;
        const Mat3 R = pose.matrix();
        std::vector<int> nn;
        for (size_t i = 0; i < n; ++i) {
            const float* sp = pt_at(src, i, stride);
            if (gate) gate[i] = 0;
            if (nn_out) for (int j = 0; j < 5; ++j) nn_out[i * 5 + j] = -1;
            if (opt.skip_nonfinite && !finite3(sp)) continue;  // deviation D1
            const Vec3 q{sp[0], sp[1], sp[2]};
            const Vec3 qs = pose * q;
            FindNearstPoints(Vec3f{static_cast<float>(qs.x), static_cast<float>(qs.y), static_cast<float>(qs.z)}, 5,
                             opt.nn_mode, nn);
            if (nn_out) for (size_t j = 0; j < nn.size() && j < 5; ++j) nn_out[i * 5 + j] = nn[j];
            if (nn.size() != 5) continue;  // (:115)
            std::vector<Vec3> nn_eigen;
            for (int j = 0; j < 5; ++j) nn_eigen.emplace_back(target[nn[j]].x, target[nn[j]].y, target[nn[j]].z);
            Vec3 d, p0;
            if (!FitLine(nn_eigen, p0, d, opt.max_line_distance)) { if (gate) gate[i] = 1; continue; }  // (:123)
            effective_num++;
            const Vec3 e = cross(d, qs - p0);  // SO3::hat(d) * (qs - p0) (:130)
            if (std::sqrt(dot(e, e)) > opt.max_line_distance) { if (gate) gate[i] = 2; continue; }
            if (gate) gate[i] = 3;
            // J = [ -hat(d) R hat(q) , hat(d) ]  (:139-140)
            const Mat3 hd = hat(d);
            const Mat3 A = mul(mul(hd, R), hat(q));
            double J[3][6];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) { J[r][c] = -A.m[r][c]; J[r][3 + c] = hd.m[r][c]; }
            const double ev[3] = {e.x, e.y, e.z};
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) H(a, b) += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
                B.v[a] += -(J[0][a] * ev[0] + J[1][a] * ev[1] + J[2][a] * ev[2]);
            }
            inliers++;
            total_res += dot(e, e);
        }
        res.n_effective = static_cast<int64_t>(effective_num);
        res.n_inlier = static_cast<int64_t>(inliers);
        res.sum_sq_res = total_res;
        if (effective_num < static_cast<size_t>(opt.min_effective_pts)) return false;
        if (lu6(H, nullptr) == 0) return false;
        return true;
    }

    bool HB(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& B, oracle_result& res,
            uint8_t* gate, int32_t* nn_out) const {
        if (opt.method == ORACLE_ICP_P2P) return HB_P2P(src, n, stride, pose, H, B, res, gate, nn_out);
        if (opt.method == ORACLE_ICP_P2LINE) return HB_P2Line(src, n, stride, pose, H, B, res, gate, nn_out);
        return HB_P2Plane(src, n, stride, pose, H, B, res, gate, nn_out);
    }

    // AlignP2P (icp_registration.cpp:267-303) / AlignP2Line (:305-343) / AlignP2Plane (:345-381)
    void Align(const float* src, size_t n, size_t stride, const SE3& init, SE3& result, oracle_result& res,
               double* trace) const {
        SE3 pose = init;
        res = oracle_result{};
        res.pose_written = 1;
        if (trace) pose.to7(trace);
        for (int iter = 0; iter < opt.max_iteration; ++iter) {
            Mat6 H;
            Vec6 err;
            res.iters = iter + 1;
            const bool ok = HB(src, n, stride, pose, H, err, res, nullptr, nullptr);
            res.degenerate = ok ? 0 : 1;
            if (ok) {
                Mat6 Hinv;
                lu6(H, &Hinv);
                if (opt.method == ORACLE_ICP_P2P)
                    for (double& v : Hinv.a) v = v / 16;  // quirk Q6 (icp_registration.cpp:287)
                const Vec6 dx = mul(Hinv, err);
                pose.right_mul_exp(Vec3{dx.v[0], dx.v[1], dx.v[2]});
                pose.t = pose.t + Vec3{dx.v[3], dx.v[4], dx.v[5]};
                res.updates++;
                if (trace) pose.to7(trace + (iter + 1) * 7);
                if (dx.norm() < opt.eps) { res.converged = 1; break; }
            } else if (trace) {
                pose.to7(trace + (iter + 1) * 7);
            }
        }
        if (trace)
            for (int it = res.iters + 1; it <= opt.max_iteration; ++it) pose.to7(trace + it * 7);
        result = pose;
    }
};

// ---------------------------------------------------------------------------------------------
// NdtRegistration (direct)
// ---------------------------------------------------------------------------------------------
struct Key3 {
    int x, y, z;
    bool operator==(const Key3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct HashKey3 {  // hash_vec<3> (eigen_types.h:104-107): int arithmetic wraps, then % 10000000, then size_t
    size_t operator()(const Key3& v) const {
        const uint32_t a = static_cast<uint32_t>(v.x) * 73856093u;
        const uint32_t b = static_cast<uint32_t>(v.y) * 471943u;
        const uint32_t c = static_cast<uint32_t>(v.z) * 83492791u;
        const int32_t h = static_cast<int32_t>(a ^ b ^ c);
        return static_cast<size_t>(h % 10000000);
    }
};
struct NdtVoxelData {  // ndt_registration.hpp:46-67 (direct-NDT members only)
    std::vector<size_t> idx_;
    Vec3 mu_;
    Mat3 sigma_;
    Mat3 info_;
};

struct Ndt {
    oracle_ndt_options opt;
    double inv_voxel_size = 1.0;  // recomputed from voxel_size_ in the ctor (ndt_registration.cpp:25)
    std::vector<Vec3f> target;
    std::unordered_map<Key3, NdtVoxelData, HashKey3> grids;
    std::vector<Key3> nearby;

    void Init() {
        inv_voxel_size = 1.0 / opt.voxel_size;
        nearby.clear();
        // GenerateNearbyGrids (ndt_registration.cpp:51-63)
        if (!opt.nearby6) nearby.push_back({0, 0, 0});
        else nearby = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}};
    }
    Key3 KeyOf(const Vec3& p) const {  // (pt * inv_voxel_size_).cast<int>(): truncation toward zero (quirk Q9)
        return {static_cast<int>(p.x * inv_voxel_size), static_cast<int>(p.y * inv_voxel_size),
                static_cast<int>(p.z * inv_voxel_size)};
    }
    // SetDirectNdtTargetCloud (ndt_registration.cpp:87-148)
    void SetTarget(const float* xyz, size_t n, size_t stride) {
        target.resize(n);
        for (size_t i = 0; i < n; ++i) { const float* p = pt_at(xyz, i, stride); target[i] = {p[0], p[1], p[2]}; }
        grids.clear();
        for (size_t idx = 0; idx < n; ++idx) {
            if (opt.skip_nonfinite && !finite3(&target[idx].x)) continue;
            const Vec3 pt{target[idx].x, target[idx].y, target[idx].z};
            grids[KeyOf(pt)].idx_.emplace_back(idx);
        }
        for (auto it = grids.begin(); it != grids.end();) {
            NdtVoxelData& v = it->second;
            if (v.idx_.size() > static_cast<size_t>(opt.min_pts_in_voxel)) {
                // math::ComputeMeanAndCov (math_utils.h:55-72)
                const size_t len = v.idx_.size();
                Vec3 sum;
                for (size_t i : v.idx_) sum = sum + Vec3{target[i].x, target[i].y, target[i].z};
                v.mu_ = {sum.x / len, sum.y / len, sum.z / len};
                Mat3 cov;
                for (size_t i : v.idx_) {
                    const Vec3 d = Vec3{target[i].x, target[i].y, target[i].z} - v.mu_;
                    const double dv[3] = {d.x, d.y, d.z};
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) cov.m[r][c] = cov.m[r][c] + dv[r] * dv[c];
                }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) cov.m[r][c] = cov.m[r][c] / (len - 1);
                v.sigma_ = cov;
                // JacobiSVD(sigma, FullU|FullV); clamp; info = V * diag(1/lambda) * U^T (ndt_registration.cpp:118-130)
                std::vector<double> A(9), sig, V, U;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) A[r * 3 + c] = cov.m[r][c];
                jacobi_svd(A, 3, 3, sig, V, &U);
                // a numerically zero singular value has no defined left vector (Eigen returns noise there):
                // take u = v, i.e. treat the PSD covariance as exactly symmetric
                for (int j = 0; j < 3; ++j)
                    if (!(sig[j] > 1e-12 * sig[0]))
                        for (int r = 0; r < 3; ++r) U[r * 3 + j] = V[r * 3 + j];
                double lambda[3] = {sig[0], sig[1], sig[2]};
                if (lambda[1] < lambda[0] * 1e-3) lambda[1] = lambda[0] * 1e-3;
                if (lambda[2] < lambda[0] * 1e-3) lambda[2] = lambda[0] * 1e-3;
                const double inv_l[3] = {1.0 / lambda[0], 1.0 / lambda[1], 1.0 / lambda[2]};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += V[r * 3 + k] * inv_l[k] * U[c * 3 + k];
                        v.info_.m[r][c] = s;
                    }
                ++it;
            } else {
                it = grids.erase(it);  // ndt_registration.cpp:136-142
            }
        }
    }

    // Loop body of AlignNdt (ndt_registration.cpp:399-433); hits[i] = number of gated-in voxels of point i
    void HB(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& err, oracle_result& res,
            uint8_t* hits) const {
        size_t effective_num = 0, inl = 0;
        double total_res = 0;
        const Mat3 R = pose.matrix();
        for (size_t i = 0; i < n; ++i) {
            const float* sp = pt_at(src, i, stride);
            if (hits) hits[i] = 0;
            if (opt.skip_nonfinite && !finite3(sp)) continue;
            const Vec3 q{sp[0], sp[1], sp[2]};
            const Vec3 qs = pose * q;
            const Key3 key = KeyOf(qs);
            for (const Key3& off : nearby) {
                const Key3 k{key.x + off.x, key.y + off.y, key.z + off.z};
                auto it = grids.find(k);
                if (it == grids.end()) continue;
                const NdtVoxelData& v = it->second;
                const Vec3 e = qs - v.mu_;
                const Vec3 ie = mul(v.info_, e);
                const double r2 = dot(e, ie);
                if (std::isnan(r2) || r2 > opt.res_outlier_th) continue;
                // J = [ -R*hat(q) , I ]; H += J^T J; err += -J^T e  — info NOT applied (quirk Q8)
                const Mat3 Rh = mul(R, hat(q));
                double J[3][6];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) { J[r][c] = -Rh.m[r][c]; J[r][3 + c] = (r == c) ? 1.0 : 0.0; }
                const double ev[3] = {e.x, e.y, e.z};
                for (int a = 0; a < 6; ++a) {
                    for (int b = 0; b < 6; ++b) H(a, b) += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
                    err.v[a] += -(J[0][a] * ev[0] + J[1][a] * ev[1] + J[2][a] * ev[2]);
                }
                total_res += dot(e, e);
                inl++;
                if (hits) hits[i]++;
            }
            effective_num++;  // per point, unconditionally (ndt_registration.cpp:432)
        }
        res.n_effective = static_cast<int64_t>(effective_num);
        res.n_inlier = static_cast<int64_t>(inl);
        res.sum_sq_res = total_res;
    }

    // AlignNdt (ndt_registration.cpp:374-464).  Returns false on the det(H)==0 early return, in which case
    // result is NOT written (quirk Q11).
    bool Align(const float* src, size_t n, size_t stride, const SE3& init, SE3& result, oracle_result& res,
               double* trace) const {
        SE3 pose = init;
        res = oracle_result{};
        if (trace) pose.to7(trace);
        for (int iter = 0; iter < opt.max_iteration; ++iter) {
            Mat6 H;
            Vec6 err;
            res.iters = iter + 1;
            HB(src, n, stride, pose, H, err, res, nullptr);
            if (lu6(H, nullptr) == 0) { res.degenerate = 1; res.pose_written = 0; return false; }
            if (res.n_effective < opt.min_effective_pts) {
                res.degenerate = 1;
                if (trace) pose.to7(trace + (iter + 1) * 7);
                continue;
            }
            res.degenerate = 0;
            Mat6 Hinv;
            lu6(H, &Hinv);
            const Vec6 dx = mul(Hinv, err);
            pose.right_mul_exp(Vec3{dx.v[0], dx.v[1], dx.v[2]});
            pose.t = pose.t + Vec3{dx.v[3], dx.v[4], dx.v[5]};
            res.updates++;
            if (trace) pose.to7(trace + (iter + 1) * 7);
            if (dx.norm() < opt.eps) { res.converged = 1; break; }
        }
        if (trace)
            for (int it = res.iters + 1; it <= opt.max_iteration; ++it) pose.to7(trace + it * 7);
        result = pose;
        res.pose_written = 1;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// NdtRegistration (incremental): SetIncNdtTargetCloud / UpdateVoxel / AlignIncNdt
// (ndt_registration.cpp:150-236, 262-372), literally: std::list LRU + unordered_map of list iterators.
// ---------------------------------------------------------------------------------------------
struct IncVoxel {  // NdtVoxelData members the incremental path uses (ndt_registration.hpp:46-67)
    std::vector<Vec3> pts_;
    bool ndt_estimated_ = false;
    int num_pts_ = 0;
    int n_last = 0;  // (oracle only) points of the last update, for the parity probe
    Vec3 mu_;
    Mat3 sigma_;
    Mat3 info_;
};
struct IncNdt {
    oracle_ndt_options opt;
    size_t capacity = 100000;  // NdtOptions::capacity_
    double inv_voxel_size = 1.0;
    using KeyAndData = std::pair<Key3, IncVoxel>;
    std::list<KeyAndData> data_;
    std::unordered_map<Key3, std::list<KeyAndData>::iterator, HashKey3> inc_grids_;
    std::vector<Key3> nearby;
    bool flag_first_scan_ = true;

    void Init() {
        inv_voxel_size = 1.0 / opt.voxel_size;
        nearby.clear();
        if (!opt.nearby6) nearby.push_back({0, 0, 0});
        else nearby = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}};
    }
    Key3 KeyOf(const Vec3& p) const {
        return {static_cast<int>(p.x * inv_voxel_size), static_cast<int>(p.y * inv_voxel_size),
                static_cast<int>(p.z * inv_voxel_size)};
    }
    static Mat3 Inverse3(const Mat3& A) {  // Eigen's fixed-size 3x3 inverse: cofactors / determinant
        Mat3 C;
        C.m[0][0] = A.m[1][1] * A.m[2][2] - A.m[1][2] * A.m[2][1];
        C.m[0][1] = A.m[0][2] * A.m[2][1] - A.m[0][1] * A.m[2][2];
        C.m[0][2] = A.m[0][1] * A.m[1][2] - A.m[0][2] * A.m[1][1];
        C.m[1][0] = A.m[1][2] * A.m[2][0] - A.m[1][0] * A.m[2][2];
        C.m[1][1] = A.m[0][0] * A.m[2][2] - A.m[0][2] * A.m[2][0];
        C.m[1][2] = A.m[0][2] * A.m[1][0] - A.m[0][0] * A.m[1][2];
        C.m[2][0] = A.m[1][0] * A.m[2][1] - A.m[1][1] * A.m[2][0];
        C.m[2][1] = A.m[0][1] * A.m[2][0] - A.m[0][0] * A.m[2][1];
        C.m[2][2] = A.m[0][0] * A.m[1][1] - A.m[0][1] * A.m[1][0];
        const double det = A.m[0][0] * C.m[0][0] + A.m[0][1] * C.m[1][0] + A.m[0][2] * C.m[2][0];
        for (auto& row : C.m) for (double& v : row) v = v / det;
        return C;
    }
    // UpdateVoxel (:185-236).  flag_first_scan_ is true whenever this runs (it is set back to true at :181), so only
    // the first branch is reachable; the others are kept out rather than restated untested.
    void UpdateVoxel(IncVoxel& v) {
        if (flag_first_scan_) {
            if (v.pts_.size() > 1) {
                Vec3 sum{0, 0, 0};
                for (const Vec3& p : v.pts_) sum = sum + p;
                const double len = static_cast<double>(v.pts_.size());
                v.mu_ = Vec3{sum.x / len, sum.y / len, sum.z / len};
                Mat3 cov;
                for (auto& row : cov.m) for (double& x : row) x = 0;
                for (const Vec3& p : v.pts_) {
                    const Vec3 d = p - v.mu_;
                    const double dv[3] = {d.x, d.y, d.z};
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) cov.m[r][c] = cov.m[r][c] + dv[r] * dv[c];
                }
                for (auto& row : cov.m) for (double& x : row) x = x / (len - 1);
                v.sigma_ = cov;
                Mat3 A = cov;
                for (int k = 0; k < 3; ++k) A.m[k][k] += 1e-3;
                v.info_ = Inverse3(A);  // (:189)
            } else {
                v.mu_ = v.pts_[0];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) v.info_.m[r][c] = (r == c) ? 1e2 : 0.0;
            }
            v.ndt_estimated_ = true;
            v.n_last = static_cast<int>(v.pts_.size());
            v.pts_.clear();
            return;
        }
    }
    // SetIncNdtTargetCloud (:150-183)
    void AddCloud(const float* xyz, size_t n, size_t stride) {
        auto less3 = [](const Key3& a, const Key3& b) {
            return a.x < b.x || (a.x == b.x && a.y < b.y) || (a.x == b.x && a.y == b.y && a.z < b.z);
        };
        std::set<Key3, decltype(less3)> active_voxels(less3);
        for (size_t i = 0; i < n; ++i) {
            const float* p = pt_at(xyz, i, stride);
            if (opt.skip_nonfinite && !finite3(p)) continue;
            const Vec3 pt{p[0], p[1], p[2]};
            const Key3 key = KeyOf(pt);
            auto iter = inc_grids_.find(key);
            if (iter == inc_grids_.end()) {
                IncVoxel v;
                v.pts_.emplace_back(pt);
                v.num_pts_ = 1;
                data_.push_front({key, v});
                inc_grids_.insert({key, data_.begin()});
                if (data_.size() >= capacity) {
                    inc_grids_.erase(data_.back().first);
                    data_.pop_back();
                }
            } else {
                IncVoxel& v = iter->second->second;
                v.pts_.emplace_back(pt);
                if (!v.ndt_estimated_) v.num_pts_++;
                data_.splice(data_.begin(), data_, iter->second);
                iter->second = data_.begin();
            }
            active_voxels.emplace(key);
        }
        for (const Key3& key : active_voxels) {
            auto it = inc_grids_.find(key);  // the reference's operator[] would insert a null iterator for an evicted key
            if (it == inc_grids_.end() || it->second->second.pts_.empty()) continue;
            UpdateVoxel(it->second->second);
        }
        flag_first_scan_ = true;
    }
    // the two loops of one AlignIncNdt iteration (:289-347); hits[i] = gated-in voxels of point i
    void HB(const float* src, size_t n, size_t stride, const SE3& pose, Mat6& H, Vec6& err, oracle_result& res, uint8_t* hits) const {
        int effective_num = 0;
        double total_res = 0;
        const Mat3 R = pose.matrix();
        for (size_t i = 0; i < n; ++i) {
            const float* sp = pt_at(src, i, stride);
            if (hits) hits[i] = 0;
            if (opt.skip_nonfinite && !finite3(sp)) continue;
            const Vec3 q{sp[0], sp[1], sp[2]};
            const Vec3 qs = pose * q;
            const Key3 key = KeyOf(qs);
            for (const Key3& off : nearby) {
                const Key3 k{key.x + off.x, key.y + off.y, key.z + off.z};
                auto it = inc_grids_.find(k);
                if (it == inc_grids_.end() || !it->second->second.ndt_estimated_) continue;
                const IncVoxel& v = it->second->second;
                const Vec3 e = qs - v.mu_;
                const Vec3 ie = mul(v.info_, e);
                const double r2 = dot(e, ie);
                if (std::isnan(r2) || r2 > opt.res_outlier_th) continue;
                const Mat3 Rh = mul(R, hat(q));
                double J[3][6];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) { J[r][c] = -Rh.m[r][c]; J[r][3 + c] = (r == c) ? 1.0 : 0.0; }
                // H += J^T info J; err += -J^T info e  (:345-346)
                double IJ[3][6];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 6; ++c) IJ[r][c] = v.info_.m[r][0] * J[0][c] + v.info_.m[r][1] * J[1][c] + v.info_.m[r][2] * J[2][c];
                const double iev[3] = {ie.x, ie.y, ie.z};
                for (int a = 0; a < 6; ++a) {
                    for (int b = 0; b < 6; ++b) H(a, b) += J[0][a] * IJ[0][b] + J[1][a] * IJ[1][b] + J[2][a] * IJ[2][b];
                    err.v[a] += -(J[0][a] * iev[0] + J[1][a] * iev[1] + J[2][a] * iev[2]);
                }
                total_res += r2;
                effective_num++;
                if (hits) hits[i]++;
            }
        }
        res.n_effective = effective_num;
        res.n_inlier = effective_num;
        res.sum_sq_res = total_res;
    }
    // AlignIncNdt (:262-372)
    bool Align(const float* src, size_t n, size_t stride, const SE3& init, SE3& result, oracle_result& res, double* trace) const {
        SE3 pose = init;
        res = oracle_result{};
        res.pose_written = 1;
        if (trace) pose.to7(trace);
        bool ok = true;
        for (int iter = 0; iter < opt.max_iteration; ++iter) {
            Mat6 H;
            Vec6 err;
            res.iters = iter + 1;
            HB(src, n, stride, pose, H, err, res, nullptr);
            if (res.n_effective < opt.min_effective_pts) { res.degenerate = 1; ok = false; break; }  // result_pose = pose; return false
            res.degenerate = 0;
            Mat6 Hinv;
            lu6(H, &Hinv);
            const Vec6 dx = mul(Hinv, err);
            pose.right_mul_exp(Vec3{dx.v[0], dx.v[1], dx.v[2]});
            pose.t = pose.t + Vec3{dx.v[3], dx.v[4], dx.v[5]};
            res.updates++;
            if (trace) pose.to7(trace + (iter + 1) * 7);
            if (dx.norm() < opt.eps) { res.converged = 1; break; }
        }
        if (trace)
            for (int it = res.iters + (ok && !res.converged ? 1 : 0); it <= opt.max_iteration; ++it) pose.to7(trace + it * 7);
        result = pose;
        return ok;
    }
};

}  // namespace oracle

using namespace oracle;

struct oracle_icp { Icp impl; };
struct oracle_ndt { Ndt impl; };
struct oracle_inc_ndt { IncNdt impl; };

extern "C" {

void oracle_icp_default_options(oracle_icp_options* o) {
    o->max_iteration = 20; o->max_nn_distance = 1.0; o->max_plane_distance = 0.1; o->max_line_distance = 0.5;
    o->min_effective_pts = 10; o->eps = 1e-2; o->method = ORACLE_ICP_P2P; o->nn_mode = ORACLE_NN_LITERAL_ANN;
    o->skip_nonfinite = 0;
}
void oracle_ndt_default_options(oracle_ndt_options* o) {
    o->max_iteration = 20; o->voxel_size = 1.0; o->min_effective_pts = 10; o->min_pts_in_voxel = 3; o->eps = 1e-2;
    o->res_outlier_th = 20.0; o->nearby6 = 1; o->skip_nonfinite = 0;
}

oracle_icp* oracle_icp_create(const oracle_icp_options* o) {
    auto* h = new oracle_icp;
    h->impl.opt = *o;
    return h;
}
void oracle_icp_destroy(oracle_icp* h) { delete h; }
int oracle_icp_set_target(oracle_icp* h, const float* xyz, size_t n, size_t stride) {
    h->impl.target.resize(n);
    for (size_t i = 0; i < n; ++i) { const float* p = pt_at(xyz, i, stride); h->impl.target[i] = {p[0], p[1], p[2]}; }
    return h->impl.tree.BuildTree(h->impl.target) ? 0 : -1;
}
size_t oracle_icp_tree_leaves(const oracle_icp* h) { return h->impl.tree.size(); }
int oracle_icp_knn(oracle_icp* h, const float* q, size_t nq, size_t stride, int k, int nn_mode, int32_t* idx_out) {
    std::vector<int> nn;
    for (size_t i = 0; i < nq; ++i) {
        const float* p = pt_at(q, i, stride);
        h->impl.FindNearstPoints(Vec3f{p[0], p[1], p[2]}, k, nn_mode, nn);
        for (int j = 0; j < k; ++j) idx_out[i * k + j] = j < static_cast<int>(nn.size()) ? nn[j] : -1;
    }
    return 0;
}
int oracle_icp_compute_hb(oracle_icp* h, const float* src, size_t n, size_t stride, const double* pose7, double* H36,
                          double* B6, oracle_result* res, uint8_t* gate, int32_t* nn_out) {
    Mat6 H;
    Vec6 B;
    oracle_result r{};
    const bool ok = h->impl.HB(src, n, stride, SE3::from7(pose7), H, B, r, gate, nn_out);
    r.degenerate = ok ? 0 : 1;
    std::memcpy(H36, H.a, sizeof(H.a));
    std::memcpy(B6, B.v, sizeof(B.v));
    if (res) *res = r;
    return ok ? 1 : 0;
}
int oracle_icp_align(oracle_icp* h, const float* src, size_t n, size_t stride, const double* pose_in, double* pose_out,
                     float* out_xyz, oracle_result* res, double* trace) {
    SE3 result;
    oracle_result r{};
    h->impl.Align(src, n, stride, SE3::from7(pose_in), result, r, trace);
    result.to7(pose_out);
    if (out_xyz) transform_cloud(src, n, stride, result, out_xyz);
    if (res) *res = r;
    return 1;  // ScanMatch always returns true (icp_registration.cpp:243)
}
// S independent ScanMatch calls spread over `threads` host threads.  Each call is the reference's own
// single-threaded loop (the reference has no parallel path); the kd-tree is shared read-only.
int oracle_icp_align_batch(oracle_icp* h, const float* srcs, const int64_t* offsets, size_t stride, const double* poses_in,
                           size_t S, double* poses_out, oracle_result* results, int threads) {
    if (threads <= 0) threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    threads = static_cast<int>(std::min<size_t>(threads, std::max<size_t>(S, 1)));
    auto work = [&](int tid) {
        for (size_t s = tid; s < S; s += threads) {
            const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(srcs) + offsets[s] * stride);
            SE3 result;
            oracle_result r{};
            h->impl.Align(src, static_cast<size_t>(offsets[s + 1] - offsets[s]), stride, SE3::from7(poses_in + s * 7), result, r, nullptr);
            result.to7(poses_out + s * 7);
            if (results) results[s] = r;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    return threads;
}
int oracle_fit_plane(const double* pts, int n, double* coeffs4, double eps) {
    std::vector<Vec3> d;
    for (int i = 0; i < n; ++i) d.emplace_back(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
    return FitPlane(d, coeffs4, eps) ? 1 : 0;
}
int oracle_bfnn(const float* map, size_t n, size_t map_stride, const float* q, size_t nq, size_t q_stride, int k,
                int32_t* idx_out) {
    for (size_t i = 0; i < nq; ++i) {
        const float* p = pt_at(q, i, q_stride);
        bfnn_one(map, n, map_stride, Vec3f{p[0], p[1], p[2]}, k, idx_out + i * k);
    }
    return 0;
}

oracle_ndt* oracle_ndt_create(const oracle_ndt_options* o) {
    auto* h = new oracle_ndt;
    h->impl.opt = *o;
    h->impl.Init();
    return h;
}
void oracle_ndt_destroy(oracle_ndt* h) { delete h; }
int oracle_ndt_set_target(oracle_ndt* h, const float* xyz, size_t n, size_t stride) {
    h->impl.SetTarget(xyz, n, stride);
    return 0;
}
size_t oracle_ndt_num_voxels(const oracle_ndt* h) { return h->impl.grids.size(); }
int oracle_ndt_get_voxels(const oracle_ndt* h, int32_t* keys, double* mu, double* info, int32_t* npts) {
    std::vector<const std::pair<const Key3, NdtVoxelData>*> v;
    for (auto& kv : h->impl.grids) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](auto* a, auto* b) {
        if (a->first.x != b->first.x) return a->first.x < b->first.x;
        if (a->first.y != b->first.y) return a->first.y < b->first.y;
        return a->first.z < b->first.z;
    });
    for (size_t i = 0; i < v.size(); ++i) {
        keys[i * 3] = v[i]->first.x; keys[i * 3 + 1] = v[i]->first.y; keys[i * 3 + 2] = v[i]->first.z;
        mu[i * 3] = v[i]->second.mu_.x; mu[i * 3 + 1] = v[i]->second.mu_.y; mu[i * 3 + 2] = v[i]->second.mu_.z;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) info[i * 9 + r * 3 + c] = v[i]->second.info_.m[r][c];
        if (npts) npts[i] = static_cast<int32_t>(v[i]->second.idx_.size());
    }
    return 0;
}
int oracle_ndt_compute_hb(oracle_ndt* h, const float* src, size_t n, size_t stride, const double* pose7, double* H36,
                          double* B6, oracle_result* res, uint8_t* hits) {
    Mat6 H;
    Vec6 B;
    oracle_result r{};
    h->impl.HB(src, n, stride, SE3::from7(pose7), H, B, r, hits);
    std::memcpy(H36, H.a, sizeof(H.a));
    std::memcpy(B6, B.v, sizeof(B.v));
    r.degenerate = (lu6(H, nullptr) == 0) ? 1 : 0;
    if (res) *res = r;
    return 1;
}
int oracle_ndt_align(oracle_ndt* h, const float* src, size_t n, size_t stride, const double* pose_in,
                     double* pose_inout, float* out_xyz, oracle_result* res, double* trace) {
    SE3 result = SE3::from7(pose_inout);  // caller's value survives the early return (quirk Q11)
    oracle_result r{};
    h->impl.Align(src, n, stride, SE3::from7(pose_in), result, r, trace);
    result.to7(pose_inout);
    if (out_xyz) transform_cloud(src, n, stride, result, out_xyz);
    if (res) *res = r;
    return 1;  // ScanMatch always returns true (ndt_registration.cpp:260)
}

/* ---- incremental NDT ---- */
oracle_inc_ndt* oracle_inc_ndt_create(const oracle_ndt_options* o, size_t capacity) {
    auto* h = new oracle_inc_ndt;
    h->impl.opt = *o;
    h->impl.capacity = capacity;
    h->impl.Init();
    return h;
}
void oracle_inc_ndt_destroy(oracle_inc_ndt* h) { delete h; }
int oracle_inc_ndt_add_cloud(oracle_inc_ndt* h, const float* xyz, size_t n, size_t stride) {
    h->impl.AddCloud(xyz, n, stride);
    return 0;
}
size_t oracle_inc_ndt_num_voxels(const oracle_inc_ndt* h) { return h->impl.data_.size(); }
int oracle_inc_ndt_get_voxels(const oracle_inc_ndt* h, int32_t* keys, double* mu, double* info, int32_t* npts) {
    std::vector<const IncNdt::KeyAndData*> v;
    for (auto& kv : h->impl.data_) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](auto* a, auto* b) {
        if (a->first.x != b->first.x) return a->first.x < b->first.x;
        if (a->first.y != b->first.y) return a->first.y < b->first.y;
        return a->first.z < b->first.z;
    });
    for (size_t i = 0; i < v.size(); ++i) {
        keys[i * 3] = v[i]->first.x; keys[i * 3 + 1] = v[i]->first.y; keys[i * 3 + 2] = v[i]->first.z;
        mu[i * 3] = v[i]->second.mu_.x; mu[i * 3 + 1] = v[i]->second.mu_.y; mu[i * 3 + 2] = v[i]->second.mu_.z;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) info[i * 9 + r * 3 + c] = v[i]->second.info_.m[r][c];
        if (npts) npts[i] = v[i]->second.n_last;
    }
    return 0;
}
int oracle_inc_ndt_compute_hb(oracle_inc_ndt* h, const float* src, size_t n, size_t stride, const double* pose7, double* H36,
                              double* B6, oracle_result* res, uint8_t* hits) {
    Mat6 H;
    Vec6 B;
    oracle_result r{};
    h->impl.HB(src, n, stride, SE3::from7(pose7), H, B, r, hits);
    std::memcpy(H36, H.a, sizeof(H.a));
    std::memcpy(B6, B.v, sizeof(B.v));
    r.degenerate = r.n_effective < h->impl.opt.min_effective_pts ? 1 : 0;
    r.pose_written = 1;
    if (res) *res = r;
    return 1;
}
int oracle_inc_ndt_align(oracle_inc_ndt* h, const float* src, size_t n, size_t stride, const double* pose_in, double* pose_out,
                         float* out_xyz, oracle_result* res, double* trace) {
    SE3 result;
    oracle_result r{};
    h->impl.Align(src, n, stride, SE3::from7(pose_in), result, r, trace);
    result.to7(pose_out);
    if (out_xyz) transform_cloud(src, n, stride, result, out_xyz);
    if (res) *res = r;
    return 1;
}

/* ---- cloud pre-filters (PCL 1.8 algorithms restated; PCL is not under /root/reference) ---- */
/* pcl::removeNaNFromPointCloud (filter.hpp): finite points, order kept */
size_t oracle_filter_remove_nan(const float* src, size_t n, size_t stride, float* out) {
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        if (!finite3(p)) continue;
        std::memcpy(pt_at(out, k++, stride), p, stride);
    }
    return k;
}
/* pcl::CropBox, identity transform, negative = false (crop_box.hpp): keep min <= p <= max */
size_t oracle_filter_crop_box(const float* src, size_t n, size_t stride, const float* mn, const float* mx, float* out) {
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        if (!finite3(p)) continue;
        if (p[0] < mn[0] || p[1] < mn[1] || p[2] < mn[2] || p[0] > mx[0] || p[1] > mx[1] || p[2] > mx[2]) continue;
        std::memcpy(pt_at(out, k++, stride), p, stride);
    }
    return k;
}
/* pcl::VoxelGrid::applyFilter (voxel_grid.hpp), cubic leaf, downsample_all_data: voxel index from float floor(p *
 * inverse_leaf) - min_b, points sorted by voxel index (stable here: PCL's std::sort leaves the order inside a voxel
 * unspecified), float32 mean of every float word, voxels in ascending index.  When the index space exceeds an int
 * ("Leaf size is too small for the input dataset. Integer indices would overflow.") PCL returns the cloud unfiltered. */
size_t oracle_filter_voxel_grid(const float* src, size_t n, size_t stride, float leaf, float* out) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        if (!finite3(p)) continue;
        any = true;
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    }
    if (!any) return 0;
    const float inv = 1.0f / leaf;
    int min_b[3];
    long long div[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
        div[a] = static_cast<long long>(static_cast<int>(std::floor(mx[a] * inv))) - min_b[a] + 1;
    }
    if (div[0] * div[1] * div[2] > 0x7fffffffll) {  // output = *input
        for (size_t i = 0; i < n; ++i) std::memcpy(pt_at(out, i, stride), pt_at(src, i, stride), stride);
        return n;
    }
    const int mul[3] = {1, static_cast<int>(div[0]), static_cast<int>(div[0] * div[1])};
    std::vector<std::pair<unsigned int, unsigned int>> iv;  // (voxel index, point index)
    for (size_t i = 0; i < n; ++i) {
        const float* p = pt_at(src, i, stride);
        if (!finite3(p)) continue;
        const int i0 = static_cast<int>(std::floor(p[0] * inv) - static_cast<float>(min_b[0]));
        const int i1 = static_cast<int>(std::floor(p[1] * inv) - static_cast<float>(min_b[1]));
        const int i2 = static_cast<int>(std::floor(p[2] * inv) - static_cast<float>(min_b[2]));
        iv.emplace_back(static_cast<unsigned int>(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), static_cast<unsigned int>(i));
    }
    std::stable_sort(iv.begin(), iv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    const size_t words = std::min<size_t>(stride / 4, 8);
    size_t k = 0;
    for (size_t a = 0; a < iv.size();) {
        size_t b = a;
        float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        while (b < iv.size() && iv[b].first == iv[a].first) {
            const float* p = pt_at(src, iv[b].second, stride);
            for (size_t w = 0; w < words; ++w) sum[w] = sum[w] + p[w];
            ++b;
        }
        float* o = pt_at(out, k++, stride);
        const float fn = static_cast<float>(b - a);
        for (size_t w = 0; w < stride / 4; ++w) o[w] = w < words ? sum[w] / fn : 0.0f;
        a = b;
    }
    return k;
}

void oracle_transform_cloud(const float* src, size_t n, size_t stride, const double* pose7, float* out_xyz) {
    transform_cloud(src, n, stride, SE3::from7(pose7), out_xyz);
}
void oracle_transform_cloud_d(const float* src, size_t n, size_t stride, const double* pose7, float* out_xyz) {
    transform_cloud_d(src, n, stride, SE3::from7(pose7), out_xyz);
}
void oracle_pose_update(double* pose7, const double* dx6) {
    SE3 T = SE3::from7(pose7);
    T.right_mul_exp(Vec3{dx6[0], dx6[1], dx6[2]});
    T.t = T.t + Vec3{dx6[3], dx6[4], dx6[5]};
    T.to7(pose7);
}
void oracle_pose_matrix(const double* pose7, double* R9) {
    const Mat3 R = SE3::from7(pose7).matrix();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R9[r * 3 + c] = R.m[r][c];
}

}  // extern "C"

// TEST-ONLY: compiles include/locreg_adapter.hpp against stand-ins for the PCL / Sophus / Eigen types it touches
// (none of those libraries exist in this image) and drives it the way Loc::Update / Lio::AddCloud drive a
// MatchingInterface: SetInputTarget(map) then ScanMatch(scan, predict, out_cloud, out_pose).
// Exit code 0: ran on a GPU and recovered the pose; 3: no CUDA device (the adapter threw, as designed: no CPU
// fallback); anything else: failure.  The stand-in types mirror the members listed in the adapter's header comment.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

namespace LocUtils {
struct alignas(16) PointType {  // pcl::PointXYZI: x y z pad | intensity pad pad pad = 32 bytes
    float x, y, z, pad0;
    float intensity, pad1, pad2, pad3;
};
static_assert(sizeof(PointType) == 32, "pcl::PointXYZI stride");
struct Header { uint32_t seq = 0; uint64_t stamp = 0; };
struct PointCloudType {
    Header header;
    std::vector<PointType> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
};
using CloudPtr = std::shared_ptr<PointCloudType>;
struct SE3 {  // Sophus::SE3d memory layout: unit quaternion (x y z w) then translation
    double v[7] = {0, 0, 0, 1, 0, 0, 0};
    double* data() { return v; }
    const double* data() const { return v; }
};
struct Mat6d { double v[36]; double* data() { return v; } };
struct Vec6d { double v[6]; double* data() { return v; } };

// shape of LocUtils::MatchingInterface (matching_interface.h:13-54)
class MatchingInterface {
   public:
    virtual ~MatchingInterface() = default;
    virtual bool SetInputTarget(const CloudPtr&) { return true; }
    virtual bool CaculateMatrixHAndB(const CloudPtr&, const SE3&, Mat6d&, Vec6d&) { return true; }
    virtual bool ScanMatch(const CloudPtr&, const SE3&, CloudPtr&, SE3&) { return true; }
    virtual bool SetInputTarget(const CloudPtr&, const CloudPtr&) { return true; }
    virtual bool ScanMatch(const CloudPtr&, const CloudPtr&, const SE3&, CloudPtr&, SE3&) { return true; }
    virtual float GetFitnessScore() = 0;
};
enum class IcpMethod { P2P, P2LINE, P2PLANE, PCLICP };
struct IcpOptions {
    int max_iteration_ = 20;
    double max_nn_distance_ = 1.0, max_plane_distance_ = 0.1, max_line_distance_ = 0.5;
    int min_effective_pts_ = 10;
    double eps_ = 1e-2, euc_fitness_eps_ = 0.36;
    bool use_initial_translation_ = true, use_ann = false;
    IcpMethod method_ = IcpMethod::P2P;
};
enum class NdtNearbyType { CENTER, NEARBY6 };
enum class NdtMethod { PCL_NDT, DIRECT_NDT, INCREMENTAL_NDT };
struct NdtOptions {
    int max_iteration_ = 20;
    double voxel_size_ = 1.0, inv_voxel_size_ = 1.0;
    int min_effective_pts_ = 10, min_pts_in_voxel_ = 3, max_pts_in_voxel_ = 50;
    double eps_ = 1e-2, res_outlier_th_ = 20.0;
    bool remove_centroid_ = false;
    size_t capacity_ = 100000;
    NdtNearbyType nearby_type_ = NdtNearbyType::NEARBY6;
    NdtMethod method_ = NdtMethod::DIRECT_NDT;
};
}  // namespace LocUtils

#define LOCREG_ADAPTER_STANDIN 1
#include "../../include/locreg_adapter.hpp"

using namespace LocUtils;

// three mutually orthogonal noiseless planes (a room corner), sampled on a 0.1 m lattice
static CloudPtr make_room() {
    auto c = std::make_shared<PointCloudType>();
    for (int i = 0; i < 60; ++i)
        for (int j = 0; j < 60; ++j) {
            const float a = 0.1f * i, b = 0.1f * j;
            c->points.push_back(PointType{a, b, 0, 0, 1, 0, 0, 0});
            c->points.push_back(PointType{a, 0, b, 0, 2, 0, 0, 0});
            c->points.push_back(PointType{0, a, b, 0, 3, 0, 0, 0});
        }
    c->width = static_cast<uint32_t>(c->points.size());
    return c;
}

int main() {
    std::shared_ptr<MatchingInterface> match;
    try {
        IcpOptions opt;
        opt.method_ = IcpMethod::P2PLANE;
        match = std::make_shared<CudaIcpRegistration>(opt);  // the one line Loc / Lio change (loc.cpp:41, lio.cpp:30)
    } catch (const std::exception& e) {
        std::printf("adapter: %s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") ? 3 : 1;
    }
    CloudPtr map = make_room();
    match->SetInputTarget(map);
    // the scan = a subset of the room seen from a sensor displaced by (0.03, -0.02, 0.025)
    auto scan = std::make_shared<PointCloudType>();
    for (size_t i = 7; i < map->points.size(); i += 11) {
        PointType p = map->points[i];
        if (p.x < 0.6f && p.y < 0.6f && p.z < 0.6f) continue;  // keep away from the edges
        if (p.x > 5.3f || p.y > 5.3f || p.z > 5.3f) continue;
        p.x -= 0.03f; p.y += 0.02f; p.z -= 0.025f;
        scan->points.push_back(p);
    }
    scan->width = static_cast<uint32_t>(scan->points.size());
    SE3 predict, result;
    CloudPtr aligned(new PointCloudType);
    if (!match->ScanMatch(scan, predict, aligned, result)) return 1;
    const double* t = result.data() + 4;
    std::printf("adapter: %zu scan points, t = (%.5f %.5f %.5f), fitness %.1f\n", scan->points.size(), t[0], t[1], t[2],
                match->GetFitnessScore());
    if (std::fabs(t[0] - 0.03) > 1e-3 || std::fabs(t[1] + 0.02) > 1e-3 || std::fabs(t[2] - 0.025) > 1e-3) return 1;
    if (aligned->points.size() != scan->points.size() || aligned->points[5].intensity != scan->points[5].intensity) return 1;
    Mat6d H;
    Vec6d B;
    if (!match->CaculateMatrixHAndB(scan, result, H, B)) return 1;
    NdtOptions nopt;
    std::shared_ptr<MatchingInterface> ndt = std::make_shared<CudaNdtRegistration>(nopt);
    ndt->SetInputTarget(map);
    SE3 nres;
    ndt->ScanMatch(scan, predict, aligned, nres);
    std::printf("adapter: NDT t = (%.4f %.4f %.4f)\n", nres.data()[4], nres.data()[5], nres.data()[6]);
    // the callers' map state on the device: Loc's crop (whole room inside the box) and Lio's key-frame window give the
    // same registration as the plain SetInputTarget above
    auto icp = std::dynamic_pointer_cast<CudaIcpRegistration>(match);
    icp->SetGlobalMap(map);
    const float half[3] = {50.0f, 50.0f, 50.0f};
    if (icp->ResetLocalMap(3.0f, 3.0f, 3.0f, half) != map->points.size()) return 1;
    SE3 again;
    icp->ScanMatch(scan, predict, aligned, again);
    for (int i = 0; i < 7; ++i)
        if (again.data()[i] != result.data()[i]) return 1;
    SE3 identity;
    if (icp->AddKeyFrame(map, identity, 10, 0.0f) != map->points.size()) return 1;  // leaf 0: NoFilter
    icp->ScanMatch(scan, predict, aligned, again);
    for (int i = 0; i < 7; ++i)
        if (again.data()[i] != result.data()[i]) return 1;
    std::printf("adapter: Loc / Lio map state on the device reproduces the pose\n");
    // the two callers' loops (CudaLocTracker / CudaLioTracker): first step from the identity prediction = the pose above,
    // second step from the constant-velocity prediction lands on it again
    {
        IcpOptions o2;
        o2.method_ = IcpMethod::P2PLANE;
        auto reg = std::make_shared<CudaIcpRegistration>(o2);
        CudaLocTracker loc(reg, map, SE3(), 50.0f, 10.0);
        if (loc.LocalMapSize() != map->points.size() || loc.Resets() != 1) return 1;
        CloudPtr out;
        SE3 p1 = loc.Update(scan, out);
        for (int i = 0; i < 7; ++i)
            if (p1.data()[i] != result.data()[i]) return 1;
        SE3 p2 = loc.Update(scan, out);
        for (int i = 4; i < 7; ++i)
            if (std::fabs(p2.data()[i] - result.data()[i]) > 1e-3) return 1;
        if (std::fabs(loc.Predict().data()[4] - (2 * p2.data()[4] - p1.data()[4])) > 1e-4) return 1;  // (almost) pure translation: 2 b - a
    }
    {
        IcpOptions o2;
        o2.method_ = IcpMethod::P2PLANE;
        auto reg = std::make_shared<CudaIcpRegistration>(o2);
        CudaLioTracker lio(reg, 3, 0.02, 10.0, 0.0f);  // key frame every 2 cm, NoFilter
        bool kf = false;
        lio.AddCloud(map, map, &kf);
        if (!kf || lio.LocalMapSize() != map->points.size()) return 1;
        SE3 p1 = lio.AddCloud(scan, scan, &kf);
        for (int i = 0; i < 7; ++i)
            if (p1.data()[i] != result.data()[i]) return 1;
        if (!kf || lio.KeyFrames() != 2 || lio.LocalMapSize() != map->points.size() + scan->points.size()) return 1;
        SE3 p2 = lio.AddCloud(scan, scan, &kf);  // not a key frame: it has not moved since the last one
        if (kf || lio.KeyFrames() != 2) return 1;
        for (int i = 4; i < 7; ++i)
            if (std::fabs(p2.data()[i] - result.data()[i]) > 1e-3) return 1;
    }
    std::printf("adapter: CudaLocTracker / CudaLioTracker loops ok\n");
    return 0;
}

// Device-resident NDT voxel grid (owner of the arrays behind NdtMapView) and its build pipeline.
#pragma once
#include <vector>

#include "device_utils.cuh"
#include "ndt_point.cuh"

namespace locreg {

class DeviceNdtMap {
   public:
    DeviceNdtMap() = default;
    ~DeviceNdtMap();
    DeviceNdtMap(const DeviceNdtMap&) = delete;
    DeviceNdtMap& operator=(const DeviceNdtMap&) = delete;

    void build(const void* d_xyz, size_t n, size_t stride, double voxel_size, int min_pts_in_voxel, cudaStream_t stream);
    const NdtMapView& view() const { return view_; }
    bool empty() const { return view_.n_voxels == 0; }
    size_t bytes() const { return bytes_; }
    // parity probe: voxels sorted by (kx,ky,kz)
    void download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info, std::vector<int>& npts,
                  cudaStream_t stream) const;

   private:
    void release();
    NdtSlot* slots_ = nullptr;
    NdtVoxel* voxels_ = nullptr;
    unsigned int cap_ = 0;
    NdtMapView view_{};
    size_t bytes_ = 0;
};

}  // namespace locreg

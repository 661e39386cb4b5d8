// Incremental NDT voxel cache (NdtRegistration::SetIncNdtTargetCloud / UpdateVoxel, ndt_registration.cpp:150-236).
//
// The reference keeps an LRU of voxels - std::list<{key, NdtVoxelData}> + unordered_map<key, list iterator>,
// capacity_ = 100000 - that is updated point by point, in cloud order: a new key is pushed to the front (and the tail
// evicted once the list holds capacity_ entries), a known key gets the point appended and moves to the front; the
// voxels the cloud touched are then re-estimated from the points THIS cloud gave them.  That bookkeeping is
// sequential control logic over a few 1e4 keys per cloud, so it stays on the host exactly as the reference writes it
// (std::list + std::unordered_map, same order of operations, hence the same eviction victims); the arithmetic -
// per-voxel mean / covariance / information matrix, and the whole alignment - runs on the device, on the same
// {slot table, voxel record} layout as the direct NDT grid, which is re-published after every cloud.
#pragma once
#include <list>
#include <unordered_map>
#include <vector>

#include "device_utils.cuh"
#include "ndt_point.cuh"

namespace locreg {

class DeviceIncNdtMap {
   public:
    DeviceIncNdtMap() = default;
    ~DeviceIncNdtMap();
    DeviceIncNdtMap(const DeviceIncNdtMap&) = delete;
    DeviceIncNdtMap& operator=(const DeviceIncNdtMap&) = delete;

    void configure(double voxel_size, size_t capacity);
    // SetIncNdtTargetCloud: h_xyz = the cloud in HOST memory (keys and the LRU are host work), d_xyz = its device copy
    void add_cloud(const void* h_xyz, const void* d_xyz, size_t n, size_t stride, cudaStream_t stream);
    const NdtMapView& view() const { return view_; }
    size_t size() const { return data_.size(); }
    // parity probe: voxels sorted by (kx,ky,kz); npts = points of the last cloud that touched the voxel
    void download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info, std::vector<int>& npts,
                  cudaStream_t stream) const;

   private:
    struct Entry {
        unsigned long long key;          // ndt_pack(kx, ky, kz)
        int vid;                         // row of the voxel record on the device
        int n_last;                      // points the last cloud put here
        std::vector<unsigned int> pts;   // NdtVoxelData::pts_: point indices of the cloud being added
    };
    std::list<Entry> data_;                                                  // data_ (front = most recent)
    std::unordered_map<unsigned long long, std::list<Entry>::iterator> grids_;  // inc_grids_
    std::vector<int> free_vids_;
    size_t capacity_ = 100000;
    NdtSlot* slots_ = nullptr;
    NdtVoxel* voxels_ = nullptr;
    unsigned int cap_slots_ = 0;
    NdtMapView view_{};
};

}  // namespace locreg

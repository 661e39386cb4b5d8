// Incremental NDT voxel cache (NdtRegistration::SetIncNdtTargetCloud / UpdateVoxel, ndt_registration.cpp:150-236),
// kept and updated ENTIRELY ON THE DEVICE: nothing but the cloud's pointer reaches add_cloud, nothing is copied back.
//
// The reference keeps an LRU of voxels - std::list<{key, NdtVoxelData}> + unordered_map<key, list iterator>,
// capacity_ = 100000 - that is updated point by point, in cloud order: a new key is pushed to the front (and the tail
// evicted once the list holds capacity_ entries), a known key gets the point appended and moves to the front; the
// voxels the cloud touched are then re-estimated from the points THIS cloud gave them.  A GPU cannot walk a std::list
// point by point; it does not have to.  LRU is a stack algorithm: with C = capacity_ - 1 entries retained after every
// insertion,
//   * an access is a HIT iff the key was accessed before and fewer than C distinct other keys were accessed since;
//   * the cache content is the C most recently accessed distinct keys;
//   * a voxel is re-estimated from the points it received since its last MISS inside the cloud (a re-inserted voxel
//     starts empty), or from all of its points of the cloud when every access was a hit
// - statements about the access SEQUENCE, not about a list, and every one of them is a count or a prefix sum:
//   runs       maximal stretches of consecutive points with the same voxel key (one access each; a scan line enters
//              and leaves a voxel a few times): head flags + scan
//   groups     the runs of one key (scratch hash table), ordered by run index; the key's old cache entry, if any
//   time line  old entry of LRU rank r (0 = oldest of m) sits at time r - m, run j at time j; prev[j] = time of the
//              previous access to run j's key
//   hit/miss   run j with prev = i is a miss iff the old entries newer than i plus the runs p in (i, j) with
//              prev[p] < i (first accesses to their key inside the window) number C or more; only windows of C or more
//              accesses need the count (a warp per such run)
//   survivors  the untouched old entries in their old order, then the cloud's keys by last access, less the
//              max(0, total - C) oldest of that sequence: two flag + scan + scatter compactions give the new LRU order
//   statistics one thread per surviving touched voxel over its runs from the last miss on (k_inc_stats), in arrival
//              order (the sums are order dependent in the last bits)
//   publish    the {key -> voxel record} table the alignment kernels probe is rebuilt from the new order.
// tests/test_inc_lru_model.py pins this formulation (as numpy) to the oracle's literal std::list on adversarial
// sequences; tests/test_gpu_parity.py pins the kernels (voxel dump after every cloud, tiny capacities).
#pragma once
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
    // SetIncNdtTargetCloud: d_xyz = the cloud in DEVICE memory.  Queued on `stream`; no host synchronisation.
    void add_cloud(const void* d_xyz, size_t n, size_t stride, cudaStream_t stream);
    const NdtMapView& view() const { return view_; }
    // voxels in the cache (one 4-byte read-back; synchronises `stream`)
    size_t size(cudaStream_t stream) const;
    // parity probe: voxels sorted by (kx,ky,kz); npts = points of the last cloud that touched the voxel
    void download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info, std::vector<int>& npts,
                  cudaStream_t stream) const;

   private:
    void release();
    void reserve_scratch(size_t n);

    size_t capacity_ = 100000;     // NdtOptions::capacity_ (the list never holds more than capacity_ - 1 entries)
    // persistent state
    NdtSlot* slots_ = nullptr;                // published table: key -> (vid, points of the last update)
    NdtVoxel* voxels_ = nullptr;              // [capacity_] voxel records
    unsigned long long* ent_key_ = nullptr;   // [capacity_] key of record vid, kNdtEmpty = free
    unsigned int* ent_cnt_ = nullptr;         // [capacity_] points of the last update
    unsigned int* order_[2] = {nullptr, nullptr};  // [capacity_] vids, oldest first (double buffered)
    unsigned int* rank_of_ = nullptr;         // [capacity_] position of vid in the order
    unsigned int* ctr_ = nullptr;             // device scalars (see inc_ndt.cu)
    unsigned int cap_slots_ = 0;
    int cur_ = 0;
    // scratch, grown with the largest cloud seen
    void* scratch_ = nullptr;
    size_t scratch_bytes_ = 0;
    NdtMapView view_{};
};

}  // namespace locreg

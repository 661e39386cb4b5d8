// Incremental NDT voxel cache: host LRU + device statistics (see device_inc_ndt.cuh).
#include <algorithm>
#include <cmath>

#include "device_inc_ndt.cuh"

namespace locreg {

// one thread per touched voxel: statistics of its points of the current cloud, in arrival order
__global__ void k_inc_ndt_stats(const unsigned int* __restrict__ group_start, const int* __restrict__ group_vid, unsigned int n_groups,
                                const unsigned int* __restrict__ members, const void* __restrict__ xyz, size_t stride,
                                NdtVoxel* voxels) {
    const unsigned int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    NdtVoxel v;
    inc_ndt_voxel_stats(members + group_start[g], group_start[g + 1] - group_start[g], xyz, stride, v);
    voxels[group_vid[g]] = v;
}

DeviceIncNdtMap::~DeviceIncNdtMap() {
    if (slots_) cudaFree(slots_);
    if (voxels_) cudaFree(voxels_);
}

void DeviceIncNdtMap::configure(double voxel_size, size_t capacity) {
    view_ = NdtMapView{};
    view_.inv_voxel = 1.0 / voxel_size;
    capacity_ = std::max<size_t>(capacity, 2);
    data_.clear();
    grids_.clear();
    free_vids_.clear();
    for (int v = static_cast<int>(capacity_) - 1; v >= 0; --v) free_vids_.push_back(v);
    if (slots_) cudaFree(slots_);
    if (voxels_) cudaFree(voxels_);
    slots_ = nullptr; voxels_ = nullptr;
    cap_slots_ = 1024;
    while (cap_slots_ < capacity_ * 4) cap_slots_ <<= 1;
    LR_CUDA(cudaMalloc(&slots_, static_cast<size_t>(cap_slots_) * sizeof(NdtSlot)));
    LR_CUDA(cudaMalloc(&voxels_, capacity_ * sizeof(NdtVoxel)));
}

void DeviceIncNdtMap::add_cloud(const void* h_xyz, const void* d_xyz, size_t n, size_t stride, cudaStream_t stream) {
    // ---- the reference's loop (ndt_registration.cpp:152-174), on keys
    std::vector<std::list<Entry>::iterator> touched;  // in order of first touch; may hold evicted (dangling) entries: see below
    std::vector<unsigned long long> touched_keys;
    for (size_t i = 0; i < n; ++i) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(h_xyz) + i * stride);
        if (!finite3(p[0], p[1], p[2])) continue;  // deviation D1
        const int kx = ndt_trunc(static_cast<double>(p[0]) * view_.inv_voxel), ky = ndt_trunc(static_cast<double>(p[1]) * view_.inv_voxel),
                  kz = ndt_trunc(static_cast<double>(p[2]) * view_.inv_voxel);
        if (!ndt_key_ok(kx, ky, kz)) continue;
        const unsigned long long key = ndt_pack(kx, ky, kz);
        auto it = grids_.find(key);
        if (it == grids_.end()) {
            Entry e;
            e.key = key; e.vid = free_vids_.back(); e.n_last = 0;
            free_vids_.pop_back();
            e.pts.push_back(static_cast<unsigned int>(i));
            data_.push_front(std::move(e));
            grids_.insert({key, data_.begin()});
            touched_keys.push_back(key);
            if (data_.size() >= capacity_) {  // evict the least recently touched voxel (:161-165)
                free_vids_.push_back(data_.back().vid);
                grids_.erase(data_.back().key);
                data_.pop_back();
            }
        } else {
            if (it->second->pts.empty()) touched_keys.push_back(key);  // first touch by this cloud
            it->second->pts.push_back(static_cast<unsigned int>(i));
            data_.splice(data_.begin(), data_, it->second);
            it->second = data_.begin();
        }
    }
    // ---- UpdateVoxel for the voxels this cloud touched and that are still cached (the reference would dereference a
    // null iterator for a voxel evicted within the same call, :177-178; that needs a cloud touching >= capacity_ voxels)
    std::vector<unsigned int> group_start{0}, members;
    std::vector<int> group_vid;
    for (unsigned long long key : touched_keys) {
        auto it = grids_.find(key);
        if (it == grids_.end() || it->second->pts.empty()) continue;
        Entry& e = *it->second;
        members.insert(members.end(), e.pts.begin(), e.pts.end());
        group_start.push_back(static_cast<unsigned int>(members.size()));
        group_vid.push_back(e.vid);
        e.n_last = static_cast<int>(e.pts.size());
        e.pts.clear();
    }
    const unsigned int n_groups = static_cast<unsigned int>(group_vid.size());
    if (n_groups) {
        unsigned int *d_start = nullptr, *d_members = nullptr;
        int* d_vid = nullptr;
        LR_CUDA(cudaMallocAsync(&d_start, group_start.size() * sizeof(unsigned int), stream));
        LR_CUDA(cudaMallocAsync(&d_members, members.size() * sizeof(unsigned int), stream));
        LR_CUDA(cudaMallocAsync(&d_vid, group_vid.size() * sizeof(int), stream));
        LR_CUDA(cudaMemcpyAsync(d_start, group_start.data(), group_start.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
        LR_CUDA(cudaMemcpyAsync(d_members, members.data(), members.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
        LR_CUDA(cudaMemcpyAsync(d_vid, group_vid.data(), group_vid.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        LR_LAUNCH(k_inc_ndt_stats, (n_groups + 127) / 128, 128, 0, stream, d_start, d_vid, n_groups, d_members, d_xyz, stride, voxels_);
        LR_CUDA(cudaFreeAsync(d_start, stream));
        LR_CUDA(cudaFreeAsync(d_members, stream));
        LR_CUDA(cudaFreeAsync(d_vid, stream));
    }
    // ---- re-publish the slot table (key -> voxel record) for the alignment kernels
    std::vector<NdtSlot> hs(cap_slots_, NdtSlot{kNdtEmpty, -1, 0u});
    for (const Entry& e : data_) {
        unsigned int h = ndt_hash(e.key) & (cap_slots_ - 1);
        while (hs[h].key != kNdtEmpty) h = (h + 1) & (cap_slots_ - 1);
        hs[h].key = e.key; hs[h].vid = e.vid; hs[h].count = static_cast<unsigned int>(e.n_last);
    }
    LR_CUDA(cudaMemcpyAsync(slots_, hs.data(), hs.size() * sizeof(NdtSlot), cudaMemcpyHostToDevice, stream));
    LR_CUDA(cudaStreamSynchronize(stream));  // hs and the group arrays are pageable
    view_.slots = slots_; view_.voxels = voxels_; view_.slot_mask = cap_slots_ - 1;
    view_.n_voxels = static_cast<unsigned int>(data_.size());
}

void DeviceIncNdtMap::download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info, std::vector<int>& npts,
                               cudaStream_t stream) const {
    std::vector<NdtVoxel> hv(capacity_);
    if (voxels_) LR_CUDA(cudaMemcpyAsync(hv.data(), voxels_, capacity_ * sizeof(NdtVoxel), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    struct Rec { int k[3]; int vid; int cnt; };
    std::vector<Rec> recs;
    for (const Entry& e : data_) {
        Rec r;
        ndt_unpack(e.key, r.k[0], r.k[1], r.k[2]);
        r.vid = e.vid; r.cnt = e.n_last;
        recs.push_back(r);
    }
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
        if (a.k[0] != b.k[0]) return a.k[0] < b.k[0];
        if (a.k[1] != b.k[1]) return a.k[1] < b.k[1];
        return a.k[2] < b.k[2];
    });
    keys.clear(); mu.clear(); info.clear(); npts.clear();
    for (const Rec& r : recs) {
        keys.insert(keys.end(), r.k, r.k + 3);
        mu.insert(mu.end(), hv[r.vid].mu, hv[r.vid].mu + 3);
        info.insert(info.end(), hv[r.vid].info, hv[r.vid].info + 9);
        npts.push_back(r.cnt);
    }
}

}  // namespace locreg

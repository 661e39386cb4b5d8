// NDT voxel grid build kernels (K5).  Pipeline:
//   1 insert   point -> trunc voxel key -> slot (atomicCAS), slot.count++
//   2 scan     exclusive scan of slot counts -> member list offsets
//   3 scatter  point indices into per-voxel member lists
//   4 stats    one thread per voxel with > min_pts_in_voxel members: sort the member list by index
//              (so the sums run in the reference's order), two-pass mean/cov, eigen, clamped inverse
#include <algorithm>
#include <numeric>

#include "device_ndt.cuh"
#include "voxel_build.cuh"

namespace locreg {

__global__ void k_ndt_clear(NdtSlot* slots, unsigned int cap) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) slots[s] = NdtSlot{kNdtEmpty, -1, 0u};
}
__global__ void k_ndt_insert(const void* __restrict__ xyz, size_t n, size_t stride, double inv_voxel, NdtSlot* slots,
                             unsigned int slot_mask, unsigned int* pt_slot, unsigned int* counters) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) ndt_insert_body<DeviceAtomics>(i, xyz, stride, inv_voxel, slots, slot_mask, pt_slot, counters);
}
__global__ void k_ndt_counts(const NdtSlot* __restrict__ slots, unsigned int cap, unsigned int* out) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) out[s] = slots[s].count;
}
__global__ void k_ndt_scatter(size_t n, const unsigned int* __restrict__ pt_slot, unsigned int* cursor, unsigned int* members) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned int s = pt_slot[i];
        if (s != 0xFFFFFFFFu) members[atomicAdd(&cursor[s], 1u)] = static_cast<unsigned int>(i);
    }
}
__global__ void k_ndt_stats(NdtSlot* slots, unsigned int cap, const unsigned int* __restrict__ start, unsigned int* members,
                            const void* __restrict__ xyz, size_t stride, int min_pts, NdtVoxel* voxels,
                            unsigned int* n_voxels) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    const unsigned int cnt = slots[s].count;
    if (slots[s].key == kNdtEmpty || !(static_cast<long long>(cnt) > static_cast<long long>(min_pts))) return;
    unsigned int* idx = members + start[s];
    for (unsigned int a = 1; a < cnt; ++a) {  // insertion sort: member lists are tens of entries
        const unsigned int v = idx[a];
        unsigned int b = a;
        while (b > 0 && idx[b - 1] > v) { idx[b] = idx[b - 1]; --b; }
        idx[b] = v;
    }
    const unsigned int vid = atomicAdd(n_voxels, 1u);
    ndt_voxel_stats(idx, cnt, xyz, stride, voxels[vid]);
    slots[s].vid = static_cast<int>(vid);
}

DeviceNdtMap::~DeviceNdtMap() { release(); }
void DeviceNdtMap::release() {
    if (slots_) cudaFree(slots_);
    if (voxels_) cudaFree(voxels_);
    slots_ = nullptr; voxels_ = nullptr; cap_ = 0; view_ = NdtMapView{}; bytes_ = 0;
}

void DeviceNdtMap::build(const void* d_xyz, size_t n, size_t stride, double voxel_size, int min_pts_in_voxel,
                         cudaStream_t stream) {
    release();
    view_.inv_voxel = 1.0 / voxel_size;
    if (n == 0) return;
    if (n >= (1ull << 31)) throw std::invalid_argument("target cloud too large (>= 2^31 points)");
    const unsigned int T = 256;
    const unsigned int gridN = static_cast<unsigned int>((n + T - 1) / T);
    unsigned int *pt_slot = nullptr, *counters = nullptr, *start = nullptr, *cursor = nullptr, *members = nullptr;
    LR_CUDA(cudaMallocAsync(&pt_slot, n * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&counters, 8 * sizeof(unsigned int), stream));
    unsigned int cap = 1024;
    while (cap < n / 4 + 1024) cap <<= 1;
    unsigned int h_counters[8];
    while (true) {
        LR_CUDA(cudaMalloc(&slots_, static_cast<size_t>(cap) * sizeof(NdtSlot)));
        LR_LAUNCH(k_ndt_clear, (cap + T - 1) / T, T, 0, stream, slots_, cap);
        LR_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), stream));
        LR_LAUNCH(k_ndt_insert, gridN, T, 0, stream, d_xyz, n, stride, view_.inv_voxel, slots_, cap - 1, pt_slot, counters);
        LR_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, stream));
        LR_CUDA(cudaStreamSynchronize(stream));
        if (h_counters[1] || static_cast<size_t>(h_counters[0]) * 2 > cap) {
            cudaFree(slots_);
            slots_ = nullptr;
            if (cap >= (1u << 30)) throw std::runtime_error("NDT hash table overflow");
            cap <<= 2;
            continue;
        }
        break;
    }
    cap_ = cap;
    const unsigned int n_slots_used = h_counters[0];
    LR_CUDA(cudaMallocAsync(&start, (static_cast<size_t>(cap) + 1) * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&cursor, static_cast<size_t>(cap) * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&members, n * sizeof(unsigned int), stream));
    LR_LAUNCH(k_ndt_counts, (cap + T - 1) / T, T, 0, stream, slots_, cap, start);
    exclusive_scan_u32(start, start, cap, nullptr, stream);
    LR_CUDA(cudaMemcpyAsync(cursor, start, static_cast<size_t>(cap) * sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
    LR_LAUNCH(k_ndt_scatter, gridN, T, 0, stream, n, pt_slot, cursor, members);
    LR_CUDA(cudaMalloc(&voxels_, std::max<size_t>(n_slots_used, 1) * sizeof(NdtVoxel)));
    LR_CUDA(cudaMemsetAsync(counters + 4, 0, sizeof(unsigned int), stream));
    LR_LAUNCH(k_ndt_stats, (cap + T - 1) / T, T, 0, stream, slots_, cap, start, members, d_xyz, stride, min_pts_in_voxel,
              voxels_, counters + 4);
    unsigned int nv = 0;
    LR_CUDA(cudaMemcpyAsync(&nv, counters + 4, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    LR_CUDA(cudaFreeAsync(pt_slot, stream));
    LR_CUDA(cudaFreeAsync(counters, stream));
    LR_CUDA(cudaFreeAsync(start, stream));
    LR_CUDA(cudaFreeAsync(cursor, stream));
    LR_CUDA(cudaFreeAsync(members, stream));
    view_.slots = slots_; view_.voxels = voxels_; view_.slot_mask = cap - 1; view_.n_voxels = nv;
    bytes_ = static_cast<size_t>(cap) * sizeof(NdtSlot) + static_cast<size_t>(nv) * sizeof(NdtVoxel);
}

void DeviceNdtMap::download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info,
                            std::vector<int>& npts, cudaStream_t stream) const {
    std::vector<NdtSlot> hs(cap_);
    std::vector<NdtVoxel> hv(view_.n_voxels);
    if (cap_) LR_CUDA(cudaMemcpyAsync(hs.data(), slots_, static_cast<size_t>(cap_) * sizeof(NdtSlot), cudaMemcpyDeviceToHost, stream));
    if (view_.n_voxels)
        LR_CUDA(cudaMemcpyAsync(hv.data(), voxels_, static_cast<size_t>(view_.n_voxels) * sizeof(NdtVoxel), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    struct Rec { int k[3]; int vid; int cnt; };
    std::vector<Rec> recs;
    for (const NdtSlot& s : hs)
        if (s.key != kNdtEmpty && s.vid >= 0) {
            Rec r;
            ndt_unpack(s.key, r.k[0], r.k[1], r.k[2]);
            r.vid = s.vid; r.cnt = static_cast<int>(s.count);
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

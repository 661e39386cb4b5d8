// Device-resident voxel-hash map: owner of the arrays a VoxelMapView points at, and the host side of
// the build pipeline described in voxel_build.cuh.
#pragma once
#include "device_utils.cuh"
#include "voxel_map.cuh"

namespace locreg {

class DeviceVoxelMap {
   public:
    DeviceVoxelMap() = default;
    ~DeviceVoxelMap();
    DeviceVoxelMap(const DeviceVoxelMap&) = delete;
    DeviceVoxelMap& operator=(const DeviceVoxelMap&) = delete;

    // d_xyz: device pointer to n points, `stride` bytes apart.  Synchronises the stream (the slot
    // table is sized from the number of occupied blocks, read back once).
    // keep_pos (optional): receives a device array (cudaMallocAsync on `stream`, caller frees) mapping every input
    // index to its canonical position in pts (undefined for dropped points).
    // dups (optional): with flags == nullptr on entry it receives this build's duplicate flags (one byte per input point,
    // cudaMallocAsync on `stream`, caller frees) and their number; with flags set, the build takes them as given.
    struct DupFlags { unsigned char* flags = nullptr; unsigned int count = 0; };
    void build(const void* d_xyz, size_t n, size_t stride, float cell, bool want_lists, cudaStream_t stream,
               unsigned int** keep_pos = nullptr, DupFlags* dups = nullptr);
    // Turns this map into a coarse level of `fine`: every entry's w (original index) is replaced by the point's
    // canonical position in fine's pts, which is what the search reports on every level.
    void attach_to(const VoxelMapView& fine, const unsigned int* fine_pos_of_index, cudaStream_t stream);
    // Block pyramid over this (fine) level for the ball-query stage 2 (voxel_map.cuh); no synchronisation.
    void build_pyramid(cudaStream_t stream);
    const PyrView& pyramid() const { return pyr_view_; }
    const VoxelMapView& view() const { return view_; }
    bool empty() const { return view_.n_pts == 0; }
    void clear() { release(); }  // forgets the map, keeps the memory
    size_t bytes() const { return bytes_ + pyr_bytes_; }
    unsigned int n_cells() const { return n_cells_; }
    unsigned int n_blocks() const { return n_blocks_; }
    unsigned int n_lists() const { return n_lists_; }

   private:
    void release();
    // Device arrays keep their capacity across builds: Loc re-crops its local map every few hundred scans, and a
    // cudaMalloc / cudaFree pair of the 0.4 GB list array costs more than all the build kernels together.
    template <class T>
    struct Grow {
        T* p = nullptr;
        size_t cap = 0;  // elements
        T* ensure(size_t count) {
            if (count > cap) {
                if (p) cudaFree(p);
                p = nullptr; cap = 0;
                const size_t want = count + count / 8 + 64;
                LR_CUDA(cudaMalloc(&p, want * sizeof(T)));
                cap = want;
            }
            return p;
        }
        void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    };
    Grow<VoxelSlot> slots_buf_;
    Grow<unsigned int> cell_start_buf_;
    Grow<float4> pts_buf_;
    Grow<NbrSlot> nbr_buf_;
    Grow<PyrSlot> pyr_buf_;
    PyrView pyr_view_{};
    size_t pyr_bytes_ = 0;
    VoxelSlot* slots_ = nullptr;
    unsigned int* cell_start_ = nullptr;
    float4* pts_ = nullptr;
    NbrSlot* nbr_ = nullptr;
    VoxelMapView view_{};
    size_t bytes_ = 0;
    unsigned int n_cells_ = 0, n_blocks_ = 0, n_lists_ = 0;
    size_t n_list_entries_ = 0;
    size_t last_n_ = 0;                      // size of the last cloud built and the table capacities that held it
    unsigned int last_cap_ = 0, last_nbr_cap_ = 0;
};

// Level 0 (cell, lists) + the mid level (cells kMidFactor times larger, lists; nullptr: none) + kCoarseLevels coarser
// levels (block tables only) of the ICP search index.
void build_icp_maps(DeviceVoxelMap& fine, DeviceVoxelMap* coarse, DeviceVoxelMap* mid, const void* d_xyz, size_t n, size_t stride,
                    float cell, bool want_lists, cudaStream_t stream);

}  // namespace locreg

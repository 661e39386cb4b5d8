// Voxel-hash map build kernels (K1) and the scan primitive.  See voxel_build.cuh for the pipeline.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_map.cuh"
#include "voxel_build.cuh"

namespace locreg {

thread_local long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* warp_sums, unsigned int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
        unsigned int winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += t;
        }
        warp_sums[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    block_total = warp_sums[32];
    return inc - v + warp_sums[warp];
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const unsigned int* __restrict__ in, size_t n,
                                                                 unsigned int* __restrict__ tile_sums) {
    __shared__ unsigned int warp_sums[33];
    const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile + static_cast<size_t>(threadIdx.x) * kScanItems;
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in[base + k];
    unsigned int total;
    block_exclusive_scan(s, warp_sums, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const unsigned int* __restrict__ in, unsigned int* out, size_t n,
                                                             const unsigned int* __restrict__ tile_offsets) {
    __shared__ unsigned int warp_sums[33];
    const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile + static_cast<size_t>(threadIdx.x) * kScanItems;
    unsigned int v[kScanItems];
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    unsigned int total;
    unsigned int run = block_exclusive_scan(s, warp_sums, total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void k_scan_total(const unsigned int* tile_sums, const unsigned int* tile_offsets, size_t last_tile, unsigned int* total) {
    *total = tile_sums[last_tile] + (tile_offsets ? tile_offsets[last_tile] : 0u);
}

void exclusive_scan_u32(const unsigned int* in, unsigned int* out, size_t n, unsigned int* total, cudaStream_t stream) {
    if (n == 0) {
        if (total) LR_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned int), stream));
        return;
    }
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    unsigned int* sums = nullptr;
    unsigned int* offs = nullptr;
    LR_CUDA(cudaMallocAsync(&sums, tiles * sizeof(unsigned int), stream));
    LR_LAUNCH(k_scan_tile_sums, static_cast<unsigned int>(tiles), kScanThreads, 0, stream, in, n, sums);
    if (tiles > 1) {
        LR_CUDA(cudaMallocAsync(&offs, tiles * sizeof(unsigned int), stream));
        exclusive_scan_u32(sums, offs, tiles, nullptr, stream);
    }
    if (total) LR_LAUNCH(k_scan_total, 1, 1, 0, stream, sums, offs, tiles - 1, total);
    LR_LAUNCH(k_scan_tiles, static_cast<unsigned int>(tiles), kScanThreads, 0, stream, in, out, n, offs);
    LR_CUDA(cudaFreeAsync(sums, stream));
    if (offs) LR_CUDA(cudaFreeAsync(offs, stream));
}

// ------------------------------------------------------------------------------------------------
// build kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_slots_clear(VoxelSlot* slots, unsigned int cap) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) slots[s] = VoxelSlot{kEmptyKey, 0ull, 0u, 0u, 0u, 0u};
}

// bounds: [0..2] min cell coord, [3..5] max cell coord
__global__ void k_build_insert(const void* __restrict__ xyz, size_t n, size_t stride, float inv_cell, VoxelSlot* slots,
                               unsigned int slot_mask, unsigned int* pt_slot, unsigned char* pt_bit,
                               unsigned int* counters, int* bounds) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int f[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    int g[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
    if (i < n) {
        int c[3];
        if (build_insert_body<DeviceAtomics>(i, xyz, stride, inv_cell, slots, slot_mask, pt_slot, pt_bit, counters, c)) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { f[a] = c[a]; g[a] = c[a]; }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int lo = __reduce_min_sync(0xffffffffu, f[a]);
        const int hi = __reduce_max_sync(0xffffffffu, g[a]);
        if ((threadIdx.x & 31) == 0) {
            if (lo != 0x7fffffff) atomicMin(&bounds[a], lo);
            if (hi != -0x7fffffff) atomicMax(&bounds[3 + a], hi);
        }
    }
}

__global__ void k_slot_popc(const VoxelSlot* __restrict__ slots, unsigned int cap, unsigned int* out) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) out[s] = slots[s].key != kEmptyKey ? static_cast<unsigned int>(__popcll(slots[s].mask)) : 0u;
}
__global__ void k_slot_set_base(VoxelSlot* slots, unsigned int cap, const unsigned int* __restrict__ base) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) slots[s].cell_base = base[s];
}
__global__ void k_build_count(size_t n, const VoxelSlot* __restrict__ slots, unsigned int* pt_slot,
                              const unsigned char* __restrict__ pt_bit, unsigned int* cell_count) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) build_count_body<DeviceAtomics>(i, slots, pt_slot, pt_bit, cell_count);
}
__global__ void k_build_scatter(const void* __restrict__ xyz, size_t n, size_t stride, const unsigned int* __restrict__ pt_cell,
                                unsigned int* cursor, float4* pts, unsigned int* pt_pos) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) build_scatter_body<DeviceAtomics>(i, xyz, stride, pt_cell, cursor, pts, pt_pos);
}
__global__ void k_build_dup_flag(size_t n, const unsigned int* __restrict__ pt_cell, const unsigned int* __restrict__ pt_pos,
                                 const unsigned int* __restrict__ cell_start, const float4* __restrict__ pts,
                                 unsigned char* dup, unsigned int* n_dup) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool d = false;
    if (i < n) {
        d = build_is_duplicate(i, pt_cell, pt_pos, cell_start, pts);
        dup[i] = d ? 1 : 0;
    }
    const unsigned int c = __popc(__ballot_sync(0xffffffffu, d));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_dup, c);
}
__global__ void k_build_dup_apply(size_t n, const unsigned char* __restrict__ dup, const unsigned int* __restrict__ pt_pos, float4* pts) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n && dup[i]) pts[pt_pos[i]].x = __int_as_float(0x7fc00000);
}

__global__ void k_nbr_clear(NbrSlot* nbr, unsigned int cap) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) nbr[s] = NbrSlot{kEmptyKey, 0u, 0u};
}
// ---- neighbourhood lists by GATHER ----------------------------------------------------------------------------------------
// The lists used to be built from the points' side: every point probed the table of list cells 27 times to count itself
// and 27 times more to write itself - 54 atomics and 27 scattered 16 B stores per map point.  From the lists' side the
// same arrays cost a fraction: (1) every OCCUPIED CELL marks the 27 cells whose neighbourhood it belongs to (cells are
// 2-3x fewer than points); (2) one warp per list cell, lane o < 27 looking up neighbour cell o in the block table (a
// read-only 32 B record, L2 resident), a shuffle scan of the 27 counts, ONE atomic to reserve the list's range, and the
// lanes copy their cells' points into consecutive places: coalesced stores, a deterministic order inside every list
// (cell by cell, canonical position ascending), no second pass.  Masked duplicates (quirk Q3) are left out.
__device__ __forceinline__ void unpack_coord(unsigned long long key, int& x, int& y, int& z) {
    x = static_cast<int>((key >> 42) & 0x1FFFFFull) - kCoordBias;
    y = static_cast<int>((key >> 21) & 0x1FFFFFull) - kCoordBias;
    z = static_cast<int>(key & 0x1FFFFFull) - kCoordBias;
}
// counters: [0] lists created, [1] overflow flag
__global__ void k_nbr_mark(const VoxelSlot* __restrict__ slots, unsigned int cap, NbrSlot* nbr, unsigned int nbr_mask, unsigned int* counters) {
    const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t s = t >> 6;
    const int bit = static_cast<int>(t & 63);
    if (s >= cap) return;
    const VoxelSlot sl = slots[s];
    if (sl.key == kEmptyKey || !((sl.mask >> bit) & 1ull)) return;
    int bx, by, bz;
    unpack_coord(sl.key, bx, by, bz);
    const int cx = (bx << 2) + (bit & 3), cy = (by << 2) + ((bit >> 2) & 3), cz = (bz << 2) + (bit >> 4);
    for (int o = 0; o < 27; ++o) {
        const unsigned long long key = pack_cell(cx + (o % 3) - 1, cy + ((o / 3) % 3) - 1, cz + (o / 9) - 1);
        unsigned int h = hash_block(key) & nbr_mask;
        unsigned int probes = 0;
        while (true) {
            const unsigned long long k = atomicCAS(&nbr[h].key, kEmptyKey, key);
            if (k == kEmptyKey) { atomicAdd(&counters[0], 1u); break; }
            if (k == key) break;
            h = (h + 1) & nbr_mask;
            if (++probes > nbr_mask) { counters[1] = 1u; return; }
        }
    }
}
// one warp per slot of the list table; list_cursor counts the list entries handed out so far
__global__ void k_nbr_gather(VoxelMapView map, NbrSlot* nbr, unsigned int nbr_cap, float4* pts, unsigned int n_kept, unsigned int* list_cursor) {
    const unsigned int w = static_cast<unsigned int>((static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const unsigned int lane = threadIdx.x & 31;
    if (w >= nbr_cap) return;
    const unsigned long long key = nbr[w].key;
    if (key == kEmptyKey) return;
    int lx, ly, lz;
    unpack_coord(key, lx, ly, lz);
    unsigned int beg = 0, end = 0;
    if (lane < 27) {
        const int cx = lx + static_cast<int>(lane % 3) - 1, cy = ly + static_cast<int>((lane / 3) % 3) - 1, cz = lz + static_cast<int>(lane / 9) - 1;
        const VoxelSlot* s = find_block(map, cx >> 2, cy >> 2, cz >> 2);
        if (s != nullptr) {
            const int bit = (cx & 3) | ((cy & 3) << 2) | ((cz & 3) << 4);
            const unsigned long long occ = s->mask;
            if ((occ >> bit) & 1ull) {
                const unsigned int cid = s->cell_base + static_cast<unsigned int>(__popcll(occ & ((1ull << bit) - 1ull)));
                beg = map.cell_start[cid];
                end = map.cell_start[cid + 1];
            }
        }
    }
    unsigned int live = 0;
    for (unsigned int i = beg; i < end; ++i) live += pts[i].x == pts[i].x ? 1u : 0u;  // a masked duplicate is NaN
    unsigned int inc = live;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= static_cast<unsigned int>(off)) inc += v;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned int start = 0;
    if (lane == 0) start = n_kept + atomicAdd(list_cursor, total);
    start = __shfl_sync(0xffffffffu, start, 0);
    unsigned int at = start + inc - live;
    for (unsigned int i = beg; i < end; ++i) {
        const float4 p = pts[i];
        if (p.x == p.x) pts[at++] = make_float4(p.x, p.y, p.z, __int_as_float(static_cast<int>(i)));  // w = canonical position
    }
    if (lane == 0) { nbr[w].start = start; nbr[w].count = total; }
}

// ------------------------------------------------------------------------------------------------
DeviceVoxelMap::~DeviceVoxelMap() {
    slots_buf_.free(); cell_start_buf_.free(); pts_buf_.free(); nbr_buf_.free(); pyr_buf_.free();
}
void DeviceVoxelMap::release() {  // forgets the map, keeps the memory
    slots_ = nullptr; cell_start_ = nullptr; pts_ = nullptr; nbr_ = nullptr;
    view_ = VoxelMapView{};
    bytes_ = 0; n_cells_ = 0; n_blocks_ = 0; n_lists_ = 0; n_list_entries_ = 0;
    pyr_view_ = PyrView{}; pyr_bytes_ = 0;
}

static unsigned int next_pow2(size_t v) {
    unsigned int p = 1024;
    while (p < v && p < (1u << 31)) p <<= 1;
    return p;
}

void DeviceVoxelMap::build(const void* d_xyz, size_t n, size_t stride, float cell, bool want_lists, cudaStream_t stream,
                           unsigned int** keep_pos, DupFlags* dups) {
    release();
    if (keep_pos) *keep_pos = nullptr;
    const bool dup_given = dups != nullptr && dups->flags != nullptr;
    if (n == 0) return;
    if (n >= (1ull << 31)) throw std::invalid_argument("target cloud too large (>= 2^31 points)");
    const float inv_cell = 1.0f / cell;
    const unsigned int T = 256;
    const unsigned int gridN = static_cast<unsigned int>((n + T - 1) / T);
    unsigned int *pt_slot = nullptr, *pt_pos = nullptr, *counters = nullptr, *tmp = nullptr, *cursor = nullptr;
    unsigned char *pt_bit = nullptr, *dup = nullptr;
    int* bounds = nullptr;
    LR_CUDA(cudaMallocAsync(&pt_slot, n * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&pt_pos, n * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&pt_bit, n, stream));
    if (dup_given) dup = dups->flags;
    else LR_CUDA(cudaMallocAsync(&dup, n, stream));
    LR_CUDA(cudaMallocAsync(&counters, 8 * sizeof(unsigned int), stream));
    LR_CUDA(cudaMallocAsync(&bounds, 6 * sizeof(int), stream));
    // table sizes start from the last build's when the cloud is of similar size (Loc's re-crops, Lio's key frames): a
    // guess that proves too small costs a whole insert pass and a synchronisation
    const bool similar = last_n_ != 0 && n <= 2 * last_n_ && 2 * n >= last_n_;
    unsigned int cap = next_pow2(n / 4 + 1024);
    if (similar && last_cap_ > cap) cap = last_cap_;
    unsigned int h_counters[8];
    int h_bounds[6];
    while (true) {
        slots_ = slots_buf_.ensure(cap);
        LR_LAUNCH(k_slots_clear, (cap + T - 1) / T, T, 0, stream, slots_, cap);
        LR_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), stream));
        const int init_bounds[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, -0x7fffffff, -0x7fffffff, -0x7fffffff};
        LR_CUDA(cudaMemcpyAsync(bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, stream));
        LR_LAUNCH(k_build_insert, gridN, T, 0, stream, d_xyz, n, stride, inv_cell, slots_, cap - 1, pt_slot, pt_bit,
                  counters, bounds);
        LR_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, stream));
        LR_CUDA(cudaMemcpyAsync(h_bounds, bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, stream));
        LR_CUDA(cudaStreamSynchronize(stream));
        if (h_counters[1] || static_cast<size_t>(h_counters[0]) * 2 > cap) {  // too loaded: grow and redo
            slots_ = nullptr;
            if (cap >= (1u << 31)) throw std::runtime_error("voxel hash table overflow");
            cap = cap << 2 ? cap << 2 : (1u << 31);
            continue;
        }
        break;
    }
    n_blocks_ = h_counters[0];
    const unsigned int n_kept = h_counters[2];
    if (n_kept == 0) {  // every point was non-finite
        slots_ = nullptr;
        cudaFreeAsync(pt_slot, stream); cudaFreeAsync(pt_pos, stream); cudaFreeAsync(pt_bit, stream);
        if (!dup_given) cudaFreeAsync(dup, stream);
        cudaFreeAsync(counters, stream); cudaFreeAsync(bounds, stream);
        return;
    }
    // 2 rank: slot.cell_base = exclusive scan of popcount(mask)
    LR_CUDA(cudaMallocAsync(&tmp, static_cast<size_t>(cap) * sizeof(unsigned int), stream));
    LR_LAUNCH(k_slot_popc, (cap + T - 1) / T, T, 0, stream, slots_, cap, tmp);
    exclusive_scan_u32(tmp, tmp, cap, counters + 4, stream);
    LR_LAUNCH(k_slot_set_base, (cap + T - 1) / T, T, 0, stream, slots_, cap, tmp);
    LR_CUDA(cudaMemcpyAsync(&n_cells_, counters + 4, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    LR_CUDA(cudaFreeAsync(tmp, stream));
    // 3-4 histogram + scan -> cell_start
    cell_start_ = cell_start_buf_.ensure(static_cast<size_t>(n_cells_) + 1);
    LR_CUDA(cudaMemsetAsync(cell_start_, 0, (static_cast<size_t>(n_cells_) + 1) * sizeof(unsigned int), stream));
    LR_LAUNCH(k_build_count, gridN, T, 0, stream, n, slots_, pt_slot, pt_bit, cell_start_);
    exclusive_scan_u32(cell_start_, cell_start_, static_cast<size_t>(n_cells_) + 1, nullptr, stream);
    // 5 scatter
    LR_CUDA(cudaMallocAsync(&cursor, static_cast<size_t>(n_cells_) * sizeof(unsigned int), stream));
    LR_CUDA(cudaMemcpyAsync(cursor, cell_start_, static_cast<size_t>(n_cells_) * sizeof(unsigned int),
                            cudaMemcpyDeviceToDevice, stream));
    // neighbourhood lists live behind the sorted points in the same array (28 float4 per point in total);
    // they are skipped when they would not fit comfortably (the search then always takes the block path)
    size_t free_b = 0, total_b = 0;
    LR_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t list_entries = static_cast<size_t>(n_kept) * 27;
    const bool lists = want_lists && (static_cast<size_t>(n_kept) + list_entries) < 0xF0000000ull &&
                       (list_entries + n_kept) * sizeof(float4) < free_b / 2;
    pts_ = pts_buf_.ensure(static_cast<size_t>(n_kept) + (lists ? list_entries : 0));
    LR_LAUNCH(k_build_scatter, gridN, T, 0, stream, d_xyz, n, stride, pt_slot, cursor, pts_, pt_pos);
    // 6 dedupe (quirk Q3).  Which input points are duplicates does not depend on the cell size: a coarser level takes the
    // flags of the fine one instead of comparing every point with the thousands that share its (large) cell.
    unsigned int n_dup = 0;
    if (dup_given) {
        n_dup = dups->count;
    } else {
        LR_CUDA(cudaMemsetAsync(counters + 5, 0, sizeof(unsigned int), stream));
        LR_LAUNCH(k_build_dup_flag, gridN, T, 0, stream, n, pt_slot, pt_pos, cell_start_, pts_, dup, counters + 5);
    }
    LR_LAUNCH(k_build_dup_apply, gridN, T, 0, stream, n, dup, pt_pos, pts_);
    if (!dup_given) {
        LR_CUDA(cudaMemcpyAsync(&n_dup, counters + 5, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        LR_CUDA(cudaStreamSynchronize(stream));
    }
    // 7 neighbourhood lists (gathered per list cell, see k_nbr_gather)
    unsigned int nbr_cap = 0;
    if (lists) {
        nbr_cap = next_pow2(static_cast<size_t>(n_cells_) * 8);
        if (similar && last_nbr_cap_ > nbr_cap) nbr_cap = last_nbr_cap_;
        view_.slots = slots_; view_.cell_start = cell_start_; view_.pts = pts_; view_.slot_mask = cap - 1;  // what find_block needs
        while (true) {
            nbr_ = nbr_buf_.ensure(nbr_cap);
            LR_LAUNCH(k_nbr_clear, (nbr_cap + T - 1) / T, T, 0, stream, nbr_, nbr_cap);
            LR_CUDA(cudaMemsetAsync(counters, 0, 3 * sizeof(unsigned int), stream));
            LR_LAUNCH(k_nbr_mark, static_cast<unsigned int>((static_cast<size_t>(cap) * 64 + T - 1) / T), T, 0, stream, slots_, cap, nbr_,
                      nbr_cap - 1, counters);
            LR_CUDA(cudaMemcpyAsync(h_counters, counters, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            LR_CUDA(cudaStreamSynchronize(stream));
            if (h_counters[1] || static_cast<size_t>(h_counters[0]) * 2 > nbr_cap) {
                nbr_ = nullptr;
                if (nbr_cap >= (1u << 30)) throw std::runtime_error("neighbourhood table overflow");
                nbr_cap <<= 2;
                continue;
            }
            break;
        }
        n_lists_ = h_counters[0];
        n_list_entries_ = static_cast<size_t>(n_kept - n_dup) * 27;  // every surviving point, once per cell of its 3x3x3 box
        LR_LAUNCH(k_nbr_gather, static_cast<unsigned int>((static_cast<size_t>(nbr_cap) * 32 + T - 1) / T), T, 0, stream, view_, nbr_, nbr_cap,
                  pts_, n_kept, counters + 2);
    }
    LR_CUDA(cudaFreeAsync(cursor, stream));
    LR_CUDA(cudaFreeAsync(pt_slot, stream));
    if (keep_pos) *keep_pos = pt_pos; else LR_CUDA(cudaFreeAsync(pt_pos, stream));
    LR_CUDA(cudaFreeAsync(pt_bit, stream));
    if (dups != nullptr && !dup_given) { dups->flags = dup; dups->count = n_dup; }  // handed to the caller (who frees it)
    else if (!dup_given) LR_CUDA(cudaFreeAsync(dup, stream));
    LR_CUDA(cudaFreeAsync(counters, stream));
    LR_CUDA(cudaFreeAsync(bounds, stream));
    last_n_ = n; last_cap_ = cap; last_nbr_cap_ = nbr_cap;
    view_.slots = slots_; view_.cell_start = cell_start_; view_.pts = pts_;
    view_.canon = pts_; view_.w_is_pos = 0;
    view_.nbr_slots = nbr_; view_.nbr_mask = nbr_cap ? nbr_cap - 1 : 0;
    view_.slot_mask = cap - 1; view_.n_pts = n_kept; view_.n_unique = n_kept - n_dup;
    view_.inv_cell = inv_cell; view_.cell = cell;
    for (int a = 0; a < 3; ++a) { view_.cmin[a] = h_bounds[a]; view_.cmax[a] = h_bounds[3 + a]; }
    bytes_ = static_cast<size_t>(cap) * sizeof(VoxelSlot) + (static_cast<size_t>(n_cells_) + 1) * 4 +
             static_cast<size_t>(n_kept) * sizeof(float4) * (lists ? 28 : 1) + static_cast<size_t>(nbr_cap) * sizeof(NbrSlot);
}

// ---- block pyramid (voxel_map.cuh: PyrView) ---------------------------------------------------------------------------
__global__ void k_pyr_clear(PyrSlot* slots, size_t n) {
    const size_t s = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s < n) slots[s] = PyrSlot{kEmptyKey, 0ull};
}
__global__ void k_pyr_from_blocks(const VoxelSlot* __restrict__ slots, unsigned int cap, PyrSlot* up, unsigned int up_mask, unsigned int* counters) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && slots[s].key != kEmptyKey) build_pyr_body<DeviceAtomics>(slots[s].key, up, up_mask, counters);
}
__global__ void k_pyr_from_pyr(const PyrSlot* __restrict__ slots, unsigned int cap, PyrSlot* up, unsigned int up_mask, unsigned int* counters) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && slots[s].key != kEmptyKey) build_pyr_body<DeviceAtomics>(slots[s].key, up, up_mask, counters);
}
void DeviceVoxelMap::build_pyramid(cudaStream_t stream) {
    pyr_view_ = PyrView{};
    pyr_bytes_ = 0;
    if (view_.n_pts == 0) return;
    const int P = pyr_levels_for(view_.cmin, view_.cmax);
    // a level never has more nodes than the level below, and the fine level has n_blocks_ of them: load <= 0.5 everywhere
    const unsigned int cap = next_pow2(static_cast<size_t>(n_blocks_) * 2);
    PyrSlot* base = pyr_buf_.ensure(static_cast<size_t>(cap) * P);
    unsigned int* counters = nullptr;
    LR_CUDA(cudaMallocAsync(&counters, 2 * sizeof(unsigned int), stream));
    LR_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), stream));
    const unsigned int T = 256;
    const size_t total = static_cast<size_t>(cap) * P;
    LR_LAUNCH(k_pyr_clear, static_cast<unsigned int>((total + T - 1) / T), T, 0, stream, base, total);
    const unsigned int fine_cap = view_.slot_mask + 1;
    LR_LAUNCH(k_pyr_from_blocks, (fine_cap + T - 1) / T, T, 0, stream, view_.slots, fine_cap, base, cap - 1, counters);
    for (int l = 1; l < P; ++l)
        LR_LAUNCH(k_pyr_from_pyr, (cap + T - 1) / T, T, 0, stream, base + static_cast<size_t>(cap) * (l - 1), cap,
                  base + static_cast<size_t>(cap) * l, cap - 1, counters);
    LR_CUDA(cudaFreeAsync(counters, stream));
    for (int l = 0; l < P; ++l) { pyr_view_.slots[l] = base + static_cast<size_t>(cap) * l; pyr_view_.mask[l] = cap - 1; }
    pyr_view_.levels = P;
    pyr_bytes_ = total * sizeof(PyrSlot);
}

__global__ void k_coarse_remap(float4* pts, unsigned int n, const unsigned int* __restrict__ fine_pos_of_index) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) pts[j].w = __int_as_float(static_cast<int>(fine_pos_of_index[__float_as_int(pts[j].w)]));
}
// list entries carry the level's OWN canonical position; after k_coarse_remap that point's w is its fine position
__global__ void k_coarse_remap_lists(float4* pts, unsigned int n_pts, size_t n_entries) {
    const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j < n_entries) pts[n_pts + j].w = pts[__float_as_int(pts[n_pts + j].w)].w;
}
void DeviceVoxelMap::attach_to(const VoxelMapView& fine, const unsigned int* fine_pos_of_index, cudaStream_t stream) {
    if (view_.n_pts == 0) return;
    LR_LAUNCH(k_coarse_remap, (view_.n_pts + 255) / 256, 256, 0, stream, pts_, view_.n_pts, fine_pos_of_index);
    if (nbr_ != nullptr && n_list_entries_ != 0)
        LR_LAUNCH(k_coarse_remap_lists, static_cast<unsigned int>((n_list_entries_ + 255) / 256), 256, 0, stream, pts_, view_.n_pts,
                  n_list_entries_);
    view_.canon = fine.pts;
    view_.w_is_pos = 1;
}

void build_icp_maps(DeviceVoxelMap& fine, DeviceVoxelMap* coarse, DeviceVoxelMap* mid, const void* d_xyz, size_t n, size_t stride,
                    float cell, bool want_lists, cudaStream_t stream) {
    unsigned int* pos = nullptr;
    DeviceVoxelMap::DupFlags dups;
    fine.build(d_xyz, n, stride, cell, want_lists, stream, &pos, &dups);
    // the block pyramid is only built for the experimental ball-query kernel (LOCREG_PYR_KERNEL=1, locreg.cu)
    static const bool want_pyramid = getenv("LOCREG_PYR_KERNEL") != nullptr && atoi(getenv("LOCREG_PYR_KERNEL")) == 1;
    if (want_pyramid) fine.build_pyramid(stream);
    DeviceVoxelMap::DupFlags* dp = dups.flags ? &dups : nullptr;
    if (mid) {
        mid->build(d_xyz, n, stride, cell * kMidFactor, want_lists, stream, nullptr, dp);
        if (pos && mid->view().nbr_slots != nullptr) mid->attach_to(fine.view(), pos, stream);
        else mid->clear();  // no lists (memory): stage 2 takes the corner-list path
    }
    float c = cell;
    for (int l = 0; l < kCoarseLevels; ++l) {
        c *= kCoarseFactor;
        coarse[l].build(d_xyz, n, stride, c, false, stream, nullptr, dp);
        if (pos) coarse[l].attach_to(fine.view(), pos, stream);
    }
    if (pos) LR_CUDA(cudaFreeAsync(pos, stream));
    if (dups.flags) LR_CUDA(cudaFreeAsync(dups.flags, stream));
}

}  // namespace locreg

// ICP ScanMatch of ONE scan in ONE cooperative launch (BASELINE config 2: single-scan tracking).
//
// The per-iteration pipeline of icp_pipeline.cuh costs a single 27 k-point scan five to six launches per Gauss-Newton
// iteration - 51 launches for ten iterations, each a 5-15 us kernel of ~106 blocks with the GPU idle in between.  Here the
// whole loop of IcpRegistration::AlignP2P / AlignP2Line / AlignP2Plane (icp_registration.cpp:267-381) is one persistent
// kernel: the same tile bodies (icp_nn_tile, icp_rings_warps, icp_fit_group, icp_post_tile - so the same neighbours, planes
// and partial rows as the batch path), separated by grid barriers instead of launches:
//
//   per iteration:  search stage 1 on the block's tiles            -> unfinished queries to the queue
//                   grid barrier; queue not empty: one queued query per warp of the WHOLE grid, grid barrier
//                   plane fit + residuals + normal equations        -> one partial row per tile
//                   grid barrier; every block sums the rows in a fixed order and solves the 6x6 system itself: the
//                   new pose is in every block's shared memory without another exchange.
//
// Three barriers in an iteration with many queued queries (the first one or two); with few, the block that searched a
// tile finishes its queued queries itself and there is ONE barrier per iteration.
#pragma once
#include <cooperative_groups.h>

#include "icp_pipeline.cuh"

namespace locreg {

namespace cg = cooperative_groups;

// Sum of the partial rows [0, n_tiles) by the whole block in one round trip: thread (g, e) = (threadIdx.x / 32, % 32) adds
// entry e of the rows g, g + 8, g + 16, ... in that order (independent coalesced loads), the eight group sums are then
// added in group order.  A fixed order for a fixed scan - not the order of k_icp_solve, which walks the tiles with 30
// lanes: the persistent kernel's sums differ from the batch path's in the last bits.
__device__ __forceinline__ void persist_sum_rows(const double* __restrict__ partials, unsigned int n_tiles, double* stage, double* acc32) {
    const unsigned int e = threadIdx.x & 31u, g = threadIdx.x >> 5;
    double v = 0;
    for (unsigned int t = g; t < n_tiles; t += kTile / 32) v += __ldcg(partials + static_cast<size_t>(t) * kPartialDoubles + e);
    stage[g * 32 + e] = v;
    __syncthreads();
    if (threadIdx.x < 30) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < kTile / 32; ++k) s += stage[k * 32 + threadIdx.x];
        acc32[threadIdx.x] = s;
    }
    __syncthreads();
}

// partials: 2 * n_tiles rows, double-buffered by iteration parity (a fast block starts the next evaluation while a slow
// one still sums the rows of this one).  track_from: first iteration that runs the tracked search.
#ifndef LR_PERSIST_LOCAL_PER_TILE
#define LR_PERSIST_LOCAL_PER_TILE 16  // queued queries per tile (previous iteration, average) up to which a block serves its own
#endif
#ifndef LR_PERSIST_MIN_BLOCKS
#define LR_PERSIST_MIN_BLOCKS 1  // 174 registers, no spills: 0.436 ms per ScanMatch against 0.468 at 2 blocks per SM (128 registers), 0.49 at 3, 0.52 at 4
#endif
template <int METHOD>
__global__ void __launch_bounds__(kTile, LR_PERSIST_MIN_BLOCKS)
k_icp_persist(VoxelMapView map, CoarseLevels coarse, IcpParams prm, BatchView bv, AlignState* state, unsigned int n_tiles,
              unsigned int* nn_pos, unsigned char* plane_valid, KnnTrack* track, RingQueue queue, double* plane_cache,
              unsigned char* plane_stat, double* partials, int track_from, unsigned long long* dbg) {
    constexpr int K = METHOD == kIcpP2P ? 1 : 5;
    cg::grid_group grid = cg::this_grid();
    __shared__ AlignState s_state;  // this block's copy: every block takes the same steps from the same sums
    __shared__ double stage[(kTile / 32) * 32];
    __shared__ double acc32[32];
    if (threadIdx.x == 0) s_state = *state;
    __syncthreads();
    const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // dbg (tools only): %globaltimer of block 0 at the phase boundaries of every iteration, 6 stamps each
#define LR_STAMP(k)                                                                        \
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {                                      \
        unsigned long long t_;                                                             \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
        dbg[it * 6 + (k)] = t_;                                                            \
    }
    unsigned char* pv = METHOD == kIcpP2Plane ? plane_valid : nullptr;
    // Three (count, cursor) pairs of the stage-2 queue, used in rotation: iteration `it` appends to pair it % 3 and reads
    // its count after a grid barrier; pair (it + 1) % 3 - last read two barriers ago - is zeroed meanwhile.
    unsigned int* const counters = queue.count;
    __shared__ unsigned int s_pend[kTile];
    __shared__ unsigned int s_npend;
    unsigned int last_n = 0xFFFFFFFFu;  // queries the previous iteration queued (none yet: take the grid-wide form)
    for (int it = 0; it < prm.max_iteration && !s_state.stop; ++it) {
        const int mode = it == 0 ? kNnTwoPass : (it >= track_from ? (kNnSeeds | kNnTrack) : kNnSeeds);
        const RingQueue q_it{counters + 2 * (it % 3), queue.entries};
        if (blockIdx.x == 0 && threadIdx.x == 0) { counters[2 * ((it + 1) % 3)] = 0u; counters[2 * ((it + 1) % 3) + 1] = 0u; }
        double* rows = partials + static_cast<size_t>(it & 1) * n_tiles * kPartialDoubles;
        // Few queued queries (a handful per tile from the second or third iteration on): the block that searched the tile
        // finishes them itself, a warp each, and goes straight on to the tile's planes and normal equations - ONE grid
        // barrier per iteration instead of three, and no block waits for the slowest tile more than once.
        const bool local = last_n <= static_cast<unsigned int>(LR_PERSIST_LOCAL_PER_TILE) * n_tiles;
        LR_STAMP(0)
        if (local) {
            for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (threadIdx.x == 0) s_npend = 0u;
                __syncthreads();
                if (mode & kNnTrack) icp_nn_tile<K, true>(tile, map, bv, &s_state, 0, mode, nn_pos, pv, track, q_it, s_pend, &s_npend);
                else icp_nn_tile<K, false>(tile, map, bv, &s_state, 0, mode, nn_pos, pv, track, q_it, s_pend, &s_npend);
                __syncthreads();
                const unsigned int n_p = s_npend;
                for (unsigned int e = threadIdx.x >> 5; e < n_p; e += kTile / 32)
                    icp_rings_query<K>(make_uint2(s_pend[e], 0u), map, coarse, bv, &s_state, nn_pos, track);
                __syncthreads();
                if (METHOD == kIcpP2Plane) {
                    icp_fit_group<1>(tile, map, prm, bv, &s_state, 0, nn_pos, plane_valid, plane_cache, plane_stat, 1u);
                    __syncthreads();
                }
                icp_post_tile<METHOD>(tile, map, prm, bv, &s_state, 0, nn_pos, rows, nullptr, nullptr, plane_cache, plane_stat);
                __syncthreads();
            }
            LR_STAMP(1) LR_STAMP(2) LR_STAMP(3) LR_STAMP(4)
            __threadfence();
            grid.sync();
            last_n = __ldcg(q_it.count);
        } else {
            for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (mode & kNnTrack) icp_nn_tile<K, true>(tile, map, bv, &s_state, 0, mode, nn_pos, pv, track, q_it);
                else icp_nn_tile<K, false>(tile, map, bv, &s_state, 0, mode, nn_pos, pv, track, q_it);
                __syncthreads();
            }
            LR_STAMP(1)
            __threadfence();
            grid.sync();
            LR_STAMP(2)
            const unsigned int n_queued = __ldcg(q_it.count);
            last_n = n_queued;
            if (n_queued != 0u) {  // the same value in every block: nobody appends between the two barriers
                icp_rings_warps<K>(n_queued, warp, n_warps, map, coarse, bv, &s_state, nn_pos, track, q_it);
                __threadfence();
                grid.sync();
            }
            LR_STAMP(3)
            for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (METHOD == kIcpP2Plane) {
                    icp_fit_group<1>(tile, map, prm, bv, &s_state, 0, nn_pos, plane_valid, plane_cache, plane_stat, 1u);
                    __syncthreads();
                }
                icp_post_tile<METHOD>(tile, map, prm, bv, &s_state, 0, nn_pos, rows, nullptr, nullptr, plane_cache, plane_stat);
                __syncthreads();
            }
            LR_STAMP(4)
            __threadfence();
            grid.sync();
        }
        LR_STAMP(5)
        persist_sum_rows(rows, n_tiles, stage, acc32);
        if (threadIdx.x == 0) {
            result_from_acc(s_state.res, acc32);
            Pose T;
            pose_load(T, s_state.pose);
            s_state.res.iters += 1;
            const int outcome = icp_gn_update<METHOD>(acc32, static_cast<unsigned int>(acc32[28]), prm, T);
            if (apply_outcome(outcome, s_state.res) || s_state.res.iters >= prm.max_iteration) s_state.stop = 1;
            pose_store(T, s_state.pose);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *state = s_state;
#undef LR_STAMP
}

}  // namespace locreg

// Point-to-plane ICP, tracked iterations: search + residual + normal equations in ONE pass over the scan.
//
// Once the pose has nearly settled a Gauss-Newton iteration moves a query by millimetres, and the KnnTrack margin of
// its last full search (voxel_map.cuh) proves that its five nearest map points are still the same SET.  The plane
// through five points does not depend on their order (beyond rounding, 1e-16 relative), so for such a point the plane
// cached by the last fit is still THE plane: the iteration needs the point (16 B), its margin record (16 B) and the
// plane (32 B + 1 B status) - no neighbour gather, no sort, no store.  The three-kernel pipeline of icp_pipeline.cuh
// read the point twice and its five neighbours once for the same result.  Here:
//
//   k_icp_track_p2plane   per tile of 256 points, a pure streaming pass: margin check -> residual row from the cached
//                         plane -> per-warp Gram matrix -> partial row A of the tile.  Points outside their margin
//                         (3 % by the tenth iteration) are appended to a global queue, one atomic per tile - left in
//                         the block they would keep it on its SM waiting for one warp with a few busy lanes.
//   k_icp_rescan          one queued point per thread, every lane busy: exact tracked search.  The same ordered
//                         neighbours as before: the cached plane stays (flag 2 = residual outstanding); otherwise the
//                         point needs a new fit (flag 0); unfinished searches also go to the stage-2 queue.
//   k_icp_fit_queue       after stage 2: the queued points whose neighbours changed (1 %), compacted per block so that
//                         every lane runs a fit -> plane cache, flag 2.
//   k_icp_pending         per group of tiles: the flag-2 points in point order, residual row from the cached plane,
//                         fixed-order sum per tile -> partial row B.
//
// k_icp_solve adds the A rows and then the B rows of a scan in tile order: the sums are reproducible and do not
// depend on how many scans share the launch (a batch equals single ScanMatch calls bit for bit).
// Reference: IcpRegistration::CaculateMatrixHAndBP2Plane (icp_registration.cpp:161-213).
#pragma once
#include "icp_pipeline.cuh"

namespace locreg {

// per-slot flags of a tile's residual rows
constexpr unsigned char kRowEff = 1, kRowInl = 2;

struct FlagSink {  // the Acc policy of icp_p2plane_residual: one row in shared memory + two flag bits
    double* row_out;
    unsigned char flags;
    __device__ __forceinline__ void row(const double (&J)[6], double r) {
#pragma unroll
        for (int i = 0; i < 6; ++i) row_out[i] = J[i];
        row_out[6] = r;
    }
    __device__ __forceinline__ void inc_eff() { flags |= kRowEff; }
    __device__ __forceinline__ void inc_inl() { flags |= kRowInl; }
};

// Per-warp Gram matrix of 32 staged rows [J | r] on the fp64 tensor cores (see k_icp_post) and the tile's partial row.
// rows: kTile * kRowStride (+ 1 pad) doubles, flags: one byte per slot.  All threads of the block call it after the
// barrier that completes `rows`.
__device__ __forceinline__ void tile_gram_store(const double* rows, const unsigned char* flags, double (*gram)[64], int (*counts)[2],
                                                double* __restrict__ partial_row) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned char f = flags[threadIdx.x];
    const unsigned int eff_mask = __ballot_sync(0xffffffffu, (f & kRowEff) != 0);
    const unsigned int inl_mask = __ballot_sync(0xffffffffu, (f & kRowInl) != 0);
    const double* wrows = rows + warp * 32 * kRowStride + (lane & 3) * kRowStride + (lane >> 2);
    double g0 = 0.0, g1 = 0.0;  // G[lane >> 2][2 * (lane & 3) + {0, 1}]
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const double a = lane < 28 ? wrows[4 * ks * kRowStride] : 0.0;  // (column 7 does not exist, see k_icp_post)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(g0), "+d"(g1)
                     : "d"(a), "d"(a));
    }
    gram[warp][(lane >> 2) * 8 + 2 * (lane & 3)] = g0;
    gram[warp][(lane >> 2) * 8 + 2 * (lane & 3) + 1] = g1;
    if (lane == 0) {
        counts[warp][0] = __popc(eff_mask);
        counts[warp][1] = __popc(inl_mask);
    }
    __syncthreads();
    if (threadIdx.x < 30) {
        constexpr unsigned long long kTriRow = tri_table(false), kTriCol = tri_table(true);
        const int e = threadIdx.x;
        double t = 0;
        if (e < 28) {
            int ea = 6, eb = 6;
            if (e < 21) {
                ea = static_cast<int>(kTriRow >> (3 * e)) & 7;
                eb = static_cast<int>(kTriCol >> (3 * e)) & 7;
            } else if (e < 27) {
                ea = e - 21;
            }
#pragma unroll
            for (int w = 0; w < kTile / 32; ++w) t += gram[w][ea * 8 + eb];
            if (e >= 21 && e < 27) t = -t;
        } else {
            int c = 0;
#pragma unroll
            for (int w = 0; w < kTile / 32; ++w) c += counts[w][e - 28];
            t = static_cast<double>(c);
        }
        partial_row[threadIdx.x] = t;
    }
}

// plane_valid[] states of a point in the tracked iterations
constexpr unsigned char kPlaneStale = 0;     // neighbours changed: fit + residual outstanding (k_icp_pending)
constexpr unsigned char kPlaneCurrent = 1;   // plane valid, residual accumulated (or nothing to do)
constexpr unsigned char kPlaneResidual = 2;  // plane valid, residual outstanding (k_icp_pending)

// One tracked search of scratch row `srow` (point sp, pose T): neighbours, margin record and flag are updated in place.
// Returns false when stage 2 has to finish the search.
__device__ __forceinline__ bool track_rescan_point(const VoxelMapView& map, const Pose& T, const float4 sp, size_t srow,
                                                   unsigned int* __restrict__ nn_pos, unsigned char* __restrict__ plane_valid, KnnTrack* track) {
    constexpr int K = 5;
    unsigned int* out = nn_pos + srow * K;
    unsigned int seeds[K];
#pragma unroll
    for (int j = 0; j < K; ++j) seeds[j] = out[j];
    double wx, wy, wz;
    pose_apply(T, static_cast<double>(sp.x), static_cast<double>(sp.y), static_cast<double>(sp.z), wx, wy, wz);
    KnnResult<K> nn;
    KnnTrack tr;
    const bool done = knn_query_fast_track<K>(map, static_cast<float>(wx), static_cast<float>(wy), static_cast<float>(wz), nn, seeds, tr);
    bool same = done;
    same = same && knn_same_set<K>(seeds, nn.pos);
    track[srow] = tr;
    if (same) {
        // the same ordered neighbours as the last fit saw: its plane is what a new fit would return bit for bit
        plane_valid[srow] = kPlaneResidual;
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = nn.pos[j];
        plane_valid[srow] = kPlaneStale;
    }
    return done;
}

#ifndef LR_TRACK_MIN_BLOCKS
#define LR_TRACK_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(kTile, LR_TRACK_MIN_BLOCKS)
k_icp_track_p2plane(VoxelMapView map, IcpParams prm, BatchView bv, const AlignState* __restrict__ states, int ignore_stop,
                    const KnnTrack* __restrict__ track, RingQueue rescan, const double* __restrict__ plane_cache,
                    const unsigned char* __restrict__ plane_stat, double* __restrict__ partials) {
    __shared__ Pose T;
    __shared__ double rows[kTile * kRowStride + 1];
    __shared__ double gram[kTile / 32][64];
    __shared__ int counts[kTile / 32][2];
    __shared__ unsigned int blk_scan, blk_base;
    __shared__ unsigned char scan_list[kTile];
    __shared__ unsigned char flags[kTile];
    const TileCoord tc = locate_tile(bv, blockIdx.x);
    if (!tc.valid) return;
    const AlignState* st = states + tc.scan;
    if (st->stop && !ignore_stop) return;
    if (threadIdx.x == 0) {
        pose_load(T, st->pose);
        blk_scan = 0u;
    }
    // everything a point inside its margin needs is requested at once, before the barrier the pose needs
    const bool mine = threadIdx.x < tc.count;
    const unsigned int p = tc.first + (mine ? threadIdx.x : 0u);
    const size_t row = static_cast<size_t>(tc.out_base + p);
    float4 sp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    KnnTrack t;
    t.qx = t.qy = t.qz = 0.0f; t.margin = -1.0f;
    double4 pl = make_double4(0, 0, 0, 0);
    unsigned char pst = kPlaneNone;
    if (mine) {
        sp = bv.src[tc.src_base + p];
        t = track[row];
        pl = reinterpret_cast<const double4*>(plane_cache)[row];
        pst = plane_stat[row];
    }
    {
        double* r = rows + threadIdx.x * kRowStride;
#pragma unroll
        for (int i = 0; i < kRowStride; ++i) r[i] = 0.0;  // points without a residual contribute zero rows
        if (threadIdx.x == 0) rows[kTile * kRowStride] = 0.0;
    }
    __syncthreads();
    bool need_scan = false;
    unsigned char fl = 0;
    if (mine && finite3(sp.x, sp.y, sp.z) && map.n_pts != 0) {
        const double qx = sp.x, qy = sp.y, qz = sp.z;
        double wx, wy, wz;
        pose_apply(T, qx, qy, qz, wx, wy, wz);
        if (knn_track_holds(t, static_cast<float>(wx), static_cast<float>(wy), static_cast<float>(wz))) {
            const double n[4] = {pl.x, pl.y, pl.z, pl.w};
            FlagSink sink{rows + threadIdx.x * kRowStride, 0};
            icp_p2plane_residual(prm, T, qx, qy, qz, wx, wy, wz, pst, n, sink);
            fl = sink.flags;
        } else {
            need_scan = true;
        }
    }
    flags[threadIdx.x] = fl;
    {
        const unsigned int lane = threadIdx.x & 31;
        const unsigned int mask = __ballot_sync(0xffffffffu, need_scan);
        unsigned int base = 0;
        if (lane == 0 && mask != 0u) base = atomicAdd(&blk_scan, static_cast<unsigned int>(__popc(mask)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (need_scan) scan_list[base + __popc(mask & ((1u << lane) - 1u))] = static_cast<unsigned char>(threadIdx.x);
    }
    __syncthreads();
    const unsigned int n_scan = blk_scan;
    if (n_scan != 0u) {  // one atomic per tile; the entries of a tile stay together (neighbouring rays, the same lists)
        if (threadIdx.x == 0) blk_base = atomicAdd(rescan.count, n_scan);
        __syncthreads();
        if (threadIdx.x < n_scan)
            rescan.entries[blk_base + threadIdx.x] = make_uint2(static_cast<unsigned int>(tc.out_base + tc.first + scan_list[threadIdx.x]), tc.scan);
    }
    tile_gram_store(rows, flags, gram, counts, partials + static_cast<size_t>(blockIdx.x) * kPartialDoubles);
}

// The queued points of k_icp_track_p2plane, one per thread (grid-stride: the queue length is only known on the device).
// Unfinished searches are appended to the stage-2 queue warp by warp.
__global__ void __launch_bounds__(128, 8)
k_icp_rescan(VoxelMapView map, BatchView bv, const AlignState* __restrict__ states, unsigned int* __restrict__ nn_pos,
             unsigned char* __restrict__ plane_valid, KnnTrack* track, RingQueue queue, RingQueue rescan) {
    const unsigned int n = *rescan.count;
    const unsigned int lane = threadIdx.x & 31;
    for (unsigned int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned int e = base + threadIdx.x;
        bool done = true;
        uint2 q = make_uint2(0u, 0u);
        if (e < n) {
            q = rescan.entries[e];
            const size_t src_idx = bv.offsets ? static_cast<size_t>(q.x) : static_cast<size_t>(q.x) - static_cast<size_t>(q.y) * bv.n_single;
            Pose T;
            pose_load(T, states[q.y].pose);
            done = track_rescan_point(map, T, bv.src[src_idx], q.x, nn_pos, plane_valid, track);
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, !done);
        if (mask != 0u) {
            unsigned int qb = 0;
            if (lane == 0) qb = atomicAdd(queue.count, static_cast<unsigned int>(__popc(mask)));
            qb = __shfl_sync(0xffffffffu, qb, 0);
            if (!done) queue.entries[qb + __popc(mask & ((1u << lane) - 1u))] = q;
        }
    }
}

// After both search stages: the queued points whose neighbours changed get a new plane.  A block takes kFitChunk queue
// entries at a time, compacts the stale ones in shared memory and fits them with every lane busy (a fit is ~500
// dependent fp64 instructions).  Flag 0 -> 2.
constexpr int kFitChunk = 1024;
__global__ void __launch_bounds__(128, 4)
k_icp_fit_queue(VoxelMapView map, IcpParams prm, const unsigned int* __restrict__ nn_pos, unsigned char* plane_valid, double* plane_cache,
                unsigned char* plane_stat, RingQueue rescan) {
    __shared__ unsigned int list[kFitChunk];
    __shared__ unsigned int list_n;
    const unsigned int n = *rescan.count;
    const unsigned int lane = threadIdx.x & 31;
    for (unsigned int base = blockIdx.x * kFitChunk; base < n; base += gridDim.x * kFitChunk) {
        if (threadIdx.x == 0) list_n = 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kFitChunk / 128; ++k) {
            const unsigned int e = base + k * 128 + threadIdx.x;
            unsigned int row = 0;
            bool stale = false;
            if (e < n) {
                row = rescan.entries[e].x;
                stale = plane_valid[row] == kPlaneStale;
            }
            const unsigned int mask = __ballot_sync(0xffffffffu, stale);
            unsigned int lb = 0;
            if (lane == 0 && mask != 0u) lb = atomicAdd(&list_n, static_cast<unsigned int>(__popc(mask)));
            lb = __shfl_sync(0xffffffffu, lb, 0);
            if (stale) list[lb + __popc(mask & ((1u << lane) - 1u))] = row;
        }
        __syncthreads();
        const unsigned int m = list_n;
        for (unsigned int i = threadIdx.x; i < m; i += 128) {
            const size_t row = list[i];
            KnnResult<5> nn;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                nn.pos[j] = nn_pos[row * 5 + j];
                nn.d2[j] = 0.0f;  // not needed by the fit
            }
            double pl[4] = {0, 0, 0, 0};
            plane_stat[row] = icp_fit_plane(map, prm, nn, pl);
            reinterpret_cast<double4*>(plane_cache)[row] = make_double4(pl[0], pl[1], pl[2], pl[3]);
            plane_valid[row] = kPlaneResidual;
        }
        __syncthreads();
    }
}

// The flag-2 points of `group` consecutive tiles: residual rows from the cached planes, summed per tile in point order
// into the tile's partial row B (partials_b + tile * kPartialDoubles; zeros for a tile without any).  Flag 2 -> 1.
// Launched after both search stages and the fits: also re-arms the stage-2 queue counters for the next evaluation.
constexpr int kPendGroup = 16;
#ifndef LR_PEND_MIN_BLOCKS
#define LR_PEND_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kTile, LR_PEND_MIN_BLOCKS)
k_icp_pending(IcpParams prm, BatchView bv, const AlignState* __restrict__ states, int ignore_stop, unsigned char* plane_valid,
              const double* __restrict__ plane_cache, const unsigned char* __restrict__ plane_stat, unsigned int group,
              double* __restrict__ partials_b, unsigned int* ring_count) {
    __shared__ TileCoord tcs[kPendGroup];
    __shared__ Pose Ts[kPendGroup];
    __shared__ unsigned short warp_cnt[kPendGroup][kTile / 32];
    __shared__ unsigned int seg_begin[kPendGroup + 1];
    __shared__ double rows[kTile * kRowStride];
    __shared__ unsigned char rflags[kTile];
    __shared__ double acc[kPendGroup][32];
    if (blockIdx.x == 0 && threadIdx.x == 0) { ring_count[0] = 0u; ring_count[1] = 0u; }
    if (threadIdx.x < group) {
        TileCoord c = locate_tile_thread(bv, blockIdx.x * group + threadIdx.x);
        if (c.valid && states[c.scan].stop && !ignore_stop) c.valid = false;
        tcs[threadIdx.x] = c;
        if (c.valid) pose_load(Ts[threadIdx.x], states[c.scan].pose);
    }
    for (unsigned int i = threadIdx.x; i < kPendGroup * 32u; i += kTile) (&acc[0][0])[i] = 0.0;
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the thread's flags of all tiles first (independent loads), then deterministic positions: tile-major, point order
    unsigned int need_bits = 0u;
#pragma unroll
    for (unsigned int g = 0; g < static_cast<unsigned int>(kPendGroup); ++g) {
        if (g < group && tcs[g].valid && threadIdx.x < tcs[g].count) {
            const size_t row = static_cast<size_t>(tcs[g].out_base + tcs[g].first + threadIdx.x);
            need_bits |= (plane_valid[row] != kPlaneCurrent ? 1u : 0u) << g;
        }
    }
    for (unsigned int g = 0; g < group; ++g) {
        const unsigned int mask = __ballot_sync(0xffffffffu, (need_bits >> g) & 1u);
        if (lane == 0) warp_cnt[g][warp] = static_cast<unsigned short>(__popc(mask));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (unsigned int g = 0; g < group; ++g) {
            seg_begin[g] = run;
            for (int w = 0; w < kTile / 32; ++w) run += warp_cnt[g][w];
        }
        seg_begin[group] = run;
    }
    __syncthreads();
    const unsigned int n = seg_begin[group];
    // chunk by chunk of kTile flagged points: a point belongs to the thread that holds its flag and finds its place in
    // the chunk by position, so no list is materialised
    for (unsigned int c0 = 0; c0 < n; c0 += kTile) {
        for (unsigned int g = 0; g < group; ++g) {
            const bool need = (need_bits >> g) & 1u;
            const unsigned int mask = __ballot_sync(0xffffffffu, need);
            if (!need) continue;
            unsigned int pos = seg_begin[g] + __popc(mask & ((1u << lane) - 1u));
            for (unsigned int w = 0; w < warp; ++w) pos += warp_cnt[g][w];
            if (pos < c0 || pos >= c0 + kTile) continue;
            const unsigned int i = pos - c0;
            const TileCoord& tc = tcs[g];
            const unsigned int ps = tc.first + threadIdx.x;
            const size_t row = static_cast<size_t>(tc.out_base + ps);
            const double4 pc = reinterpret_cast<const double4*>(plane_cache)[row];
            const unsigned char pst = plane_stat[row];
            const float4 sp = bv.src[tc.src_base + ps];
            plane_valid[row] = kPlaneCurrent;
            double* r = rows + i * kRowStride;
#pragma unroll
            for (int k = 0; k < kRowStride; ++k) r[k] = 0.0;
            FlagSink sink{r, 0};
            if (finite3(sp.x, sp.y, sp.z)) {
                const double plv[4] = {pc.x, pc.y, pc.z, pc.w};
                const double qx = sp.x, qy = sp.y, qz = sp.z;
                double wx, wy, wz;
                pose_apply(Ts[g], qx, qy, qz, wx, wy, wz);
                icp_p2plane_residual(prm, Ts[g], qx, qy, qz, wx, wy, wz, pst, plv, sink);
            }
            rflags[i] = sink.flags;
        }
        __syncthreads();
        // fixed-order sums: warp w owns the tiles g = w, w + 8, ...; lane e < 28 owns one entry, lanes 28 / 29 the counts
        const unsigned int c1 = min(c0 + kTile, n);
        for (unsigned int g = warp; g < group; g += kTile / 32) {
            const unsigned int b = max(seg_begin[g], c0), e = min(seg_begin[g + 1], c1);
            if (b >= e) continue;
            constexpr unsigned long long kTriRow = tri_table(false), kTriCol = tri_table(true);
            int ea = 6, eb = 6;
            if (lane < 21) {
                ea = static_cast<int>(kTriRow >> (3 * lane)) & 7;
                eb = static_cast<int>(kTriCol >> (3 * lane)) & 7;
            } else if (lane < 27) {
                ea = static_cast<int>(lane) - 21;
            }
            double s = acc[g][lane];
            if (lane < 28) {
                for (unsigned int i = b; i < e; ++i) {
                    const double* r = rows + (i - c0) * kRowStride;
                    s += r[ea] * r[eb];
                }
            } else if (lane < 30) {
                const unsigned char bit = lane == 28 ? kRowEff : kRowInl;
                for (unsigned int i = b; i < e; ++i) s += (rflags[i - c0] & bit) ? 1.0 : 0.0;
            }
            acc[g][lane] = s;
        }
        __syncthreads();
    }
    // partial rows B (B[a] = -sum J[a] r)
    for (unsigned int i = threadIdx.x; i < group * 32u; i += kTile) {
        const unsigned int g = i >> 5, e = i & 31u;
        const unsigned int tile = blockIdx.x * group + g;
        if (!tcs[g].valid || e >= 30u) continue;
        const double v = acc[g][e];
        partials_b[static_cast<size_t>(tile) * kPartialDoubles + e] = (e >= 21u && e < 27u) ? -v : v;
    }
}

}  // namespace locreg

// Stage 1 of the exact k-NN search with the candidate lists of a tile STAGED IN SHARED MEMORY.
//
// k_icp_nn (icp_pipeline.cuh) gives every query its own stream of dependent global loads: hash probe, then the ~32
// entries of its cell's neighbourhood list four at a time - ten round trips to L2 / HBM per query, which is what bounds
// the kernel (ncu: 6.6 warps per issue slot waiting on the long scoreboard).  But the 256 points of a tile are
// consecutive rays of one scan line: they fall into ~35 distinct cells.  Here the tile
//   1. transforms its points and de-duplicates their cell keys in a shared-memory hash table,
//   2. probes the map's cell table once per DISTINCT cell,
//   3. copies the distinct lists into shared memory with all 256 threads issuing independent coalesced 16 B loads
//      (the whole candidate set of the tile in ONE round trip),
//   4. and only then lets every query scan its list - from shared memory, where a dependent access costs ~30 cycles.
// Lists that do not fit the staging buffer are scanned from global memory as before (dense corners of the map).
// The candidates a query sees, their order and the k-best bookkeeping are those of knn_query_fast /
// knn_query_fast_track (voxel_map.cuh): same result bit for bit (tests/test_gpu_parity.py runs both).
// Replaces KdTree::GetClosestPoint (kdtree.cpp:147-236) for the points of one tile.
#pragma once
#include "icp_pipeline.cuh"

namespace locreg {

#ifndef LR_STAGE_CAP
#define LR_STAGE_CAP 2048  // staged list entries per tile (32 KB)
#endif
constexpr unsigned int kStageCap = LR_STAGE_CAP;
constexpr unsigned int kStageTab = 512;  // de-duplication table slots (256 keys at most: load <= 0.5)

// ---- branch-free k-best selection ------------------------------------------------------------------------------------
// knn_offer (voxel_map.cuh) is cheap when a candidate is rejected and ~65 divergent instructions when it is inserted;
// with 32 unrelated queries per warp SOME lane inserts at almost every step of a poorly seeded scan, so the warp pays
// for the insertion path nearly every candidate at a fifth of its lanes.  Here every candidate of every lane goes
// through the same 5-slot compare-exchange chain on the DISTANCE alone (1 FSETP + 4 SEL per slot, NaN distances of
// masked duplicates fall straight through), the value that drops off the end feeds d6 = the smallest distance left
// outside the set, and the exact order on ties - (dis2, original index), which needs the index of both points - is
// restored afterwards: a result is only ambiguous when two of its distances, or its last one and d6, are EQUAL as
// floats, and then the query is simply redone with knn_offer (rare: jittered coordinates hardly ever tie).
template <int K>
struct KnnSel {
    float d[K];
    unsigned int p[K];
    float d6;
};
template <int K>
__device__ __forceinline__ void sel_init(KnnSel<K>& s) {
#pragma unroll
    for (int j = 0; j < K; ++j) { s.d[j] = INFINITY; s.p[j] = kNoPos; }
    s.d6 = INFINITY;
}
template <int K>
__device__ __forceinline__ void sel_push(KnnSel<K>& s, float d, unsigned int p) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const bool lt = d < s.d[j];  // false for NaN: a masked duplicate never enters
        const float nd = lt ? d : s.d[j], cd = lt ? s.d[j] : d;
        const unsigned int np = lt ? p : s.p[j], cp = lt ? s.p[j] : p;
        s.d[j] = nd; s.p[j] = np;
        d = cd; p = cp;
    }
    s.d6 = fminf(s.d6, d);  // what fell off the end (rejected or evicted); fminf drops NaN
}
// true when the order / membership of the selection does not depend on the index tie-break
template <int K>
__device__ __forceinline__ bool sel_unambiguous(const KnnSel<K>& s) {
    bool ok = !(s.d[K - 1] == s.d6);
#pragma unroll
    for (int j = 0; j + 1 < K; ++j) ok = ok && !(s.d[j] == s.d[j + 1] && s.p[j + 1] != kNoPos);
    return ok;
}

LR_HD void knn_find_list_key(const VoxelMapView& m, unsigned long long key, unsigned int& beg, unsigned int& cnt) {
    unsigned int h = hash_block(key) & m.nbr_mask;
    beg = 0; cnt = 0;
    while (true) {
        const NbrSlot s = m.nbr_slots[h];
        if (s.key == key) { beg = s.start; cnt = s.count; return; }
        if (s.key == kEmptyKey) return;
        h = (h + 1) & m.nbr_mask;
    }
}

// MODE 0: unseeded, threshold pre-pass (first Gauss-Newton iteration); 1: seeded with the previous iteration's
// neighbours; 2: seeded and tracked (records the margin of every finished search, KnnTrack).
// The map must have neighbourhood lists (map.nbr_slots != nullptr).
#ifndef LR_STAGED_MIN_BLOCKS
#define LR_STAGED_MIN_BLOCKS 4
#endif
template <int K, int MODE>
__global__ void __launch_bounds__(kTile, LR_STAGED_MIN_BLOCKS)
k_icp_nn_staged(VoxelMapView map, BatchView bv, const AlignState* __restrict__ states, int ignore_stop, unsigned int* __restrict__ nn_pos,
                unsigned char* __restrict__ plane_valid, KnnTrack* track, RingQueue queue) {
    __shared__ Pose T;
    __shared__ float4 s_cand[kStageCap];
    __shared__ unsigned long long s_key[kStageTab];
    __shared__ unsigned short s_slot_id[kStageTab];
    __shared__ unsigned long long s_lkey[kTile];   // key of distinct list i
    __shared__ unsigned int s_beg[kTile], s_cnt[kTile];
    __shared__ unsigned int s_off[kTile + 1];      // staging offset of list i (kStageCap: not staged)
    __shared__ unsigned int s_nlists;
    __shared__ unsigned int blk_pending, blk_done;
    __shared__ unsigned int staged[kTile];
    const TileCoord tc = locate_tile(bv, blockIdx.x);
    if (!tc.valid) return;
    const AlignState* st = states + tc.scan;
    if (st->stop && !ignore_stop) return;
    if (threadIdx.x == 0) {
        pose_load(T, st->pose);
        blk_pending = 0u; blk_done = 0u; s_nlists = 0u;
    }
    for (unsigned int i = threadIdx.x; i < kStageTab; i += kTile) s_key[i] = kEmptyKey;
    const bool in_tile = threadIdx.x < tc.count;
    const unsigned int p = tc.first + (in_tile ? threadIdx.x : 0u);
    const float4 sp = bv.src[tc.src_base + p];
    const size_t row = static_cast<size_t>(tc.out_base + p);
    unsigned int* out = nn_pos + row * K;
    unsigned int seeds[K];
#pragma unroll
    for (int j = 0; j < K; ++j) seeds[j] = (MODE != 0 && in_tile) ? out[j] : kNoPos;
    __syncthreads();
    // ---- 1. transform, cell key, de-duplication
    const bool valid = in_tile && finite3(sp.x, sp.y, sp.z) && map.n_pts != 0;
    float qx = 0.0f, qy = 0.0f, qz = 0.0f;
    KnnCellFrame c{};
    bool listed = false;
    unsigned int my_slot = 0;
    if (valid) {
        double wx, wy, wz;
        pose_apply(T, static_cast<double>(sp.x), static_cast<double>(sp.y), static_cast<double>(sp.z), wx, wy, wz);
        qx = static_cast<float>(wx); qy = static_cast<float>(wy); qz = static_cast<float>(wz);
        c = knn_frame(map, qx, qy, qz);
        listed = c.R0 == 1;  // the 3x3x3 box around the query's cell touches the map: the list scan applies
        if (listed) {
            const unsigned long long key = pack_cell(c.fx, c.fy, c.fz);
            unsigned int h = hash_block(key) & (kStageTab - 1);
            while (true) {
                const unsigned long long prev = atomicCAS(&s_key[h], kEmptyKey, key);
                if (prev == kEmptyKey) {
                    const unsigned int id = atomicAdd(&s_nlists, 1u);
                    s_slot_id[h] = static_cast<unsigned short>(id);
                    s_lkey[id] = key;
                    break;
                }
                if (prev == key) break;
                h = (h + 1) & (kStageTab - 1);
            }
            my_slot = h;
        }
    }
    // the seeds' points are requested now and looked at after the staging (their round trip hides behind it)
    KnnResult<K> res;
    knn_init(res);
    float4 sd[K];
    bool seeds_all = MODE != 0 && valid;
#pragma unroll
    for (int j = 0; j < K; ++j) seeds_all = seeds_all && seeds[j] < map.n_pts;
    if (seeds_all) {
#pragma unroll
        for (int j = 0; j < K; ++j) sd[j] = map.pts[seeds[j]];
    }
    __syncthreads();
    // ---- 2. one probe of the map's cell table per distinct cell
    const unsigned int n_lists = s_nlists;
    if (threadIdx.x < n_lists) {
        unsigned int beg = 0, cnt = 0;
        knn_find_list_key(map, s_lkey[threadIdx.x], beg, cnt);
        s_beg[threadIdx.x] = beg;
        s_cnt[threadIdx.x] = cnt;
    }
    __syncthreads();
    // ---- 3. staging offsets: lists are staged in id order while they fit (a list is staged whole or not at all; what
    //         does not fit is scanned from global memory)
    if (threadIdx.x < 32) {
        const unsigned int lane = threadIdx.x;
        constexpr unsigned int per = kTile / 32;
        unsigned int cnts[per], sum = 0;
#pragma unroll
        for (unsigned int k = 0; k < per; ++k) {
            const unsigned int i = lane * per + k;
            cnts[k] = i < n_lists ? s_cnt[i] : 0u;
            sum += cnts[k];
        }
        unsigned int inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= static_cast<unsigned int>(off)) inc += v;
        }
        unsigned int run = inc - sum;
#pragma unroll
        for (unsigned int k = 0; k < per; ++k) {
            const unsigned int i = lane * per + k;
            const unsigned int incl = run + cnts[k];
            if (i < n_lists) {
                const bool fits = incl <= kStageCap;  // prefix sums grow: the staged lists are a prefix of the id order
                s_off[i] = fits ? run : kStageCap;
            }
            run = incl;
        }
    }
    __syncthreads();
    // ---- 4. cooperative copy: 32 groups of 8 lanes, a group per staged list, each lane every eighth entry (independent
    //         coalesced 16 B loads: the tile's whole candidate set is in flight at once)
    for (unsigned int l = threadIdx.x >> 3; l < n_lists; l += kTile / 8) {
        const unsigned int off = s_off[l];
        if (off >= kStageCap) continue;
        const unsigned int cnt = s_cnt[l];
        const float4* __restrict__ src = map.pts + s_beg[l];
        for (unsigned int i = threadIdx.x & 7u; i < cnt; i += 8) s_cand[off + i] = src[i];
    }
    __syncthreads();
    // ---- 5. every query scans its list: branch-free selection over the list, then over the seeds that are not in it
    bool done = true, same = false;
    KnnTrack tr;
    tr.qx = qx; tr.qy = qy; tr.qz = qz; tr.margin = -1.0f;
    if (valid) {
        done = false;
        unsigned int id = 0, cnt = 0, off = kStageCap;
        if (listed) {
            id = s_slot_id[my_slot];
            cnt = s_cnt[id];
            off = s_off[id];
        }
        KnnSel<K> sel;
        sel_init(sel);
        if (off < kStageCap) {
            const float4* lst = s_cand + off;
#pragma unroll 4
            for (unsigned int i = 0; i < cnt; ++i) {
                const float4 p = lst[i];
                sel_push(sel, dis2_f32(qx, qy, qz, p.x, p.y, p.z), static_cast<unsigned int>(float_as_int(p.w)));
            }
        } else if (listed) {
            const float4* lst = map.pts + s_beg[id];
#pragma unroll 4
            for (unsigned int i = 0; i < cnt; ++i) {
                const float4 p = lst[i];
                sel_push(sel, dis2_f32(qx, qy, qz, p.x, p.y, p.z), static_cast<unsigned int>(float_as_int(p.w)));
            }
        }
        if (MODE != 0) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const unsigned int sj = seeds[j];
                if (sj >= map.n_pts) continue;
                bool member = false;
#pragma unroll
                for (int k = 0; k < K; ++k) member = member || sel.p[k] == sj;
                const float4 pt = seeds_all ? sd[j] : map.pts[sj];
                // (a seed the list scan already rejected is rejected again and leaves d6 as it is)
                if (!member) sel_push(sel, dis2_f32(qx, qy, qz, pt.x, pt.y, pt.z), sj);
            }
        }
        if (sel_unambiguous(sel)) {
#pragma unroll
            for (int j = 0; j < K; ++j) { res.d2[j] = sel.d[j]; res.pos[j] = sel.p[j]; }
        } else {  // exact float ties: the (dis2, original index) order decides - redo this query with knn_offer
            knn_init(res);
            knn_seed<K>(map, qx, qy, qz, seeds, res);
            sel.d6 = INFINITY;
            if (listed) {
                const float4* lst = off < kStageCap ? s_cand + off : map.pts + s_beg[id];
                for (unsigned int i = 0; i < cnt; ++i) {
                    const float4 p = lst[i];
                    knn_offer_track(map.pts, res, dis2_f32(qx, qy, qz, p.x, p.y, p.z), static_cast<unsigned int>(float_as_int(p.w)), sel.d6);
                }
            }
        }
        if (listed) {
            done = knn_list_final<K>(map, c, res);
            if (MODE == 2 && done) tr.margin = knn_track_margin<K>(map, c, res, sel.d6, qx, qy, qz);
        }
        same = done && MODE != 0;
        same = same && knn_same_set<K>(seeds, res.pos);
    }
    if (in_tile) {
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = res.pos[j];
        if (plane_valid && !same) plane_valid[row] = 0;  // k_icp_fit sets it again
        if (track) track[row] = tr;  // margin -1 unless this was a tracked search that ended here
    }
    tile_queue_append(!done, static_cast<unsigned int>(row), tc.scan, &blk_pending, &blk_done, staged, queue);
}

}  // namespace locreg

// Warp-cooperative stage 2 of the exact k-NN search (device only): ONE queued query per warp.
//
// The serial stage 2 (voxel_map.cuh: knn_query_corners / knn_query_rings, kept as the host-testable statement of the
// algorithm and as the in-place finish of single-scan jobs) walks 1-8 corner lists, ~30 hash-probed blocks and a few
// hundred candidate points per query, one dependent memory access after the other: ~50 us of pure latency per query
// and, with 32 unrelated queries per warp, 4-5 active lanes on average.  Here the 32 lanes of a warp share one query:
// they stride over the candidates of a list, or take one block of a shell each, keep a private k-best set bounded
// by the (replicated) global k-th distance, and merge the private sets into the global one with warp reductions.
// The visiting order differs from the serial code, the result does not: it is the k smallest elements of the same
// candidate set under the same strict total order (dis2, original index).
#pragma once
#include "voxel_map.cuh"

namespace locreg {

constexpr unsigned int kFullMask = 0xffffffffu;

// Merges the lanes' private sets into the replicated global set `res` (identical in all lanes on entry and on exit).
// Private sets are sorted ascending and only hold candidates that were <= the global k-th distance when they were
// accepted.  Each round takes the smallest private head of the warp ((dis2, index) order), inserts it into the global
// set in every lane (unless it is already a member: a point met again on another level) and pops it from its owner.
template <int K>
__device__ __forceinline__ void warp_merge(const float4* __restrict__ canon, KnnResult<K>& res, KnnResult<K>& priv) {
    const unsigned int lane = threadIdx.x & 31;
    while (true) {
        const unsigned int bits = priv.pos[0] != kNoPos ? __float_as_uint(priv.d2[0]) : 0x7f800001u;  // > +inf: no head
        const unsigned int best = __reduce_min_sync(kFullMask, bits);
        if (best > __float_as_uint(res.d2[K - 1])) break;  // nothing left that can enter (also: no heads at all)
        const bool tied = bits == best;
        const unsigned int idx = tied ? static_cast<unsigned int>(knn_index_of(canon, priv.pos[0])) : 0x7fffffffu;
        const unsigned int best_idx = __reduce_min_sync(kFullMask, idx);
        const int src = __ffs(__ballot_sync(kFullMask, tied && idx == best_idx)) - 1;
        const unsigned int wpos = __shfl_sync(kFullMask, priv.pos[0], src);
        const float wd2 = __uint_as_float(best);
        bool member = false;
#pragma unroll
        for (int j = 0; j < K; ++j) member = member || (res.pos[j] == wpos);
        if (!member) {
            // heads arrive in ascending (dis2, index) order: the first one that cannot enter ends the merge
            if (!knn_accepts(canon, res, wd2, wpos)) break;
            knn_insert(canon, res, wd2, wpos);
        }
        if (lane == static_cast<unsigned int>(src)) {  // pop the head that was dealt with
#pragma unroll
            for (int j = 0; j + 1 < K; ++j) { priv.d2[j] = priv.d2[j + 1]; priv.pos[j] = priv.pos[j + 1]; }
            priv.d2[K - 1] = INFINITY;
            priv.pos[K - 1] = kNoPos;
        }
    }
}

// warp_merge that also reports d6 = the smallest dis2 of any candidate left OUTSIDE the global set: what the merge
// evicts from it, and the heads that stay behind in the private sets (sorted: a lane's head is its smallest).
template <int K>
__device__ __forceinline__ float warp_merge_track(const float4* __restrict__ canon, KnnResult<K>& res, KnnResult<K>& priv) {
    const unsigned int lane = threadIdx.x & 31;
    float evicted = INFINITY;
    while (true) {
        const unsigned int bits = priv.pos[0] != kNoPos ? __float_as_uint(priv.d2[0]) : 0x7f800001u;  // > +inf: no head
        const unsigned int best = __reduce_min_sync(kFullMask, bits);
        if (best > __float_as_uint(res.d2[K - 1])) break;
        const bool tied = bits == best;
        const unsigned int idx = tied ? static_cast<unsigned int>(knn_index_of(canon, priv.pos[0])) : 0x7fffffffu;
        const unsigned int best_idx = __reduce_min_sync(kFullMask, idx);
        const int src = __ffs(__ballot_sync(kFullMask, tied && idx == best_idx)) - 1;
        const unsigned int wpos = __shfl_sync(kFullMask, priv.pos[0], src);
        const float wd2 = __uint_as_float(best);
        bool member = false;
#pragma unroll
        for (int j = 0; j < K; ++j) member = member || (res.pos[j] == wpos);
        if (!member) {
            if (!knn_accepts(canon, res, wd2, wpos)) break;
            evicted = fminf(evicted, res.d2[K - 1]);  // INFINITY while the set is not full
            knn_insert(canon, res, wd2, wpos);
        }
        if (lane == static_cast<unsigned int>(src)) {
#pragma unroll
            for (int j = 0; j + 1 < K; ++j) { priv.d2[j] = priv.d2[j + 1]; priv.pos[j] = priv.pos[j + 1]; }
            priv.d2[K - 1] = INFINITY;
            priv.pos[K - 1] = kNoPos;
        }
    }
    // (a member met again that is still a head counts as left outside: d6 can only come out too small, never too large)
    const float left = priv.pos[0] != kNoPos ? priv.d2[0] : INFINITY;
    const unsigned int m = __reduce_min_sync(kFullMask, __float_as_uint(fminf(left, evicted)));
    return __uint_as_float(m);
}
// warp_scan_list with the bookkeeping of knn_scan_list's tracked form: returns d6 over the whole list.
template <int K>
__device__ __forceinline__ float warp_scan_list_track(const float4* __restrict__ pts, const float4* __restrict__ canon, KnnResult<K>& res,
                                                      float qx, float qy, float qz, unsigned int beg, unsigned int cnt) {
    const unsigned int lane = threadIdx.x & 31;
    KnnResult<K> priv;
    knn_init(priv);
    const float bound = res.d2[K - 1];
    float rej = INFINITY;
    for (unsigned int i = lane; i < cnt; i += 32) {
        const float4 p = pts[beg + i];
        const float d2 = dis2_f32(qx, qy, qz, p.x, p.y, p.z);
        if (d2 <= bound) knn_offer_track(canon, priv, d2, static_cast<unsigned int>(float_as_int(p.w)), rej);
        else rej = fminf(rej, d2);  // fminf drops the NaN of a masked duplicate
    }
    const float d6 = warp_merge_track<K>(canon, res, priv);
    return fminf(d6, __uint_as_float(__reduce_min_sync(kFullMask, __float_as_uint(rej))));
}

// One contiguous neighbourhood list (entries carry the canonical position in w), lanes striding over it.
template <int K>
__device__ __forceinline__ void warp_scan_list(const float4* __restrict__ pts, const float4* __restrict__ canon, KnnResult<K>& res,
                                               float qx, float qy, float qz, unsigned int beg, unsigned int cnt) {
    const unsigned int lane = threadIdx.x & 31;
    KnnResult<K> priv;
    knn_init(priv);
    const float bound = res.d2[K - 1];
    for (unsigned int i = lane; i < cnt; i += 32) {
        const float4 p = pts[beg + i];
        const float d2 = dis2_f32(qx, qy, qz, p.x, p.y, p.z);
        if (d2 <= bound) knn_offer(canon, priv, d2, static_cast<unsigned int>(float_as_int(p.w)));
    }
    warp_merge<K>(canon, res, priv);
}

// Stage 2a (see knn_query_corners): same corner selection, lists scanned cooperatively.
template <int K>
__device__ __forceinline__ bool warp_query_corners(const VoxelMapView& m, float qx, float qy, float qz, KnnResult<K>& res) {
    const KnnCellFrame c = knn_frame(m, qx, qy, qz);
    const unsigned int need = knn_corner_mask(m, c, res.d2[K - 1]);
    for (int o = 0; o < 8; ++o) {
        if (!((need >> o) & 1u)) continue;
        unsigned int beg = 0, cnt = 0;
        knn_find_list(m, c.fx + ((o & 1) ? 1 : -1), c.fy + ((o & 2) ? 1 : -1), c.fz + ((o & 4) ? 1 : -1), beg, cnt);
        warp_scan_list<K>(m.pts, m.canon, res, qx, qy, qz, beg, cnt);
    }
    return knn_corners_final<K>(m, c, res);
}

// Stage 2b (see knn_query_rings).  Fine level (a handful of points per cell): the blocks of a shell are dealt out to
// the lanes, each lane walks its block's cells and points on its own.  Coarse levels (tens to thousands of points
// per cell): the lanes still probe one block each, but the cells that survive are then visited one after the other
// by the whole warp, lanes striding over the cell's points - a lane alone would chew on one big cell while 31 idle.
template <int K>
__device__ __forceinline__ bool warp_query_rings(const VoxelMapView& m, float qx, float qy, float qz, KnnResult<K>& res,
                                                 int boxes_done, int max_R) {
    const unsigned int lane = threadIdx.x & 31;
    const KnnCellFrame c = knn_frame(m, qx, qy, qz);
    bool first = boxes_done == 0;
    for (int R = first ? c.R0 : boxes_done + 1;; ++R) {
        if (R > max_R) return false;
        const KnnShell sh = knn_shell(m, c, R, first);
        KnnResult<K> priv;
        knn_init(priv);
        const float bound = res.d2[K - 1];
        const int nb = sh.nbx * sh.nby * sh.nbz;
        if (!m.w_is_pos) {
            for (int j = static_cast<int>(lane); j < nb; j += 32) {
                const int jx = j % sh.nbx, jy = (j / sh.nbx) % sh.nby, jz = j / (sh.nbx * sh.nby);
                knn_shell_block<K>(m, c, sh, (sh.clx >> 2) + jx, (sh.cly >> 2) + jy, (sh.clz >> 2) + jz, qx, qy, qz, bound, priv);
            }
        } else {
            const float magR = c.mag + static_cast<float>(R);
            for (int j0 = 0; j0 < nb; j0 += 32) {
                const int j = j0 + static_cast<int>(lane);
                unsigned long long todo = 0ull, occ = 0ull;
                unsigned int base = 0;
                int ox = 0, oy = 0, oz = 0;
                if (j < nb) {
                    const int jx = j % sh.nbx, jy = (j / sh.nbx) % sh.nby, jz = j / (sh.nbx * sh.nby);
                    const int bx = (sh.clx >> 2) + jx, by = (sh.cly >> 2) + jy, bz = (sh.clz >> 2) + jz;
                    ox = bx << 2; oy = by << 2; oz = bz << 2;
                    todo = knn_shell_block_cells(m, sh, bx, by, bz, occ, base);
                }
                unsigned int active = __ballot_sync(kFullMask, todo != 0ull);
                while (active) {
                    const int src = __ffs(active) - 1;
                    active &= active - 1;
                    unsigned long long t = __shfl_sync(kFullMask, todo, src);
                    const unsigned long long o = __shfl_sync(kFullMask, occ, src);
                    const unsigned int b = __shfl_sync(kFullMask, base, src);
                    const int sox = __shfl_sync(kFullMask, ox, src), soy = __shfl_sync(kFullMask, oy, src), soz = __shfl_sync(kFullMask, oz, src);
                    while (t) {
                        const int bit = __ffsll(static_cast<long long>(t)) - 1;
                        t &= t - 1;
                        if (bound < INFINITY &&
                            knn_cell_min_d2(m, c, sox + (bit & 3), soy + ((bit >> 2) & 3), soz + (bit >> 4), magR) > bound)
                            continue;
                        const unsigned int cid = b + __popcll(o & ((1ull << bit) - 1ull));
                        const unsigned int beg = m.cell_start[cid], end = m.cell_start[cid + 1];
                        for (unsigned int i = beg + lane; i < end; i += 32) {
                            const float4 p = m.pts[i];
                            const float d2 = dis2_f32(qx, qy, qz, p.x, p.y, p.z);
                            if (d2 <= bound) knn_offer(m.canon, priv, d2, static_cast<unsigned int>(float_as_int(p.w)));
                        }
                    }
                }
            }
        }
        warp_merge<K>(m.canon, res, priv);
        first = false;
        if (knn_shell_final<K>(m, c, sh, res)) return true;
    }
}

// Escalation limits of the warp-cooperative search (the thread-per-query form has its own, voxel_map.cuh): a warp walks
// the cells of a mid-level shell one after the other, so it leaves that level sooner.
#ifndef LR_WARP_MID_SHELLS
#define LR_WARP_MID_SHELLS 1  // measured 0..3 (stage 2 of a single scan, first iteration: 57 / 55 / 65 / 71 us; the thread form's 4: 85 us)
#endif
#ifndef LR_WARP_COARSE_SHELLS
#define LR_WARP_COARSE_SHELLS 16  // measured 8..24: flat from 12 on
#endif
constexpr int kWarpMidShells = LR_WARP_MID_SHELLS, kWarpCoarseShells = LR_WARP_COARSE_SHELLS;
// Everything after stage 1 for one query, by one warp (see knn_query_finish for the escalation logic).
// margin (optional): receives the KnnTrack margin when the search ends with the mid level's list (the stage-2 form of
// knn_query_fast_track: the list holds every point of the mid box, so the scan knows d6), else -1.
template <int K>
__device__ __forceinline__ void warp_query_finish(const VoxelMapView& m, const CoarseLevels& coarse, float qx, float qy, float qz,
                                                  KnnResult<K>& res, float* margin = nullptr) {
    const unsigned int lane = threadIdx.x & 31;
    const bool have_mid = coarse.mid.n_pts != 0;
    if (margin) *margin = -1.0f;
    if (have_mid) {  // knn_query_mid, lists and shells scanned cooperatively
        const VoxelMapView& md = coarse.mid;
        const bool have_coarse = coarse.lv[0].n_pts != 0;
        const KnnCellFrame c = knn_frame(md, qx, qy, qz);
        int boxes_done = 0;
        if (knn_uses_list(md, c)) {
            unsigned int beg = 0, cnt = 0;
            knn_find_list(md, c.fx, c.fy, c.fz, beg, cnt);
            if (margin && res.pos[K - 1] != kNoPos) {
                const float d6 = warp_scan_list_track<K>(md.pts, md.canon, res, qx, qy, qz, beg, cnt);
                if (knn_list_final<K>(md, c, res)) {
                    *margin = knn_track_margin<K>(md, c, res, d6, qx, qy, qz);
                    return;
                }
            } else {
                warp_scan_list<K>(md.pts, md.canon, res, qx, qy, qz, beg, cnt);
                if (knn_list_final<K>(md, c, res)) return;
            }
            boxes_done = 1;
        }
        if (!(have_coarse && c.R0 > kWarpMidShells) &&
            warp_query_rings<K>(md, qx, qy, qz, res, boxes_done, have_coarse ? kWarpMidShells : kBruteForceShell))
            return;
    }
    if (!have_mid) {
        const KnnCellFrame c = knn_frame(m, qx, qy, qz);
        int boxes_done = knn_uses_list(m, c) ? 1 : 0;
        if (boxes_done == 1 && res.pos[K - 1] != kNoPos) {
            if (warp_query_corners<K>(m, qx, qy, qz, res)) return;
            boxes_done = 2;
        }
        const bool have_coarse = coarse.lv[0].n_pts != 0;
        if (!(have_coarse && c.R0 > kFineShells) &&
            warp_query_rings<K>(m, qx, qy, qz, res, boxes_done, have_coarse ? kFineShells : kBruteForceShell))
            return;
    }
    for (int l = 0; l < kCoarseLevels; ++l) {
        const VoxelMapView& cl = coarse.lv[l];
        if (cl.n_pts == 0) break;
        const bool last = l + 1 == kCoarseLevels || coarse.lv[l + 1].n_pts == 0;
        if (!last && knn_frame(cl, qx, qy, qz).R0 > kWarpCoarseShells) continue;
        if (warp_query_rings<K>(cl, qx, qy, qz, res, 0, last ? kBruteForceShell : kWarpCoarseShells)) return;
    }
    // linear scan, lanes striding over the whole map
    knn_init(res);
    KnnResult<K> priv;
    knn_init(priv);
    for (unsigned int i = lane; i < m.n_pts; i += 32) {
        const float4 p = m.pts[i];
        knn_offer(m.pts, priv, dis2_f32(qx, qy, qz, p.x, p.y, p.z), i);
    }
    warp_merge<K>(m.pts, res, priv);
}

}  // namespace locreg

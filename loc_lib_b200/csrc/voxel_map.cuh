// GPU voxel-hash map for exact k-NN (kernels K1/K2 of SURVEY.md §2.3).
//
// Replaces the reference's pointer kd-tree (KdTree::BuildTree / GetClosestPoint, kdtree.cpp:10-31,
// 147-236) with a structure that suits a B200: all map points live in ONE float4 array sorted by
// voxel (w carries the caller's original index), grouped into 4x4x4-cell blocks.  A block is one
// 32-byte open-addressing hash slot {key, 64-bit cell-occupancy mask, first cell id}, so a query
// touches one 32 B sector per block, skips empty cells with bit tricks instead of probes, and reads
// candidate points as contiguous 16 B vectors.  On top of that every cell whose 3x3x3 neighbourhood holds a
// point owns a NEIGHBOURHOOD LIST: all points of those 27 cells copied contiguously (each map point is stored
// 27 times).  The common query is then ONE 16 B probe of a cell-keyed table plus ONE contiguous float4 scan
// (~25 candidates) with no per-cell work and little divergence; the block structure remains the exact fallback
// for queries whose k-th neighbour lies beyond the guaranteed radius (far-off points of the first iteration).
// Layout in HBM (1 M-point map, cell 0.5 m; sized for 180 GB, not for the L2):
//   slots       32 B x capacity (power of two, load <= 0.5)      ~ 2 MB     block table (fallback search)
//   cell_start  4 B x (occupied cells + 1)                        ~ 1.6 MB
//   pts         16 B x N                                           16 MB     sorted by cell; positions [0, N)
//   nbr_slots   16 B x capacity                                   ~ 64 MB    cell -> (start, count) of its list
//   lists       16 B x 27 N                                        432 MB    positions [N, 28 N) of the same array
//
// NN contract: the k nearest points in float32
//   dis2 = dx*dx + (dy*dy + dz*dz)   (Eigen 3.3 association, no FMA; kdtree.h:94)
// under the total order (dis2, original index) ascending, over the de-duplicated point set
// (quirk Q3: of exactly coincident points only the lowest index survives, kdtree.cpp:76-81).
// The search visits cells in Chebyshev shells around the query's cell and stops once the k-th best
// dis2 is strictly below a conservative lower bound of dis2 for every unvisited point, so the
// result is exact, not approximate (see DESIGN.md "Termination proof").
#pragma once
#include "common.cuh"

namespace locreg {

struct __attribute__((aligned(32))) VoxelSlot {
    unsigned long long key;   // packed block coordinate, kEmptyKey if unused
    unsigned long long mask;  // bit (z&3)<<4 | (y&3)<<2 | (x&3) set iff that cell holds points
    unsigned int cell_base;   // id of the block's first occupied cell in cell_start[]
    unsigned int pad0, pad1, pad2;
};
static_assert(sizeof(VoxelSlot) == 32, "slot must be one 32 B sector");

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kCoordBias = 1 << 20;        // block coords are biased into 21 bits
constexpr float kCellClamp = 1048000.0f;   // |fine cell index| clamp (fits 21 bits, exact in float)
constexpr int kBruteForceShell = 24;       // beyond this Chebyshev radius fall back to a linear scan

struct __attribute__((aligned(16))) NbrSlot {
    unsigned long long key;  // packed fine-cell coordinate, kEmptyKey if unused
    unsigned int start;      // first entry of the cell's neighbourhood list, as a position in pts[]
    unsigned int count;
};
static_assert(sizeof(NbrSlot) == 16, "neighbourhood slot must be 16 B");

struct VoxelMapView {
    const VoxelSlot* slots;
    const unsigned int* cell_start;
    const float4* pts;        // [0, n_pts): points sorted by cell; [n_pts, ...): neighbourhood lists
    const NbrSlot* nbr_slots; // nullptr: no neighbourhood lists (fallback search only)
    unsigned int nbr_mask;
    unsigned int slot_mask;  // capacity - 1
    unsigned int n_pts;      // points stored (non-finite inputs are dropped)
    unsigned int n_unique;   // points that survive de-duplication (= KdTree::size())
    float inv_cell;
    float cell;
    int cmin[3], cmax[3];    // inclusive bounds of occupied fine-cell coordinates
};

LR_HD unsigned long long pack_block(int bx, int by, int bz) {
    return (static_cast<unsigned long long>(static_cast<unsigned int>(bx + kCoordBias)) << 42) |
           (static_cast<unsigned long long>(static_cast<unsigned int>(by + kCoordBias)) << 21) |
           static_cast<unsigned long long>(static_cast<unsigned int>(bz + kCoordBias));
}
LR_HD unsigned long long pack_cell(int cx, int cy, int cz) {  // |c| <= kCellClamp < 2^20
    return (static_cast<unsigned long long>(static_cast<unsigned int>(cx + kCoordBias)) << 42) |
           (static_cast<unsigned long long>(static_cast<unsigned int>(cy + kCoordBias)) << 21) |
           static_cast<unsigned long long>(static_cast<unsigned int>(cz + kCoordBias));
}
LR_HD unsigned int hash_block(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return static_cast<unsigned int>(k);
}
// cell coordinate of a float coordinate; monotone non-decreasing in x (needed by the bound proof)
LR_HD float cell_coord_f(float x, float inv_cell) {
    float u = LR_FMUL(x, inv_cell);
    u = fminf(fmaxf(u, -kCellClamp), kCellClamp);
    return u;
}
LR_HD int cell_of(float u) { return static_cast<int>(floorf(u)); }

LR_HD const VoxelSlot* find_block(const VoxelMapView& m, int bx, int by, int bz) {
    const unsigned long long key = pack_block(bx, by, bz);
    unsigned int h = hash_block(key) & m.slot_mask;
    while (true) {
        const VoxelSlot* s = m.slots + h;
        const unsigned long long k = s->key;
        if (k == key) return s;
        if (k == kEmptyKey) return nullptr;
        h = (h + 1) & m.slot_mask;
    }
}

// 4-bit per-axis masks -> 64-bit cell masks (bit = z<<4 | y<<2 | x)
LR_HD unsigned long long spread_x(unsigned int mx) { return 0x1111111111111111ull * mx; }
LR_HD unsigned long long spread_y(unsigned int my) {
    const unsigned int t = ((my & 1u) * 0xFu) | ((my & 2u) * (0xF0u >> 1)) | ((my & 4u) * (0xF00u >> 2)) |
                           ((my & 8u) * (0xF000u >> 3));
    return 0x0001000100010001ull * t;
}
LR_HD unsigned long long spread_z(unsigned int mz) {
    return ((mz & 1u) ? 0x000000000000FFFFull : 0ull) | ((mz & 2u) ? 0x00000000FFFF0000ull : 0ull) |
           ((mz & 4u) ? 0x0000FFFF00000000ull : 0ull) | ((mz & 8u) ? 0xFFFF000000000000ull : 0ull);
}
// bits of a block (origin o = 4*b) whose cell coordinate lies in [lo, hi]
LR_HD unsigned int axis_range_mask(int o, int lo, int hi) {
    int a = lo - o, b = hi - o;
    a = a < 0 ? 0 : a;
    b = b > 3 ? 3 : b;
    if (a > b) return 0u;
    return ((1u << (b + 1)) - 1u) & ~((1u << a) - 1u);
}
// bits of a block whose cell coordinate equals lo or hi
LR_HD unsigned int axis_edge_mask(int o, int lo, int hi) {
    unsigned int m = 0;
    if (lo - o >= 0 && lo - o <= 3) m |= 1u << (lo - o);
    if (hi - o >= 0 && hi - o <= 3) m |= 1u << (hi - o);
    return m;
}

template <int K>
struct KnnResult {
    float d2[K];
    int idx[K];           // caller's original index, 0x7fffffff = empty
    unsigned int pos[K];  // position in the sorted pts[] (to re-read coordinates)
};

template <int K>
LR_HD void knn_init(KnnResult<K>& r) {
#pragma unroll
    for (int j = 0; j < K; ++j) { r.d2[j] = INFINITY; r.idx[j] = 0x7fffffff; r.pos[j] = 0u; }
}
template <int K>
LR_HD void knn_offer(KnnResult<K>& r, float d2, int idx, unsigned int pos) {
    if (d2 < r.d2[K - 1] || (d2 == r.d2[K - 1] && idx < r.idx[K - 1])) {
        r.d2[K - 1] = d2;
        r.idx[K - 1] = idx;
        r.pos[K - 1] = pos;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            const bool sw = r.d2[j] < r.d2[j - 1] || (r.d2[j] == r.d2[j - 1] && r.idx[j] < r.idx[j - 1]);
            if (sw) {
                const float td = r.d2[j]; r.d2[j] = r.d2[j - 1]; r.d2[j - 1] = td;
                const int ti = r.idx[j]; r.idx[j] = r.idx[j - 1]; r.idx[j - 1] = ti;
                const unsigned int tp = r.pos[j]; r.pos[j] = r.pos[j - 1]; r.pos[j - 1] = tp;
            }
        }
    }
}
template <int K>
LR_HD int knn_count(const KnnResult<K>& r) {
    int c = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) c += (r.idx[j] != 0x7fffffff) ? 1 : 0;
    return c;
}

LR_HD float dis2_f32(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = LR_FSUB(qx, px), dy = LR_FSUB(qy, py), dz = LR_FSUB(qz, pz);
    return LR_FADD(LR_FMUL(dx, dx), LR_FADD(LR_FMUL(dy, dy), LR_FMUL(dz, dz)));
}

// Conservative (never over-estimating) distance in metres for a gap of `g` cell units when the
// coordinates involved are about `mag` cell units from the origin: absorbs the rounding of
// x*inv_cell on both the map point and the query, and of the float dis2 itself.
LR_HD float safe_gap(float g, float mag, float cell) {
    const float e = g - 4.8e-7f * (mag + 4.0f);
    return e > 0.0f ? e * cell * 0.99999f : 0.0f;
}

// Exact k-NN of (qx,qy,qz).  The query must be finite.
template <int K>
LR_HD void knn_query(const VoxelMapView& m, float qx, float qy, float qz, KnnResult<K>& res) {
    knn_init(res);
    if (m.n_pts == 0) return;
    const float ux = cell_coord_f(qx, m.inv_cell), uy = cell_coord_f(qy, m.inv_cell), uz = cell_coord_f(qz, m.inv_cell);
    const int fx = cell_of(ux), fy = cell_of(uy), fz = cell_of(uz);
    const float frx = ux - static_cast<float>(fx), fry = uy - static_cast<float>(fy), frz = uz - static_cast<float>(fz);
    const float mag = fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz));
    // smallest Chebyshev radius whose box touches the occupied bounds
    int R = 1;
    {
        const int ex = fx < m.cmin[0] ? m.cmin[0] - fx : (fx > m.cmax[0] ? fx - m.cmax[0] : 0);
        const int ey = fy < m.cmin[1] ? m.cmin[1] - fy : (fy > m.cmax[1] ? fy - m.cmax[1] : 0);
        const int ez = fz < m.cmin[2] ? m.cmin[2] - fz : (fz > m.cmax[2] ? fz - m.cmax[2] : 0);
        const int e = ex > ey ? (ex > ez ? ex : ez) : (ey > ez ? ey : ez);
        if (e > R) R = e;
    }
    bool first = true;
    if (R == 1 && m.nbr_slots != nullptr) {
        // fast path: the whole box [f-1, f+1]^3 is one contiguous list
        const unsigned long long key = pack_cell(fx, fy, fz);
        unsigned int h = hash_block(key) & m.nbr_mask;
        unsigned int beg = 0, cnt = 0;
        while (true) {
            const NbrSlot s = m.nbr_slots[h];
            if (s.key == key) { beg = s.start; cnt = s.count; break; }
            if (s.key == kEmptyKey) break;
            h = (h + 1) & m.nbr_mask;
        }
        for (unsigned int i = beg; i < beg + cnt; ++i) {
            const float4 p = m.pts[i];
            knn_offer(res, dis2_f32(qx, qy, qz, p.x, p.y, p.z), float_as_int(p.w), i);
        }
        first = false;
        if (res.idx[K - 1] != 0x7fffffff) {
            float mf = fminf(frx, 1.0f - frx);
            mf = fminf(mf, fminf(fry, 1.0f - fry));
            mf = fminf(mf, fminf(frz, 1.0f - frz));
            const float g = safe_gap(1.0f + mf, mag + 1.0f, m.cell);
            if (res.d2[K - 1] < g * g * 0.99999f) return;
        }
        if (fx - 1 <= m.cmin[0] && fx + 1 >= m.cmax[0] && fy - 1 <= m.cmin[1] && fy + 1 >= m.cmax[1] &&
            fz - 1 <= m.cmin[2] && fz + 1 >= m.cmax[2])
            return;
        R = 2;
    }
    while (true) {
        if (R > kBruteForceShell) {
            // pathological query (> kBruteForceShell cells from every candidate seen so far): linear scan
            knn_init(res);
            for (unsigned int i = 0; i < m.n_pts; ++i) {
                const float4 p = m.pts[i];
                knn_offer(res, dis2_f32(qx, qy, qz, p.x, p.y, p.z), float_as_int(p.w), i);
            }
            return;
        }
        const int lox = fx - R, hix = fx + R, loy = fy - R, hiy = fy + R, loz = fz - R, hiz = fz + R;
        // clip enumeration to the occupied bounds
        const int clx = lox > m.cmin[0] ? lox : m.cmin[0], chx = hix < m.cmax[0] ? hix : m.cmax[0];
        const int cly = loy > m.cmin[1] ? loy : m.cmin[1], chy = hiy < m.cmax[1] ? hiy : m.cmax[1];
        const int clz = loz > m.cmin[2] ? loz : m.cmin[2], chz = hiz < m.cmax[2] ? hiz : m.cmax[2];
        for (int bz = clz >> 2; bz <= (chz >> 2); ++bz) {
            const int oz = bz << 2;
            const unsigned int rz = axis_range_mask(oz, clz, chz);
            const unsigned int iz = first ? 0u : (rz & ~axis_edge_mask(oz, loz, hiz));
            for (int by = cly >> 2; by <= (chy >> 2); ++by) {
                const int oy = by << 2;
                const unsigned int ry = axis_range_mask(oy, cly, chy);
                const unsigned int iy = first ? 0u : (ry & ~axis_edge_mask(oy, loy, hiy));
                for (int bx = clx >> 2; bx <= (chx >> 2); ++bx) {
                    const int ox = bx << 2;
                    const unsigned int rx = axis_range_mask(ox, clx, chx);
                    const unsigned int ix = first ? 0u : (rx & ~axis_edge_mask(ox, lox, hix));
                    // cells of this block inside the box but not strictly interior (= the new shell)
                    if (!first && ix == rx && iy == ry && iz == rz) continue;  // block entirely interior
                    const VoxelSlot* s = find_block(m, bx, by, bz);
                    if (s == nullptr) continue;
                    unsigned long long want = spread_x(rx) & spread_y(ry) & spread_z(rz);
                    if (!first) want &= ~(spread_x(ix) & spread_y(iy) & spread_z(iz));
                    const unsigned long long occ = s->mask;
                    unsigned long long todo = occ & want;
                    const unsigned int base = s->cell_base;
                    while (todo) {
                        const int bit = ffs64(todo) - 1;
                        todo &= todo - 1;
                        const int cx = ox + (bit & 3), cy = oy + ((bit >> 2) & 3), cz = oz + (bit >> 4);
                        if (res.idx[K - 1] != 0x7fffffff) {
                            // prune: conservative min distance from the query to this cell
                            const float gx = cx > fx ? static_cast<float>(cx - fx) - frx
                                                     : (cx < fx ? static_cast<float>(fx - cx - 1) + frx : 0.0f);
                            const float gy = cy > fy ? static_cast<float>(cy - fy) - fry
                                                     : (cy < fy ? static_cast<float>(fy - cy - 1) + fry : 0.0f);
                            const float gz = cz > fz ? static_cast<float>(cz - fz) - frz
                                                     : (cz < fz ? static_cast<float>(fz - cz - 1) + frz : 0.0f);
                            const float sx = safe_gap(gx, mag + R, m.cell), sy = safe_gap(gy, mag + R, m.cell),
                                        sz = safe_gap(gz, mag + R, m.cell);
                            const float md2 = (sx * sx + sy * sy + sz * sz) * 0.99999f;
                            if (md2 > res.d2[K - 1]) continue;
                        }
                        const unsigned int cid = base + popc64(occ & ((1ull << bit) - 1ull));
                        const unsigned int beg = m.cell_start[cid], end = m.cell_start[cid + 1];
                        for (unsigned int i = beg; i < end; ++i) {
                            const float4 p = m.pts[i];
                            knn_offer(res, dis2_f32(qx, qy, qz, p.x, p.y, p.z), float_as_int(p.w), i);
                        }
                    }
                }
            }
        }
        first = false;
        // every unvisited point lies outside the box [f-R, f+R]: at least R + min(frac, 1-frac) cell
        // units away along some axis
        if (res.idx[K - 1] != 0x7fffffff) {
            float mf = fminf(frx, 1.0f - frx);
            mf = fminf(mf, fminf(fry, 1.0f - fry));
            mf = fminf(mf, fminf(frz, 1.0f - frz));
            const float g = safe_gap(static_cast<float>(R) + mf, mag + R, m.cell);
            if (res.d2[K - 1] < g * g * 0.99999f) return;
        }
        if (lox <= m.cmin[0] && hix >= m.cmax[0] && loy <= m.cmin[1] && hiy >= m.cmax[1] && loz <= m.cmin[2] &&
            hiz >= m.cmax[2])
            return;  // the whole map has been visited
        ++R;
    }
}

}  // namespace locreg

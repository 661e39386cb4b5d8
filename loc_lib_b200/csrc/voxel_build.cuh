// Per-element bodies of the voxel-hash map build (kernel K1 of SURVEY.md §2.3), replacing
// KdTree::BuildTree / Insert / FindSplitAxisAndThresh (kdtree.cpp:10-31,58-123).
//
// Pipeline (each step one kernel over the map points or the hash slots; see voxel_map.cu):
//   1 insert   point -> fine cell -> 4x4x4 block key -> open-addressing slot (atomicCAS), OR the cell bit
//   2 rank     exclusive scan of popcount(mask) over slots -> slot.cell_base   (cells numbered block by block)
//   3 count    per point: cell id = cell_base + rank(bit); atomicAdd cell histogram
//   4 scan     exclusive scan of the histogram -> cell_start[]
//   5 scatter  counting-sort points into pts[] as float4 (x, y, z, original index)
//   6 dedupe   quirk Q3: a point that coincides exactly with a lower-index point of its cell is masked
//              (x := NaN), which makes every dis2 against it NaN and therefore never selected
//   7 lists    neighbourhood lists: every surviving point is inserted into the lists of the 27 cells around
//              its own (count pass, scan, scatter pass) - see voxel_map.cuh
// The order of points inside a cell depends on atomic arrival order, but no result does: the k-NN
// total order (dis2, original index) is independent of storage order.
#pragma once
#include "voxel_map.cuh"

namespace locreg {

constexpr unsigned int kDropped = 0xFFFFFFFFu;

struct HostAtomics {  // serial stand-ins for tests/hostsim
    static unsigned long long cas64(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
        const unsigned long long old = *p;
        if (old == cmp) *p = v;
        return old;
    }
    static void or64(unsigned long long* p, unsigned long long v) { *p |= v; }
    static unsigned int add32(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p += v; return o; }
};
#if defined(__CUDACC__)
struct DeviceAtomics {
    static __device__ unsigned long long cas64(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
        return atomicCAS(p, cmp, v);
    }
    static __device__ void or64(unsigned long long* p, unsigned long long v) { atomicOr(p, v); }
    static __device__ unsigned int add32(unsigned int* p, unsigned int v) { return atomicAdd(p, v); }
};
#endif

LR_HD const float* point_ptr(const void* base, size_t i, size_t stride) {
    return reinterpret_cast<const float*>(static_cast<const char*>(base) + i * stride);
}

// Step 1.  Returns true if the point was kept; f[] receives its fine cell coordinate.
// counters: [0] blocks inserted, [1] overflow flag, [2] points kept
template <class A>
LR_HD bool build_insert_body(size_t i, const void* xyz, size_t stride, float inv_cell, VoxelSlot* slots,
                             unsigned int slot_mask, unsigned int* pt_slot, unsigned char* pt_bit,
                             unsigned int* counters, int* f) {
    const float* p = point_ptr(xyz, i, stride);
    const float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) { pt_slot[i] = kDropped; return false; }
    f[0] = cell_of(cell_coord_f(x, inv_cell));
    f[1] = cell_of(cell_coord_f(y, inv_cell));
    f[2] = cell_of(cell_coord_f(z, inv_cell));
    const unsigned long long key = pack_block(f[0] >> 2, f[1] >> 2, f[2] >> 2);
    const int bit = ((f[2] & 3) << 4) | ((f[1] & 3) << 2) | (f[0] & 3);
    unsigned int h = hash_block(key) & slot_mask;
    unsigned int probes = 0;
    while (true) {
        const unsigned long long k = A::cas64(&slots[h].key, kEmptyKey, key);
        if (k == kEmptyKey) { A::add32(&counters[0], 1u); break; }
        if (k == key) break;
        h = (h + 1) & slot_mask;
        if (++probes > slot_mask) { counters[1] = 1u; pt_slot[i] = kDropped; return false; }
    }
    A::or64(&slots[h].mask, 1ull << bit);
    pt_slot[i] = h;
    pt_bit[i] = static_cast<unsigned char>(bit);
    A::add32(&counters[2], 1u);
    return true;
}

// Block pyramid (voxel_map.cuh): the node with packed coordinate `child_key` registers itself with its parent on the
// next level (coordinate >> 2 per axis).  counters: [0] nodes created on the parent level, [1] overflow flag
template <class A>
LR_HD void build_pyr_body(unsigned long long child_key, PyrSlot* slots, unsigned int slot_mask, unsigned int* counters) {
    const int x = static_cast<int>((child_key >> 42) & 0x1FFFFFull) - kCoordBias, y = static_cast<int>((child_key >> 21) & 0x1FFFFFull) - kCoordBias,
              z = static_cast<int>(child_key & 0x1FFFFFull) - kCoordBias;
    const unsigned long long key = pack_block(x >> 2, y >> 2, z >> 2);
    const int bit = ((z & 3) << 4) | ((y & 3) << 2) | (x & 3);
    unsigned int h = hash_block(key) & slot_mask;
    unsigned int probes = 0;
    while (true) {
        const unsigned long long k = A::cas64(&slots[h].key, kEmptyKey, key);
        if (k == kEmptyKey) { A::add32(&counters[0], 1u); break; }
        if (k == key) break;
        h = (h + 1) & slot_mask;
        if (++probes > slot_mask) { counters[1] = 1u; return; }
    }
    A::or64(&slots[h].mask, 1ull << bit);
}
// number of pyramid levels for occupied cell bounds [cmin, cmax]: the smallest P >= 1 whose top nodes (2^(2P+2) cells
// per axis) span the bounds with at most two nodes per axis
inline int pyr_levels_for(const int* cmin, const int* cmax) {
    for (int P = 1; P <= kPyrMaxLevels; ++P) {
        const int s = 2 * P + 2;
        bool ok = true;
        for (int a = 0; a < 3; ++a) ok = ok && ((cmax[a] >> s) - (cmin[a] >> s)) <= 1;
        if (ok) return P;
    }
    return kPyrMaxLevels;  // (cannot happen within the 21-bit coordinate clamp; the search copes with more top nodes)
}

// Step 3.  pt_slot[i] is rewritten in place with the point's cell id.
template <class A>
LR_HD void build_count_body(size_t i, const VoxelSlot* slots, unsigned int* pt_slot, const unsigned char* pt_bit,
                            unsigned int* cell_count) {
    const unsigned int s = pt_slot[i];
    if (s == kDropped) return;
    const int bit = pt_bit[i];
    const unsigned int cid = slots[s].cell_base + popc64(slots[s].mask & ((1ull << bit) - 1ull));
    pt_slot[i] = cid;
    A::add32(&cell_count[cid], 1u);
}

// Step 5.  cursor[] starts as a copy of cell_start[0..ncells).
template <class A>
LR_HD void build_scatter_body(size_t i, const void* xyz, size_t stride, const unsigned int* pt_cell,
                              unsigned int* cursor, float4* pts, unsigned int* pt_pos) {
    const unsigned int cid = pt_cell[i];
    if (cid == kDropped) return;
    const float* p = point_ptr(xyz, i, stride);
    const unsigned int pos = A::add32(&cursor[cid], 1u);
    pts[pos] = make_float4(p[0], p[1], p[2], int_as_float(static_cast<int>(i)));
    pt_pos[i] = pos;
}

// Step 6a: is input point i an exact duplicate of a lower-index point of its cell?
LR_HD bool build_is_duplicate(size_t i, const unsigned int* pt_cell, const unsigned int* pt_pos,
                              const unsigned int* cell_start, const float4* pts) {
    const unsigned int cid = pt_cell[i];
    if (cid == kDropped) return false;
    const float4 me = pts[pt_pos[i]];
    const unsigned int beg = cell_start[cid], end = cell_start[cid + 1];
    for (unsigned int j = beg; j < end; ++j) {
        const float4 o = pts[j];
        if (float_as_int(o.w) < static_cast<int>(i) && o.x == me.x && o.y == me.y && o.z == me.z) return true;
    }
    return false;
}

// Step 7a/7c.  The 27 cells whose neighbourhood contains point i each get the point (a copy of its coordinates
// with w = the point's canonical position pt_pos[i] in the sorted array, which is what the search reports).  pass 0: create the cell's
// slot if needed and count; pass 1: append the point to the cell's list (cursor[] starts at 0 per slot).
// counters: [0] lists created, [1] overflow flag
template <class A>
LR_HD void build_nbr_body(size_t i, int pass, const void* xyz, size_t stride, float inv_cell, const unsigned int* pt_cell,
                          const unsigned int* pt_pos, const unsigned char* dup, NbrSlot* nbr, unsigned int nbr_mask,
                          unsigned int* cursor, float4* pts, unsigned int* counters) {
    if (pt_cell[i] == kDropped || dup[i]) return;
    const float* p = point_ptr(xyz, i, stride);
    const int fx = cell_of(cell_coord_f(p[0], inv_cell)), fy = cell_of(cell_coord_f(p[1], inv_cell)),
              fz = cell_of(cell_coord_f(p[2], inv_cell));
    for (int o = 0; o < 27; ++o) {
        const unsigned long long key = pack_cell(fx + (o % 3) - 1, fy + ((o / 3) % 3) - 1, fz + (o / 9) - 1);
        unsigned int h = hash_block(key) & nbr_mask;
        unsigned int probes = 0;
        bool ok = true;
        while (true) {
            if (pass == 0) {
                const unsigned long long k = A::cas64(&nbr[h].key, kEmptyKey, key);
                if (k == kEmptyKey) { A::add32(&counters[0], 1u); break; }
                if (k == key) break;
            } else if (nbr[h].key == key) {
                break;
            }
            h = (h + 1) & nbr_mask;
            if (++probes > nbr_mask) { counters[1] = 1u; ok = false; break; }
        }
        if (!ok) return;
        if (pass == 0) {
            A::add32(&nbr[h].count, 1u);
        } else {
            const unsigned int pos = nbr[h].start + A::add32(&cursor[h], 1u);
            pts[pos] = make_float4(p[0], p[1], p[2], int_as_float(static_cast<int>(pt_pos[i])));  // w = canonical position
        }
    }
}

}  // namespace locreg

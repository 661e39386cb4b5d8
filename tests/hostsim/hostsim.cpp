// TEST-ONLY serial harness ("hostsim").  NOT product code, never linked into liblocreg.so.
//
// The development container has no GPU, so the per-element bodies of the CUDA kernels
// (loc_lib_b200/csrc/*.cuh, written as __host__ __device__ inline functions) are compiled here with
// g++ and driven by plain loops.  The CPU test-suite uses it to check the kernel *logic* (voxel-hash
// build, exact k-NN termination, plane fit, gates, Gauss-Newton update) against the oracle before
// GPU time is spent; the `-m gpu` tests then check the real kernels through the C ABI.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

unsigned long long g_knn_stats[16] = {0};  // LR_STATS counters (voxel_map.cuh)
#include "../../loc_lib_b200/csrc/icp_point.cuh"
#include "../../loc_lib_b200/csrc/ndt_point.cuh"
#include "../../loc_lib_b200/csrc/voxel_build.cuh"

using namespace locreg;

struct HsMap {
    std::vector<NbrSlot> nbr;
    std::vector<VoxelSlot> slots;
    std::vector<unsigned int> cell_start;
    std::vector<float4> pts;
    std::vector<unsigned int> pos_of_index;  // input index -> canonical position
    VoxelMapView view;
    HsMap* coarse_map[kCoarseLevels] = {};   // coarser levels (no lists), or nullptr
    CoarseLevels coarse{};                   // lv[l].n_pts == 0: level absent
    HsMap* mid_map = nullptr;                // mid level (lists), or nullptr
    std::vector<PyrSlot> pyr;                // block pyramid of the fine level, levels back to back
    ~HsMap() { for (HsMap* c : coarse_map) delete c; delete mid_map; }
};

static unsigned int next_pow2(unsigned int v) {
    unsigned int p = 1;
    while (p < v) p <<= 1;
    return p;
}

extern "C" {

static HsMap* build_level(const float* xyz, size_t n, size_t stride, float cell, unsigned int capacity_hint) {
    auto* m = new HsMap;
    const float inv_cell = 1.0f / cell;
    unsigned int cap = (capacity_hint && capacity_hint != 0xFFFFFFFFu) ? next_pow2(capacity_hint) : next_pow2(static_cast<unsigned int>(n / 4 + 1024));
    std::vector<unsigned int> pt_slot(n), pt_pos(n, 0);
    std::vector<unsigned char> pt_bit(n);
    unsigned int counters[3];
    int cmin[3], cmax[3];
    while (true) {
        m->slots.assign(cap, VoxelSlot{kEmptyKey, 0ull, 0u, 0u, 0u, 0u});
        counters[0] = counters[1] = counters[2] = 0;
        for (int a = 0; a < 3; ++a) { cmin[a] = 0x7fffffff; cmax[a] = -0x7fffffff; }
        for (size_t i = 0; i < n; ++i) {
            int f[3];
            if (build_insert_body<HostAtomics>(i, xyz, stride, inv_cell, m->slots.data(), cap - 1, pt_slot.data(),
                                               pt_bit.data(), counters, f))
                for (int a = 0; a < 3; ++a) { cmin[a] = std::min(cmin[a], f[a]); cmax[a] = std::max(cmax[a], f[a]); }
        }
        if (counters[1] || counters[0] * 2u > cap) { cap *= 4; continue; }
        break;
    }
    unsigned int ncells = 0;
    for (unsigned int s = 0; s < cap; ++s) {
        m->slots[s].cell_base = ncells;
        if (m->slots[s].key != kEmptyKey) ncells += popc64(m->slots[s].mask);
    }
    std::vector<unsigned int> cnt(ncells + 1, 0);
    for (size_t i = 0; i < n; ++i) build_count_body<HostAtomics>(i, m->slots.data(), pt_slot.data(), pt_bit.data(), cnt.data());
    m->cell_start.assign(ncells + 1, 0);
    unsigned int run = 0;
    for (unsigned int c = 0; c < ncells; ++c) { m->cell_start[c] = run; run += cnt[c]; }
    m->cell_start[ncells] = run;
    std::vector<unsigned int> cursor(m->cell_start.begin(), m->cell_start.end());
    m->pts.assign(run, float4{0, 0, 0, 0});
    for (size_t i = 0; i < n; ++i) build_scatter_body<HostAtomics>(i, xyz, stride, pt_slot.data(), cursor.data(), m->pts.data(), pt_pos.data());
    std::vector<unsigned char> dup(n, 0);
    unsigned int ndup = 0;
    for (size_t i = 0; i < n; ++i) {
        dup[i] = build_is_duplicate(i, pt_slot.data(), pt_pos.data(), m->cell_start.data(), m->pts.data());
        ndup += dup[i];
    }
    for (size_t i = 0; i < n; ++i)
        if (dup[i]) m->pts[pt_pos[i]].x = NAN;
    unsigned int nbr_cap = 0;
    if (capacity_hint != 0xFFFFFFFFu) {  // neighbourhood lists (capacity_hint == ~0 disables them)
        nbr_cap = next_pow2(ncells * 8 + 1024);
        while (true) {
            m->nbr.assign(nbr_cap, NbrSlot{kEmptyKey, 0u, 0u});
            counters[0] = counters[1] = 0;
            for (size_t i = 0; i < n; ++i)
                build_nbr_body<HostAtomics>(i, 0, xyz, stride, inv_cell, pt_slot.data(), pt_pos.data(), dup.data(), m->nbr.data(), nbr_cap - 1, nullptr, nullptr, counters);
            if (counters[1] || counters[0] * 2u > nbr_cap) { nbr_cap *= 4; continue; }
            break;
        }
        unsigned int acc = run;
        for (unsigned int s = 0; s < nbr_cap; ++s) { m->nbr[s].start = acc; acc += m->nbr[s].count; }
        m->pts.resize(acc, float4{0, 0, 0, 0});
        std::vector<unsigned int> ncur(nbr_cap, 0);
        for (size_t i = 0; i < n; ++i)
            build_nbr_body<HostAtomics>(i, 1, xyz, stride, inv_cell, pt_slot.data(), pt_pos.data(), dup.data(), m->nbr.data(), nbr_cap - 1, ncur.data(), m->pts.data(), counters);
    }
    VoxelMapView& v = m->view;
    v.slots = m->slots.data(); v.cell_start = m->cell_start.data(); v.pts = m->pts.data();
    v.nbr_slots = nbr_cap ? m->nbr.data() : nullptr; v.nbr_mask = nbr_cap ? nbr_cap - 1 : 0;
    v.slot_mask = cap - 1; v.n_pts = run; v.n_unique = run - ndup; v.inv_cell = inv_cell; v.cell = cell;
    for (int a = 0; a < 3; ++a) { v.cmin[a] = cmin[a]; v.cmax[a] = cmax[a]; }
    v.canon = m->pts.data(); v.w_is_pos = 0;
    m->pos_of_index = pt_pos;
    return m;
}

// capacity_hint: 0 = defaults; 0xFFFFFFFF = no neighbourhood lists; 0xFFFFFFFE = no lists and no coarse level;
// 0xFFFFFFFD = lists on the fine level only (no mid level)
HsMap* hs_map_create(const float* xyz, size_t n, size_t stride, float cell, unsigned int capacity_hint) {
    const bool want_coarse = capacity_hint != 0xFFFFFFFEu;
    HsMap* m = build_level(xyz, n, stride, cell, capacity_hint == 0xFFFFFFFEu ? 0xFFFFFFFFu : (capacity_hint == 0xFFFFFFFDu ? 0u : capacity_hint));
    // the mid level (lists) exists whenever the fine level has lists; capacity_hint 0xFFFFFFFD: fine lists, no mid level
    if (want_coarse && m->view.n_pts && m->view.nbr_slots != nullptr && capacity_hint != 0xFFFFFFFDu) {
        HsMap* c = build_level(xyz, n, stride, cell * kMidFactor, 0);
        const unsigned int np = c->view.n_pts;
        for (unsigned int j = 0; j < np; ++j)  // what DeviceVoxelMap::attach_to does
            c->pts[j].w = int_as_float(static_cast<int>(m->pos_of_index[float_as_int(c->pts[j].w)]));
        for (size_t j = np; j < c->pts.size(); ++j) c->pts[j].w = c->pts[float_as_int(c->pts[j].w)].w;
        c->view.canon = m->pts.data(); c->view.w_is_pos = 1;
        m->mid_map = c;
        m->coarse.mid = c->view;
    }
    float cc = cell;
    for (int l = 0; want_coarse && m->view.n_pts && l < kCoarseLevels; ++l) {
        cc *= kCoarseFactor;
        HsMap* c = build_level(xyz, n, stride, cc, 0xFFFFFFFFu);
        for (unsigned int j = 0; j < c->view.n_pts; ++j)  // what DeviceVoxelMap::attach_to does
            c->pts[j].w = int_as_float(static_cast<int>(m->pos_of_index[float_as_int(c->pts[j].w)]));
        c->view.canon = m->pts.data(); c->view.w_is_pos = 1;
        m->coarse_map[l] = c;
        m->coarse.lv[l] = c->view;
    }
    if (m->view.n_pts) {  // block pyramid (what DeviceVoxelMap::build_pyramid does)
        const int P = pyr_levels_for(m->view.cmin, m->view.cmax);
        unsigned int nb = 0;
        for (const VoxelSlot& s : m->slots) nb += s.key != kEmptyKey;
        const unsigned int cap = std::max(1024u, next_pow2(nb * 2));
        m->pyr.assign(static_cast<size_t>(cap) * P, PyrSlot{kEmptyKey, 0ull});
        unsigned int counters[2] = {0, 0};
        for (const VoxelSlot& s : m->slots)
            if (s.key != kEmptyKey) build_pyr_body<HostAtomics>(s.key, m->pyr.data(), cap - 1, counters);
        for (int l = 1; l < P; ++l)
            for (unsigned int j = 0; j < cap; ++j) {
                const PyrSlot& s = m->pyr[static_cast<size_t>(cap) * (l - 1) + j];
                if (s.key != kEmptyKey) build_pyr_body<HostAtomics>(s.key, m->pyr.data() + static_cast<size_t>(cap) * l, cap - 1, counters);
            }
        for (int l = 0; l < P; ++l) { m->coarse.pyr.slots[l] = m->pyr.data() + static_cast<size_t>(cap) * l; m->coarse.pyr.mask[l] = cap - 1; }
        m->coarse.pyr.levels = P;
        m->coarse.pyr_mode = 0;
    }
    return m;
}
// 0: shells only (default), 1: pyramid after the mid level's list, 2: pyramid for all of stage 2
void hs_map_set_pyr_mode(HsMap* m, int mode) { m->coarse.pyr_mode = mode; }
void hs_map_destroy(HsMap* m) { delete m; }
// search-stage counters since the last call (see LR_STAT in voxel_map.cuh); reading clears them
void hs_knn_stats(unsigned long long* out16) {
    for (int i = 0; i < 16; ++i) { out16[i] = g_knn_stats[i]; g_knn_stats[i] = 0; }
}
void hs_map_stats(const HsMap* m, uint32_t* out) {  // capacity, n_pts, n_unique, n_cells
    out[0] = m->view.slot_mask + 1; out[1] = m->view.n_pts; out[2] = m->view.n_unique;
    out[3] = static_cast<uint32_t>(m->cell_start.size() - 1);
}

void hs_knn(const HsMap* m, const float* q, size_t nq, size_t stride, int k, int32_t* idx_out) {
    for (size_t i = 0; i < nq; ++i) {
        const float* p = point_ptr(q, i, stride);
        const bool ok = finite3(p[0], p[1], p[2]);
        if (k == 1) {
            KnnResult<1> r;
            knn_query<1>(m->view, m->coarse, ok, p[0], p[1], p[2], r);
            idx_out[i] = r.pos[0] != kNoPos ? knn_index_of(m->view.pts, r.pos[0]) : -1;
        } else {
            KnnResult<5> r;
            knn_query<5>(m->view, m->coarse, ok, p[0], p[1], p[2], r);
            for (int j = 0; j < 5; ++j) idx_out[i * 5 + j] = r.pos[j] != kNoPos ? knn_index_of(m->view.pts, r.pos[j]) : -1;
        }
    }
}

// k-NN of q seeded with the k-NN of seed_q (a nearby but different query): must equal the unseeded result
void hs_knn_seeded(const HsMap* m, const float* q, const float* seed_q, size_t nq, size_t stride, int32_t* idx_out) {
    for (size_t i = 0; i < nq; ++i) {
        const float* p = point_ptr(q, i, stride);
        const float* sq = point_ptr(seed_q, i, stride);
        KnnResult<5> s, r;
        knn_query<5>(m->view, m->coarse, finite3(sq[0], sq[1], sq[2]), sq[0], sq[1], sq[2], s);
        knn_query<5>(m->view, m->coarse, finite3(p[0], p[1], p[2]), p[0], p[1], p[2], r, s.pos);
        for (int j = 0; j < 5; ++j) idx_out[i * 5 + j] = r.pos[j] != kNoPos ? knn_index_of(m->view.pts, r.pos[j]) : -1;
    }
}

// The tracked search of k_icp_nn<K, true> over a SEQUENCE of query sets (steps x nq points, the same points moving a
// little from step to step): step 0 searches from scratch, later steps try the K-gathers shortcut first
// (knn_track_try), else search with the bookkeeping (knn_query_fast_track), else finish through stage 2.
// idx_out: steps x nq x K original indices; returns the number of searches the shortcut replaced.
}  // extern "C"
template <int K>
static size_t knn_tracked_t(const HsMap* m, const float* q, size_t steps, size_t nq, int32_t* idx_out) {
    std::vector<unsigned int> pos(nq * K, kNoPos);
    std::vector<KnnTrack> track(nq, KnnTrack{0.0f, 0.0f, 0.0f, -1.0f});
    size_t skipped = 0;
    for (size_t s = 0; s < steps; ++s)
        for (size_t i = 0; i < nq; ++i) {
            const float* p = q + (s * nq + i) * 3;
            KnnResult<K> r;
            knn_init(r);
            if (finite3(p[0], p[1], p[2]) && m->view.n_pts != 0) {
                if (s == 0) {
                    knn_query<K>(m->view, m->coarse, true, p[0], p[1], p[2], r);
                    track[i].margin = -1.0f;
                } else if (knn_track_try<K>(m->view, p[0], p[1], p[2], &pos[i * K], track[i], r)) {
                    ++skipped;
                } else if (!knn_query_fast_track<K>(m->view, p[0], p[1], p[2], r, &pos[i * K], track[i])) {
                    knn_query_finish<K>(m->view, m->coarse, p[0], p[1], p[2], r);
                }
            }
            for (int j = 0; j < K; ++j) {
                pos[i * K + j] = r.pos[j];
                idx_out[(s * nq + i) * K + j] = r.pos[j] != kNoPos ? knn_index_of(m->view.pts, r.pos[j]) : -1;
            }
        }
    return skipped;
}
extern "C" {
size_t hs_knn_tracked(const HsMap* m, const float* q, size_t steps, size_t nq, int k, int32_t* idx_out) {
    return k == 1 ? knn_tracked_t<1>(m, q, steps, nq, idx_out) : knn_tracked_t<5>(m, q, steps, nq, idx_out);
}

}  // extern "C"

// prm: max_nn_distance, max_plane_distance, plane_fit_eps, eps, max_iteration, min_effective_pts, max_line_distance
static IcpParams make_params(const double* prm) {
    IcpParams p;
    p.max_line_distance = prm[6];
    p.max_nn_distance = prm[0]; p.max_plane_distance = prm[1]; p.plane_fit_eps = prm[2]; p.eps = prm[3];
    p.max_iteration = static_cast<int>(prm[4]); p.min_effective_pts = static_cast<int>(prm[5]);
    return p;
}
// seeds (optional): n * k neighbour positions carried from one Gauss-Newton iteration to the next, exactly as the
// kernels reuse their nn_pos scratch
template <int METHOD>
static void hb_impl(const HsMap* m, const IcpParams& p, const float* src, size_t n, size_t stride, const Pose& T,
                    Accum& acc, uint8_t* gate, int32_t* nn_out, unsigned int* seeds = nullptr) {
    accum_zero(acc);
    const int k = METHOD == kIcpP2P ? 1 : 5;
    for (size_t i = 0; i < n; ++i) {
        const float* s = point_ptr(src, i, stride);
        int nn[5];
        const unsigned char g = icp_point<METHOD>(m->view, m->coarse, p, T, s[0], s[1], s[2], acc, nn, seeds ? seeds + i * k : nullptr);
        if (gate) gate[i] = g;
        if (nn_out) for (int j = 0; j < k; ++j) nn_out[i * k + j] = nn[j];
    }
}
static void unpack(const Accum& acc, double* H36, double* B6) {
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) { H36[c * 6 + r] = acc.v[hidx(r, c)]; H36[r * 6 + c] = acc.v[hidx(r, c)]; }
    for (int i = 0; i < 6; ++i) B6[i] = acc.v[21 + i];
}

extern "C" {

void hs_icp_hb(const HsMap* m, int method, const double* prm, const float* src, size_t n, size_t stride,
               const double* pose7, double* H36, double* B6, int64_t* counts, double* sum_sq, uint8_t* gate,
               int32_t* nn_out) {
    const IcpParams p = make_params(prm);
    Pose T;
    pose_load(T, pose7);
    Accum acc;
    if (method == kIcpP2P) hb_impl<kIcpP2P>(m, p, src, n, stride, T, acc, gate, nn_out);
    else if (method == kIcpP2Line) hb_impl<kIcpP2Line>(m, p, src, n, stride, T, acc, gate, nn_out);
    else hb_impl<kIcpP2Plane>(m, p, src, n, stride, T, acc, gate, nn_out);
    unpack(acc, H36, B6);
    counts[0] = acc.n_eff; counts[1] = acc.n_inl;
    *sum_sq = acc.v[27];
}

// returns iterations executed; status[0]=updates, [1]=converged, [2]=degenerate(last)
int hs_icp_align(const HsMap* m, int method, const double* prm, const float* src, size_t n, size_t stride,
                 const double* pose_in, double* pose_out, int32_t* status) {
    const IcpParams p = make_params(prm);
    Pose T;
    pose_load(T, pose_in);
    int iters = 0;
    status[0] = status[1] = status[2] = 0;
    std::vector<unsigned int> seeds(n * 5, kNoPos);  // iteration 0 starts unseeded
    for (int it = 0; it < p.max_iteration; ++it) {
        Accum acc;
        int r;
        if (method == kIcpP2P) {
            hb_impl<kIcpP2P>(m, p, src, n, stride, T, acc, nullptr, nullptr, seeds.data());
            r = icp_gn_update<kIcpP2P>(acc.v, acc.n_eff, p, T);
        } else if (method == kIcpP2Line) {
            hb_impl<kIcpP2Line>(m, p, src, n, stride, T, acc, nullptr, nullptr, seeds.data());
            r = icp_gn_update<kIcpP2Line>(acc.v, acc.n_eff, p, T);
        } else {
            hb_impl<kIcpP2Plane>(m, p, src, n, stride, T, acc, nullptr, nullptr, seeds.data());
            r = icp_gn_update<kIcpP2Plane>(acc.v, acc.n_eff, p, T);
        }
        iters = it + 1;
        status[2] = r == 0;
        if (r) status[0]++;
        if (r == 2) { status[1] = 1; break; }
    }
    pose_store(T, pose_out);
    return iters;
}

void hs_plane_svd5(const double* pts15, double* coef4) {
    double P[5][3];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j) P[i][j] = pts15[i * 3 + j];
    double c[4];
    plane_svd5(P, c);
    for (int i = 0; i < 4; ++i) coef4[i] = c[i];
}
int hs_plane_fit5_fast(const double* pts15, double* coef4) {
    double P[5][3];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j) P[i][j] = pts15[i * 3 + j];
    double c[4] = {0, 0, 0, 0};
    const bool ok = plane_fit5_fast(P, c);
    for (int i = 0; i < 4; ++i) coef4[i] = c[i];
    return ok ? 1 : 0;
}
void hs_sym3_eigen(const double* S6, double* lam3, double* Q9) { sym3_eigen(S6, lam3, Q9); }
int hs_gn_solve6(const double* Hu21, const double* b6, double* dx6) { return gn_solve6(Hu21, b6, dx6) ? 1 : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// NDT
// ---------------------------------------------------------------------------------------------
#include <algorithm>

struct HsNdt {
    std::vector<NdtSlot> slots;
    std::vector<NdtVoxel> voxels;
    NdtMapView view;
};

static NdtParams make_ndt_params(const double* prm) {  // res_outlier_th, eps, max_iteration, min_effective_pts, min_pts_in_voxel, n_nearby
    NdtParams p;
    p.res_outlier_th = prm[0]; p.eps = prm[1]; p.max_iteration = static_cast<int>(prm[2]);
    p.min_effective_pts = static_cast<int>(prm[3]); p.min_pts_in_voxel = static_cast<int>(prm[4]);
    p.n_nearby = static_cast<int>(prm[5]);
    return p;
}
static void ndt_hb_impl(const HsNdt* m, const NdtParams& p, const float* src, size_t n, size_t stride, const Pose& T,
                        Accum& acc, uint8_t* hits) {
    accum_zero(acc);
    for (size_t i = 0; i < n; ++i) {
        const float* s = point_ptr(src, i, stride);
        const unsigned char h = ndt_point(m->view, p, T, s[0], s[1], s[2], acc);
        if (hits) hits[i] = h;
    }
}

extern "C" {

HsNdt* hs_ndt_create(const float* xyz, size_t n, size_t stride, double voxel_size, int min_pts) {
    auto* m = new HsNdt;
    const double inv = 1.0 / voxel_size;
    unsigned int cap = 1024;
    while (cap < n / 4 + 1024) cap <<= 1;
    std::vector<unsigned int> pt_slot(n);
    unsigned int counters[2];
    while (true) {
        m->slots.assign(cap, NdtSlot{kNdtEmpty, -1, 0u});
        counters[0] = counters[1] = 0;
        for (size_t i = 0; i < n; ++i) ndt_insert_body<HostAtomics>(i, xyz, stride, inv, m->slots.data(), cap - 1, pt_slot.data(), counters);
        if (counters[1] || counters[0] * 2u > cap) { cap <<= 2; continue; }
        break;
    }
    std::vector<unsigned int> start(cap + 1, 0);
    for (unsigned int s = 0; s < cap; ++s) start[s + 1] = start[s] + m->slots[s].count;
    std::vector<unsigned int> cursor(start.begin(), start.end() - 1), members(n);
    for (size_t i = 0; i < n; ++i)
        if (pt_slot[i] != 0xFFFFFFFFu) members[cursor[pt_slot[i]]++] = static_cast<unsigned int>(i);
    for (unsigned int s = 0; s < cap; ++s) {
        const unsigned int cnt = m->slots[s].count;
        if (m->slots[s].key == kNdtEmpty || !(static_cast<long long>(cnt) > static_cast<long long>(min_pts))) continue;
        unsigned int* idx = members.data() + start[s];
        std::sort(idx, idx + cnt);
        NdtVoxel v;
        ndt_voxel_stats(idx, cnt, xyz, stride, v);
        m->slots[s].vid = static_cast<int>(m->voxels.size());
        m->voxels.push_back(v);
    }
    m->view.slots = m->slots.data(); m->view.voxels = m->voxels.data(); m->view.slot_mask = cap - 1;
    m->view.n_voxels = static_cast<unsigned int>(m->voxels.size()); m->view.inv_voxel = inv;
    return m;
}
void hs_ndt_destroy(HsNdt* m) { delete m; }

// Incremental NDT: a voxel table built from a dump (keys nv*3, point lists per voxel as in one SetIncNdtTargetCloud
// call: member indices + group starts) so that inc_ndt_voxel_stats and inc_ndt_point run on the host.
HsNdt* hs_inc_ndt_create(const int32_t* keys, const uint32_t* group_start, const uint32_t* members, size_t nv,
                         const float* xyz, size_t stride, double voxel_size) {
    auto* m = new HsNdt;
    unsigned int cap = 1024;
    while (cap < nv * 4 + 16) cap <<= 1;
    m->slots.assign(cap, NdtSlot{kNdtEmpty, -1, 0u});
    m->voxels.resize(nv);
    for (size_t v = 0; v < nv; ++v) {
        inc_ndt_voxel_stats(members + group_start[v], group_start[v + 1] - group_start[v], xyz, stride, m->voxels[v]);
        const unsigned long long key = ndt_pack(keys[v * 3], keys[v * 3 + 1], keys[v * 3 + 2]);
        unsigned int h = ndt_hash(key) & (cap - 1);
        while (m->slots[h].key != kNdtEmpty) h = (h + 1) & (cap - 1);
        m->slots[h] = NdtSlot{key, static_cast<int>(v), group_start[v + 1] - group_start[v]};
    }
    m->view.slots = m->slots.data(); m->view.voxels = m->voxels.data(); m->view.slot_mask = cap - 1;
    m->view.n_voxels = static_cast<unsigned int>(nv); m->view.inv_voxel = 1.0 / voxel_size;
    return m;
}
void hs_inc_ndt_hb(const HsNdt* m, const double* prm, const float* src, size_t n, size_t stride, const double* pose7,
                   double* H36, double* B6, int64_t* counts, double* sum_sq, uint8_t* hits) {
    const NdtParams p = make_ndt_params(prm);
    Pose T;
    pose_load(T, pose7);
    Accum acc;
    accum_zero(acc);
    for (size_t i = 0; i < n; ++i) {
        const float* s = point_ptr(src, i, stride);
        const unsigned char h = inc_ndt_point(m->view, p, T, s[0], s[1], s[2], acc);
        if (hits) hits[i] = h;
    }
    unpack(acc, H36, B6);
    counts[0] = acc.n_eff; counts[1] = acc.n_inl;
    *sum_sq = acc.v[27];
}
size_t hs_ndt_num_voxels(const HsNdt* m) { return m->voxels.size(); }
void hs_ndt_get_voxels(const HsNdt* m, int32_t* keys, double* mu, double* info, int32_t* npts) {
    struct Rec { int k[3]; int vid; int cnt; };
    std::vector<Rec> recs;
    for (const NdtSlot& s : m->slots)
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
    for (size_t i = 0; i < recs.size(); ++i) {
        for (int a = 0; a < 3; ++a) { keys[i * 3 + a] = recs[i].k[a]; mu[i * 3 + a] = m->voxels[recs[i].vid].mu[a]; }
        for (int a = 0; a < 9; ++a) info[i * 9 + a] = m->voxels[recs[i].vid].info[a];
        npts[i] = recs[i].cnt;
    }
}
void hs_ndt_hb(const HsNdt* m, const double* prm, const float* src, size_t n, size_t stride, const double* pose7,
               double* H36, double* B6, int64_t* counts, double* sum_sq, uint8_t* hits) {
    const NdtParams p = make_ndt_params(prm);
    Pose T;
    pose_load(T, pose7);
    Accum acc;
    ndt_hb_impl(m, p, src, n, stride, T, acc, hits);
    unpack(acc, H36, B6);
    counts[0] = acc.n_eff; counts[1] = acc.n_inl;
    *sum_sq = acc.v[27];
}
// status: [0]=updates [1]=converged [2]=degenerate [3]=pose_written
int hs_ndt_align(const HsNdt* m, const double* prm, const float* src, size_t n, size_t stride, const double* pose_in,
                 double* pose_inout, int32_t* status) {
    const NdtParams p = make_ndt_params(prm);
    Pose T;
    pose_load(T, pose_in);
    int iters = 0;
    status[0] = status[1] = status[2] = 0; status[3] = 1;
    for (int it = 0; it < p.max_iteration; ++it) {
        Accum acc;
        ndt_hb_impl(m, p, src, n, stride, T, acc, nullptr);
        const int r = ndt_gn_update(acc.v, acc.n_eff, p, T);
        iters = it + 1;
        status[2] = (r == 0 || r == 3);
        if (r == 1 || r == 2) status[0]++;
        if (r == 2) { status[1] = 1; break; }
        if (r == 3) { status[3] = 0; break; }
    }
    if (status[3]) pose_store(T, pose_inout);
    return iters;
}

}  // extern "C"

// Incremental NDT voxel cache on the device (see device_inc_ndt.cuh for the formulation).
#include <algorithm>
#include <climits>
#include <cmath>

#include "device_inc_ndt.cuh"

namespace locreg {

namespace {

constexpr int kNoPrev = INT_MIN;  // prev[]: the key was never accessed before (not in the cache, not in the cloud)
constexpr unsigned int kT = 256;

// device scalars (ctr_)
enum : int {
    kCtrM = 0,       // entries in the cache
    kCtrRuns = 1,    // runs of the cloud
    kCtrGroups = 2,  // distinct keys of the cloud
    kCtrOldGroups = 3,  // ... that have an entry in the cache
    kCtrEvictOld = 4,   // untouched old entries to evict (the oldest ones)
    kCtrEvictGrp = 5,   // keys of the cloud to evict (those accessed last the longest ago)
    kCtrGroupBase = 6,  // position of the first surviving key of the cloud in the new order
    kCtrNewM = 7,       // entries after the cloud
    kCtrScratch = 8,    // scan totals nobody reads
    kCtrCount = 16
};

unsigned int next_pow2(size_t v) {
    unsigned int p = 1024;
    while (p < v) p <<= 1;
    return p;
}
unsigned int blocks_for(size_t n) { return static_cast<unsigned int>(std::max<size_t>((n + kT - 1) / kT, 1)); }

// ---- runs -------------------------------------------------------------------------------------------------------------
// voxel key of every point ((pt * inv_voxel_size).cast<int>(), ndt_registration.cpp:154), kNdtEmpty = no access
// (non-finite point: deviation D1; key outside the packable range)
__global__ void k_inc_keys(const void* __restrict__ xyz, size_t n, size_t stride, double inv_voxel, unsigned long long* __restrict__ pkey) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(static_cast<const char*>(xyz) + i * stride);
    unsigned long long key = kNdtEmpty;
    if (finite3(p[0], p[1], p[2])) {
        const int kx = ndt_trunc(LR_DMUL(static_cast<double>(p[0]), inv_voxel)), ky = ndt_trunc(LR_DMUL(static_cast<double>(p[1]), inv_voxel)),
                  kz = ndt_trunc(LR_DMUL(static_cast<double>(p[2]), inv_voxel));
        if (ndt_key_ok(kx, ky, kz)) key = ndt_pack(kx, ky, kz);
    }
    pkey[i] = key;
}
// a run starts where a valid point follows a point with another key (an invalid point separates runs too: two runs of
// the same key in a row are harmless - the second is a hit that moves the front entry to the front)
__global__ void k_inc_heads(const unsigned long long* __restrict__ pkey, size_t n, unsigned int* __restrict__ head) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = pkey[i];
    head[i] = (k != kNdtEmpty && (i == 0 || pkey[i - 1] != k)) ? 1u : 0u;
}
// rid = exclusive scan of head: the run a head starts; a run ends where the next point has another key
__global__ void k_inc_runs(const unsigned long long* __restrict__ pkey, const unsigned int* __restrict__ head,
                           const unsigned int* __restrict__ rid, size_t n, unsigned long long* __restrict__ run_key,
                           unsigned int* __restrict__ run_first, unsigned int* __restrict__ run_last) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = pkey[i];
    if (k == kNdtEmpty) return;
    const unsigned int r = rid[i] + head[i] - 1u;  // runs started up to and including i, less one
    if (head[i]) { run_key[r] = k; run_first[r] = static_cast<unsigned int>(i); }
    if (i + 1 == n || pkey[i + 1] != k) run_last[r] = static_cast<unsigned int>(i);
}

// ---- groups -----------------------------------------------------------------------------------------------------------
// scratch table: key -> slot; per slot the last run and the number of runs of the key
__global__ void k_inc_group(const unsigned long long* __restrict__ run_key, const unsigned int* __restrict__ ctr,
                            unsigned long long* t_key, unsigned int t_mask, unsigned int* __restrict__ run_slot,
                            unsigned int* g_last, unsigned int* g_nruns) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ctr[kCtrRuns]) return;
    const unsigned long long key = run_key[j];
    unsigned int h = ndt_hash(key) & t_mask;
    while (true) {
        const unsigned long long k = atomicCAS(&t_key[h], kNdtEmpty, key);
        if (k == kNdtEmpty || k == key) break;
        h = (h + 1) & t_mask;
    }
    run_slot[j] = h;
    atomicMax(&g_last[h], j);
    atomicAdd(&g_nruns[h], 1u);
}
// the cache entry of every key of the cloud (published table of the previous cloud), marked as touched
__global__ void k_inc_lookup(const unsigned long long* __restrict__ t_key, unsigned int t_cap, const NdtSlot* __restrict__ slots,
                             unsigned int slot_mask, int* __restrict__ g_vid, unsigned int* touched, unsigned int* ctr) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t_cap) return;
    const unsigned long long key = t_key[s];
    if (key == kNdtEmpty) return;
    int vid = -1;
    unsigned int h = ndt_hash(key) & slot_mask;
    while (true) {
        const NdtSlot sl = slots[h];
        if (sl.key == key) { vid = sl.vid; break; }
        if (sl.key == kNdtEmpty) break;
        h = (h + 1) & slot_mask;
    }
    g_vid[s] = vid;
    atomicAdd(&ctr[kCtrGroups], 1u);
    if (vid >= 0) {
        touched[vid] = 1u;
        atomicAdd(&ctr[kCtrOldGroups], 1u);
    }
}
__global__ void k_inc_scatter(const unsigned int* __restrict__ run_slot, const unsigned int* __restrict__ ctr,
                              const unsigned int* __restrict__ g_start, unsigned int* g_cursor, unsigned int* __restrict__ members) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ctr[kCtrRuns]) return;
    const unsigned int s = run_slot[j];
    members[g_start[s] + atomicAdd(&g_cursor[s], 1u)] = j;
}
// one thread per key: its runs in run order (insertion sort: a handful to a few hundred), each run's previous access on
// the unified time line, and the runs as (first point, length) for the statistics
__global__ void k_inc_sort_prev(const unsigned long long* __restrict__ t_key, unsigned int t_cap, const unsigned int* __restrict__ g_start,
                                const unsigned int* __restrict__ g_nruns, unsigned int* members, const int* __restrict__ g_vid,
                                const unsigned int* __restrict__ rank_of, const unsigned int* __restrict__ ctr,
                                const unsigned int* __restrict__ run_first, const unsigned int* __restrict__ run_last,
                                IncRun* __restrict__ runs_sorted, int* __restrict__ prev) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t_cap || t_key[s] == kNdtEmpty) return;
    const unsigned int beg = g_start[s], cnt = g_nruns[s];
    unsigned int* idx = members + beg;
    for (unsigned int a = 1; a < cnt; ++a) {
        const unsigned int v = idx[a];
        unsigned int b = a;
        while (b > 0 && idx[b - 1] > v) { idx[b] = idx[b - 1]; --b; }
        idx[b] = v;
    }
    const int vid = g_vid[s];
    int pv = vid >= 0 ? static_cast<int>(rank_of[vid]) - static_cast<int>(ctr[kCtrM]) : kNoPrev;
    for (unsigned int a = 0; a < cnt; ++a) {
        const unsigned int j = idx[a];
        prev[j] = pv;
        pv = static_cast<int>(j);
        runs_sorted[beg + a] = IncRun{run_first[j], run_last[j] - run_first[j] + 1u};
    }
}

// ---- hit or miss --------------------------------------------------------------------------------------------------------
// One warp per run.  Run j whose key was last accessed at time i is a miss iff C or more distinct keys were accessed in
// between: the -1 - i old entries newer than i (when i < 0), plus the runs p of the window (i, j) that are the first
// access to their key inside it (prev[p] < i).  Windows shorter than C need no count.
__global__ void k_inc_classify(const int* __restrict__ prev, const unsigned int* __restrict__ ctr, unsigned int C,
                               unsigned int* __restrict__ miss) {
    const unsigned int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (j >= ctr[kCtrRuns]) return;
    const int i = prev[j];
    unsigned int is_miss = 0u;
    if (i == kNoPrev) {
        is_miss = 1u;
    } else if (static_cast<long long>(j) - static_cast<long long>(i) - 1 >= static_cast<long long>(C)) {
        const unsigned int older = i < 0 ? static_cast<unsigned int>(-1 - i) : 0u;
        unsigned int distinct = older;
        const unsigned int lo = i < 0 ? 0u : static_cast<unsigned int>(i) + 1u;
        for (unsigned int p0 = lo; p0 < j && distinct < C; p0 += 32) {
            const unsigned int p = p0 + lane;
            distinct += __popc(__ballot_sync(0xffffffffu, p < j && prev[p] < i));
        }
        is_miss = distinct >= C ? 1u : 0u;
    }
    if (lane == 0) miss[j] = is_miss;
}

// ---- survivors and the new LRU order --------------------------------------------------------------------------------------
__global__ void k_inc_plan(unsigned int* ctr, unsigned int C) {
    const unsigned int m = ctr[kCtrM], d = ctr[kCtrGroups], d_old = ctr[kCtrOldGroups];
    const unsigned int untouched = m - d_old, total = untouched + d;
    const unsigned int evict = total > C ? total - C : 0u;
    const unsigned int e_old = evict < untouched ? evict : untouched;
    ctr[kCtrEvictOld] = e_old;
    ctr[kCtrEvictGrp] = evict - e_old;
    ctr[kCtrGroupBase] = untouched - e_old;
    ctr[kCtrNewM] = total - evict;
}
__global__ void k_inc_old_flags(const unsigned int* __restrict__ order, const unsigned int* __restrict__ ctr,
                                const unsigned int* __restrict__ touched, unsigned int cap, unsigned int* __restrict__ uflag) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cap) return;
    uflag[r] = (r < ctr[kCtrM] && !touched[order[r]]) ? 1u : 0u;
}
// untouched old entries: the oldest kCtrEvictOld are dropped (their records become free), the rest keep their order
__global__ void k_inc_keep_old(const unsigned int* __restrict__ order, const unsigned int* __restrict__ ctr,
                               const unsigned int* __restrict__ uflag, const unsigned int* __restrict__ upos, unsigned int cap,
                               unsigned int* __restrict__ new_order, unsigned long long* ent_key) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cap || !uflag[r]) return;
    const unsigned int vid = order[r], e_old = ctr[kCtrEvictOld];
    if (upos[r] < e_old) ent_key[vid] = kNdtEmpty;
    else new_order[upos[r] - e_old] = vid;
}
__global__ void k_inc_last_flags(const unsigned int* __restrict__ run_slot, const unsigned int* __restrict__ g_last,
                                 const unsigned int* __restrict__ ctr, size_t n, unsigned int* __restrict__ islast) {
    const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    islast[j] = (j < ctr[kCtrRuns] && g_last[run_slot[j]] == j) ? 1u : 0u;
}
// keys of the cloud by last access: the first kCtrEvictGrp are dropped, the rest follow the old entries in the new order;
// needflag marks the survivors that have no record yet
__global__ void k_inc_keep_groups(const unsigned int* __restrict__ run_slot, const unsigned int* __restrict__ islast,
                                  const unsigned int* __restrict__ lpos, const unsigned int* __restrict__ ctr, size_t n,
                                  const int* __restrict__ g_vid, unsigned int* __restrict__ g_keep, unsigned int* __restrict__ g_newpos,
                                  unsigned long long* ent_key, unsigned int* __restrict__ new_order, unsigned int* __restrict__ needflag) {
    const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned int need = 0u;
    if (islast[j]) {
        const unsigned int s = run_slot[j], e_grp = ctr[kCtrEvictGrp];
        const int vid = g_vid[s];
        if (lpos[j] < e_grp) {
            g_keep[s] = 0u;
            if (vid >= 0) ent_key[vid] = kNdtEmpty;
        } else {
            const unsigned int pos = ctr[kCtrGroupBase] + lpos[j] - e_grp;
            g_keep[s] = 1u;
            g_newpos[s] = pos;
            if (vid >= 0) new_order[pos] = static_cast<unsigned int>(vid);
            else need = 1u;
        }
    }
    needflag[j] = need;
}
__global__ void k_inc_free_flags(const unsigned long long* __restrict__ ent_key, unsigned int cap, unsigned int* __restrict__ fflag) {
    const unsigned int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < cap) fflag[v] = ent_key[v] == kNdtEmpty ? 1u : 0u;
}
__global__ void k_inc_free_list(const unsigned int* __restrict__ fflag, const unsigned int* __restrict__ fpos, unsigned int cap,
                                unsigned int* __restrict__ free_list) {
    const unsigned int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < cap && fflag[v]) free_list[fpos[v]] = v;
}
// the i-th new surviving key takes the i-th free record
__global__ void k_inc_assign(const unsigned int* __restrict__ run_slot, const unsigned int* __restrict__ needflag,
                             const unsigned int* __restrict__ npos, size_t n, const unsigned int* __restrict__ free_list,
                             const unsigned long long* __restrict__ run_key, int* __restrict__ g_vid,
                             const unsigned int* __restrict__ g_newpos, unsigned long long* ent_key, unsigned int* __restrict__ new_order) {
    const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n || !needflag[j]) return;
    const unsigned int s = run_slot[j], vid = free_list[npos[j]];
    g_vid[s] = static_cast<int>(vid);
    ent_key[vid] = run_key[j];
    new_order[g_newpos[s]] = vid;
}

// ---- statistics -----------------------------------------------------------------------------------------------------------
// one thread per surviving key of the cloud: UpdateVoxel (:185-236, first branch) over the points it received since its
// last miss, in arrival order
__global__ void k_inc_stats(const unsigned long long* __restrict__ t_key, unsigned int t_cap, const unsigned int* __restrict__ g_keep,
                            const int* __restrict__ g_vid, const unsigned int* __restrict__ g_start, const unsigned int* __restrict__ g_nruns,
                            const unsigned int* __restrict__ members, const unsigned int* __restrict__ miss,
                            const IncRun* __restrict__ runs_sorted, const void* __restrict__ xyz, size_t stride, NdtVoxel* voxels,
                            unsigned int* ent_cnt) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t_cap || t_key[s] == kNdtEmpty || !g_keep[s]) return;
    const unsigned int beg = g_start[s], cnt = g_nruns[s];
    unsigned int from = 0;
    for (unsigned int a = cnt; a > 0; --a)
        if (miss[members[beg + a - 1]]) { from = a - 1; break; }
    unsigned int npts = 0;
    for (unsigned int a = from; a < cnt; ++a) npts += runs_sorted[beg + a].len;
    NdtVoxel v;
    inc_ndt_voxel_stats_of(IncMembersRuns{runs_sorted + beg + from, cnt - from}, npts, xyz, stride, v);
    const int vid = g_vid[s];
    voxels[vid] = v;
    ent_cnt[vid] = npts;
}

// ---- publish ----------------------------------------------------------------------------------------------------------------
__global__ void k_inc_clear_slots(NdtSlot* slots, unsigned int cap) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) slots[s] = NdtSlot{kNdtEmpty, -1, 0u};
}
__global__ void k_inc_publish(const unsigned int* __restrict__ new_order, unsigned int* ctr, const unsigned long long* __restrict__ ent_key,
                              const unsigned int* __restrict__ ent_cnt, unsigned int cap, NdtSlot* slots, unsigned int slot_mask,
                              unsigned int* __restrict__ rank_of) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int m_new = ctr[kCtrNewM];
    if (r == 0) ctr[kCtrM] = m_new;  // (nobody reads kCtrM in this kernel)
    if (r >= cap || r >= m_new) return;
    const unsigned int vid = new_order[r];
    rank_of[vid] = r;
    const unsigned long long key = ent_key[vid];
    unsigned int h = ndt_hash(key) & slot_mask;
    while (atomicCAS(&slots[h].key, kNdtEmpty, key) != kNdtEmpty) h = (h + 1) & slot_mask;
    slots[h].vid = static_cast<int>(vid);
    slots[h].count = ent_cnt[vid];
}

}  // namespace

DeviceIncNdtMap::~DeviceIncNdtMap() {
    release();
    if (scratch_) cudaFree(scratch_);
}
void DeviceIncNdtMap::release() {
    void* ptrs[] = {slots_, voxels_, ent_key_, ent_cnt_, order_[0], order_[1], rank_of_, ctr_};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    slots_ = nullptr; voxels_ = nullptr; ent_key_ = nullptr; ent_cnt_ = nullptr; order_[0] = order_[1] = nullptr;
    rank_of_ = nullptr; ctr_ = nullptr;
}

void DeviceIncNdtMap::configure(double voxel_size, size_t capacity) {
    release();
    view_ = NdtMapView{};
    view_.inv_voxel = 1.0 / voxel_size;
    capacity_ = std::max<size_t>(capacity, 2);
    if (capacity_ >= (1ull << 30)) throw std::invalid_argument("incremental NDT capacity too large");
    cap_slots_ = next_pow2(capacity_ * 4);
    cur_ = 0;
    LR_CUDA(cudaMalloc(&slots_, static_cast<size_t>(cap_slots_) * sizeof(NdtSlot)));
    LR_CUDA(cudaMalloc(&voxels_, capacity_ * sizeof(NdtVoxel)));
    LR_CUDA(cudaMalloc(&ent_key_, capacity_ * sizeof(unsigned long long)));
    LR_CUDA(cudaMalloc(&ent_cnt_, capacity_ * sizeof(unsigned int)));
    LR_CUDA(cudaMalloc(&order_[0], capacity_ * sizeof(unsigned int)));
    LR_CUDA(cudaMalloc(&order_[1], capacity_ * sizeof(unsigned int)));
    LR_CUDA(cudaMalloc(&rank_of_, capacity_ * sizeof(unsigned int)));
    LR_CUDA(cudaMalloc(&ctr_, kCtrCount * sizeof(unsigned int)));
    // (the default stream: configure runs once, before any cloud; cudaMemset is ordered before later work of any stream
    // of a handle created with blocking streams - and add_cloud's first kernels follow a host-side return of this call)
    LR_CUDA(cudaMemset(slots_, 0xFF, static_cast<size_t>(cap_slots_) * sizeof(NdtSlot)));  // key = kNdtEmpty, vid = -1
    LR_CUDA(cudaMemset(ent_key_, 0xFF, capacity_ * sizeof(unsigned long long)));
    LR_CUDA(cudaMemset(ent_cnt_, 0, capacity_ * sizeof(unsigned int)));
    LR_CUDA(cudaMemset(ctr_, 0, kCtrCount * sizeof(unsigned int)));
    LR_CUDA(cudaDeviceSynchronize());
    view_.slots = slots_; view_.voxels = voxels_; view_.slot_mask = cap_slots_ - 1;
    view_.n_voxels = static_cast<unsigned int>(capacity_);  // an upper bound; size() reads the count back
}

void DeviceIncNdtMap::reserve_scratch(size_t bytes) {
    if (bytes <= scratch_bytes_) return;
    if (scratch_) cudaFree(scratch_);
    scratch_ = nullptr; scratch_bytes_ = 0;
    const size_t want = bytes + bytes / 4;
    LR_CUDA(cudaMalloc(&scratch_, want));
    scratch_bytes_ = want;
}

void DeviceIncNdtMap::add_cloud(const void* d_xyz, size_t n, size_t stride, cudaStream_t stream) {
    if (n == 0) return;
    if (n >= (1ull << 30)) throw std::invalid_argument("cloud too large for the incremental NDT cache (>= 2^30 points)");
    const unsigned int cap = static_cast<unsigned int>(capacity_), C = cap - 1u;
    const unsigned int t_cap = next_pow2(2 * n);
    // ---- carve the scratch block
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = at; at += (bytes + 255) & ~static_cast<size_t>(255); return o; };
    const size_t o_pkey = take(n * 8), o_runkey = take(n * 8), o_runs = take(n * sizeof(IncRun)), o_tkey = take(static_cast<size_t>(t_cap) * 8);
    const size_t o_head = take(n * 4), o_rid = take(n * 4), o_first = take(n * 4), o_last = take(n * 4), o_slot = take(n * 4),
                 o_prev = take(n * 4), o_miss = take(n * 4), o_islast = take(n * 4), o_lpos = take(n * 4), o_need = take(n * 4),
                 o_npos = take(n * 4), o_members = take(n * 4);
    const size_t o_glast = take(static_cast<size_t>(t_cap) * 4), o_gnruns = take(static_cast<size_t>(t_cap) * 4),
                 o_gstart = take(static_cast<size_t>(t_cap) * 4), o_gcursor = take(static_cast<size_t>(t_cap) * 4),
                 o_gvid = take(static_cast<size_t>(t_cap) * 4), o_gkeep = take(static_cast<size_t>(t_cap) * 4),
                 o_gnewpos = take(static_cast<size_t>(t_cap) * 4);
    const size_t o_touched = take(static_cast<size_t>(cap) * 4), o_uflag = take(static_cast<size_t>(cap) * 4), o_upos = take(static_cast<size_t>(cap) * 4),
                 o_fflag = take(static_cast<size_t>(cap) * 4), o_fpos = take(static_cast<size_t>(cap) * 4), o_free = take(static_cast<size_t>(cap) * 4);
    reserve_scratch(at);
    char* base = static_cast<char*>(scratch_);
    auto* pkey = reinterpret_cast<unsigned long long*>(base + o_pkey);
    auto* run_key = reinterpret_cast<unsigned long long*>(base + o_runkey);
    auto* runs_sorted = reinterpret_cast<IncRun*>(base + o_runs);
    auto* t_key = reinterpret_cast<unsigned long long*>(base + o_tkey);
    auto u32 = [&](size_t o) { return reinterpret_cast<unsigned int*>(base + o); };
    unsigned int *head = u32(o_head), *rid = u32(o_rid), *run_first = u32(o_first), *run_last = u32(o_last), *run_slot = u32(o_slot),
                 *miss = u32(o_miss), *islast = u32(o_islast), *lpos = u32(o_lpos), *needflag = u32(o_need), *npos = u32(o_npos),
                 *members = u32(o_members), *g_last = u32(o_glast), *g_nruns = u32(o_gnruns), *g_start = u32(o_gstart),
                 *g_cursor = u32(o_gcursor), *g_keep = u32(o_gkeep), *g_newpos = u32(o_gnewpos), *touched = u32(o_touched),
                 *uflag = u32(o_uflag), *upos = u32(o_upos), *fflag = u32(o_fflag), *fpos = u32(o_fpos), *free_list = u32(o_free);
    int* prev = reinterpret_cast<int*>(base + o_prev);
    int* g_vid = reinterpret_cast<int*>(base + o_gvid);
    unsigned int* order = order_[cur_];
    unsigned int* new_order = order_[cur_ ^ 1];
    const unsigned int gN = blocks_for(n), gT = blocks_for(t_cap), gC = blocks_for(cap);

    // ---- runs
    LR_LAUNCH(k_inc_keys, gN, kT, 0, stream, d_xyz, n, stride, view_.inv_voxel, pkey);
    LR_LAUNCH(k_inc_heads, gN, kT, 0, stream, pkey, n, head);
    exclusive_scan_u32(head, rid, n, ctr_ + kCtrRuns, stream);
    LR_LAUNCH(k_inc_runs, gN, kT, 0, stream, pkey, head, rid, n, run_key, run_first, run_last);
    // ---- groups
    LR_CUDA(cudaMemsetAsync(t_key, 0xFF, static_cast<size_t>(t_cap) * 8, stream));
    LR_CUDA(cudaMemsetAsync(g_last, 0, static_cast<size_t>(t_cap) * 4, stream));
    LR_CUDA(cudaMemsetAsync(g_nruns, 0, static_cast<size_t>(t_cap) * 4, stream));
    LR_CUDA(cudaMemsetAsync(g_cursor, 0, static_cast<size_t>(t_cap) * 4, stream));
    LR_CUDA(cudaMemsetAsync(g_keep, 0, static_cast<size_t>(t_cap) * 4, stream));
    LR_CUDA(cudaMemsetAsync(touched, 0, static_cast<size_t>(cap) * 4, stream));
    LR_CUDA(cudaMemsetAsync(ctr_ + kCtrGroups, 0, 2 * sizeof(unsigned int), stream));
    LR_LAUNCH(k_inc_group, gN, kT, 0, stream, run_key, ctr_, t_key, t_cap - 1, run_slot, g_last, g_nruns);
    LR_LAUNCH(k_inc_lookup, gT, kT, 0, stream, t_key, t_cap, slots_, cap_slots_ - 1, g_vid, touched, ctr_);
    exclusive_scan_u32(g_nruns, g_start, t_cap, nullptr, stream);
    LR_LAUNCH(k_inc_scatter, gN, kT, 0, stream, run_slot, ctr_, g_start, g_cursor, members);
    LR_LAUNCH(k_inc_sort_prev, gT, kT, 0, stream, t_key, t_cap, g_start, g_nruns, members, g_vid, rank_of_, ctr_, run_first, run_last,
              runs_sorted, prev);
    // ---- hit or miss: a warp per run
    LR_LAUNCH(k_inc_classify, blocks_for(n * 32), kT, 0, stream, prev, ctr_, C, miss);
    // ---- survivors, new order, records for the new keys
    LR_LAUNCH(k_inc_plan, 1, 1, 0, stream, ctr_, C);
    LR_LAUNCH(k_inc_old_flags, gC, kT, 0, stream, order, ctr_, touched, cap, uflag);
    exclusive_scan_u32(uflag, upos, cap, ctr_ + kCtrScratch, stream);
    LR_LAUNCH(k_inc_keep_old, gC, kT, 0, stream, order, ctr_, uflag, upos, cap, new_order, ent_key_);
    LR_LAUNCH(k_inc_last_flags, gN, kT, 0, stream, run_slot, g_last, ctr_, n, islast);
    exclusive_scan_u32(islast, lpos, n, ctr_ + kCtrScratch, stream);
    LR_LAUNCH(k_inc_keep_groups, gN, kT, 0, stream, run_slot, islast, lpos, ctr_, n, g_vid, g_keep, g_newpos, ent_key_, new_order, needflag);
    LR_LAUNCH(k_inc_free_flags, gC, kT, 0, stream, ent_key_, cap, fflag);
    exclusive_scan_u32(fflag, fpos, cap, ctr_ + kCtrScratch, stream);
    LR_LAUNCH(k_inc_free_list, gC, kT, 0, stream, fflag, fpos, cap, free_list);
    exclusive_scan_u32(needflag, npos, n, ctr_ + kCtrScratch, stream);
    LR_LAUNCH(k_inc_assign, gN, kT, 0, stream, run_slot, needflag, npos, n, free_list, run_key, g_vid, g_newpos, ent_key_, new_order);
    // ---- statistics of the touched survivors, then the table the alignment kernels probe
    LR_LAUNCH(k_inc_stats, gT, kT, 0, stream, t_key, t_cap, g_keep, g_vid, g_start, g_nruns, members, miss, runs_sorted, d_xyz, stride,
              voxels_, ent_cnt_);
    LR_LAUNCH(k_inc_clear_slots, blocks_for(cap_slots_), kT, 0, stream, slots_, cap_slots_);
    LR_LAUNCH(k_inc_publish, gC, kT, 0, stream, new_order, ctr_, ent_key_, ent_cnt_, cap, slots_, cap_slots_ - 1, rank_of_);
    cur_ ^= 1;
}

size_t DeviceIncNdtMap::size(cudaStream_t stream) const {
    if (!ctr_) return 0;
    unsigned int m = 0;
    LR_CUDA(cudaMemcpyAsync(&m, ctr_ + kCtrM, sizeof(m), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    return m;
}

void DeviceIncNdtMap::download(std::vector<int>& keys, std::vector<double>& mu, std::vector<double>& info, std::vector<int>& npts,
                               cudaStream_t stream) const {
    keys.clear(); mu.clear(); info.clear(); npts.clear();
    const size_t m = size(stream);
    if (m == 0) return;
    std::vector<NdtVoxel> hv(capacity_);
    std::vector<unsigned long long> hk(capacity_);
    std::vector<unsigned int> hc(capacity_), ho(m);
    LR_CUDA(cudaMemcpyAsync(hv.data(), voxels_, capacity_ * sizeof(NdtVoxel), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaMemcpyAsync(hk.data(), ent_key_, capacity_ * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaMemcpyAsync(hc.data(), ent_cnt_, capacity_ * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaMemcpyAsync(ho.data(), order_[cur_], m * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    struct Rec { int k[3]; unsigned int vid; };
    std::vector<Rec> recs;
    for (unsigned int vid : ho) {
        Rec r;
        ndt_unpack(hk[vid], r.k[0], r.k[1], r.k[2]);
        r.vid = vid;
        recs.push_back(r);
    }
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
        if (a.k[0] != b.k[0]) return a.k[0] < b.k[0];
        if (a.k[1] != b.k[1]) return a.k[1] < b.k[1];
        return a.k[2] < b.k[2];
    });
    for (const Rec& r : recs) {
        keys.insert(keys.end(), r.k, r.k + 3);
        mu.insert(mu.end(), hv[r.vid].mu, hv[r.vid].mu + 3);
        info.insert(info.end(), hv[r.vid].info, hv[r.vid].info + 9);
        npts.push_back(static_cast<int>(hc[r.vid]));
    }
}

}  // namespace locreg

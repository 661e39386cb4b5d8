// Device versions of the cloud pre-filters that sit directly in front of every SetInputTarget / ScanMatch call in the
// reference: pcl::removeNaNFromPointCloud (point_cloud_utils.h:13-20), pcl::CropBox (BoxFilter,
// box_filter.cpp:24-32) and pcl::VoxelGrid (VoxelFilter, voxel_filter.cpp:20-26; scan 1.0 m / map 0.5 m leaves at
// lio.cpp:236,248, loc.cpp:218).  PCL itself is not under /root/reference: these restate its published algorithms
// (PCL 1.8 filters/voxel_grid.hpp, crop_box.hpp, filter.hpp) - parity is against oracle/filters of this repo.
//   * keep-filters (NaN removal, crop box): flag -> exclusive scan -> scatter, order preserved;
//   * voxel grid: voxel index exactly as PCL computes it (float floor of p * inverse_leaf minus min_b), counting sort
//     over the dense index space, one thread per occupied voxel averaging ALL float words of its points in float32,
//     in point-index order (PCL sorts with an unstable std::sort, so its own summation order is unspecified), output
//     in ascending voxel index like PCL.
#include <algorithm>
#include <cfloat>
#include <vector>

#include "device_utils.cuh"
#include "filters.cuh"

namespace locreg {

__global__ void k_flag_finite(const unsigned char* __restrict__ raw, size_t n, size_t stride, unsigned int* flag) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + i * stride);
    flag[i] = finite3(p[0], p[1], p[2]) ? 1u : 0u;
}
// pcl::CropBox with the identity transform, negative = false: a point is kept iff min <= p <= max on every axis
// (crop_box.hpp: "if (pt.x < min[0] || ... || pt.x > max[0] ...) -> outside"); non-finite points are dropped.
__global__ void k_flag_box(const unsigned char* __restrict__ raw, size_t n, size_t stride, float3 lo, float3 hi, unsigned int* flag) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + i * stride);
    const bool out = !finite3(p[0], p[1], p[2]) || p[0] < lo.x || p[1] < lo.y || p[2] < lo.z || p[0] > hi.x || p[1] > hi.y || p[2] > hi.z;
    flag[i] = out ? 0u : 1u;
}
__global__ void k_compact(const unsigned char* __restrict__ raw, size_t n, size_t stride, const unsigned int* __restrict__ flag,
                          const unsigned int* __restrict__ pos, unsigned char* out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const unsigned int* s = reinterpret_cast<const unsigned int*>(raw + i * stride);
    unsigned int* d = reinterpret_cast<unsigned int*>(out + static_cast<size_t>(pos[i]) * stride);
    for (size_t w = 0; w < stride / 4; ++w) d[w] = s[w];
}

// order-preserving float <-> uint map for atomicMin / atomicMax on floats
__device__ __forceinline__ unsigned int f2ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float ord2f(unsigned int o) {
    const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
__global__ void k_minmax(const unsigned char* __restrict__ raw, size_t n, size_t stride, unsigned int* mm) {  // mm[0..2] min, [3..5] max
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    unsigned int lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    if (i < n) {
        const float* p = reinterpret_cast<const float*>(raw + i * stride);
        if (finite3(p[0], p[1], p[2]))
            for (int a = 0; a < 3; ++a) { lo[a] = f2ord(p[a]); hi[a] = lo[a]; }
    }
    for (int a = 0; a < 3; ++a) {
        const unsigned int l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
        if ((threadIdx.x & 31) == 0) { atomicMin(&mm[a], l); atomicMax(&mm[3 + a], h); }
    }
}
struct GridSpec {
    float inv_leaf;
    int min_b[3];
    int div_mul[3];
};
// int ijk = static_cast<int>(floor(p * inverse_leaf) - static_cast<float>(min_b))   (voxel_grid.hpp)
__device__ __forceinline__ unsigned int voxel_index(const GridSpec& g, const float* p) {
    const int i0 = static_cast<int>(floorf(__fmul_rn(p[0], g.inv_leaf)) - static_cast<float>(g.min_b[0]));
    const int i1 = static_cast<int>(floorf(__fmul_rn(p[1], g.inv_leaf)) - static_cast<float>(g.min_b[1]));
    const int i2 = static_cast<int>(floorf(__fmul_rn(p[2], g.inv_leaf)) - static_cast<float>(g.min_b[2]));
    return static_cast<unsigned int>(i0 * g.div_mul[0] + i1 * g.div_mul[1] + i2 * g.div_mul[2]);
}
// Voxel grid through a BITMAP of the index space (1 bit per voxel: 256 MB at PCL's own limit of 2^31 voxels, where dense
// per-voxel arrays would take tens of GB): occupied voxels are marked, a voxel's place in the output - PCL emits the
// voxels in ascending index - is the number of set bits below it (prefix sum over the bitmap's words + popcount), and
// everything else is sized by the number of points.
__global__ void k_vg_mark(const unsigned char* __restrict__ raw, size_t n, size_t stride, GridSpec g, unsigned int* pt_idx,
                          unsigned int* bitmap) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + i * stride);
    if (!finite3(p[0], p[1], p[2])) { pt_idx[i] = 0xFFFFFFFFu; return; }
    const unsigned int idx = voxel_index(g, p);
    pt_idx[i] = idx;
    atomicOr(&bitmap[idx >> 5], 1u << (idx & 31u));
}
__global__ void k_vg_popc(const unsigned int* __restrict__ bitmap, size_t words, unsigned int* wcount) {
    const size_t w = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w < words) wcount[w] = static_cast<unsigned int>(__popc(bitmap[w]));
}
// pt_idx[i] becomes the rank of the point's voxel among the occupied ones; vcount[rank]++
__global__ void k_vg_rank(size_t n, unsigned int* pt_idx, const unsigned int* __restrict__ bitmap, const unsigned int* __restrict__ wprefix,
                          unsigned int* vcount) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int idx = pt_idx[i];
    if (idx == 0xFFFFFFFFu) return;
    const unsigned int r = wprefix[idx >> 5] + static_cast<unsigned int>(__popc(bitmap[idx >> 5] & ((1u << (idx & 31u)) - 1u)));
    pt_idx[i] = r;
    atomicAdd(&vcount[r], 1u);
}
__global__ void k_vg_scatter(size_t n, const unsigned int* __restrict__ pt_rank, const unsigned int* __restrict__ start,
                             unsigned int* cursor, unsigned int* members) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || pt_rank[i] == 0xFFFFFFFFu) return;
    members[start[pt_rank[i]] + atomicAdd(&cursor[pt_rank[i]], 1u)] = static_cast<unsigned int>(i);
}
// one thread per occupied voxel (rank r < *n_out): float32 mean of every float word over its points in point order
__global__ void k_vg_centroid(const unsigned char* __restrict__ raw, size_t stride, const unsigned int* __restrict__ n_out,
                              const unsigned int* __restrict__ start, unsigned int* members, unsigned char* out) {
    const size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= *n_out) return;
    const unsigned int beg = start[r], cnt = start[r + 1] - beg;
    unsigned int* idx = members + beg;
    for (unsigned int a = 1; a < cnt; ++a) {  // point-index order: the atomics above arrive in any order
        const unsigned int v = idx[a];
        unsigned int b = a;
        while (b > 0 && idx[b - 1] > v) { idx[b] = idx[b - 1]; --b; }
        idx[b] = v;
    }
    const size_t words = stride / 4 < 8 ? stride / 4 : 8;
    float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned int j = 0; j < cnt; ++j) {
        const float* p = reinterpret_cast<const float*>(raw + static_cast<size_t>(idx[j]) * stride);
        for (size_t w = 0; w < words; ++w) sum[w] = __fadd_rn(sum[w], p[w]);
    }
    float* o = reinterpret_cast<float*>(out + r * stride);
    const float fn = static_cast<float>(cnt);
    for (size_t w = 0; w < words; ++w) o[w] = __fdiv_rn(sum[w], fn);
    for (size_t w = words; w < stride / 4; ++w) o[w] = 0.0f;
}

namespace {
struct Tmp {  // stream-ordered scratch
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit Tmp(cudaStream_t st) : s(st) {}
    template <class T> T* get(size_t count) {
        void* p = nullptr;
        LR_CUDA(cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), s));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    ~Tmp() { for (void* p : ptrs) cudaFreeAsync(p, s); }
};
size_t run_keep(const unsigned char* d_raw, size_t n, size_t stride, unsigned int* flag, unsigned char* d_out, cudaStream_t stream, Tmp& tmp) {
    unsigned int* pos = tmp.get<unsigned int>(n);
    unsigned int* total = tmp.get<unsigned int>(1);
    exclusive_scan_u32(flag, pos, n, total, stream);
    const unsigned int grid = static_cast<unsigned int>((n + 255) / 256);
    LR_LAUNCH(k_compact, grid, 256, 0, stream, d_raw, n, stride, flag, pos, d_out);
    unsigned int kept = 0;
    LR_CUDA(cudaMemcpyAsync(&kept, total, sizeof(kept), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    return kept;
}
}  // namespace

size_t filter_remove_nan(const unsigned char* d_raw, size_t n, size_t stride, unsigned char* d_out, cudaStream_t stream) {
    if (n == 0) return 0;
    Tmp tmp(stream);
    unsigned int* flag = tmp.get<unsigned int>(n);
    LR_LAUNCH(k_flag_finite, static_cast<unsigned int>((n + 255) / 256), 256, 0, stream, d_raw, n, stride, flag);
    return run_keep(d_raw, n, stride, flag, d_out, stream, tmp);
}
size_t filter_crop_box(const unsigned char* d_raw, size_t n, size_t stride, const float* min3, const float* max3, unsigned char* d_out,
                       cudaStream_t stream) {
    if (n == 0) return 0;
    Tmp tmp(stream);
    unsigned int* flag = tmp.get<unsigned int>(n);
    LR_LAUNCH(k_flag_box, static_cast<unsigned int>((n + 255) / 256), 256, 0, stream, d_raw, n, stride,
              make_float3(min3[0], min3[1], min3[2]), make_float3(max3[0], max3[1], max3[2]), flag);
    return run_keep(d_raw, n, stride, flag, d_out, stream, tmp);
}

size_t filter_voxel_grid(const unsigned char* d_raw, size_t n, size_t stride, float leaf, unsigned char* d_out, cudaStream_t stream) {
    if (n == 0) return 0;
    Tmp tmp(stream);
    const unsigned int gridN = static_cast<unsigned int>((n + 255) / 256);
    unsigned int* mm = tmp.get<unsigned int>(6);
    const unsigned int init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    unsigned int h_mm[6];
    LR_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    LR_LAUNCH(k_minmax, gridN, 256, 0, stream, d_raw, n, stride, mm);
    LR_CUDA(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    if (h_mm[0] == 0xFFFFFFFFu) return 0;  // no finite point
    GridSpec g;
    g.inv_leaf = 1.0f / leaf;
    long long div[3];
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(h_mm[a]), hi = ord2f(h_mm[3 + a]);
        g.min_b[a] = static_cast<int>(floorf(lo * g.inv_leaf));
        div[a] = static_cast<long long>(static_cast<int>(floorf(hi * g.inv_leaf))) - g.min_b[a] + 1;
    }
    // PCL gives up when the voxel index would overflow an int ("Leaf size is too small for the input dataset. Integer
    // indices would overflow.") and returns the cloud UNFILTERED (voxel_grid.hpp: output = *input): one far outlier in a
    // key frame must not stop Lio
    if (div[0] * div[1] * div[2] > 0x7fffffffll) {
        LR_CUDA(cudaMemcpyAsync(d_out, d_raw, n * stride, cudaMemcpyDeviceToDevice, stream));
        LR_CUDA(cudaStreamSynchronize(stream));
        return n;
    }
    g.div_mul[0] = 1; g.div_mul[1] = static_cast<int>(div[0]); g.div_mul[2] = static_cast<int>(div[0] * div[1]);
    const size_t cells = static_cast<size_t>(div[0] * div[1] * div[2]);
    const size_t words = (cells + 31) / 32;
    unsigned int* pt_idx = tmp.get<unsigned int>(n);
    unsigned int* bitmap = tmp.get<unsigned int>(words);
    unsigned int* wprefix = tmp.get<unsigned int>(words);
    unsigned int* vstart = tmp.get<unsigned int>(n + 1);  // occupied voxels <= points
    unsigned int* cursor = tmp.get<unsigned int>(n);
    unsigned int* members = tmp.get<unsigned int>(n);
    unsigned int* total = tmp.get<unsigned int>(1);
    LR_CUDA(cudaMemsetAsync(bitmap, 0, words * sizeof(unsigned int), stream));
    LR_CUDA(cudaMemsetAsync(vstart, 0, (n + 1) * sizeof(unsigned int), stream));
    LR_CUDA(cudaMemsetAsync(cursor, 0, n * sizeof(unsigned int), stream));
    LR_LAUNCH(k_vg_mark, gridN, 256, 0, stream, d_raw, n, stride, g, pt_idx, bitmap);
    LR_LAUNCH(k_vg_popc, static_cast<unsigned int>((words + 255) / 256), 256, 0, stream, bitmap, words, wprefix);
    exclusive_scan_u32(wprefix, wprefix, words, total, stream);
    LR_LAUNCH(k_vg_rank, gridN, 256, 0, stream, n, pt_idx, bitmap, wprefix, vstart);
    exclusive_scan_u32(vstart, vstart, n + 1, nullptr, stream);
    LR_LAUNCH(k_vg_scatter, gridN, 256, 0, stream, n, pt_idx, vstart, cursor, members);
    LR_LAUNCH(k_vg_centroid, gridN, 256, 0, stream, d_raw, stride, total, vstart, members, d_out);
    unsigned int n_out = 0;
    LR_CUDA(cudaMemcpyAsync(&n_out, total, sizeof(n_out), cudaMemcpyDeviceToHost, stream));
    LR_CUDA(cudaStreamSynchronize(stream));
    return n_out;
}

}  // namespace locreg

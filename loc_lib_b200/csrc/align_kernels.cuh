// Fused registration kernels (K4/K6/K7/K8/K9 of SURVEY.md §2.3).  NDT's per-point work is seven independent hash
// probes, light enough to fuse with the reduction and the Gauss-Newton update; ICP runs the three-kernel
// pipeline of icp_pipeline.cuh instead.  The "Problem" policy supplies the per-point body and the update.
//
//   k_eval            one H/B evaluation over a scan (grid-stride), per-block partial sums
//   k_finalize        fixed-order sum of the partials + Gauss-Newton update (1 block)
//   k_align_persist   whole ScanMatch in ONE cooperative launch: evaluate, grid barrier, every block
//                     redundantly reduces + solves (bitwise identical), next iteration — no host round-trip
//   k_align_batch     one CTA per scan / pose hypothesis: the entire Gauss-Newton loop runs inside the CTA
//                     with __syncthreads only (batch offline mapping and global relocalisation)
//   k_transform       pcl::transformPointCloud
#pragma once
#include <cooperative_groups.h>

#include "device_utils.cuh"
#include "icp_pipeline.cuh"
#include "ndt_point.cuh"

#ifndef LR_MIN_BLOCKS
#define LR_MIN_BLOCKS 2  // resident 256-thread CTAs per SM the registration kernels are compiled for
#endif

namespace locreg {
namespace cg = cooperative_groups;

// ---- problem policies -------------------------------------------------------------------------
// A warp processes a chunk of up to 32 consecutive source points: lane j owns point base + j.
// chunk(): count <= 32 points starting at src[base]; gate / nn_out (debug probe) may be nullptr.
struct NdtProblem {
    NdtMapView map;
    NdtParams prm;
    __device__ __forceinline__ void chunk(const Pose& T, const float4* __restrict__ src, unsigned int base,
                                          unsigned int count, SmemAccum& acc, unsigned char* gate, int* nn_out) const {
        (void)nn_out;
        const unsigned int lane = threadIdx.x & 31;
        if (lane < count) {
            const float4 sp = src[base + lane];
            const unsigned char h = ndt_point(map, prm, T, sp.x, sp.y, sp.z, acc);
            if (gate) gate[base + lane] = h;
        }
    }
    __device__ __forceinline__ int update(const double* acc30, Pose& T) const {
        return ndt_gn_update(acc30, static_cast<unsigned int>(acc30[28]), prm, T);
    }
    __device__ __forceinline__ int max_iteration() const { return prm.max_iteration; }
};

// Incremental NDT (AlignIncNdt): same voxel table layout, information-weighted residuals.
struct IncNdtProblem {
    NdtMapView map;
    NdtParams prm;
    __device__ __forceinline__ void chunk(const Pose& T, const float4* __restrict__ src, unsigned int base,
                                          unsigned int count, SmemAccum& acc, unsigned char* gate, int* nn_out) const {
        (void)nn_out;
        const unsigned int lane = threadIdx.x & 31;
        if (lane < count) {
            const float4 sp = src[base + lane];
            const unsigned char h = inc_ndt_point(map, prm, T, sp.x, sp.y, sp.z, acc);
            if (gate) gate[base + lane] = h;
        }
    }
    __device__ __forceinline__ int update(const double* acc30, Pose& T) const {
        return inc_ndt_gn_update(acc30, static_cast<unsigned int>(acc30[28]), prm, T);
    }
    __device__ __forceinline__ int max_iteration() const { return prm.max_iteration; }
};

// Walks the chunks of a scan assigned to this warp: chunk c covers points [c*ppw, min(n, (c+1)*ppw)).
template <class Problem>
__device__ __forceinline__ void eval_scan(const Problem& pb, const Pose& T, const float4* __restrict__ src, unsigned int n,
                                          unsigned int ppw, unsigned int first_warp, unsigned int warp_stride, SmemAccum& acc,
                                          unsigned char* gate, int* nn_out) {
    const unsigned int n_chunks = (n + ppw - 1) / ppw;
    for (unsigned int c = first_warp; c < n_chunks; c += warp_stride) {
        const unsigned int base = c * ppw;
        const unsigned int count = n - base < ppw ? n - base : ppw;
        pb.chunk(T, src, base, count, acc, gate, nn_out);
    }
}

// ---- single evaluation (compute_hb, debug probe, multi-launch loop) -----------------------------
template <class Problem>
__global__ void __launch_bounds__(256, LR_MIN_BLOCKS) k_eval(Problem pb, const float4* __restrict__ src, unsigned int n, unsigned int ppw,
                                              const AlignState* __restrict__ state, double* __restrict__ partials,
                                              unsigned char* gate, int* nn_out) {
    extern __shared__ double dyn_acc[];
    __shared__ Pose T;
    __shared__ double red[8 * kPartialDoubles];
    if (state->stop) return;
    if (threadIdx.x == 0) pose_load(T, state->pose);
    __syncthreads();
    SmemAccum acc = smem_accum_init(dyn_acc);
    eval_scan(pb, T, src, n, ppw, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5), acc, gate,
              nn_out);
    block_reduce_accum(acc, dyn_acc, red, partials + static_cast<size_t>(blockIdx.x) * kPartialDoubles);
}

// Sum of `nblocks` partial rows in a fixed order: 8 groups of rows summed sequentially, then the 8
// group sums.  Needs 256 threads and 8*32 doubles of shared memory; result in out30[0..29] (shared).
__device__ __forceinline__ void reduce_partials(const double* __restrict__ partials, unsigned int nblocks, double* sh8x32,
                                                double* out30) {
    const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;  // blockDim.x == 256
    double s = 0;
    if (col < 30)
        for (unsigned int b = grp; b < nblocks; b += 8) s += __ldcg(partials + static_cast<size_t>(b) * kPartialDoubles + col);
    sh8x32[grp * 32 + col] = s;
    __syncthreads();
    if (threadIdx.x < 30) {
        double t = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += sh8x32[g * 32 + threadIdx.x];
        out30[threadIdx.x] = t;
    }
    __syncthreads();
}

template <class Problem>
__global__ void __launch_bounds__(256) k_finalize(Problem pb, const double* __restrict__ partials, unsigned int nblocks,
                                                  AlignState* state, int do_update, double* acc_out) {
    __shared__ double sh[8 * 32];
    __shared__ double acc30[32];
    if (state->stop) return;
    reduce_partials(partials, nblocks, sh, acc30);
    if (threadIdx.x == 0) {
        result_from_acc(state->res, acc30);
        if (acc_out)
            for (int i = 0; i < 30; ++i) acc_out[i] = acc30[i];
        if (do_update) {
            Pose T;
            pose_load(T, state->pose);
            state->res.iters += 1;
            const int outcome = pb.update(acc30, T);
            if (apply_outcome(outcome, state->res)) state->stop = 1;
            pose_store(T, state->pose);
        }
    }
}

// ---- persistent cooperative ScanMatch ----------------------------------------------------------
// partials: 2 * gridDim.x rows (double-buffered by iteration parity so a fast block cannot overwrite
// rows a slow block is still summing).
template <class Problem>
__global__ void __launch_bounds__(256, LR_MIN_BLOCKS) k_align_persist(Problem pb, const float4* __restrict__ src, unsigned int n, unsigned int ppw,
                                                       AlignState* state, double* partials, int final_eval) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double dyn_acc[];
    __shared__ Pose T;
    __shared__ double red[8 * kPartialDoubles];
    __shared__ double acc30[32];
    __shared__ DevResult res;
    __shared__ int stop;
    if (threadIdx.x == 0) {
        pose_load(T, state->pose);
        res = state->res;
        stop = 0;
    }
    __syncthreads();
    const int max_it = pb.max_iteration();
    // Loop trips 0..max_it-1 are the reference's Gauss-Newton iterations; with final_eval one more
    // evaluation (no update) runs at the final pose.  Every block takes the same branches because
    // every block reduces the same rows in the same order.
    int it = 0;
    bool final_pass = max_it <= 0;
    while (!(final_pass && !final_eval)) {
        SmemAccum acc = smem_accum_init(dyn_acc);
        eval_scan(pb, T, src, n, ppw, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5), acc,
                  nullptr, nullptr);
        double* rows = partials + static_cast<size_t>(it & 1) * gridDim.x * kPartialDoubles;
        block_reduce_accum(acc, dyn_acc, red, rows + static_cast<size_t>(blockIdx.x) * kPartialDoubles);
        grid.sync();
        reduce_partials(rows, gridDim.x, red, acc30);
        if (threadIdx.x == 0) {
            result_from_acc(res, acc30);
            if (!final_pass) {
                res.iters += 1;
                if (apply_outcome(pb.update(acc30, T), res)) stop = 1;
            }
        }
        __syncthreads();
        if (final_pass) break;
        ++it;
        if (!stop && it < max_it) continue;
        if (final_eval && res.pose_written) { final_pass = true; continue; }
        break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pose_store(T, state->pose);
        state->res = res;
    }
}

// ---- batch: one CTA per scan / hypothesis -------------------------------------------------------
// offsets == nullptr: every item registers the same scan src[0..n_single) (relocalisation).
template <class Problem>
__global__ void __launch_bounds__(256, LR_MIN_BLOCKS) k_align_batch(Problem pb, const float4* __restrict__ src,
                                                     const long long* __restrict__ offsets, unsigned int n_single,
                                                     const double* __restrict__ poses_in, double* poses_out,
                                                     DevResult* results, unsigned int S, unsigned int* work_counter,
                                                     int final_eval, int zero_t) {
    extern __shared__ double dyn_acc[];
    __shared__ Pose T;
    __shared__ double red[8 * kPartialDoubles];
    __shared__ double acc30[32];
    __shared__ DevResult res;
    __shared__ int stop;
    __shared__ unsigned int item;
    const int max_it = pb.max_iteration();
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) item = atomicAdd(work_counter, 1u);
        __syncthreads();
        const unsigned int s = item;
        if (s >= S) break;
        const long long beg = offsets ? offsets[s] : 0;
        const unsigned int n = offsets ? static_cast<unsigned int>(offsets[s + 1] - beg) : n_single;
        const float4* pts = src + beg;
        if (threadIdx.x == 0) {
            double p7[7];
            for (int i = 0; i < 7; ++i) p7[i] = (zero_t && i >= 4) ? 0.0 : poses_in[static_cast<size_t>(s) * 7 + i];
            pose_load(T, p7);
            res = DevResult{0, 0, 0, 0, 0, 0, 0.0, 1, 0};
            stop = 0;
        }
        __syncthreads();
        int it = 0;
        bool final_pass = max_it <= 0;
        while (!(final_pass && !final_eval)) {
            SmemAccum acc = smem_accum_init(dyn_acc);
            eval_scan(pb, T, pts, n, 32u, threadIdx.x >> 5, blockDim.x >> 5, acc, nullptr, nullptr);
            block_reduce_accum(acc, dyn_acc, red, acc30);
            if (threadIdx.x == 0) {
                result_from_acc(res, acc30);
                if (!final_pass) {
                    res.iters += 1;
                    if (apply_outcome(pb.update(acc30, T), res)) stop = 1;
                }
            }
            __syncthreads();
            if (final_pass) break;
            ++it;
            if (!stop && it < max_it) continue;
            if (final_eval && res.pose_written) { final_pass = true; continue; }
            break;
        }
        if (threadIdx.x == 0) {
            if (res.pose_written) pose_store(T, poses_out + static_cast<size_t>(s) * 7);
            if (results) results[s] = res;
        }
    }
}

// ---- relocalisation score + argmin --------------------------------------------------------------
// key = (float32 score bits << 32) | index; score = sum_sq / n_inlier, +inf if degenerate / no inlier.
// The index in the key is index_base + s * index_stride (a rank's share of a strided deal of the hypotheses).
__global__ void k_score_argmin(const DevResult* __restrict__ results, unsigned int S, unsigned int index_base, unsigned int index_stride,
                               double* scores, unsigned long long* best_key) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = ~0ull;
    if (s < S) {
        const DevResult r = results[s];
        double sc = INFINITY;
        if (!r.degenerate && r.n_inlier > 0 && r.pose_written) sc = r.sum_sq_res / static_cast<double>(r.n_inlier);
        if (scores) scores[s] = sc;
        const float f = static_cast<float>(sc);
        key = (static_cast<unsigned long long>(__float_as_uint(f)) << 32) | static_cast<unsigned long long>(index_base + s * index_stride);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xffffffffu, key, off);
        key = o < key ? o : key;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(best_key, key);
}

// ---- pcl::transformPointCloud (icp_registration.cpp:241) ------------------------------------------
// out is a byte copy of the raw input cloud (same stride); only x,y,z of finite points are rewritten:
// x' = m00*x + m01*y + m02*z + m03 in float32, left to right, no FMA (PCL 1.8 transforms.hpp).
__global__ void k_transform(const unsigned char* __restrict__ raw_in, unsigned char* raw_out, size_t n, size_t stride,
                            const double* __restrict__ pose7) {
    __shared__ float m[12];
    if (threadIdx.x == 0) {
        Pose T;
        pose_load(T, pose7);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) m[r * 4 + c] = static_cast<float>(T.R[r * 3 + c]);
        }
        m[3] = static_cast<float>(T.tx); m[7] = static_cast<float>(T.ty); m[11] = static_cast<float>(T.tz);
    }
    __syncthreads();
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(raw_in + i * stride);
        float* o = reinterpret_cast<float*>(raw_out + i * stride);
        const float x = p[0], y = p[1], z = p[2];
        const size_t words = stride / 4;
        for (size_t w = 3; w < words; ++w) o[w] = p[w];
        if (finite3(x, y, z)) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
                o[r] = LR_FADD(LR_FADD(LR_FADD(LR_FMUL(m[r * 4], x), LR_FMUL(m[r * 4 + 1], y)), LR_FMUL(m[r * 4 + 2], z)), m[r * 4 + 3]);
        } else {
            o[0] = x; o[1] = y; o[2] = z;
        }
    }
}

// pcl::transformPointCloud<PointT, double> (Lio's key frames, lio.cpp:243,278: pose.matrix() is a Matrix4d and is not cast):
// m00*x + m01*y + m02*z + m03 in DOUBLE, left to right, no FMA, one cast to float.
__global__ void k_transform_d(const unsigned char* __restrict__ raw_in, unsigned char* raw_out, size_t n, size_t stride,
                              const double* __restrict__ pose7) {
    __shared__ double m[12];
    if (threadIdx.x == 0) {
        Pose T;
        pose_load(T, pose7);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) m[r * 4 + c] = T.R[r * 3 + c];
        }
        m[3] = T.tx; m[7] = T.ty; m[11] = T.tz;
    }
    __syncthreads();
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(raw_in + i * stride);
        float* o = reinterpret_cast<float*>(raw_out + i * stride);
        const float x = p[0], y = p[1], z = p[2];
        const size_t words = stride / 4;
        for (size_t w = 3; w < words; ++w) o[w] = p[w];
        if (finite3(x, y, z)) {
            const double xd = x, yd = y, zd = z;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                o[r] = static_cast<float>(LR_DADD(LR_DADD(LR_DADD(LR_DMUL(m[r * 4], xd), LR_DMUL(m[r * 4 + 1], yd)), LR_DMUL(m[r * 4 + 2], zd)), m[r * 4 + 3]));
        } else {
            o[0] = x; o[1] = y; o[2] = z;
        }
    }
}

// raw strided cloud -> float4 (x, y, z, 0)
__global__ void k_pack_float4(const unsigned char* __restrict__ raw, size_t n, size_t stride, float4* out) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(raw + i * stride);
        out[i] = make_float4(p[0], p[1], p[2], 0.0f);
    }
}

}  // namespace locreg

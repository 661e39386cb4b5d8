// Device-only helpers: error handling, exclusive scan, block reduction of the GN accumulator.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "common.cuh"

namespace locreg {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define LR_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            throw ::locreg::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                      ":" + std::to_string(__LINE__) + ")");                           \
    } while (0)

// Launch counter so callers can report how many of our kernels ran (bench.py "gpu_launches").
extern thread_local long long g_launch_count;
#define LR_LAUNCH(kernel, grid, block, smem, stream, ...)  \
    do {                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ++::locreg::g_launch_count;                        \
        LR_CUDA(cudaGetLastError());                       \
    } while (0)

// Exclusive prefix sum of n uint32 (out may alias in).  If total != nullptr, *total receives the sum.
// Three-phase reduce/scan/propagate with 4096-element tiles, recursing on the tile sums.
void exclusive_scan_u32(const unsigned int* in, unsigned int* out, size_t n, unsigned int* total, cudaStream_t stream);

constexpr int kPartialDoubles = 32;  // 28 accumulator doubles + n_eff + n_inl (+2 pad): one 256 B row per block

// Per-thread Gauss-Newton accumulators live in dynamic shared memory: row i (i < kAccDoubles) holds element i of
// every thread of the block, base[i * blockDim.x + threadIdx.x].
constexpr int kBlockThreads = 256;
constexpr size_t kAccSmemBytes = static_cast<size_t>(kAccDoubles) * kBlockThreads * sizeof(double);

__device__ __forceinline__ SmemAccum smem_accum_init(double* dyn) {
    SmemAccum a;
    a.base = dyn + threadIdx.x;
    a.stride = kBlockThreads;
    a.n_eff = 0;
    a.n_inl = 0;
#pragma unroll
    for (int i = 0; i < kAccDoubles; ++i) a.base[i * kBlockThreads] = 0.0;
    return a;
}

// Sum of the block's accumulators -> out[0..29] (28 sums, n_eff, n_inl).  Warp w reduces rows w, w+8, ...: lanes
// read 32 consecutive threads' values (conflict free), shuffle-reduce, and walk the 8 column groups in order, so
// the result is reproducible for a fixed launch shape.  `red` needs 8 * 2 doubles for the counters.
__device__ __forceinline__ void block_reduce_accum(const SmemAccum& a, const double* dyn, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double ce = static_cast<double>(a.n_eff), ci = static_cast<double>(a.n_inl);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ce += __shfl_down_sync(0xffffffffu, ce, off);
        ci += __shfl_down_sync(0xffffffffu, ci, off);
    }
    if (lane == 0) { red[warp * 2] = ce; red[warp * 2 + 1] = ci; }
    __syncthreads();  // all threads' accumulators are final
    for (int row = warp; row < kAccDoubles; row += kBlockThreads / 32) {
        double s = 0;
#pragma unroll
        for (int g = 0; g < kBlockThreads / 32; ++g) {
            double v = dyn[row * kBlockThreads + g * 32 + lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            s += v;
        }
        if (lane == 0) out[row] = s;
    }
    if (threadIdx.x < 2) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) s += red[w * 2 + threadIdx.x];
        out[28 + threadIdx.x] = s;
    }
    __syncthreads();
}

}  // namespace locreg

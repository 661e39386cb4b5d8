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

// Sum an Accum over the block; the result lands in out[0..29] of thread 0's view (shared memory `red`
// must hold (blockDim.x/32) * kPartialDoubles doubles).  Fixed order: lanes by shuffle tree, warps
// sequentially, so results are reproducible for a fixed launch shape.
__device__ __forceinline__ void block_reduce_accum(const Accum& a, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double v[30];
#pragma unroll
    for (int i = 0; i < kAccDoubles; ++i) v[i] = a.v[i];
    v[28] = static_cast<double>(a.n_eff);
    v[29] = static_cast<double>(a.n_inl);
#pragma unroll
    for (int i = 0; i < 30; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 30; ++i) red[warp * kPartialDoubles + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < 30) {
        double s = 0;
        for (int w = 0; w < nwarps; ++w) s += red[w * kPartialDoubles + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

}  // namespace locreg

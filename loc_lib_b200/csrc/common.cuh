// Shared host/device helpers for the locreg kernels (sm_100a).
//
// Every per-point routine is written as an LR_HD inline function so the same source compiles
//   * under nvcc into the sm_100a kernels (the product), and
//   * under g++ into tests/hostsim (a serial, test-only harness that lets the CPU test-suite
//     check the kernel logic against the oracle in a container without a GPU).
// The hostsim build is never linked into liblocreg.so and is not reachable from the C ABI.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LR_HD __host__ __device__ __forceinline__
#define LR_HD_NOINLINE __host__ __device__ __noinline__
#define LR_D __device__ __forceinline__
#else
#define LR_HD inline
#define LR_HD_NOINLINE inline
#define LR_D inline
#endif

#if defined(__CUDA_ARCH__)
// float32 ops that must not be contracted into FMAs: the NN contract is bit-exact on
// dis2 = dx*dx + (dy*dy + dz*dz) as the reference's SSE2 build evaluates it (kdtree.h:94).
#define LR_FMUL(a, b) __fmul_rn((a), (b))
#define LR_FADD(a, b) __fadd_rn((a), (b))
#define LR_FSUB(a, b) __fsub_rn((a), (b))
#define LR_DMUL(a, b) __dmul_rn((a), (b))
#define LR_DADD(a, b) __dadd_rn((a), (b))
#define LR_DSUB(a, b) __dsub_rn((a), (b))
#else
#define LR_FMUL(a, b) ((a) * (b))
#define LR_FADD(a, b) ((a) + (b))
#define LR_FSUB(a, b) ((a) - (b))
#define LR_DMUL(a, b) ((a) * (b))
#define LR_DADD(a, b) ((a) + (b))
#define LR_DSUB(a, b) ((a) - (b))
#endif

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct int3 { int x, y, z; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace locreg {

LR_HD float int_as_float(int v) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(v);
#else
    union { int i; float f; } u; u.i = v; return u.f;
#endif
}
LR_HD int float_as_int(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(v);
#else
    union { int i; float f; } u; u.f = v; return u.i;
#endif
}
LR_HD int popc64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
LR_HD int ffs64(unsigned long long v) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
    return __ffsll(static_cast<long long>(v));
#else
    return __builtin_ffsll(static_cast<long long>(v));
#endif
}
LR_HD bool finite3(float x, float y, float z) {
    // true iff all three are finite (x - x == 0 fails for inf and nan)
    return (x - x == 0.0f) && (y - y == 0.0f) && (z - z == 0.0f);
}

// Per-scan Gauss-Newton accumulator: upper triangle of H (21), B (6), sum of squared gated
// residuals, and the two counters of the reference loops (effective_num, inliers).
constexpr int kAccDoubles = 28;  // 21 + 6 + 1
struct Accum {  // register / host flavour
    double v[kAccDoubles];
    unsigned int n_eff;
    unsigned int n_inl;
    LR_HD void add(int i, double x) { v[i] += x; }
    LR_HD void inc_eff(unsigned int c = 1u) { n_eff += c; }
    LR_HD void inc_inl(unsigned int c = 1u) { n_inl += c; }
    // one residual row: H += J^T J (upper triangle), B += -J^T r, sum of squares += r^2
    LR_HD void row(const double (&J)[6], double r) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) v[k++] += J[i] * J[j];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) v[21 + i] += -J[i] * r;
        v[27] += r * r;
    }
};
LR_HD void accum_zero(Accum& a) {
#pragma unroll
    for (int i = 0; i < kAccDoubles; ++i) a.v[i] = 0.0;
    a.n_eff = 0;
    a.n_inl = 0;
}
// Shared-memory flavour used by the kernels: element i of thread t lives at base[i * stride + t], so the 28
// running sums cost no registers while the k-NN search is in flight and every access is bank-conflict free.
struct SmemAccum {
    double* base;  // already offset by the thread index
    unsigned int stride;
    unsigned int n_eff;
    unsigned int n_inl;
    LR_HD void add(int i, double x) { base[i * stride] += x; }
    LR_HD void inc_eff(unsigned int c = 1u) { n_eff += c; }
    LR_HD void inc_inl(unsigned int c = 1u) { n_inl += c; }
};
// index of H(r,c), r <= c, in the packed upper triangle (row-major)
LR_HD constexpr int hidx(int r, int c) { return r * 6 - (r * (r - 1)) / 2 + (c - r); }

}  // namespace locreg

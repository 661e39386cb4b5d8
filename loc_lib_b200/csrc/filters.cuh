// Device cloud pre-filters (filters.cu): raw strided clouds in, compacted / downsampled raw clouds out (same stride).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace locreg {
// each returns the number of points written to d_out (capacity: n points)
size_t filter_remove_nan(const unsigned char* d_raw, size_t n, size_t stride, unsigned char* d_out, cudaStream_t stream);
size_t filter_crop_box(const unsigned char* d_raw, size_t n, size_t stride, const float* min3, const float* max3, unsigned char* d_out,
                       cudaStream_t stream);
size_t filter_voxel_grid(const unsigned char* d_raw, size_t n, size_t stride, float leaf, unsigned char* d_out, cudaStream_t stream);
}  // namespace locreg

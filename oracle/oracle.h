/* ORACLE C API (test infrastructure, NOT product code).
 *
 * CPU restatement of maotian123/loc_lib's scan-to-map registration inner loop,
 * used only as the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py cpu_baseline / --impl reference).  Each entry point names the reference
 * function it restates.  Pose layout everywhere: 7 doubles [qx qy qz qw tx ty tz]
 * (= Sophus::SE3d::data()).  H is 6x6 column-major, B is 6x1, f64.
 * Parity status: unpinned by the reference (it ships no golden vectors); pinned by
 * analytic KATs + numpy cross-checks in tests/.
 */
#ifndef LOCREG_ORACLE_H
#define LOCREG_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_ICP_P2P = 0, ORACLE_ICP_P2LINE = 1, ORACLE_ICP_P2PLANE = 2 }; /* IcpMethod, icp_registration.hpp:15-20 */
enum {
    ORACLE_NN_LITERAL_ANN = 0, /* literal tree search, approximate_=true alpha=0.1 (what the reference runs, quirk Q1) */
    ORACLE_NN_LITERAL_EXACT = 1, /* literal tree search, approximate_=false (kdtree.cpp:227-235) */
    ORACLE_NN_EXACT_TIEBREAK = 2, /* exact k-NN over the tree's leaves, total order (dis2_f32, index): the GPU parity target */
    ORACLE_NN_BRUTE_FORCE = 3 /* BfnnRegistration semantics (bfnn.cpp:24-50) with the same total order */
};

typedef struct oracle_icp_options { /* IcpOptions, icp_registration.hpp:22-39 */
    int32_t max_iteration;
    double max_nn_distance;
    double max_plane_distance;
    double max_line_distance;
    int32_t min_effective_pts;
    double eps;
    int32_t method;
    int32_t nn_mode;
    int32_t skip_nonfinite; /* 1: skip non-finite source points in every method (the C-ABI's documented deviation D1) */
} oracle_icp_options;

typedef struct oracle_ndt_options { /* NdtOptions, ndt_registration.hpp:27-42 */
    int32_t max_iteration;
    double voxel_size;
    int32_t min_effective_pts;
    int32_t min_pts_in_voxel;
    double eps;
    double res_outlier_th;
    int32_t nearby6; /* 0 = CENTER, 1 = NEARBY6 */
    int32_t skip_nonfinite;
} oracle_ndt_options;

typedef struct oracle_result {
    int32_t iters;        /* GN iterations executed (loop trips) */
    int32_t updates;      /* iterations whose H/B evaluation succeeded and updated the pose */
    int32_t converged;    /* 1 if ||dx|| < eps broke the loop */
    int32_t degenerate;   /* last H/B evaluation failed (too few effective points or det(H)==0) */
    int64_t n_effective;  /* effective_num of the last evaluation */
    int64_t n_inlier;     /* residuals that entered H/B in the last evaluation */
    double sum_sq_res;    /* sum of squared gated residuals of the last evaluation */
    int32_t pose_written; /* 0 only on direct-NDT's early return (ndt_registration.cpp:435-436) */
    int32_t pad_;
} oracle_result;

void oracle_icp_default_options(oracle_icp_options* o);
void oracle_ndt_default_options(oracle_ndt_options* o);

/* ---- k-NN (KdTree, kdtree.cpp; BfnnRegistration, bfnn.cpp) ---- */
typedef struct oracle_icp oracle_icp;
oracle_icp* oracle_icp_create(const oracle_icp_options* o);
void oracle_icp_destroy(oracle_icp* h);
/* IcpRegistration::SetInputTarget (icp_registration.cpp:9-29) -> KdTree::BuildTree (kdtree.cpp:10-31) */
int oracle_icp_set_target(oracle_icp* h, const float* xyz, size_t n, size_t stride_bytes);
size_t oracle_icp_tree_leaves(const oracle_icp* h);
/* KdtreeRegistration::FindNearstPoints (kdtree.cpp:272-283); idx_out is nq*k, -1 padded */
int oracle_icp_knn(oracle_icp* h, const float* q_xyz, size_t nq, size_t stride_bytes, int k, int nn_mode, int32_t* idx_out);
/* IcpRegistration::CaculateMatrixHAndB (icp_registration.cpp:31-55).  Optional per-point outputs:
 * gate[i]: 0 skipped before fit (non-finite / <k neighbours / P2P too far), 1 fit failed,
 *          2 fit ok but residual gated out, 3 inlier.  nn_out: n*k indices (-1 padded). Returns 1 if the
 * reference's bool result is true, 0 if false. */
int oracle_icp_compute_hb(oracle_icp* h, const float* src, size_t n, size_t stride_bytes, const double* pose7,
                          double* H36, double* B6, oracle_result* res, uint8_t* gate, int32_t* nn_out);
/* IcpRegistration::ScanMatch (icp_registration.cpp:216-244).  out_xyz (same stride, may be NULL) receives
 * pcl::transformPointCloud's result; pose_trace (may be NULL) receives (max_iteration+1)*7 doubles. */
int oracle_icp_align(oracle_icp* h, const float* src, size_t n, size_t stride_bytes, const double* pose_in,
                     double* pose_out, float* out_xyz, oracle_result* res, double* pose_trace);

/* S independent ScanMatch calls (scan s = points [offsets[s], offsets[s+1]) of srcs) on `threads` host threads
 * (<= 0: all hardware threads); each call is the single-threaded reference loop.  Returns the threads used. */
int oracle_icp_align_batch(oracle_icp* h, const float* srcs, const int64_t* offsets, size_t stride_bytes,
                           const double* poses_in, size_t S, double* poses_out, oracle_result* results, int threads);

/* math::FitPlane (math_utils.h:112-136): pts = n*3 doubles; returns 1 on success */
int oracle_fit_plane(const double* pts, int n, double* coeffs4, double eps);
/* brute-force k-NN without a handle (for kd-tree cross checks) */
int oracle_bfnn(const float* map_xyz, size_t n, size_t map_stride, const float* q_xyz, size_t nq, size_t q_stride,
                int k, int32_t* idx_out);

/* ---- direct NDT (ndt_registration.cpp:87-148, 374-464) ---- */
typedef struct oracle_ndt oracle_ndt;
oracle_ndt* oracle_ndt_create(const oracle_ndt_options* o);
void oracle_ndt_destroy(oracle_ndt* h);
int oracle_ndt_set_target(oracle_ndt* h, const float* xyz, size_t n, size_t stride_bytes);
size_t oracle_ndt_num_voxels(const oracle_ndt* h);
/* voxels sorted by (kx,ky,kz): keys nv*3 int32, mu nv*3, info nv*9 (row-major), npts nv */
int oracle_ndt_get_voxels(const oracle_ndt* h, int32_t* keys, double* mu, double* info, int32_t* npts);
int oracle_ndt_compute_hb(oracle_ndt* h, const float* src, size_t n, size_t stride_bytes, const double* pose7,
                          double* H36, double* B6, oracle_result* res, uint8_t* hits);
int oracle_ndt_align(oracle_ndt* h, const float* src, size_t n, size_t stride_bytes, const double* pose_in,
                     double* pose_inout, float* out_xyz, oracle_result* res, double* pose_trace);

/* ---- incremental NDT (ndt_registration.cpp:150-236, 262-372): LRU voxel cache of `capacity` voxels; add_cloud =
 * SetIncNdtTargetCloud; npts of get_voxels = points the last cloud that touched the voxel put into it ---- */
typedef struct oracle_inc_ndt oracle_inc_ndt;
oracle_inc_ndt* oracle_inc_ndt_create(const oracle_ndt_options* o, size_t capacity);
void oracle_inc_ndt_destroy(oracle_inc_ndt* h);
int oracle_inc_ndt_add_cloud(oracle_inc_ndt* h, const float* xyz, size_t n, size_t stride_bytes);
size_t oracle_inc_ndt_num_voxels(const oracle_inc_ndt* h);
int oracle_inc_ndt_get_voxels(const oracle_inc_ndt* h, int32_t* keys, double* mu, double* info, int32_t* npts);
int oracle_inc_ndt_compute_hb(oracle_inc_ndt* h, const float* src, size_t n, size_t stride_bytes, const double* pose7,
                              double* H36, double* B6, oracle_result* res, uint8_t* hits);
int oracle_inc_ndt_align(oracle_inc_ndt* h, const float* src, size_t n, size_t stride_bytes, const double* pose_in,
                         double* pose_out, float* out_xyz, oracle_result* res, double* pose_trace);

/* cloud pre-filters (pcl::removeNaNFromPointCloud, pcl::CropBox, pcl::VoxelGrid as the reference's RemoveNanPoint /
 * BoxFilter / VoxelFilter use them); out has room for n points of the same stride; return = points written */
size_t oracle_filter_remove_nan(const float* src, size_t n, size_t stride_bytes, float* out);
size_t oracle_filter_crop_box(const float* src, size_t n, size_t stride_bytes, const float* min3, const float* max3, float* out);
size_t oracle_filter_voxel_grid(const float* src, size_t n, size_t stride_bytes, float leaf, float* out);

/* pcl::transformPointCloud (icp_registration.cpp:241, ndt_registration.cpp:258) */
void oracle_transform_cloud(const float* src, size_t n, size_t stride_bytes, const double* pose7, float* out_xyz);
/* the Scalar = double instantiation Lio::AddCloud uses for key frames (lio.cpp:243,278): double arithmetic, one cast */
void oracle_transform_cloud_d(const float* src, size_t n, size_t stride_bytes, const double* pose7, float* out_xyz);
/* pose helpers for tests: pose7 <- pose7 * (exp(w), +dt) in the reference's split update form */
void oracle_pose_update(double* pose7, const double* dx6);
void oracle_pose_matrix(const double* pose7, double* R9_rowmajor);

#ifdef __cplusplus
}
#endif
#endif

/* Deterministic synthetic world / LiDAR generator (host only).
 *
 * The reference ships no data, so BASELINE.json's configs are defined on this generator
 * (SURVEY.md §8d): piecewise-planar world, surface-sampled map cloud, analytic ray-cast scans.
 * Clouds are float4 (x, y, z, tag/ring) with 16-byte stride.  Poses are 7 doubles
 * [qx qy qz qw tx ty tz] (Sophus::SE3d::data() layout, LocUtils eigen_types.h:66).
 */
#ifndef LOCREG_SYNTH_H
#define LOCREG_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct synth_world synth_world;

synth_world* synth_world_create(double W, uint64_t seed);
void synth_world_destroy(synth_world* w);
size_t synth_world_num_boxes(const synth_world* w);
size_t synth_world_sample_map(const synth_world* w, size_t n_map, double pitch, double sigma, uint64_t seed,
                              float* out4);
size_t synth_world_scan(const synth_world* w, const double* pose7, int beams, int azimuth, uint64_t seed, float* out4);
void synth_world_scan_batch(const synth_world* w, const double* poses7, size_t S, int beams, int azimuth,
                            uint64_t seed, float* out4, int32_t* counts, int threads);
void synth_world_poses(const synth_world* w, size_t n, uint64_t seed, double* poses7);
void synth_perturb_pose(const double* gt7, uint64_t seed, double max_trans, double max_rot_rad, double* out7);

#ifdef __cplusplus
}
#endif
#endif

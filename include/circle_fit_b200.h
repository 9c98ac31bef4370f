/* circle_fit_b200.h — C ABI of the batched laser landmark detector (rigid2d::CircleFitting on the GPU).
 *
 * Stands in for rigid2d/include/rigid2d/circle_fitting.hpp:18-60 as called by the landmarks node
 * (nuslam/src/landmarks.cpp:133-141) and by the reference's own tests (nuslam/tests/circle_tests.cpp).
 * Same conventions as ekf_slam_b200.h: 0 = ok, < 0 invalid/unsupported, > 0 cudaError_t; no CPU fallback.
 */
#ifndef CIRCLE_FIT_B200_H
#define CIRCLE_FIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct circles_ctx circles_ctx;

const char* circles_last_error(void);
int circles_max_clusters(void); /* per-scan cluster slots in the detail arrays (56) */
int circles_max_beams(void);    /* 384; the reference's node feeds 360 (landmarks.cpp:65) */

/* CircleFitting::CircleFitting(), sized for batches of up to max_scans scans of n_beams readings;
 * at most max_circles centres are returned per scan.           rigid2d/src/circle_fitting.cpp:7-9 */
int circles_create(int64_t max_scans, int n_beams, int max_circles, int device, circles_ctx** out);
int circles_destroy(circles_ctx* c);

/* CircleFitting::approxCirclePositions(ranges), batched over B scans            circle_fitting.cpp:298-304
 *   ranges  [B][n_beams]  float32 as on the wire (sensor_msgs/LaserScan, widened to double on the device the
 *                         way landmarks.cpp:65-68 does) or float64 (the class's own std::vector<double>)
 *   centers [B][max_circles][2] robot-frame (x, y) of accepted circles, in cluster order
 *   counts  [B]           accepted circles per scan (may exceed max_circles; only the first max_circles are stored)
 * A scan in which no cluster survives — undefined behaviour in the reference (:54) — yields count 0. */
int circles_run_f32(circles_ctx* c, const float* ranges, int64_t B, double* centers, int32_t* counts);
int circles_run_f64(circles_ctx* c, const double* ranges, int64_t B, double* centers, int32_t* counts);
/* device-resident input, results stay on the device (see circles_device_outputs) */
int circles_run_dev_f32(circles_ctx* c, const float* d_ranges, int64_t B);
int circles_device_outputs(circles_ctx* c, void** d_centers, void** d_counts);

/* Per-cluster detail of scan `scan` of the last run — the seams the reference exposes as
 * clusteringRanges / get_point_cluster / get_r_cluster (circle_fitting.cpp:11-102).  Any pointer may be NULL.
 *   segs  [max_clusters][4] = {s1, l1, s2, l2}: beams s1..s1+l1-1 then s2..s2+l2-1 (second segment only after
 *                             the wrap-around merge, :62-70)
 *   cxr   [max_clusters][4] = {cx, cy, radius, mean inscribed angle}
 *   flags [max_clusters]      bit0 = classified as circle (:271), bit1 = no eigenvalue in (0,1000) (:187-197)
 *   xy    [n_beams][2]        cartesian beam end points (:44-45) */
int circles_last_clusters(circles_ctx* c, int64_t scan, int32_t* n_clusters, int32_t* segs, double* cxr,
                          uint8_t* flags, double* xy);

/* CircleFitting::circleRegression + classifyCircle on caller-supplied clusters (the set_xy_cluster seam,
 * circle_fitting.cpp:96-98, 104-296): flat_xy holds sum(sizes) points, cxr [n][4], flags [n] as above. */
int circles_fit_clusters(circles_ctx* c, const double* flat_xy, const int32_t* sizes, int n_clusters, double* cxr,
                         uint8_t* flags);

/* on != 0: subsequent runs produce only what approxCirclePositions() returns (accepted centres and counts).  A
 * cluster is accepted iff its mean inscribed angle is in range AND its fitted radius is < 0.2 (:264-271); the angle
 * does not depend on the fit, so clusters failing it (the wall segments) skip the SVD fit.  Accepted centres are
 * identical in both modes; circles_last_clusters() then reports NaN for unfitted clusters.  Default: off. */
int circles_set_centres_only(circles_ctx* c, int on);

int circles_sync(circles_ctx* c);
int circles_timer_start(circles_ctx* c);
int circles_timer_stop(circles_ctx* c, float* ms_out);
int circles_launch_count(circles_ctx* c, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* CIRCLE_FIT_B200_H */

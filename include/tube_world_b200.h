/* tube_world_b200.h — batched on-device restatement of the reference simulator's input side
 * (nurtlesim/src/tube_world.cpp: cmd_vel noise :193-210, wheel slip + truth kinematics :214-250, collision :316-366,
 * fake sensor :369-414, 360-beam laser :423-577, tick schedule :583-600) plus the odometry twist the SLAM node hands
 * to the filter (nuslam/src/slam.cpp:173-176).  SURVEY.md §8(f)-1: keeps sim -> filter on the GPU for the Monte-Carlo
 * batch.  Same counter-based RNG as ekf-slam-ml_b200/tracegen.py (the numpy oracle of this row), keyed on
 * (seed, filter, tick, purpose), so filter b's trace does not depend on the batch size.
 * Conventions as in ekf_slam_b200.h (0 ok, < 0 argument, > 0 cudaError_t; no CPU fallback). */
#ifndef TUBE_WORLD_B200_H
#define TUBE_WORLD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tubeworld tubeworld;

typedef struct {
    double tube_radius, border, wheel_base, wheel_radius;
    double vx_std, the_std, slip_min, slip_max, sensor_std, max_visible, range_std;
    double cmd_v, cmd_radius;
    int32_t n_slots; /* landmark slots of the filter (readings beyond the tubes stay 0, slam.cpp:259) */
    int32_t pad;
} tubeworld_params;

const char* tubeworld_last_error(void);
int tubeworld_create(int64_t n_robots, const tubeworld_params* p, const double* tubes_x, const double* tubes_y,
                     int n_tubes, uint64_t seed, int64_t first_filter, int device, tubeworld** out);
int tubeworld_destroy(tubeworld* w);
/* One SLAM-node cycle of the known-association pipeline for every robot: 11 simulator ticks, then the fake sensor
 * message and the odometry twist.  Results stay on the device (tubeworld_outputs) in exactly the layouts
 * ekf_batch_step_known_dev() takes.  The first call reports no landmark as visible (the node's first
 * measurement() call only initialises, slam.cpp:315-327). */
int tubeworld_step_known(tubeworld* w);
/* Advance `ticks` simulator ticks and ray-cast one n_beams scan per robot (float32, as on the wire). */
int tubeworld_step_scan(tubeworld* w, int ticks, int n_beams);
/* device pointers: twists [B][2], xy [B][2 n_slots], vis [B][n_slots], truth [B][3] = {x, y, theta},
 * ranges [B][n_beams] (valid after tubeworld_step_scan) */
int tubeworld_outputs(tubeworld* w, void** d_twists, void** d_xy, void** d_vis, void** d_truth, void** d_ranges);
/* host copies of the same (any pointer may be NULL) */
int tubeworld_download(tubeworld* w, double* twists, double* xy, uint8_t* vis, double* truth, float* ranges);
/* The dead-reckoning odometer the SLAM node runs beside the filter (nuslam/src/slam.cpp:96: DiffDrive::updatePose on
 * every joint_states message): the same wheel increments as the truth, no collision handling.  odom [B][3] = {x, y,
 * theta}; d_odom receives the device pointer, odom (host, optional) a copy; either may be NULL. */
int tubeworld_odometry(tubeworld* w, void** d_odom, double* odom);
int tubeworld_sync(tubeworld* w);
void* tubeworld_stream(tubeworld* w);
/* run on the consumer's CUDA stream (e.g. ekf_batch_stream()) so that sim -> filter is stream-ordered */
int tubeworld_set_stream(tubeworld* w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TUBE_WORLD_B200_H */

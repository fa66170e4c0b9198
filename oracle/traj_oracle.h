/*
 * traj_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's trajectory samplers (jrached/trajectory_generator_ros2,
 * src/trajectories/{Circle,Line,Figure8,Boomerang,Square,Rectangle,Reciprocating,Bounce,M,I,T}.cpp).  It exists only to CHECK the CUDA engine: nothing in the
 * product path (trajectory_generator_ros2_b200/, include/) may link, import or call it.  Allowed users:
 * tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity of the oracle itself is pinned by running the UNMODIFIED reference sources (oracle/_ref, built by
 * oracle/Makefile from /root/reference behind the stub headers in compat/ros2_stubs) on the same inputs:
 * tests/test_oracle_vs_ref.py (bit-exact, when oracle/_ref is present) and the committed golden vectors in
 * tests/golden/ that were generated from oracle/_ref by tests/golden/make_golden.py.  The reference ships no
 * tests or golden vectors of its own.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -march=native, no -ffast-math): the reference builds with
 * no optimisation or arch flags (CMakeLists.txt:17-22), so every a*b+c is two IEEE roundings.
 *
 * The parameter record (tgx_params), channel order and status bits are shared with include/tgx.h.
 */
#ifndef TRAJ_ORACLE_H_
#define TRAJ_ORACLE_H_

#include <stdint.h>

#include "../include/tgx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Trajectory::generateTraj for one trajectory (Circle.cpp:30-94, Line.cpp:31-89, Figure8.cpp:30-94).
 * Writes channel c of sample k to out[c*chan_stride + k] for k < cap (out may be NULL to count only).
 * Returns the sample count N (what goals.size() would be starting from an empty vector), or -1 if the
 * parameters are rejected (status BAD_PARAM / TOO_LONG).  *status receives tgx_status_bits; ph (may be NULL)
 * receives the index_msgs entries in emission order. */
int64_t orc_generate(const tgx_params* p, double* out, int64_t chan_stride, int64_t cap,
                     uint32_t* status, tgx_phases* ph, int64_t max_samples);

/* generateTraj of the constant-speed polyline family (Square.cpp:21-92, Rectangle.cpp:20-93, Reciprocating.cpp:24-60,
 * Bounce.cpp:19-52, M.cpp:13-67, I.cpp:19-75, T.cpp:19-73), same output convention as orc_generate (which also accepts
 * these types).  leg_of (may be NULL, `cap` entries) receives the leg of every sample in tgx_polyline_legs numbering
 * (-1 for the Square / Rectangle start sample): the reference's per-sample index_msgs strings are a function of it. */
int64_t orc_polyline_generate(const tgx_params* p, double* out, int64_t chan_stride, int64_t cap, uint32_t* status,
                              int16_t* leg_of, int64_t max_samples);

/* The reference's index_msgs text of sample k (of n) of a polyline trajectory whose leg is `leg`. */
int orc_polyline_msg(int type, int leg, int64_t k, int64_t n, char* buf, int cap);

/* Trajectory::generateStopTraj (Circle.cpp:132-169, Line.cpp:117-152, Figure8.cpp:130-167): brake from the
 * setpoint from[14] (tgx_channel order).  Same output convention.  Returns the number of braking samples. */
int64_t orc_stop(const tgx_params* p, const double* from, double* out, int64_t chan_stride, int64_t cap,
                 uint32_t* status, tgx_phases* ph, int64_t max_samples);

/* Trajectory::trajectoryInsideBounds (Circle.cpp:171-179, Line.cpp:154-173, Figure8.cpp:169-177).
 * box = xmin,xmax,ymin,ymax,zmin,zmax.  Returns 1 / 0. */
int orc_inside_bounds(const tgx_params* p, const double box[6]);

/* Line::get_d2 (Line.cpp:175-181). */
double orc_line_d2(const tgx_params* p);

/* Batch drivers (nthreads POSIX threads over contiguous blocks of trajectories).
 * Output element (i, c, k) at out[i*traj_stride + c*chan_stride + k]; out may be NULL (count only). */
int orc_generate_batch(const tgx_params* p, int64_t n, double* out, int64_t traj_stride,
                       int64_t chan_stride, int64_t cap, int32_t* counts, uint32_t* status,
                       int64_t max_samples, int nthreads);

/* Feasibility as BASELINE.json config 4 defines it: generate every sample, reduce
 * max_k sqrt(vx^2+vy^2+vz^2) and max_k sqrt(ax^2+ay^2+az^2); flag = max_v<=v_max && max_a<=a_max &&
 * status==0 (status includes OUTSIDE_BOUNDS when limits->check_box). */
int orc_feasibility_batch(const tgx_params* p, int64_t n, const tgx_limits* limits, uint8_t* flags,
                          double* max_v, double* max_a, int32_t* counts, uint32_t* status,
                          int64_t max_samples, int nthreads);

/* Timing leg for bench.py: generate n trajectories into per-thread scratch rows (SoA, 14 planes) and
 * return the total number of samples produced; *checksum gets a cheap sum so the work cannot be elided. */
int64_t orc_time_batch(const tgx_params* p, int64_t n, int64_t max_samples, int nthreads, double* checksum);

/* Consumer side: the Goal the node publishes on tick k of TRAJ_FOLLOWING, traj_goals_[k] with its
 * position saturated to the room box (TrajectoryGenerator.cpp:557, :602-604, saturate :773-780), for every sample of
 * one trajectory given as SoA planes samples[c * chan_stride + k].  box = xmin,xmax,ymin,ymax,zmin,zmax or NULL. */
void orc_pack_goals(const double* samples, int64_t chan_stride, int64_t n, int32_t traj, const double* box,
                    tgx_goal_record* out);

/* Node-side transitions: take-off ramp, simpleInterpolation towards a destination, landing
 * (TrajectoryGenerator.cpp:531-599, :637-764, :602-604) under perfect tracking; see tgx_transition_params. */
int64_t orc_transition(const tgx_transition_params* t, int32_t traj, const double* box, tgx_goal_record* out,
                       int64_t cap, uint32_t* status, int64_t max_samples);

/* FNV-1a-64 over n doubles (little-endian bytes, -0.0 canonicalised to +0.0), chained through `seed`
 * (pass 0xcbf29ce484222325 to start).  Used for golden checksums. */
uint64_t orc_fnv1a64(const double* x, int64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif

#endif /* TRAJ_ORACLE_H_ */

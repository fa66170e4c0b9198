/*
 * tgx.h — C-ABI of the B200 batched trajectory-evaluation engine (libtgx.so).
 *
 * This is the drop-in boundary for ONE path of jrached/trajectory_generator_ros2: sampling the
 * parametric trajectory classes (Circle / Line / Figure8 with their ramp-up / hold / ramp-down phases, Boomerang,
 * the constant-speed polyline family Square / Rectangle / Reciprocating / Bounce / M / I / T, and every class's
 * braking trajectory) into per-time-step setpoints, plus what the node does with them (clamp + publish) and the
 * node's own transition setpoints around a trajectory.  Citations are file:line in the reference tree.
 *
 *   reference interface                                         replaced by
 *   ----------------------------------------------------------  ---------------------------------------------
 *   Trajectory::generateTraj          Trajectory.hpp:33-35       tgx_plan + tgx_eval      (tgx_generate_host)
 *     Circle::generateTraj            Circle.cpp:30-94
 *     Line::generateTraj              Line.cpp:31-89
 *     Figure8::generateTraj           Figure8.cpp:30-94
 *   create{Circle,Line,Figure8}Goal   Circle.cpp:96-130, Line.cpp:91-115, Figure8.cpp:96-128   (inside tgx_eval)
 *   Trajectory::generateStopTraj      Trajectory.hpp:38-41       tgx_plan_stop + tgx_eval (tgx_stop_host)
 *     Circle/Line/Figure8 overrides   Circle.cpp:132-169, Line.cpp:117-152, Figure8.cpp:130-167
 *   Trajectory::trajectoryInsideBounds Trajectory.hpp:44-46      tgx_limits.box + TGX_ST_OUTSIDE_BOUNDS
 *     Circle/Line/Figure8 overrides   Circle.cpp:171-179, Line.cpp:154-173, Figure8.cpp:169-177
 *   index_msgs (phase announcements)  Circle.cpp:45,61-62,74,89  tgx_phases (keys + kinds; host formats text)
 *   traj_goals_[pub_index_]           TrajectoryGenerator.cpp:557  sample k of the output planes
 *   {Square,Rectangle,Reciprocating,Bounce,M,I,T}::generateTraj                tgx_plan_polyline + tgx_eval
 *                                     Square.cpp:21-92 ... T.cpp:19-73         (tgx_generate_host_legs)
 *     create<Shape>Goal               Square.cpp:94-110, Bounce.cpp:54-72      tgx_plan_samples / tgx_sample_host
 *     generateStopTraj                Square.cpp:112-137, Bounce.cpp:74-103    tgx_plan_stop + tgx_eval
 *     per-sample index_msgs           Square.cpp:61,79,88; M.cpp:57,65         tgx_polyline_legs (host formats text)
 *   pubCB, TRAJ_FOLLOWING + saturate  TrajectoryGenerator.cpp:556-561,602-604  tgx_eval_records / tgx_pack_goals
 *   pubCB, TAKING_OFF / INIT_POS(_TRAJ) / LANDING + simpleInterpolation        tgx_transitions
 *                                     TrajectoryGenerator.cpp:531-599,637-764
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every function returns an int status (TGX_OK == 0) and
 *     never throws, never calls exit(): the reference's exit(1) / RCLCPP_WARN paths become per-trajectory
 *     status bits (tgx_status_bits).
 *   - "d_" pointers are CUDA device pointers on the engine's device, "h_" pointers are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - Sample k of a trajectory is time k*dt (TrajectoryGenerator.cpp:557 walks the vector by index).
 *   - An engine handle is bound to one GPU and is not thread-safe (the reference's objects are not either,
 *     Trajectory.hpp:47).
 *   - There is no CPU fallback: every entry point that computes samples runs CUDA kernels and fails with
 *     TGX_ERR_CUDA if no device is usable.
 */
#ifndef TGX_H_
#define TGX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGX_VERSION 100          /* 0.1.0 */
#define TGX_MAX_VGOALS 8         /* goal speeds held by one record (default.yaml:42 uses 3); see TGX_VGOALS_MORE */
#define TGX_MAX_VGOALS_TOTAL 64  /* goal speeds of one Circle / Figure8 over its record and its continuation records */
/* Records a Circle / Figure8 with K goal speeds occupies in the parameter array: 1 for K <= 8, else 1 + ceil((K-8)/8). */
#define TGX_ORBIT_RECORDS(K) ((K) <= TGX_MAX_VGOALS ? 1 : 1 + ((K) - 1) / TGX_MAX_VGOALS)
#define TGX_NCHAN 14             /* numeric fields of one setpoint */
#define TGX_MAX_PHASES (2 * TGX_MAX_VGOALS + 2)  /* index_msgs entries of one trajectory, upper bound */

/* Trajectory class tag (TrajectoryGenerator.cpp:176-260 dispatches on the same three names). */
enum tgx_type {
    TGX_CIRCLE = 0,
    TGX_LINE = 1,
    TGX_FIGURE8 = 2,
    TGX_BOOMERANG = 3,    /* Line out and back (Boomerang.cpp:31-141); takes tgx_line_params */
    /* Constant-speed polyline family; all take tgx_polyline_params and are planned by
     * tgx_plan_polyline.  default.yaml:9 ships traj_type: T. */
    TGX_SQUARE = 4,        /* Square.cpp:21-87 */
    TGX_RECTANGLE = 5,     /* Rectangle.cpp:20-88 */
    TGX_RECIPROCATING = 6, /* Reciprocating.cpp:24-60 */
    TGX_BOUNCE = 7,        /* Bounce.cpp:19-52 */
    TGX_M = 8,             /* M.cpp:13-67 */
    TGX_I = 9,             /* I.cpp:19-75 */
    TGX_T = 10,            /* T.cpp:19-73 */
    /* Continuation of the Circle / Figure8 record in front of it.  The reference loops over a std::vector of any
     * length (Circle.cpp:43, Figure8.cpp:43); a trajectory with K > 8 goal speeds is written as TGX_ORBIT_RECORDS(K)
     * CONSECUTIVE records: the first one carries n_vgoals = K and the goals 0..7, continuation record q = 1, 2, ...
     * (this type) carries the goals 8q .. 8q+7 in u.orbit.v_goals (its other fields are ignored).  A continuation record
     * is an entry of the batch like any other — count 0, status 0, no samples, its output row is never touched — so
     * that trajectory i of the batch stays row i of every output; the tgx_phases entries of its row hold the index_msgs
     * entries 18q .. 18q+17 of the trajectory.  A caller that cuts a batch (chunks, shards) must not separate a record
     * from its continuations. */
    TGX_VGOALS_MORE = 30
};
#define TGX_IS_POLYLINE(type) ((type) >= TGX_SQUARE && (type) <= TGX_T)

/* Channel order of the struct-of-arrays output: the numeric fields of snapstack_msgs2/Goal in the order
 * create*Goal fills them (Circle.cpp:107-126). */
enum tgx_channel {
    TGX_PX = 0, TGX_PY, TGX_PZ,
    TGX_VX, TGX_VY, TGX_VZ,
    TGX_AX, TGX_AY, TGX_AZ,
    TGX_JX, TGX_JY, TGX_JZ,
    TGX_PSI, TGX_DPSI
};

/* Circle / Figure8 constructor arguments (Circle.hpp:30-31, Figure8.hpp:30-31). 13 doubles. */
typedef struct tgx_orbit_params {
    double r;                        /* radius, m */
    double cx, cy;                   /* centre, m */
    double t_traj;                   /* hold time per goal velocity, s */
    double accel;                    /* ramp acceleration, m/s^2 */
    double v_goals[TGX_MAX_VGOALS];  /* goal speeds, m/s; first n_vgoals are used */
} tgx_orbit_params;

/* Line constructor arguments (Line.hpp:30-31). 13 doubles. */
typedef struct tgx_line_params {
    double A[3];                     /* start point (z only enters |B-A| and the bounds test) */
    double B[3];                     /* end point */
    double a1;                       /* acceleration, m/s^2 (> 0) */
    double a3;                       /* deceleration magnitude, m/s^2 (> 0) */
    double v_goal;                   /* cruise speed = v_goals[0] (Line.cpp:43) */
    double reserved[4];
} tgx_line_params;

/* Constructor arguments of the constant-speed polyline family (Square.hpp:31-32, Rectangle.hpp, Reciprocating.hpp,
 * Bounce.hpp, M.hpp, I.hpp, T.hpp). 13 doubles.
 *
 * cos_o / sin_o: the reference rotates its waypoints with std::cos / std::sin of `orientation` (Square.cpp:37-38,
 * M.cpp:29-30, ...), and the per-side step counts ceil(distance / (v*dt)) depend on the last bit of those two
 * numbers.  libm's results are not reproducible across math libraries, so they are INPUTS here: a caller that wants
 * indexing bit-identical to a reference built against its own libm passes that libm's values and sets
 * TGX_POLY_TRIG_GIVEN in tgx_params.n_vgoals (tgx_polyline_finalize_host does exactly that; the host-buffer calls and
 * the drop-in classes use it).  Without the flag the planner uses the device's cos / sin (<= 2 ulp from libm's). */
typedef struct tgx_polyline_params {
    double t_traj;                   /* total time, s */
    double v_goal;                   /* v_goals_.empty() ? 1.0 : v_goals_[0] (Square.cpp:48, M.cpp:39, ...) */
    double decel;                    /* braking deceleration: accel_ (Square.cpp:129, Rectangle.cpp:126), a3_
                                        (Reciprocating.cpp:98); ignored by M / I / T (literal 1.0, M.cpp:101) and Bounce */
    double orientation;              /* rad; unused by Reciprocating */
    double cos_o, sin_o;             /* see above */
    double g[7];                     /* geometry, by type:
                                        SQUARE         side_length, cx, cy
                                        RECTANGLE      side_a, side_b, cx, cy
                                        RECIPROCATING  Ax, Ay, Az, Bx, By, Bz
                                        BOUNCE         cx, cy, Az, Bz
                                        M, I, T        cx, cy, length, width
                                        g[5], g[6]: z and heading arguments of tgx_plan_samples (create*Goal helpers) */
} tgx_polyline_params;

#define TGX_POLY_TRIG_GIVEN 1        /* tgx_params.n_vgoals bit 0 for polyline types: cos_o / sin_o are valid */

/* One trajectory's parameters: exactly 128 bytes, the unit of the batch parameter array. */
typedef struct tgx_params {
    int32_t type;                    /* enum tgx_type */
    int32_t n_vgoals;                /* orbit: 0..TGX_MAX_VGOALS_TOTAL (0: the one-sample trajectory the reference makes of an
                                        empty vector; > 8: see TGX_VGOALS_MORE); line: ignored; polyline: flag bits */
    double dt;                       /* Trajectory::dt_ = 1/pub_freq (TrajectoryGenerator.cpp:171-172) */
    double alt;                      /* alt_: z of every sample */
    union {
        tgx_orbit_params orbit;      /* TGX_CIRCLE, TGX_FIGURE8 */
        tgx_line_params line;        /* TGX_LINE, TGX_BOOMERANG */
        tgx_polyline_params poly;    /* TGX_SQUARE .. TGX_T */
    } u;
} tgx_params;

/* Per-trajectory status bitmask. 0 = clean. */
enum tgx_status_bits {
    TGX_ST_VGOALS_NOT_INCREASING = 1u << 0, /* RCLCPP_WARN at Circle.cpp:57-59 / Figure8.cpp:57-59 (samples still produced) */
    TGX_ST_FINAL_V_NONZERO       = 1u << 1, /* exit(1) at Circle.cpp:85-88 / Figure8.cpp:85-88 */
    TGX_ST_LINE_END_NOT_B        = 1u << 2, /* exit(1) at Line.cpp:76-79 (end point > 0.05 m from B); for a Boomerang also the
                                               return leg's "final point is not A" (Boomerang.cpp:126-129) */
    TGX_ST_LINE_D2_NEGATIVE      = 1u << 3, /* Line.cpp:165-168: cruise segment length < 0; reported, like the reference does,
                                               by the bounds check only (set together with OUTSIDE_BOUNDS) */
    TGX_ST_OUTSIDE_BOUNDS        = 1u << 4, /* trajectoryInsideBounds() == false (only if a box was given) */
    TGX_ST_BAD_PARAM             = 1u << 5, /* v<=0, accel<=0 (TrajectoryGenerator.cpp:184-195,268-277), dt<=0, r==0
                                               (a negative radius flies the mirrored circle, as in the reference),
                                               n_vgoals out of range or continuation records missing, unknown type,
                                               non-finite input: no samples */
    TGX_ST_VMAX_EXCEEDED         = 1u << 6, /* max_k |v_k| > limits.v_max  (tgx_feasibility only) */
    TGX_ST_AMAX_EXCEEDED         = 1u << 7, /* max_k |a_k| > limits.a_max  (tgx_feasibility only) */
    TGX_ST_TOO_LONG              = 1u << 8, /* sample count would exceed the engine's max_samples guard (the reference
                                               would loop for ever / exhaust memory): no samples */
    TGX_ST_TRUNCATED             = 1u << 9, /* row capacity of the output layout < sample count: tail not written */
    TGX_ST_WRONG_PLANNER         = 1u << 10 /* a polyline-family trajectory handed to tgx_plan, or a Circle / Line / Figure8 /
                                               Boomerang handed to tgx_plan_polyline: no samples (the host-buffer calls route
                                               every trajectory to its planner themselves) */
};

/* Status bits that mean "the reference would not have produced this trajectory". */
#define TGX_ST_FATAL_MASK \
    (TGX_ST_FINAL_V_NONZERO | TGX_ST_LINE_END_NOT_B | TGX_ST_BAD_PARAM | TGX_ST_TOO_LONG | TGX_ST_WRONG_PLANNER)

/* Library return codes. */
enum tgx_error {
    TGX_OK = 0,
    TGX_ERR_INVALID = 1,      /* NULL / negative / inconsistent argument */
    TGX_ERR_CUDA = 2,         /* a CUDA runtime call failed; tgx_last_cuda_error() has the text */
    TGX_ERR_ALIGNMENT = 3,    /* output layout not aligned for vector stores */
    TGX_ERR_NO_PLAN = 4,      /* tgx_eval / tgx_feasibility before tgx_plan */
    TGX_ERR_NOMEM = 5,        /* device or host allocation failed */
    TGX_ERR_CAPACITY = 6,     /* a caller-provided buffer is too small */
    TGX_ERR_COMM = 7          /* NCCL is missing or a collective failed; tgx_comm_last_error() has the text */
};

/* Room box + kinematic limits. Box follows trajectoryInsideBounds(xmin,xmax,ymin,ymax,zmin,zmax). */
typedef struct tgx_limits {
    double box[6];            /* xmin, xmax, ymin, ymax, zmin, zmax */
    double v_max;             /* m/s   (used by tgx_feasibility) */
    double a_max;             /* m/s^2 (used by tgx_feasibility) */
    int32_t check_box;        /* 0: ignore box */
    int32_t reserved;
} tgx_limits;

/* Where samples go: element (trajectory i, channel c, sample k) lives at
 *     base[ (traj_offset ? traj_offset[i] : i*traj_stride) + c*chan_stride + k ]        (units: doubles)
 * Plane-major:       chan_stride = n*row, traj_stride = row.
 * Trajectory-major:  chan_stride = row,   traj_stride = 14*row.
 * base must be 32-byte aligned and every stride/offset a multiple of 4 doubles (vector stores).
 * Samples k >= capacity are not written (status TGX_ST_TRUNCATED).  Padding: the slots that share a 32-byte sector
 * with the trajectory's last samples, k in [N_i, round_up(N_i, 4)), are zero-filled when they lie inside `capacity`
 * (a partially written sector would have to be fetched from DRAM first: one read-fill per channel per trajectory in the
 * middle of the store stream); padding k >= round_up(N_i, 4) is never written. */
typedef struct tgx_layout {
    double* d_base;
    int64_t traj_stride;
    int64_t chan_stride;
    const int64_t* d_traj_offset;   /* optional device array [n]; overrides traj_stride */
    int64_t capacity;               /* max samples per trajectory that fit */
    uint32_t channel_mask;          /* bit c set = write channel c; 0 means all 14 */
    uint32_t reserved;
} tgx_layout;

/* index_msgs of one trajectory: the reference stores announcement strings keyed by sample index
 * (Circle.cpp:45,61-62,74,89).  The engine returns (key, kind, value) triples in emission order; a later
 * entry with the same key overwrites an earlier one, as operator[] does in the reference. */
enum tgx_phase_kind {
    TGX_PH_ACCEL_TO = 0,     /* "<Shape> traj: accelerating to <value> m/s"                              */
    TGX_PH_REACHED = 1,      /* "<Shape> traj: reached <value> m/s, keeping constant v for <value2> s"   */
    TGX_PH_DECEL = 2,        /* "<Shape> traj: decelerating to 0 m/s"                                    */
    TGX_PH_STOPPED = 3,      /* "<Shape> traj: stopped"                                                  */
    TGX_PH_PRESSED_END = 4   /* "<Shape> traj: pressed END, decelerating to 0 m/s" (stop trajectories)   */
};

typedef struct tgx_phases {
    int32_t n;
    int32_t key[TGX_MAX_PHASES];
    int32_t kind[TGX_MAX_PHASES];
    double value[TGX_MAX_PHASES];    /* v_goal for ACCEL_TO / REACHED */
    double value2[TGX_MAX_PHASES];   /* hold time for REACHED (t_traj, or Line's t2) */
} tgx_phases;

/* index_msgs of a polyline-family trajectory.  The reference announces EVERY sample ("Square traj: moving along side
 * 2", Square.cpp:79; "M traj: segment 1 rev", M.cpp:57; "Reciprocating: forward", Reciprocating.cpp:47), so instead of
 * one entry per sample the engine returns the structure those strings are a function of: the samples follow a
 * periodic pattern of `n_legs` legs, leg l contributing count[l] consecutive samples per period,
 *     sample k  ->  m = k - first_special,  leg = the l with  sum(count[0..l-1]) <= m mod period < sum(count[0..l]).
 * Leg numbering: SQUARE / RECTANGLE side 0..3; RECIPROCATING 0 forward, 1 yaw flip at B, 2 reverse, 3 yaw flip at A;
 * BOUNCE 0 "ascending", 1 "descending" (the reference labels by lap parity, Bounce.cpp:41); M legs 0..3 "segment l fwd",
 * 4..7 "segment l-4 rev"; I 0..4 fwd, 5..9 rev; T 0..2 fwd, 3..5 rev.  The last sample's entry is then overwritten by
 * "<Shape> traj: completed" (Square.cpp:88, Bounce.cpp:50, M.cpp:65; not for RECIPROCATING). */
#define TGX_POLY_MAX_LEGS 10
typedef struct tgx_polyline_legs {
    int32_t n;                          /* sample count (0: rejected or empty) */
    int32_t n_legs;
    int32_t first_special;              /* 1: sample 0 precedes the pattern ("starting at corner 0", Square.cpp:60-61) */
    int32_t last_special;               /* 1: the last sample is the yaw-flip goal the reference appends after a leg
                                           that t_traj cut short (Reciprocating.cpp:50-57) */
    int32_t period;                     /* sum of count[] */
    int32_t count[TGX_POLY_MAX_LEGS];
    int32_t reserved;
} tgx_polyline_legs;

/* ---- consumer side ------------------------------------------------------------------
 * What the node publishes on tick k of TRAJ_FOLLOWING: goal_ = traj_goals_[pub_index_] (TrajectoryGenerator.cpp:557)
 * with the position saturated to the room bounds (:602-604, saturate() :773-780) — one array-of-structs record per
 * (trajectory, sample), 128 bytes, so that a fleet of nodes can be fed straight from one engine: the host side turns a
 * record into a snapstack_msgs2/Goal with a plain copy instead of gathering 14 strided planes. */
typedef struct tgx_goal_record {
    double p[3], v[3], a[3], j[3];   /* Goal.p / v / a / j */
    double psi, dpsi;
    int32_t traj;                    /* trajectory index in the batch */
    int32_t k;                       /* sample index = pub_index_ of the tick that publishes it */
    uint8_t power;                   /* Goal.power: true on every trajectory sample (Circle.cpp:127) */
    uint8_t mode_xy, mode_z;         /* Goal.MODE_POSITION_CONTROL (0): the samplers never touch them */
    uint8_t clamped;                 /* bit 0 / 1 / 2: p.x / p.y / p.z was saturated */
    uint8_t last;                    /* 1 on the trajectory's last sample.  The node never publishes that one: on the tick
                                        that loads it pub_index_ reaches size(), goal_ is overwritten by the hover goal
                                        at the vehicle's pose and the node switches to HOVERING
                                        (TrajectoryGenerator.cpp:561-572) */
    uint8_t reserved[3];
} tgx_goal_record;

/* ---- node-side transitions -------------------------------------------------------------
 * The setpoints the node generates itself around a trajectory, one per 100 Hz tick, as recurrences on its own goal_:
 *   TAKEOFF  goal_.p.z = saturate(goal_.p.z + vel_take*dt, 0, alt) until the vehicle is within 0.10 m of alt
 *            (TrajectoryGenerator.cpp:531-548)
 *   GOTO     goal_ = simpleInterpolation(goal_, dest, dest_yaw, vel, vel_yaw, dist_thresh, yaw_thresh, dt, finished)
 *            (:549-554 towards traj_goals_[0], :574-586 back to the take-off point; simpleInterpolation :637-764)
 *   LANDING  goal_.p.z -= (pose.z > ground + 0.4 ? vel_land_fast : vel_land_slow)*dt until goal_.p.z < 0 (:588-599)
 * followed on every tick by the saturation of goal_.p to the room box, which feeds back into the recurrence (:602-604).
 * The node reads the vehicle's pose in two places (the hover test of TAKEOFF, the speed choice of LANDING); the batch
 * engine assumes PERFECT TRACKING there: the pose on tick k is the goal published on tick k-1.  One record per tick. */
enum tgx_transition_kind { TGX_TR_TAKEOFF = 0, TGX_TR_GOTO = 1, TGX_TR_LANDING = 2 };

typedef struct tgx_transition_params {          /* 128 bytes */
    int32_t kind;                /* enum tgx_transition_kind */
    int32_t ticks;               /* 0: run until the phase ends by itself (hover reached / finished / landed);
                                    > 0: exactly that many ticks (a GOTO keeps holding its destination, as the node does
                                    while it waits for the operator, :549-554) */
    double dt;
    double start[3];             /* goal_.p when the phase starts */
    double start_v[2];           /* goal_.v.x, goal_.v.y (simpleInterpolation smooths the velocity reference, :662-663) */
    double start_psi;            /* goal_.psi */
    double dest[3];              /* GOTO: destination; TAKEOFF: dest[2] = alt_; LANDING: dest[2] = init_pos_.z */
    double dest_yaw;             /* GOTO only (the trip home passes start_psi: :579 uses goal_.psi itself) */
    double vel;                  /* TAKEOFF vel_take_; GOTO vel_initpos_; LANDING vel_land_fast_ */
    double vel_yaw;              /* GOTO vel_yaw_; LANDING vel_land_slow_ */
    double dist_thresh, yaw_thresh;   /* GOTO */
} tgx_transition_params;

typedef struct tgx_engine tgx_engine;

/* ---- life cycle ---------------------------------------------------------------------------------- */
int tgx_version(void);
const char* tgx_strerror(int code);
const char* tgx_last_cuda_error(void);            /* text of the last CUDA failure on this thread */
int tgx_create(tgx_engine** out, int device);     /* binds to `device`; TGX_ERR_CUDA if there is none */
int tgx_destroy(tgx_engine* e);
/* Guard against parameters for which the reference never terminates. Default 1<<24 samples. */
int tgx_set_max_samples(tgx_engine* e, int64_t max_samples);
/* Kernel shape used by tgx_eval / tgx_feasibility: one CTA evaluates a tile of (1 << tile_shift) consecutive
 * samples of one trajectory (tile_shift 9 or 10); in the vector-store kernels each thread owns `spt` adjacent samples
 * (2: 128-bit stores, 4: 256-bit stores) and (1 << tile_shift) / spt must be 128 or 256 threads.  The TMA kernels
 * (planes, records) and the reduction-only kernel choose their own CTA width and walk the tile in passes of 256
 * samples; only tile_shift matters to them.  The tuning changes the speed only: the planner cuts a trajectory into
 * closed-form segments by the replay alone (phase boundaries, at most 2048 steps per segment), never by the tile
 * size, so every tuning writes the same bytes.  Invalidates the current plan.  Default 10, 4. */
int tgx_set_tuning(tgx_engine* e, int tile_shift, int spt);
/* Planning mode.  Sample counts, phase boundaries (index_msgs keys), status bits and the speed at every segment base
 * are bit-identical to the reference in both modes; inside a segment v_k = fma(j, a*dt, v_base) is rounded once where
 * the reference rounds every step (relative difference <= j * 2^-54, observed <= 1e-13).
 *   exact_ramps = 0 (default): ramps are advanced in exact arithmetic-progression jumps of v and the angle / position
 *       state in closed form per jump; the state at segment bases then differs from the reference's running sums by
 *       those sums' own accumulated rounding (<= ~1e-12 rad, ~1e-12 m).  O(#binades) work per trajectory.
 *   exact_ramps = 1: every ramp step is replayed with the reference's operation sequence; the (v, theta | x, y) state
 *       at every segment base, and theta on every hold sample, is bit-identical to the reference.  O(N) work.
 * Braking plans (tgx_plan_stop) always use the exact replay.  Invalidates the current plan.
 * Within one mode the samples are a function of the parameters alone: which planning path the engine takes (first
 * plan, fixed slices, phase records, below) depends on the batches it has seen, the bytes it writes do not. */
int tgx_set_plan_mode(tgx_engine* e, int exact_ramps);
/* Single-replay planning (default on): once a plan has measured the largest per-trajectory segment / tile counts of
 * a batch, later plans give every trajectory a fixed slice of the tables and need no counting pass and no scans; a
 * batch that does not fit falls back to the two-replay exact-offset path automatically.  allow = 0 disables it. */
int tgx_set_slab_planning(tgx_engine* e, int allow);
/* Phase planning (default on): a batch of Circle / Figure8 trajectories with at most 2 goal speeds and at most 20
 * segments and of plain Line trajectories, none longer than 4096 samples, is planned into ONE self-contained 256-byte
 * record per trajectory (where each segment ends, the replayed state there, the constants of the parameter record;
 * orbits of more than 12 segments keep the rest in a 96-byte extension row) instead of segment tables.  One CTA
 * evaluates one trajectory: it rebuilds from the record exactly the segments the table path would have read — the same
 * bytes come out, with a tenth of the table traffic in the store-bound kernel and no tile directory, however ragged the
 * batch is.  The plan holds a copy of everything it needs: d_params may be overwritten or freed as soon as tgx_plan
 * returns, whichever path was taken.
 * Engaged automatically after a plan has seen such a batch; anything else (Boomerangs, more goal speeds, long
 * trajectories) is planned with segment tables, and so are batches whose previous plan only fed tgx_feasibility: the
 * reduction kernel is bound by instruction issue, not by HBM, and copying segments is cheaper for it than rebuilding
 * them. */
int tgx_set_phase_planning(tgx_engine* e, int allow);
int64_t tgx_phase_plan_count(const tgx_engine* e);
/* Store path of tgx_eval (default tma = 1): when the layout is regular (no per-trajectory offsets, all 14 channels,
 * rows of at least 32 samples) and no maxima are requested, CTAs of 128 threads stage 32-sample groups of all 14
 * channels in shared memory and hand each group to the TMA unit as one box of a [trajectory][channel][sample] tensor
 * map over the caller's planes; only the group that holds a row's end uses vector stores.  tma = 0 always uses the
 * vector-store kernel (one 256-bit streaming store per thread per channel).  Both write the same bytes. */
int tgx_set_store_path(tgx_engine* e, int tma);
/* How many plans so far took the single-replay / the two-replay path (either pointer may be NULL). */
int tgx_plan_path_counts(const tgx_engine* e, int64_t* slab_plans, int64_t* exact_plans);
/* Bytes of device scratch currently held by the engine (plan tables). */
int64_t tgx_scratch_bytes(const tgx_engine* e);

/* ---- counting pass: Trajectory::generateTraj's loop structure only -------------------------------- */
/* Replays the reference's scalar recurrences (v <- min(v + a*dt, v_goal), cur += dt, ...) with
 * non-contracted IEEE fp64 operations, one thread per trajectory, and writes the exact sample count and
 * status of each.  No samples are produced. d_counts / d_status may be NULL. */
int tgx_count(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
              int32_t* d_counts, uint32_t* d_status, void* stream);

/* ---- planning pass: counts + the segment / tile tables tgx_eval consumes --------------------------- */
/* Must precede tgx_eval / tgx_feasibility. Keeps device scratch inside the engine (the "current plan").
 * Synchronises `stream` once (to size the scratch). d_counts / d_status / d_phases may be NULL.
 * *total_samples (host, may be NULL) receives sum_i N_i. */
int tgx_plan(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
             int32_t* d_counts, uint32_t* d_status, tgx_phases* d_phases, int64_t* total_samples,
             void* stream);

/* Planning pass of the constant-speed polyline family (Square / Rectangle / Reciprocating / Bounce / M / I / T
 * ::generateTraj: Square.cpp:21-92, Rectangle.cpp:20-93, Reciprocating.cpp:24-60, Bounce.cpp:19-52, M.cpp:13-67,
 * I.cpp:19-75, T.cpp:19-73).  One thread per trajectory builds the waypoints and per-leg step counts
 * ceil(distance / (v*dt)) with non-contracted IEEE operations and looks the number of `t += dt` iterations up in the
 * same table the hold phases use; the result is a current plan that tgx_eval / tgx_feasibility evaluate with the
 * polyline kernel (frac = i / steps, p = start + frac * (end - start): positions are bit-identical to the reference).
 * d_counts / d_status / d_legs / total_samples may be NULL.  Synchronises `stream` once. */
int tgx_plan_polyline(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
                      int32_t* d_counts, uint32_t* d_status, tgx_polyline_legs* d_legs, int64_t* total_samples,
                      void* stream);

/* Fills u.poly.cos_o / sin_o of every polyline-family record with the HOST libm's cos / sin of `orientation` and sets
 * TGX_POLY_TRIG_GIVEN (see tgx_polyline_params); records of other types are left untouched. */
int tgx_polyline_finalize_host(tgx_params* h_params, int64_t n);

/* Braking plan (Trajectory::generateStopTraj): trajectory i brakes from the setpoint d_from[i*14 .. i*14+13]
 * (channel order tgx_channel; the reference reads goals[pub_index], Circle.cpp:140-143, Line.cpp:124-127).
 * The result is a new current plan whose samples are the braking trajectory (sample 0 = first braking
 * step, as the reference replaces the vector and resets pub_index to 0, Circle.cpp:162-164). */
int tgx_plan_stop(tgx_engine* e, const tgx_params* d_params, int64_t n, const double* d_from,
                  int32_t* d_counts, uint32_t* d_status, tgx_phases* d_phases, int64_t* total_samples,
                  void* stream);

/* Per-time evaluation from an explicit state: the public helpers createCircleGoal(v, accel, theta) (Circle.hpp:39),
 * createFigure8Goal (Figure8.hpp:39) and createLineGoal(last_x, last_y, v, accel, theta) (Line.hpp:38).
 * d_state[4*i ..] = {v, accel, s0, s1}: orbit s0 = theta; line s0 = last_x, s1 = last_y and the explicit heading
 * theta is taken from d_params[i].u.line.reserved[0].  The result is a current plan with exactly one sample per
 * trajectory; tgx_eval then writes it at k = 0. */
int tgx_plan_samples(tgx_engine* e, const tgx_params* d_params, int64_t n, const double* d_state,
                     int32_t* d_counts, uint32_t* d_status, void* stream);

/* ---- evaluation: create*Goal for every (trajectory, k) of the current plan ------------------------- */
/* One CTA per tile of the plan, fp64, struct-of-arrays planes.  Regular layouts leave through TMA (tgx_set_store_path:
 * warps stage {32 samples x 14 channels} groups in shared memory, one cp.async.bulk.tensor per group); layouts with
 * per-trajectory offsets, a channel subset or rows shorter than 32 samples, polyline plans and calls that ask for maxima
 * use 256-bit streaming vector stores, one per thread per channel.  Same bytes either way.
 * If d_max_v / d_max_a are non-NULL the per-trajectory maxima of |v_k| and |a_k| (Euclidean norm of the
 * written x,y,z components) are reduced in the same pass. */
int tgx_eval(tgx_engine* e, const tgx_layout* out, double* d_max_v, double* d_max_a, void* stream);

/* ---- generateTraj for a device-resident batch in one call: tgx_plan + tgx_eval, pipelined ------------------------ */
/* What a caller that wants the samples of a whole batch (Trajectory::generateTraj for n trajectories,
 * Trajectory.hpp:33) should call.  Planning is a latency-bound replay, evaluation a store stream; run back to back on
 * one stream they leave half of the machine idle in turn.  tgx_generate cuts the batch into chunks of `chunk`
 * trajectories (0: an eighth of the batch, between 32 Ki and 1 Mi) and alternates them between the engine and a private
 * twin on two internal streams, so that chunk c+1 is planned while chunk c is evaluated: all planning but the first
 * chunk's hides behind the store-bound kernel.  The work starts after everything queued on `stream` so far, and
 * `stream` waits for all of it; the host returns when the last chunk has been planned (its evaluation may still run).
 * Writes exactly the bytes tgx_plan + tgx_eval of the whole batch write (the samples do not depend on how a batch is
 * cut; a chunk never ends between a record and its continuation records), d_counts / d_status / d_phases (each may be
 * NULL) as tgx_plan does, and the sample total.  Circle / Line /
 * Figure8 / Boomerang records; polyline-family records get TGX_ST_WRONG_PLANNER as in tgx_plan.  Leaves no current
 * plan. */
int tgx_generate(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits, const tgx_layout* out,
                 int32_t* d_counts, uint32_t* d_status, tgx_phases* d_phases, int64_t chunk, int64_t* total_samples,
                 void* stream);
/* Evaluations of consecutive chunks are chained (two store streams at once are not faster than one); what overlaps an
 * evaluation is the next chunk's planning.
 * Profiling: with on = 1 every evaluation launch of tgx_generate / tgx_generate_feasibility is bracketed by a CUDA event
 * pair on the stream it is launched on; tgx_generate_profile waits for them and returns the sum of their durations in
 * milliseconds and their number since the last query (bench.py's roofline.achieved). */
int tgx_set_generate_profiling(tgx_engine* e, int on);
int tgx_generate_profile(tgx_engine* e, double* eval_ms, int64_t* eval_launches);
/* The same pipeline for the feasibility reduction: tgx_plan + tgx_feasibility per chunk (any output may be NULL). */
int tgx_generate_feasibility(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
                             uint8_t* d_flags, double* d_max_v, double* d_max_a, uint32_t* d_status, int64_t chunk,
                             int64_t* total_samples, void* stream);

/* ---- consumer side: clamp to the room bounds + pack to array-of-structs records ------------ */
/* Reads the struct-of-arrays planes `planes` (as written by tgx_eval; d_counts[i] samples of trajectory i are valid),
 * saturates p.x / p.y / p.z to limits->box if limits && limits->check_box (TrajectoryGenerator.cpp:602-604) and writes
 * record (i, k) to d_records[(d_rec_offset ? d_rec_offset[i] : i * rec_stride) + k] for k < min(d_counts[i],
 * rec_capacity).  One CTA per 256 samples: coalesced plane reads, a swizzled shared-memory transpose, 512-byte
 * coalesced record writes.  Does not need (or touch) the current plan. */
/* The fused form: evaluate the current plan (either family, braking plans too) straight into clamped records,
 * d_records[(d_rec_offset ? d_rec_offset[i] : i * rec_stride) + k], k < rec_capacity — every warp of the evaluation
 * kernels stages 64 consecutive records in shared memory (128-byte TMA swizzle) and sends them as two 4 KiB TMA boxes of
 * a [records][16 doubles] tensor map over d_records, so a sample costs 128 bytes of HBM traffic instead of the
 * 112 + 112 + 128 of tgx_eval followed by tgx_pack_goals.  Bit-identical to that pair.  d_records must be 16-byte
 * aligned and hold fewer than 2^31 records.  The call must know where the buffer ends: without offsets it holds
 * n * rec_stride records; WITH d_rec_offset, rec_stride is the total number of records d_records holds (> 0).  That
 * count is the extent of the tensor map, and a trajectory whose offset lies outside [0, total) writes nothing, one whose
 * row runs past the end is cut there. */
int tgx_eval_records(tgx_engine* e, const tgx_limits* limits, tgx_goal_record* d_records, int64_t rec_stride,
                     const int64_t* d_rec_offset, int64_t rec_capacity, void* stream);

int tgx_pack_goals(tgx_engine* e, const tgx_layout* planes, const int32_t* d_counts, int64_t n,
                   const tgx_limits* limits, tgx_goal_record* d_records, int64_t rec_stride,
                   const int64_t* d_rec_offset, int64_t rec_capacity, void* stream);

/* ---- node-side transitions ------------------------------------------------------------ */
/* One thread per vehicle replays the recurrence of its tgx_transition_params with the reference's own operation
 * order (no transcendental is involved: the results are bit-identical) and writes one tgx_goal_record per tick to
 * d_records[i * rec_stride + k], k < rec_capacity (record.k = tick within the phase, record.last = 1 on the tick that
 * ends it, record.power = 0 on the landing tick that cuts the motors).  d_counts[i] = number of ticks of the phase,
 * d_status[i] = 0, TGX_ST_BAD_PARAM, TGX_ST_TOO_LONG (not finished within max_samples ticks) or TGX_ST_TRUNCATED.
 * limits->box (if limits && limits->check_box) is the room box of :602-604. */
int tgx_transitions(tgx_engine* e, const tgx_transition_params* d_tparams, int64_t n, const tgx_limits* limits,
                    tgx_goal_record* d_records, int64_t rec_stride, int64_t rec_capacity, int32_t* d_counts,
                    uint32_t* d_status, void* stream);
/* Host-buffer variant: H2D, tgx_transitions, D2H of h_records[i * rec_capacity + k]. */
int tgx_transitions_host(tgx_engine* e, const tgx_transition_params* h_tparams, int64_t n, const tgx_limits* limits,
                         tgx_goal_record* h_records, int64_t rec_capacity, int32_t* h_counts, uint32_t* h_status);

/* ---- feasibility only: no sample stores ------------------------------------------------------------ */
/* Evaluates every sample of the current plan, reduces max |v|, max |a| per trajectory and sets
 * d_flags[i] = 1 iff max_v <= v_max && max_a <= a_max && status has no bit set (incl. OUTSIDE_BOUNDS).
 * d_status (may be NULL) is updated in place with VMAX/AMAX bits on top of the plan's status. */
int tgx_feasibility(tgx_engine* e, const tgx_limits* limits, uint8_t* d_flags, double* d_max_v,
                    double* d_max_a, uint32_t* d_status, void* stream);

/* ---- host-buffer convenience calls (what the drop-in C++ classes use) ------------------------------ */
/* Counts for host-resident parameters (H2D, tgx_count, D2H). */
int tgx_count_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                   int32_t* h_counts, uint32_t* h_status);

/* Full generateTraj for host-resident parameters into a host buffer with the layout
 * h_out[(i*14 + c)*capacity + k] (trajectory-major rows of `capacity` doubles).  Runs in chunks of
 * trajectories through two device staging slots: while one slot's planes travel to the host, the next chunk is
 * planned and evaluated into the other.  h_out is best page-locked (tgx_alloc_host_for); h_params, h_counts, h_status,
 * h_phases may be ordinary memory — the parameters go up in one copy at the start of the call and the per-trajectory
 * results are handed over from a pinned area of the engine at its end, so nothing pageable sits between two chunks'
 * copies (it would serialise them).  TGX_TRACE_D2H=1 in the environment prints, per call, when every chunk's
 * evaluation and copies finished and when the host queued them.  h_phases may be NULL. */
int tgx_generate_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                      double* h_out, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                      tgx_phases* h_phases);

/* Compact wire format of tgx_generate_host for callers that declare it.  The z-components of every setpoint of the
 * planar classes are literal constants in the reference (p.z = alt_, v.z = a.z = j.z = 0: Circle.cpp:109-121,
 * Line.cpp:99-108, Figure8.cpp:110-119, Square.cpp:101-107), so the host buffer holds only the TGX_NCHAN_VARYING = 10
 * planes that vary along a trajectory, in tgx_channel order with the four constant ones left out
 * (TGX_COMPACT_PLANE(c): px py vx vy ax ay jx jy psi dpsi -> 0..9):
 *     h_out10[(i*10 + q)*capacity + k]            (default)
 *     h_out10[(q*n + i)*capacity + k]             (tgx_set_host_layout(e, 1): plane-major)
 * and the consumer takes p.z from h_params[i].alt and v.z = a.z = j.z = 0 (the drop-in classes' repack into
 * std::vector<Goal> does exactly that).  A sample then costs 80 bytes of PCIe traffic AND 80 bytes of host-memory
 * writes, against 80 + 32 (constant planes written by host threads) for tgx_generate_host.  Padding (k >= N_i) is zero.
 * h_phases / h_legs (either may be NULL) as in tgx_generate_host_legs.
 * TGX_ERR_INVALID if the batch holds a TGX_BOUNCE trajectory (it moves along z: Bounce.cpp:39-41). */
#define TGX_NCHAN_VARYING 10
#define TGX_VARYING_CHANNEL_MASK 0x36dbu   /* all but TGX_PZ, TGX_VZ, TGX_AZ, TGX_JZ */
/* plane index of channel c in the compact format, -1 for the constant channels */
#define TGX_COMPACT_PLANE(c) \
    (((c) % 3 == 2 && (c) < TGX_PSI) ? -1 : ((c) < TGX_PSI ? 2 * ((c) / 3) + (c) % 3 : (c) - 4))
int tgx_generate_host_compact(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                              double* h_out10, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                              tgx_phases* h_phases, tgx_polyline_legs* h_legs);

/* tgx_generate_host that additionally returns, for polyline-family trajectories, the leg structure their per-sample
 * index_msgs are a function of (h_legs[i].n == 0 for the other families).  A batch may mix families: every chunk is
 * routed to tgx_plan and / or tgx_plan_polyline as its trajectories require, and polyline records without
 * TGX_POLY_TRIG_GIVEN get the host libm's cos / sin (see tgx_polyline_params). */
int tgx_generate_host_legs(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                           double* h_out, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                           tgx_phases* h_phases, tgx_polyline_legs* h_legs);

/* generateTraj for host-resident parameters straight into clamped array-of-structs records:
 * h_records[i * rec_capacity + k] (plan, evaluate, tgx_pack_goals, D2H of the records; chunked like tgx_generate_host).
 * limits->box, if given, is both the trajectoryInsideBounds box (status bit) and the saturation box. */
int tgx_generate_records_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                              tgx_goal_record* h_records, int64_t rec_capacity, int32_t* h_counts,
                              uint32_t* h_status);

/* Full generateStopTraj: h_from[i*14..] is the setpoint being braked from. Same output layout. */
int tgx_stop_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const double* h_from,
                  double* h_out, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                  tgx_phases* h_phases);

/* Wire format of the host-buffer calls.  The z-components of every setpoint are literal constants in the reference
 * (p.z = alt_, v.z = a.z = j.z = 0: Circle.cpp:109-121, Line.cpp:99-108, Figure8.cpp:110-119).  With
 * fill_constants_on_host = 1 (default) the device evaluates and ships only the 10 varying planes over PCIe and host
 * threads write the 4 constant rows of each trajectory; with 0 all 14 planes are evaluated and shipped.  The samples
 * (k < N_i) are the same either way.  Padding (k >= N_i, and every row of a rejected trajectory) is zero, except that
 * with the host fill the p.z row holds alt over its whole capacity (the filling threads run before the counts are
 * known). */
int tgx_set_host_fill(tgx_engine* e, int fill_constants_on_host);

/* Layout of the host buffers of tgx_generate_host / tgx_generate_host_legs / tgx_stop_host.
 *   plane_major = 0 (default): h_out[(i*14 + c)*capacity + k]   — one trajectory's 14 rows together (what the drop-in
 *       classes repack into std::vector<Goal>);
 *   plane_major = 1:           h_out[(c*n + i)*capacity + k]    — struct-of-arrays across the whole call: every plane of
 *       a chunk is one contiguous run on both sides of PCIe, so the device->host copies run at the link's plain-copy
 *       rate instead of the rate of 2-D copies with 16 KB runs (measured 52 vs 46 GB/s). */
int tgx_set_host_layout(tgx_engine* e, int plane_major);

/* One create*Goal call for host-resident arguments (see tgx_plan_samples); h_out14 receives the 14 channels. */
int tgx_sample_host(tgx_engine* e, const tgx_params* h_params, double v, double accel, double s0, double s1,
                    double* h_out14);

/* Page-locked host memory, so that the D2H copies of the calls above run asynchronously at full PCIe rate.
 * tgx_alloc_host_for prefers pages of the NUMA node the engine's GPU is attached to (where the platform says which). */
void* tgx_alloc_host(int64_t bytes);
void* tgx_alloc_host_for(tgx_engine* e, int64_t bytes);
void tgx_free_host(void* p);

/* How the host-buffer calls size themselves on this machine: the end-to-end rate is set by PCIe and by what the host's
 * memory absorbs, so the threads that write the constant planes are sized to this process's share of the host — the
 * CPUs it may run on (sched_getaffinity) divided by the processes on the node (environment LOCAL_WORLD_SIZE, or
 * TGX_LOCAL_RANKS), half of that, at most 8 (TGX_FILLER_THREADS overrides) — and bound to the CPUs of the GPU's NUMA
 * node when sysfs names one. */
typedef struct tgx_host_info_t {
    int32_t numa_node;        /* NUMA node of the GPU's PCIe slot; -1: unknown (single-node hosts, most VMs) */
    int32_t cpus_allowed;     /* CPUs this process may run on */
    int32_t local_ranks;      /* processes sharing the host */
    int32_t filler_threads;   /* host threads tgx_generate_host uses for the constant planes */
    int32_t filler_cpus;      /* CPUs they may run on */
    int32_t reserved[3];
} tgx_host_info_t;
int tgx_host_info(tgx_engine* e, tgx_host_info_t* out);

/* ---- multi-GPU: contiguous block partition, no data-path collective, one optional flag gather ------------ */
/* Rank `rank` of `world` owns trajectories [*lo, *hi) of a batch of n. */
int tgx_shard_range(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi);

/* The only exchange of the path (BASELINE.json configs[4]: "NCCL gather of feasibility flags"): every GPU reduces its
 * own shard with tgx_feasibility and the 1-byte flags are all-gathered over NVLink.  A tgx_comm wraps one NCCL
 * communicator rank; NCCL is bound at run time (dlopen of libnccl.so.2 — a process that already carries one, e.g.
 * torch's, shares it), and every entry point below returns TGX_ERR_COMM with tgx_comm_last_error() set if it is
 * missing or a call fails.
 *   one process per GPU:   rank 0 calls tgx_comm_unique_id and ships the 128 bytes to the other ranks by any means
 *                          (bench.py: torch.distributed's store), then every rank calls tgx_comm_init_rank;
 *   one process, all GPUs: tgx_comm_init_all(comms, ndev, devices) (ncclCommInitAll); the per-device
 *                          tgx_gather_flags calls are then bracketed by tgx_comm_group_start / _end. */
#define TGX_COMM_ID_BYTES 128
typedef struct tgx_comm tgx_comm;
const char* tgx_comm_last_error(void);
int tgx_comm_nccl_version(int* version);                       /* e.g. 22809 */
int tgx_comm_unique_id(char id[TGX_COMM_ID_BYTES]);
int tgx_comm_init_rank(tgx_comm** out, int world, int rank, const char id[TGX_COMM_ID_BYTES], int device);
int tgx_comm_init_all(tgx_comm** out /* [ndev] */, int ndev, const int* devices /* NULL: 0..ndev-1 */);
int tgx_comm_destroy(tgx_comm* c);
int tgx_comm_group_start(void);
int tgx_comm_group_end(void);
/* All-gather of per-shard flag vectors into the full vector: the calling rank holds d_local[hi - lo] for its
 * tgx_shard_range(n_total, rank, world) and receives d_all[n_total] (both on the communicator's device), ordered on
 * `stream`.  Equal shards: one ncclAllGather; shards that differ by one trajectory: one ncclBroadcast per shard
 * inside a group.  c == NULL or a world of 1: a device-to-device copy. */
int tgx_gather_flags(tgx_comm* c, const uint8_t* d_local, int64_t n_total, uint8_t* d_all, void* stream);

/* ---- synthetic parameters drawn on the device (BASELINE.json configs[3]-[4]) ----------------------------- */
/* Writes records first_index .. first_index + n - 1 of the config-4 Monte-Carlo distribution (circles, r ~ U[0.2, 5],
 * centre ~ U[-2, 2]^2, alt ~ U[1, 2.5], v_goal ~ U[0.2, 8], accel ~ U[0.7, 2], dt = 0.01,
 * t_traj = max(9.98 - 2 v / a, 0.5)) to d_params, record i from Philox4x32-10 with key = seed and counter = (i, draw):
 * a shard of the 10^8-trajectory sweep needs no host->device parameter copy, and any record can be reproduced on
 * the host (trajectory_generator_ros2_b200/workloads.py: montecarlo_philox, bit-identical). */
int tgx_fill_montecarlo(tgx_engine* e, uint64_t seed, int64_t first_index, int64_t n, tgx_params* d_params,
                        void* stream);

/* ---- measurement probe: the FP64 roofline denominator of the reduction-only path -------------------------------- */
/* tgx_feasibility writes 17 bytes per trajectory: it is bound by the FP64 pipe and by instruction issue, not by HBM, and
 * MEASURED_PEAKS.json has no FP64 figure, so the peak is measured where the sweep runs: a hand-written
 * kernel of independent DFMA chains (8 per thread, one full wave of 1024 threads per SM, ~20 ms per launch), best of
 * `reps` launches after one warm-up, timed with CUDA events on the default stream.  *dfma_per_s = thread-level DFMA
 * instructions per second (x 2 = FLOP/s). */
int tgx_probe_dfma(tgx_engine* e, int reps, double* dfma_per_s, double* ms_per_launch);
/* The ceiling of the host-buffer calls (tgx_generate_host*): `reps` plain device->host copies of `bytes` into the
 * caller's page-locked buffer on the engine's copy stream, timed with CUDA events; *seconds = time of all reps.
 * bench.py runs it on every rank concurrently, so the ceiling it reports is what PCIe AND the host's memory absorb
 * from N GPUs at once. */
int tgx_probe_d2h(tgx_engine* e, void* h_dst, int64_t bytes, int reps, double* seconds);

/* ---- introspection used by bench.py ---------------------------------------------------------------- */
/* Number of this library's own (hand-written) kernels launched on this engine since creation: plan_count,
 * build_cur_table, plan_fill, eval, feasibility_finalize.  The cub scans inside tgx_plan are not counted. */
int64_t tgx_launch_count(const tgx_engine* e);
/* Tiles / segments of the current plan (0 if none). */
int64_t tgx_plan_tiles(const tgx_engine* e);
int64_t tgx_plan_segments(const tgx_engine* e);

/* Debug aid: checks the planner's hoisted-reciprocal division against IEEE division on n*per_thread pseudo-random
 * operand pairs on the GPU; *mismatches must come back 0. */
int tgx_selftest_division(tgx_engine* e, int64_t n, uint64_t seed, int per_thread, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif

#endif /* TGX_H_ */

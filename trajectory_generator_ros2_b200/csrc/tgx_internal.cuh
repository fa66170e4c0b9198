// tgx_internal.cuh — device-side records shared by the planning and evaluation kernels of libtgx.
//
// Data layout in HBM:
//
//   tgx_params[n]   caller's parameter records, 128 B each (one cache line per trajectory)
//   TrajRec[n]      per-trajectory constants the evaluation kernel needs, 64 B each
//   Seg[n_seg]      "linear-v segments": maximal runs of samples inside ONE phase of ONE tile over which
//                   v_k is an arithmetic progression.  Each carries the exact (bit-for-bit replayed) state
//                   of the reference's recurrence at its base sample, so the closed form inside a segment
//                   never drifts by more than tile_size roundings from the reference's running sums.
//   Tile[n_tile]    work list: one entry per (trajectory, block of tile_size consecutive samples)
//
// The planning kernels (plan.cu) build TrajRec/Seg/Tile by replaying the reference's scalar recurrences
// (Circle.cpp:47-82, Line.cpp:46-68, Figure8.cpp:47-82) one thread per trajectory; the evaluation kernel
// (eval.cu) consumes one Tile per CTA.
#pragma once

#include <cstddef>
#include <cstdint>

#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

#include "../../include/tgx.h"

namespace tgx {

// type field of TrajRec: low byte = the evaluation formula (TGX_CIRCLE, TGX_LINE, TGX_FIGURE8; a Boomerang is planned
// as a TGX_LINE whose return leg has negative speeds).
constexpr int32_t kRecTypeMask = 0xff;

struct __align__(16) TrajRec {
    int32_t type;
    int32_t n;        // sample count (0: rejected / empty)
    double f[7];
    // orbit (Circle, Figure8): f0 = r, f1 = cx, f2 = cy, f3 = alt, f4 = dt / r, f5 = 1 / r
    // line:                    f0 = cos(theta), f1 = sin(theta), f2 = theta, f3 = alt, f4 = dt
};
static_assert(sizeof(TrajRec) == 64, "TrajRec must be 64 bytes");

constexpr int32_t kSegClampLast = 1;   // sample kb+n has v == vclamp exactly (the std::min / std::max clamp fired)
constexpr int32_t kSegForcePos = 2;    // line: a one-sample segment whose position is (s0, s1) itself: the reference
                                       // overwrites a leg's last position with B / A (Line.cpp:81-82,
                                       // Boomerang.cpp:81-82,131-132)

struct __align__(16) Seg {
    int32_t kb;       // base sample index; the segment covers samples kb+1 .. kb+n (j = k - kb in 1..n);
                      // the first segment of a generateTraj plan has kb = 0 and also serves sample 0 (j = 0)
    int32_t n;
    int32_t flags;
    int32_t pad;
    double vb;        // speed at the base sample (exact replay)
    double dv;        // signed speed increment per step: +accel*dt, 0, or -accel*dt (the rounded product)
    double vclamp;    // v_goal (ramp-up) or 0 (ramp-down)
    double s0;        // orbit: theta at the base sample;  line: x at the base sample   (exact replay)
    double s1;        // orbit: theta increment per step at the base speed (a hold's exact progression step, or
                      //        (vb/r)*dt for a ramp);  line: y at the base sample
    double acc;       // orbit: theta at the segment's LAST sample kb+n (exact replay);
                      // line: the `accel` argument createLineGoal receives on this phase (+a1, 0, -a3)
};
static_assert(sizeof(Seg) == 64, "Seg must be 64 bytes");

struct __align__(16) Tile {
    int32_t traj;       // trajectory index
    int32_t k_lo;       // first sample of the tile (multiple of tile_size)
    int32_t seg_begin;  // index of the tile's first segment in Seg[]
    int32_t nseg;       // number of segments that intersect the tile (all lie inside it)
};
static_assert(sizeof(Tile) == 16, "Tile must be 16 bytes");

// A ramp is cut into chunks of at most kRampChunk steps, each starting from the exactly replayed state, which
// bounds the closed form's rounding drift against the reference's running sums to kRampChunk half-ulps.
constexpr int kRampChunk = 256;
// Fast planning mode: a segment holds at most kRebase steps of one phase — the closed form inside a segment starts from
// the replayed state at its base, so its drift against the reference's running sums is bounded by kRebase half-ulps
// of the state (2048 x 1.1e-16 x |theta|: 2e-11 rad at theta = 100).  Segments are cut by this rule and by the phases
// alone, never by the evaluation kernel's tile size: the samples do not depend on the tuning.
constexpr int kRebase = 2048;

// Shared-memory segment table of the evaluation kernel.  Mandatory segments in one tile: every phase of a
// trajectory (2*K ramps/holds + ramp-down = 17 for K = 8) could start inside the same tile, plus ramp chunks
// (tile/256 <= 8).  Optional segments (exact-progression breaks inside holds, plan.cu) are only emitted while the tile
// holds fewer than kMaxOptionalSegPerTile segments.  The planner ENFORCES the bound: a trajectory one of whose tiles
// would need more than kMaxSegPerTile segments (dozens of speed goals inside 1024 samples) is rejected with
// TGX_ST_TOO_LONG instead of being evaluated from a truncated list.
constexpr int kMaxSegPerTile = 64;
constexpr int kMaxOptionalSegPerTile = 36;

// Where the evaluation kernel finds its plan: three dense arrays TrajRec[n], Tile[n_tile], Seg[n_seg].
//   exact-offset plans (tile_slab == 0): a CTA reads its Tile, then the record and the segments it points to (two
//       dependent rounds of loads);
//   slab plans (tile_slab > 0): trajectory i owns Tile[i*tile_slab ..] and Seg[i*seg_slab ..], so everything is found
//       from blockIdx alone (traj = blockIdx / tile_slab) and the first tile of a trajectory is staged in ONE round of
//       independent 16-byte loads (record + tile entry + the first kSlabSpecSegs segments, speculatively).
// Keeping these reads few and dense matters more than their size suggests: every DRAM read that lands in the middle
// of the kernel's write stream costs a write->read->write bus turnaround (tools/wbw: 1 % of read traffic costs 8 %
// of the write bandwidth; with cache-resident tables the same kernel writes 7.27 TB/s instead of 6.5 TB/s).
//   phase plans (phase != nullptr): batches of short orbits with at most kPhaseMaxGoals speed goals whose replay emits
//       at most kPhaseMaxSegs segments (one per ramp, one per binade of theta a hold passes through) and of plain lines
//       (at most kPhaseLineMaxSegs segments: ramp, cruise, ramp, forced end point).  The planner then writes one
//       self-contained 256-byte PhaseRec per trajectory — where each segment ends, the replayed state there, what kind
//       of segment it is, and the constants of the parameter record — instead of TrajRec + Seg + Tile records, and the
//       CTA rebuilds exactly the Seg records the table path would have read (build_phase_segment, eval.cu): same bits,
//       one round of 16 independent 16-byte loads.  The rare orbit with more than kPhaseBaseSegs segments (a slow first
//       goal speed: the hold starts at a tiny angle and passes through ten binades) keeps the rest in a PhaseExt row,
//       a second round of loads for that CTA alone.  There is no tile directory: CTA i serves trajectory i and walks
//       its samples (at most kPhaseMaxSamples) in passes, so a ragged batch costs no empty CTAs and no work list.
constexpr int kPhaseMaxGoals = 2;
constexpr int kPhaseBaseSegs = 12;     // segments held by the PhaseRec itself
constexpr int kPhaseMaxSegs = 20;      // ... and with the trajectory's PhaseExt row
constexpr int kPhaseKindHold = 2;      // orbits, 2 bits per segment: 0, 1 = ramp up to speed goal 0 / 1
constexpr int kPhaseKindDown = 3;
constexpr int kPhaseLineMaxSegs = 6;
constexpr int kPhaseLineUp = 0;        // lines, 3 bits per segment
constexpr int kPhaseLineHold = 1;
constexpr int kPhaseLineDown = 2;
constexpr int kPhaseLineForced = 3;    // the leg's last sample, position forced to B (Line.cpp:81-82)
// The constants of a phase plan's trajectory (12 doubles), shared by the record in HBM and its image in shared memory.
union PhaseConsts {
    struct {
        double r, cx, cy, alt;
        double dtr, rinv;     // dt / r and 1 / r, divided once by the planner instead of once per CTA
        double adt;           // the rounded product accel*dt the reference adds every step
        double vg[kPhaseMaxGoals];
        double w[kPhaseMaxGoals];     // (vg / r) * dt as the reference rounds it: what a hold adds to theta per step
        double spare;
    };
    struct {
        double lcos, lsin, ltheta, lalt, ldt;       // TrajRec.f[0..4] of a line
        double lvg;                                 // v_goal
        double ladt1, ladt3;                        // the rounded products a1*dt, a3*dt (Line.cpp:48, :67)
        double la1, la3;
        double lspare[2];
    };
};
struct __align__(16) PhaseRec {
    int32_t n;            // number of segments (0: rejected trajectory)
    int32_t type;         // TGX_CIRCLE / TGX_FIGURE8 / TGX_LINE
    uint32_t kinds[2];    // 2 (orbit) or 3 (line) bits per segment, 64 bits in all
    PhaseConsts c;
    int32_t key[kPhaseBaseSegs];     // key[q] = last sample of segment q (segment q starts after key[q-1], or at sample 0)
    union {
        double th[kPhaseBaseSegs];           // orbit: the replayed angle at sample key[q]
        double xy[kPhaseLineMaxSegs][2];     // line: the replayed position at the BASE sample of segment q (Seg.s0, Seg.s1)
    };
};
static_assert(sizeof(PhaseRec) == 256, "PhaseRec must be 256 bytes (two 128-byte lines)");
// Segments kPhaseBaseSegs .. kPhaseMaxSegs - 1 of trajectory i (written and read only when PhaseRec.n > kPhaseBaseSegs).
struct __align__(16) PhaseExt {
    int32_t key[kPhaseMaxSegs - kPhaseBaseSegs];
    double th[kPhaseMaxSegs - kPhaseBaseSegs];
};
static_assert(sizeof(PhaseExt) == 96, "PhaseExt must be 96 bytes");
// What the evaluation CTA keeps of both in shared memory: the 16-byte chunks of the two records land so that key[] and
// th[] each run on from the record into the extension row (stage_phase_plan, eval.cu).
struct __align__(16) PhasePlan {
    int32_t n, type;
    uint32_t kinds[2];
    PhaseConsts c;
    int32_t key[kPhaseMaxSegs];
    union {
        double th[kPhaseMaxSegs];
        double xy[kPhaseLineMaxSegs][2];
    };
};
static_assert(sizeof(PhasePlan) == sizeof(PhaseRec) + sizeof(PhaseExt), "PhasePlan is PhaseRec + PhaseExt");
static_assert(offsetof(PhaseRec, key) % 16 == 0 && offsetof(PhaseRec, th) % 16 == 0 && offsetof(PhasePlan, th) % 16 == 0 &&
              offsetof(PhaseExt, th) % 16 == 0, "the chunks of key[] / th[] must not straddle");

struct TableView {
    const TrajRec* recs;
    const Seg* segs;
    const Tile* tiles;
    int seg_slab;
    int tile_slab;                  // > 0: slab plan
    const PhaseRec* phase;          // phase plan
    const PhaseExt* phase_ext;
};

// Longest trajectory a phase plan accepts (one CTA walks all of it).
constexpr int kPhaseMaxSamples = 4096;

constexpr int kSlabSpecSegs = 4;   // segments fetched speculatively with the record in a slab plan

// Statistics of one plan, accumulated by the fill pass and read back by the host (one small D2H per plan).
struct PlanStats {
    unsigned long long total_samples;   // sum of N_i
    unsigned long long total_tiles;     // sum of ceil(N_i / tile)
    int max_nseg;                       // largest segment count of a trajectory
    int max_ntile;                      // largest tile count of a trajectory
    int overflow;                       // slab / phase mode: some trajectory did not fit
    int phase_misfit;                   // some trajectory cannot be written as a PhaseRec (see plan_fill_kernel)
    int max_n;                          // largest sample count of a trajectory
    int has_line;                       // some trajectory is a Line / Boomerang
    int kinds;                          // bit mask of the replay classes seen (replay_class(): orbit x K, line, boomerang)
    int max_seg_len;                    // longest segment of the plan
    int max_tile_segs;                  // longest segment list of a tile
};

// ---- constant-speed polyline family (Square / Rectangle / Reciprocating / Bounce / M / I / T) -------------------
// The samples of such a trajectory follow a periodic pattern of legs: leg l contributes the samples
// i = i0 .. steps of `frac = i / steps; p = start + frac * (end - start)` (Square.cpp:73-77, M.cpp:52-55), and the
// number of samples is the number of `t += dt` iterations (the same counter as a hold phase).  The planner
// (polyline.cu) writes one PolyHead + up to 10 PolyLeg + 2 special one-sample records per trajectory, kPolyRecs * 64
// bytes at a fixed stride; the evaluation CTA stages them in one round of independent 16-byte loads.
struct __align__(16) PolyLeg {
    double sx, sy;        // start of the leg (Bounce: sx = z_start)
    double dx, dy;        // end - start, the difference the reference recomputes for every sample
    double heading;       // goal.psi
    double vx, vy;        // v*cos(heading), v*sin(heading)  (Bounce: vx = vz)
    int32_t steps;        // frac = (double)i / steps
    int32_t i0;           // first i of the leg: 1 for Square / Rectangle sides, 0 otherwise
};
static_assert(sizeof(PolyLeg) == 64, "PolyLeg must be 64 bytes");

struct __align__(16) PolyHead {
    int32_t type;
    int32_t n;              // sample count (0: rejected / empty)
    int32_t n_legs;
    int32_t first_special;  // 1: sample 0 is the record in slot kPolySlotFirst
    int32_t last_special;   // 1: sample n-1 is the record in slot kPolySlotLast
    int32_t period;         // sum over legs of (steps + 1 - i0)
    int32_t pad[2];
    double c0, c1;          // planar types: c0 = alt;  Bounce: c0 = cx, c1 = cy
    double pad2[2];
};
static_assert(sizeof(PolyHead) == 64, "PolyHead must be 64 bytes");

constexpr int kPolySlotLeg0 = 1;
constexpr int kPolySlotFirst = 1 + TGX_POLY_MAX_LEGS;
constexpr int kPolySlotLast = 2 + TGX_POLY_MAX_LEGS;
constexpr int kPolyRecs = 3 + TGX_POLY_MAX_LEGS;     // 13 records of 64 bytes = 832 bytes per trajectory

// Where the polyline evaluation kernel finds its work: tile t of trajectory i is CTA i*tile_slab + t (dense batches)
// or the Tile list entry blockIdx.x (ragged batches; only Tile.traj and Tile.k_lo are used).
struct PolyView {
    const int4* recs;       // kPolyRecs * 4 int4 per trajectory
    const Tile* tiles;      // nullptr: slab addressing
    int tile_slab;
};

// TrajRec.type of a braking plan of the polyline family (Square.cpp:112-137 and its copies; Bounce.cpp:74-103): the
// position is frozen at the setpoint being braked from, the velocity and acceleration point along a fixed direction.
//   f0, f1, f2 = position;  f3 = psi;  f4, f5, f6 = unit direction (cos heading, sin heading, 0) or (0, 0, 1)
//   Seg: v_k as for any ramp-down; Seg.acc = the signed acceleration magnitude along the direction (-decel, or 0)
constexpr int32_t kRecStatic = 16;

// Device-side view of tgx_layout.
struct OutView {
    double* base;
    int64_t traj_stride;
    int64_t chan_stride;
    const int64_t* traj_offset;
    int64_t capacity;
    uint32_t channel_mask;
};

// Where the evaluation kernels write array-of-structs records (tgx_eval_records): see RecStager / RecTma in store.cuh.
// tmap: the record buffer as a 2-D tensor [rows = records][16 doubles], box = 32 records, 128-byte swizzle — the
// TMA descriptor eval_kernel's warps store their staged records through (filled by tgx_eval_records on the host).
struct RecOut {
    alignas(64) CUtensorMap tmap;
    tgx_goal_record* base;
    int64_t stride;            // records per trajectory (ignored when offset != nullptr)
    const int64_t* offset;     // optional per-trajectory record offsets
    int64_t capacity;          // records that fit per trajectory
    int64_t total;             // records the buffer holds (= the tensor map's outer extent): nothing is written beyond
    double box[6];
    int clamp;                 // saturate p to box (TrajectoryGenerator.cpp:602-604)
};

}  // namespace tgx

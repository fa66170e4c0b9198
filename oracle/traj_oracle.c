/*
 * traj_oracle.c — CPU ORACLE (test infrastructure, NOT product code; see traj_oracle.h).
 *
 * Plain-C restatement of the reference samplers.  Every function cites the reference file:line it
 * follows.  Arithmetic is written in the reference's own operation order (C left-to-right evaluation,
 * std::min / std::max argument order, pow() calls kept as pow()) so that, built with
 * -O2 -ffp-contract=off against the same libm, it is bit-identical to the reference's output.
 */
#define _GNU_SOURCE
#include "traj_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- sample sink ------------------------------------------------------------------------------------ */

typedef struct {
    double* out;          /* may be NULL */
    int64_t chan_stride;
    int64_t cap;
    int64_t n;            /* samples pushed so far (== goals.size()) */
    double last[TGX_NCHAN]; /* goals.back() */
} sink_t;

static void sink_push(sink_t* s, const double g[TGX_NCHAN]) {
    if (s->out && s->n < s->cap) {
        for (int c = 0; c < TGX_NCHAN; ++c) s->out[c * s->chan_stride + s->n] = g[c];
    }
    memcpy(s->last, g, sizeof(s->last));
    s->n += 1;
}

/* index_msgs entry: a trajectory with more than 8 goal speeds owns the tgx_phases rows of its continuation records too
 * (tgx.h: TGX_VGOALS_MORE); entry e goes to slot e % 18 of row e / 18.  g_ph_rows is set by the entry points. */
static __thread int g_ph_rows = 1;
/* records available from the one an entry point was handed (the batch drivers know; single calls trust the caller) */
static __thread int64_t g_navail = INT64_MAX;

static void phase_add(tgx_phases* ph, int64_t key, int kind, double value, double value2) {
    if (!ph) return;
    int b = 0;
    while (ph[b].n >= TGX_MAX_PHASES && b + 1 < g_ph_rows) ++b;
    ph += b;
    if (ph->n >= TGX_MAX_PHASES) return;
    ph->key[ph->n] = (int32_t)key;
    ph->kind[ph->n] = kind;
    ph->value[ph->n] = value;
    ph->value2[ph->n] = value2;
    ph->n += 1;
}

static void phases_clear(tgx_phases* ph) {
    if (!ph) return;
    for (int b = 0; b < g_ph_rows; ++b) ph[b].n = 0;
}

/* Goal speed g of an orbit record: beyond the eighth they live in the continuation records that follow it. */
static double orbit_goal_speed(const tgx_params* p, int g) {
    return g < TGX_MAX_VGOALS ? p->u.orbit.v_goals[g] : p[g >> 3].u.orbit.v_goals[g & 7];
}

/* std::min(a, b) / std::max(a, b) exactly as libstdc++ defines them. */
static double std_min(double a, double b) { return (b < a) ? b : a; }
static double std_max(double a, double b) { return (a < b) ? b : a; }

/* ---- create*Goal ------------------------------------------------------------------------------------ */

/* Circle::createCircleGoal, Circle.cpp:96-130 (accel is accepted and unused, :113-114). */
static void circle_goal(const tgx_params* p, double v, double theta, double g[TGX_NCHAN]) {
    const tgx_orbit_params* o = &p->u.orbit;
    double s = sin(theta);
    double c = cos(theta);
    double v2r = pow(v, 2) / o->r;
    double v3r2 = pow(v, 3) / pow(o->r, 2);
    double omega = v / o->r;
    g[TGX_PX] = o->cx + o->r * c;
    g[TGX_PY] = o->cy + o->r * s;
    g[TGX_PZ] = p->alt;
    g[TGX_VX] = -v * s;
    g[TGX_VY] = v * c;
    g[TGX_VZ] = 0;
    g[TGX_AX] = -v2r * c;
    g[TGX_AY] = -v2r * s;
    g[TGX_AZ] = 0;
    g[TGX_JX] = v3r2 * s;
    g[TGX_JY] = -v3r2 * c;
    g[TGX_JZ] = 0;
    g[TGX_PSI] = theta + M_PI_2;
    g[TGX_DPSI] = omega;
}

/* Figure8::createFigure8Goal, Figure8.cpp:96-128. */
static void figure8_goal(const tgx_params* p, double v, double theta, double g[TGX_NCHAN]) {
    const tgx_orbit_params* o = &p->u.orbit;
    double s = sin(theta);
    double c = cos(theta);
    double sc = s * c;
    double omega = v / o->r;
    g[TGX_PX] = o->cx + o->r * s;
    g[TGX_PY] = o->cy + o->r * s * c;
    g[TGX_PZ] = p->alt;
    g[TGX_VX] = o->r * omega * c;
    g[TGX_VY] = o->r * omega * (c * c - s * s);
    g[TGX_VZ] = 0;
    g[TGX_AX] = -o->r * omega * omega * s;
    g[TGX_AY] = -4 * o->r * omega * omega * sc;
    g[TGX_AZ] = 0;
    g[TGX_JX] = 0;
    g[TGX_JY] = 0;
    g[TGX_JZ] = 0;
    g[TGX_PSI] = atan2(g[TGX_VY], g[TGX_VX]);
    g[TGX_DPSI] = omega;
}

/* Line::createLineGoal, Line.cpp:91-115 (sin/cos re-evaluated per call on the same theta). */
static void line_goal(const tgx_params* p, double last_x, double last_y, double v, double accel,
                      double theta, double g[TGX_NCHAN]) {
    double s = sin(theta);
    double c = cos(theta);
    g[TGX_PX] = last_x + v * c * p->dt;
    g[TGX_PY] = last_y + v * s * p->dt;
    g[TGX_PZ] = p->alt;
    g[TGX_VX] = v * c;
    g[TGX_VY] = v * s;
    g[TGX_VZ] = 0;
    g[TGX_AX] = accel * c;
    g[TGX_AY] = accel * s;
    g[TGX_AZ] = 0;
    g[TGX_JX] = 0;
    g[TGX_JY] = 0;
    g[TGX_JZ] = 0;
    g[TGX_PSI] = theta;
    g[TGX_DPSI] = 0;
}

static void orbit_goal(const tgx_params* p, double v, double theta, double g[TGX_NCHAN]) {
    if (p->type == TGX_FIGURE8) figure8_goal(p, v, theta, g);
    else circle_goal(p, v, theta, g);
}

/* ---- parameter validation ----------------------------------------------------------------------------
 * Mirrors the node-side checks that precede construction (TrajectoryGenerator.cpp:184-195 "All velocities
 * must be > 0", "accel must be > 0"; :268-277 for Line) plus the conditions under which the reference's
 * loops cannot terminate or divide by zero (dt <= 0, r <= 0, non-finite input). */
static int finite_pos(double x) { return isfinite(x) && x > 0.0; }

static int params_ok(const tgx_params* p) {
    if (!finite_pos(p->dt) || !isfinite(p->alt)) return 0;
    if (p->type == TGX_CIRCLE || p->type == TGX_FIGURE8) {
        const tgx_orbit_params* o = &p->u.orbit;
        /* the reference takes a vector of any length (an empty one gives the start sample alone) and any non-zero
         * radius; the caller guarantees that the continuation records of n_vgoals > 8 follow p in memory */
        if (p->n_vgoals < 0 || p->n_vgoals > TGX_MAX_VGOALS_TOTAL) return 0;
        if (!isfinite(o->r) || o->r == 0.0 || !finite_pos(o->accel)) return 0;
        if (!isfinite(o->cx) || !isfinite(o->cy) || !isfinite(o->t_traj)) return 0;
        if (TGX_ORBIT_RECORDS(p->n_vgoals) > g_navail) return 0;
        for (int q = 1; q < TGX_ORBIT_RECORDS(p->n_vgoals); ++q)
            if (p[q].type != TGX_VGOALS_MORE) return 0;
        for (int i = 0; i < p->n_vgoals; ++i)
            if (!finite_pos(orbit_goal_speed(p, i))) return 0;
        return 1;
    }
    if (p->type == TGX_LINE || p->type == TGX_BOOMERANG) {
        const tgx_line_params* l = &p->u.line;
        for (int i = 0; i < 3; ++i)
            if (!isfinite(l->A[i]) || !isfinite(l->B[i])) return 0;
        return finite_pos(l->v_goal) && finite_pos(l->a1) && finite_pos(l->a3);
    }
    if (TGX_IS_POLYLINE(p->type)) {
        /* v > 0: "All velocities must be > 0" (TrajectoryGenerator.cpp:227-232, 306-311, ...); accel > 0 (:241-244,
         * :251-254, :275-278).  A leg of length 0 would make the reference divide 0/0 (frac = i / steps with steps == 0). */
        const tgx_polyline_params* q = &p->u.poly;
        if (!finite_pos(q->v_goal) || !isfinite(q->t_traj) || !isfinite(q->orientation)) return 0;
        for (int i = 0; i < 5; ++i)
            if (!isfinite(q->g[i])) return 0;
        switch (p->type) {
            case TGX_SQUARE: return finite_pos(q->g[0]) && finite_pos(q->decel);
            case TGX_RECTANGLE: return finite_pos(q->g[0]) && finite_pos(q->g[1]) && finite_pos(q->decel);
            case TGX_RECIPROCATING:
                return isfinite(q->g[5]) && finite_pos(q->decel) && (q->g[0] != q->g[3] || q->g[1] != q->g[4]);
            case TGX_BOUNCE: return q->g[2] != q->g[3];
            default: return finite_pos(q->g[2]) && finite_pos(q->g[3]);   /* M, I, T: length, width */
        }
    }
    return 0;
}

/* ---- Line helpers ----------------------------------------------------------------------------------- */

/* (B_ - A_).norm(): Eigen's 3-element reduction is x^2 + (y^2 + z^2), then sqrt (Line.cpp:157,176). */
static double line_length(const tgx_line_params* l) {
    double dx = l->B[0] - l->A[0];
    double dy = l->B[1] - l->A[1];
    double dz = l->B[2] - l->A[2];
    return sqrt(dx * dx + (dy * dy + dz * dz));
}

/* Line::get_d2, Line.cpp:175-181. */
double orc_line_d2(const tgx_params* p) {
    const tgx_line_params* l = &p->u.line;
    double d = line_length(l);
    double vg = l->v_goal;
    double d1 = 0.5 * vg * vg / l->a1;
    double d3 = 0.5 * vg * vg / l->a3;
    return d - d1 - d3;
}

/* ---- generateTraj ----------------------------------------------------------------------------------- */

/* Circle::generateTraj, Circle.cpp:30-94 == Figure8::generateTraj, Figure8.cpp:30-94. */
static int64_t orbit_generate(const tgx_params* p, sink_t* sk, uint32_t* status, tgx_phases* ph,
                              int64_t max_samples) {
    const tgx_orbit_params* o = &p->u.orbit;
    double g[TGX_NCHAN];
    double theta = 0;
    double v = 0;
    orbit_goal(p, v, theta, g);                       /* :41 */
    sink_push(sk, g);
    for (int i = 0; i < p->n_vgoals; ++i) {           /* :43 */
        double v_goal = orbit_goal_speed(p, i);
        phase_add(ph, sk->n - 1, TGX_PH_ACCEL_TO, v_goal, 0.0);   /* :45 */
        while (v < v_goal) {                          /* :47 */
            double v_new = std_min(v + o->accel * p->dt, v_goal);
            if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
            v = v_new;
            double omega = v / o->r;
            theta += omega * p->dt;
            orbit_goal(p, v, theta, g);
            sink_push(sk, g);
        }
        if (fabs(v - v_goal) > 0.001) *status |= TGX_ST_VGOALS_NOT_INCREASING;   /* :57-59 */
        phase_add(ph, sk->n - 1, TGX_PH_REACHED, v_goal, o->t_traj);             /* :61-62 */
        double current_t_traj = 0;
        while (current_t_traj < o->t_traj) {          /* :63 */
            if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
            double omega = v / o->r;
            theta += omega * p->dt;
            orbit_goal(p, v, theta, g);
            sink_push(sk, g);
            double t_new = current_t_traj + p->dt;
            if (t_new == current_t_traj) { *status |= TGX_ST_TOO_LONG; return -1; }
            current_t_traj = t_new;
        }
    }
    phase_add(ph, sk->n - 1, TGX_PH_DECEL, 0.0, 0.0); /* :74 */
    while (v > 0) {                                   /* :75 */
        double v_new = std_max(v - o->accel * p->dt, 0.0);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        double omega = v / o->r;
        theta += omega * p->dt;
        orbit_goal(p, v, theta, g);
        sink_push(sk, g);
    }
    if (fabs(v) > 0.001) *status |= TGX_ST_FINAL_V_NONZERO;       /* :85-88 (exit(1) in the reference) */
    phase_add(ph, sk->n - 1, TGX_PH_STOPPED, 0.0, 0.0);           /* :89 */
    return sk->n;
}

/* Line::Line (theta_, Line.cpp:24) + Line::generateTraj, Line.cpp:31-89. */
static int64_t line_generate(const tgx_params* p, sink_t* sk, uint32_t* status, tgx_phases* ph,
                             int64_t max_samples) {
    const tgx_line_params* l = &p->u.line;
    double g[TGX_NCHAN];
    double theta = atan2(l->B[1] - l->A[1], l->B[0] - l->A[0]);   /* :24 */
    double v = 0;
    line_goal(p, l->A[0], l->A[1], v, 0, theta, g);   /* :40 */
    sink_push(sk, g);
    double v_goal = l->v_goal;                        /* :43 */
    phase_add(ph, sk->n - 1, TGX_PH_ACCEL_TO, v_goal, 0.0);
    while (v < v_goal) {                              /* :46 */
        double v_new = std_min(v + l->a1 * p->dt, v_goal);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, l->a1, theta, g);
        sink_push(sk, g);
    }
    double t2 = orc_line_d2(p) / v_goal;              /* :53 (a negative d2 simply skips the cruise loop) */
    phase_add(ph, sk->n - 1, TGX_PH_REACHED, v_goal, t2);
    double current_t_traj = 0;
    while (current_t_traj < t2) {                     /* :57 */
        if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, 0, theta, g);
        sink_push(sk, g);
        double t_new = current_t_traj + p->dt;
        if (t_new == current_t_traj) { *status |= TGX_ST_TOO_LONG; return -1; }
        current_t_traj = t_new;
    }
    phase_add(ph, sk->n - 1, TGX_PH_DECEL, 0.0, 0.0); /* :64 */
    while (v > 0) {                                   /* :65 */
        double v_new = std_max(v - l->a3 * p->dt, 0.0);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, -l->a3, theta, g);
        sink_push(sk, g);
    }
    double thresh = 0.05;                             /* :71-79 (exit(1) in the reference) */
    if (fabs(l->B[0] - sk->last[TGX_PX]) > thresh || fabs(l->B[1] - sk->last[TGX_PY]) > thresh)
        *status |= TGX_ST_LINE_END_NOT_B;
    /* Force last goal pos to be equal to B, :81-82 */
    sk->last[TGX_PX] = l->B[0];
    sk->last[TGX_PY] = l->B[1];
    if (sk->out && sk->n - 1 < sk->cap) {
        sk->out[TGX_PX * sk->chan_stride + (sk->n - 1)] = l->B[0];
        sk->out[TGX_PY * sk->chan_stride + (sk->n - 1)] = l->B[1];
    }
    phase_add(ph, sk->n - 1, TGX_PH_STOPPED, 0.0, 0.0);           /* :84 */
    return sk->n;
}

/* Boomerang::generateTraj, Boomerang.cpp:31-141: Line's A -> B leg (without its "stopped" announcement), then the
 * same profile back with negative speeds, ending forced at A. */
static int64_t boomerang_generate(const tgx_params* p, sink_t* sk, uint32_t* status, tgx_phases* ph,
                                  int64_t max_samples) {
    const tgx_line_params* l = &p->u.line;
    double g[TGX_NCHAN];
    double theta = atan2(l->B[1] - l->A[1], l->B[0] - l->A[0]);   /* :24 */
    double v = 0;
    line_goal(p, l->A[0], l->A[1], v, 0, theta, g);               /* :40 */
    sink_push(sk, g);
    double v_goal = l->v_goal;
    phase_add(ph, sk->n - 1, TGX_PH_ACCEL_TO, v_goal, 0.0);       /* :44 */
    while (v < v_goal) {                                          /* :46 */
        double v_new = std_min(v + l->a1 * p->dt, v_goal);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, l->a1, theta, g);
        sink_push(sk, g);
    }
    double t2 = orc_line_d2(p) / v_goal;                          /* :53 */
    phase_add(ph, sk->n - 1, TGX_PH_REACHED, v_goal, t2);
    double current_t_traj = 0;
    while (current_t_traj < t2) {                                 /* :58 */
        if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, 0, theta, g);
        sink_push(sk, g);
        double t_new = current_t_traj + p->dt;
        if (t_new == current_t_traj) { *status |= TGX_ST_TOO_LONG; return -1; }
        current_t_traj = t_new;
    }
    phase_add(ph, sk->n - 1, TGX_PH_DECEL, 0.0, 0.0);             /* :64 */
    while (v > 0) {                                               /* :65 */
        double v_new = std_max(v - l->a3 * p->dt, 0.0);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, -l->a3, theta, g);
        sink_push(sk, g);
    }
    double thresh = 0.05;                                         /* :71-79 */
    if (fabs(l->B[0] - sk->last[TGX_PX]) > thresh || fabs(l->B[1] - sk->last[TGX_PY]) > thresh)
        *status |= TGX_ST_LINE_END_NOT_B;
    sk->last[TGX_PX] = l->B[0];                                   /* :81-82 */
    sk->last[TGX_PY] = l->B[1];
    if (sk->out && sk->n - 1 < sk->cap) {
        sk->out[TGX_PX * sk->chan_stride + (sk->n - 1)] = l->B[0];
        sk->out[TGX_PY * sk->chan_stride + (sk->n - 1)] = l->B[1];
    }
    /* ---- return leg, :85-134 ---- */
    v = 0;
    if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
    line_goal(p, l->B[0], l->B[1], -v, 0, theta, g);              /* :90 */
    sink_push(sk, g);
    phase_add(ph, sk->n - 1, TGX_PH_ACCEL_TO, v_goal, 0.0);       /* :94 */
    while (fabs(v) < fabs(v_goal)) {                              /* :97 */
        double v_new = std_max(v - l->a1 * p->dt, -v_goal);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, -l->a1, theta, g);
        sink_push(sk, g);
    }
    t2 = orc_line_d2(p) / v_goal;                                 /* :105 */
    phase_add(ph, sk->n - 1, TGX_PH_REACHED, v_goal, t2);
    current_t_traj = 0;
    while (current_t_traj < t2) {                                 /* :110 */
        if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, 0, theta, g);
        sink_push(sk, g);
        double t_new = current_t_traj + p->dt;
        if (t_new == current_t_traj) { *status |= TGX_ST_TOO_LONG; return -1; }
        current_t_traj = t_new;
    }
    phase_add(ph, sk->n - 1, TGX_PH_DECEL, 0.0, 0.0);             /* :116 */
    while (v < 0) {                                               /* :117 */
        double v_new = std_min(v + l->a3 * p->dt, 0.0);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        line_goal(p, sk->last[TGX_PX], sk->last[TGX_PY], v, l->a3, theta, g);
        sink_push(sk, g);
    }
    if (fabs(l->A[0] - sk->last[TGX_PX]) > thresh || fabs(l->A[1] - sk->last[TGX_PY]) > thresh)   /* :126-129 */
        *status |= TGX_ST_LINE_END_NOT_B;
    sk->last[TGX_PX] = l->A[0];                                   /* :131-132 */
    sk->last[TGX_PY] = l->A[1];
    if (sk->out && sk->n - 1 < sk->cap) {
        sk->out[TGX_PX * sk->chan_stride + (sk->n - 1)] = l->A[0];
        sk->out[TGX_PY * sk->chan_stride + (sk->n - 1)] = l->A[1];
    }
    phase_add(ph, sk->n - 1, TGX_PH_STOPPED, 0.0, 0.0);           /* :134 */
    return sk->n;
}


/* ---- constant-speed polyline family ---------------------------------------------------
 * Square / Rectangle / Reciprocating / Bounce / M / I / T.  `leg_of` (may be NULL) receives, per sample, the leg of
 * the periodic pattern it belongs to in tgx_polyline_legs numbering (-1 for Square / Rectangle's sample 0), which is
 * what the reference's per-sample index_msgs strings are a function of. */

typedef struct {
    int16_t* leg_of;
    int64_t leg_cap;
} legsink_t;

static void leg_push(legsink_t* ls, int64_t k, int leg) {
    if (ls && ls->leg_of && k < ls->leg_cap) ls->leg_of[k] = (int16_t)leg;
}

/* createSquareGoal (Square.cpp:94-110) == createRectangleGoal (Rectangle.cpp:95-111) == createReciprocatingGoal
 * (Reciprocating.cpp:62-79) == createMGoal (M.cpp:69-86) == createIGoal (I.cpp:77-94) == createTGoal (T.cpp:75-92).
 * goal.j is never assigned and keeps the message default 0. */
static void planar_goal(const tgx_params* p, double x, double y, double v, double accel, double heading,
                        double g[TGX_NCHAN]) {
    g[TGX_PX] = x;
    g[TGX_PY] = y;
    g[TGX_PZ] = p->alt;
    g[TGX_VX] = v * cos(heading);
    g[TGX_VY] = v * sin(heading);
    g[TGX_VZ] = 0;
    g[TGX_AX] = accel * cos(heading);
    g[TGX_AY] = accel * sin(heading);
    g[TGX_AZ] = 0;
    g[TGX_JX] = 0;
    g[TGX_JY] = 0;
    g[TGX_JZ] = 0;
    g[TGX_PSI] = heading;
    g[TGX_DPSI] = 0;
}

/* Bounce::createBounceGoal, Bounce.cpp:54-72. */
static void bounce_goal(double x, double y, double z, double vz, double heading, double g[TGX_NCHAN]) {
    memset(g, 0, TGX_NCHAN * sizeof(double));
    g[TGX_PX] = x;
    g[TGX_PY] = y;
    g[TGX_PZ] = z;
    g[TGX_VZ] = vz;
    g[TGX_PSI] = heading;
}

/* double -> int conversion of a std::ceil result; out-of-range values (undefined behaviour in the reference) are
 * reported as "too long". */
static int ceil_to_int(double x, int* ok) {
    double c = ceil(x);
    if (!(c >= -2147483648.0 && c <= 1073741824.0)) { *ok = 0; return 0; }
    return (int)c;
}

/* Square::generateTraj (Square.cpp:21-92) and Rectangle::generateTraj (Rectangle.cpp:20-93). */
static int64_t square_generate(const tgx_params* p, sink_t* sk, legsink_t* ls, uint32_t* status, int64_t max_samples) {
    const tgx_polyline_params* q = &p->u.poly;
    const int rect = p->type == TGX_RECTANGLE;
    double g[TGX_NCHAN];
    double half_a = q->g[0] / 2.0;
    double half_b = rect ? q->g[1] / 2.0 : half_a;
    double cx = rect ? q->g[2] : q->g[1], cy = rect ? q->g[3] : q->g[2];
    double cxs[4] = {-half_a, half_a, half_a, -half_a};       /* :30-33 counter-clockwise from top-left */
    double cys[4] = {half_b, half_b, -half_b, -half_b};
    double c = cos(q->orientation);                           /* :37-38 */
    double s = sin(q->orientation);
    for (int i = 0; i < 4; ++i) {                             /* :39-44 */
        double x_new = c * cxs[i] - s * cys[i];
        double y_new = s * cxs[i] + c * cys[i];
        cxs[i] = x_new + cx;
        cys[i] = y_new + cy;
    }
    double v_goal = q->v_goal;                                /* :48 */
    double total_perimeter = rect ? 2 * (q->g[0] + q->g[1]) : 4 * q->g[0];   /* :49, Rectangle.cpp:49 */
    double time_per_lap = total_perimeter / v_goal;
    int ok = 1;
    int num_laps = ceil_to_int(q->t_traj / time_per_lap, &ok);   /* :51 */
    if (!ok) { *status |= TGX_ST_TOO_LONG; return -1; }
    double dt = p->dt;
    int current_corner = 0;
    double heading0 = q->orientation + M_PI;                  /* :57 */
    double t = 0.0;
    planar_goal(p, cxs[0], cys[0], v_goal, 0, heading0, g);   /* :60 */
    leg_push(ls, sk->n, -1);
    sink_push(sk, g);
    for (int lap = 0; lap < num_laps; ++lap) {                /* :64 */
        for (int side = 0; side < 4; ++side) {
            int next_corner = (current_corner + 1) % 4;
            double sx = cxs[current_corner], sy = cys[current_corner];
            double ex = cxs[next_corner], ey = cys[next_corner];
            double ddx = ex - sx, ddy = ey - sy;
            double side_length = sqrt(ddx * ddx + ddy * ddy);   /* (end - start).norm(), :69 */
            double side_time = side_length / v_goal;
            int steps = ceil_to_int(side_time / dt, &ok);       /* :71 */
            if (!ok) { *status |= TGX_ST_TOO_LONG; return -1; }
            for (int step = 1; step <= steps && t < q->t_traj; ++step) {   /* :73 */
                if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
                double frac = (double)step / steps;
                double x_new = sx + frac * (ex - sx);
                double y_new = sy + frac * (ey - sy);
                double heading = atan2(ey - sy, ex - sx);
                planar_goal(p, x_new, y_new, v_goal, 0, heading, g);
                leg_push(ls, sk->n, side);
                sink_push(sk, g);
                double t_new = t + dt;
                if (t_new == t) { *status |= TGX_ST_TOO_LONG; return -1; }
                t = t_new;
                if (t >= q->t_traj) break;
            }
            current_corner = next_corner;
        }
        if (!(t < q->t_traj)) break;   /* later laps emit nothing */
    }
    return sk->n;
}

/* Reciprocating::Reciprocating + generateTraj, Reciprocating.cpp:13-60. */
static int64_t recip_generate(const tgx_params* p, sink_t* sk, legsink_t* ls, uint32_t* status, int64_t max_samples) {
    const tgx_polyline_params* q = &p->u.poly;
    double g[TGX_NCHAN];
    double Ax = q->g[0], Ay = q->g[1], Bx = q->g[3], By = q->g[4];
    double theta_fwd = atan2(By - Ay, Bx - Ax);               /* :17-18 */
    double theta_rev = atan2(Ay - By, Ax - Bx);
    double v_goal = q->v_goal, dt = p->dt, t = 0.0;
    int forward = 1;
    double sx = Ax, sy = Ay, ex = Bx, ey = By;
    double heading = theta_fwd;
    while (t < q->t_traj) {                                   /* :36 */
        double ddx = ex - sx, ddy = ey - sy;
        double distance = sqrt(ddx * ddx + ddy * ddy);        /* (end - start).head<2>().norm(), :38 */
        int ok = 1;
        int steps = ceil_to_int(distance / (v_goal * dt), &ok);   /* :39 */
        if (!ok || steps < 1) { *status |= TGX_ST_TOO_LONG; return -1; }
        for (int i = 0; i <= steps && t < q->t_traj; ++i) {   /* :40 */
            if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
            double frac = (double)i / steps;
            double x = sx + frac * (ex - sx);
            double y = sy + frac * (ey - sy);
            planar_goal(p, x, y, v_goal, 0, heading, g);
            leg_push(ls, sk->n, forward ? 0 : 2);
            sink_push(sk, g);
            double t_new = t + dt;
            if (t_new == t) { *status |= TGX_ST_TOO_LONG; return -1; }
            t = t_new;
        }
        forward = !forward;                                   /* :50-52 */
        { double tx = sx, ty = sy; sx = ex; sy = ey; ex = tx; ey = ty; }
        heading = forward ? theta_fwd : theta_rev;
        if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        planar_goal(p, sx, sy, 0, 0, heading, g);             /* :54-56 yaw flip at the endpoint, unconditional */
        leg_push(ls, sk->n, forward ? 3 : 1);
        sink_push(sk, g);
        t += dt;
    }
    return sk->n;
}

/* Bounce::generateTraj, Bounce.cpp:19-52. */
static int64_t bounce_generate(const tgx_params* p, sink_t* sk, legsink_t* ls, uint32_t* status, int64_t max_samples) {
    const tgx_polyline_params* q = &p->u.poly;
    double g[TGX_NCHAN];
    double v_goal = q->v_goal, dt = p->dt, t = 0.0;
    int going_up = 1;
    double z_start = q->g[2], z_end = q->g[3];
    double heading = q->orientation;                          /* :32 */
    while (t < q->t_traj) {                                   /* :34 */
        double distance = fabs(z_end - z_start);
        int ok = 1;
        int steps = ceil_to_int(distance / (v_goal * dt), &ok);   /* :36 */
        if (!ok || steps < 1) { *status |= TGX_ST_TOO_LONG; return -1; }
        for (int i = 0; i <= steps && t < q->t_traj; ++i) {   /* :37 */
            if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
            double frac = (double)i / steps;
            double z = z_start + frac * (z_end - z_start);
            double vz = (z_end > z_start) ? v_goal : -v_goal;
            bounce_goal(q->g[0], q->g[1], z, vz, heading, g);
            leg_push(ls, sk->n, going_up ? 0 : 1);
            sink_push(sk, g);
            double t_new = t + dt;
            if (t_new == t) { *status |= TGX_ST_TOO_LONG; return -1; }
            t = t_new;
        }
        going_up = !going_up;                                 /* :46-47 */
        { double tz = z_start; z_start = z_end; z_end = tz; }
    }
    return sk->n;
}

/* M::generateTraj (M.cpp:13-67), I::generateTraj (I.cpp:19-75), T::generateTraj (T.cpp:19-73). */
static int64_t letter_generate(const tgx_params* p, sink_t* sk, legsink_t* ls, uint32_t* status, int64_t max_samples) {
    const tgx_polyline_params* q = &p->u.poly;
    double g[TGX_NCHAN];
    double cx = q->g[0], cy = q->g[1], length = q->g[2], width = q->g[3];
    double bx[6], by[6], px[6], py[6];
    int np;
    if (p->type == TGX_M) {                                   /* M.cpp:20-26 */
        np = 5;
        bx[0] = -width / 2; by[0] = -length / 2;
        bx[1] = -width / 2; by[1] = length / 2;
        bx[2] = 0.0;        by[2] = -length / 2;
        bx[3] = width / 2;  by[3] = length / 2;
        bx[4] = width / 2;  by[4] = -length / 2;
    } else if (p->type == TGX_I) {                            /* I.cpp:28-35 */
        np = 6;
        bx[0] = -width / 2; by[0] = length / 2;
        bx[1] = width / 2;  by[1] = length / 2;
        bx[2] = 0.0;        by[2] = length / 2;
        bx[3] = 0.0;        by[3] = -length / 2;
        bx[4] = -width / 2; by[4] = -length / 2;
        bx[5] = width / 2;  by[5] = -length / 2;
    } else {                                                  /* T.cpp:28-33 */
        np = 4;
        bx[0] = -width / 2; by[0] = length / 2;
        bx[1] = width / 2;  by[1] = length / 2;
        bx[2] = 0.0;        by[2] = length / 2;
        bx[3] = 0.0;        by[3] = -length / 2;
    }
    double c = cos(q->orientation);                           /* M.cpp:29-30 */
    double s = sin(q->orientation);
    for (int i = 0; i < np; ++i) {                            /* M.cpp:32-37, I.cpp:40-45, T.cpp:38-43 */
        px[i] = c * bx[i] - s * by[i] + cx;
        py[i] = s * bx[i] + c * by[i] + cy;
    }
    double v_goal = q->v_goal, dt = p->dt, t = 0.0;
    int forward = 1;
    while (t < q->t_traj) {                                   /* M.cpp:44 */
        for (int seg = 0; seg < np - 1 && t < q->t_traj; ++seg) {   /* :46 */
            int a = forward ? seg : np - 1 - seg;             /* reversed copy of the point list, :45 */
            int b = forward ? seg + 1 : np - 2 - seg;
            double sx = px[a], sy = py[a], ex = px[b], ey = py[b];
            double heading = atan2(ey - sy, ex - sx);         /* :49 */
            double ddx = ex - sx, ddy = ey - sy;
            double distance = sqrt(ddx * ddx + ddy * ddy);    /* :50 */
            int ok = 1;
            int steps = ceil_to_int(distance / (v_goal * dt), &ok);   /* :51 */
            if (!ok || steps < 1) { *status |= TGX_ST_TOO_LONG; return -1; }
            for (int i = 0; i <= steps && t < q->t_traj; ++i) {   /* :52 */
                if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
                double frac = (double)i / steps;
                double x = sx + frac * (ex - sx);
                double y = sy + frac * (ey - sy);
                planar_goal(p, x, y, v_goal, 0, heading, g);
                leg_push(ls, sk->n, forward ? seg : (np - 1) + seg);
                sink_push(sk, g);
                double t_new = t + dt;
                if (t_new == t) { *status |= TGX_ST_TOO_LONG; return -1; }
                t = t_new;
            }
        }
        forward = !forward;                                   /* :61 */
    }
    return sk->n;
}

static int64_t poly_generate(const tgx_params* p, sink_t* sk, legsink_t* ls, uint32_t* status, int64_t max_samples) {
    switch (p->type) {
        case TGX_SQUARE:
        case TGX_RECTANGLE: return square_generate(p, sk, ls, status, max_samples);
        case TGX_RECIPROCATING: return recip_generate(p, sk, ls, status, max_samples);
        case TGX_BOUNCE: return bounce_generate(p, sk, ls, status, max_samples);
        default: return letter_generate(p, sk, ls, status, max_samples);
    }
}

int64_t orc_polyline_generate(const tgx_params* p, double* out, int64_t chan_stride, int64_t cap, uint32_t* status,
                              int16_t* leg_of, int64_t max_samples) {
    uint32_t st = 0;
    int64_t n;
    sink_t sk = {out, chan_stride, cap, 0, {0}};
    legsink_t ls = {leg_of, cap};
    if (!TGX_IS_POLYLINE(p->type) || !params_ok(p)) {
        st |= TGX_ST_BAD_PARAM;
        n = -1;
    } else {
        n = poly_generate(p, &sk, &ls, &st, max_samples);
    }
    if (n > cap && out) st |= TGX_ST_TRUNCATED;
    if (status) *status = st;
    return n;
}

/* The reference's index_msgs text for sample k of a polyline trajectory with n samples, given its leg
 * (Square.cpp:61,79,88; Rectangle.cpp:61,79,88; Reciprocating.cpp:47,57; Bounce.cpp:42,50; M.cpp:57,65; I.cpp:65,73;
 * T.cpp:63,71).  Returns the length written. */
int orc_polyline_msg(int type, int leg, int64_t k, int64_t n, char* buf, int cap) {
    static const char* const letter[] = {"M", "I", "T"};
    const int last = (k == n - 1);
    switch (type) {
        case TGX_SQUARE:
        case TGX_RECTANGLE: {
            const char* name = type == TGX_SQUARE ? "Square" : "Rectangle";
            if (last) return snprintf(buf, (size_t)cap, "%s traj: completed", name);
            if (leg < 0) return snprintf(buf, (size_t)cap, "%s traj: starting at corner 0", name);
            return snprintf(buf, (size_t)cap, "%s traj: moving along side %d", name, leg);
        }
        case TGX_RECIPROCATING:
            if (leg == 1 || leg == 3) return snprintf(buf, (size_t)cap, "Reciprocating: yaw flip at endpoint");
            return snprintf(buf, (size_t)cap, leg == 0 ? "Reciprocating: forward" : "Reciprocating: reverse");
        case TGX_BOUNCE:
            if (last) return snprintf(buf, (size_t)cap, "Bounce: completed");
            return snprintf(buf, (size_t)cap, leg == 0 ? "Bounce: ascending" : "Bounce: descending");
        default: {
            const char* name = letter[type - TGX_M];
            const int nseg = type == TGX_M ? 4 : (type == TGX_I ? 5 : 3);
            if (last) return snprintf(buf, (size_t)cap, "%s traj: completed", name);
            return snprintf(buf, (size_t)cap, "%s traj: segment %d %s", name, leg % nseg, leg < nseg ? "fwd" : "rev");
        }
    }
}

/* Braking trajectories of the family.  Square.cpp:112-137, Rectangle.cpp:113-137, Reciprocating.cpp:81-107,
 * M.cpp:88-113, I.cpp:96-121, T.cpp:94-119: v = |v_xy|, heading = atan2(vy, vx), position frozen at the setpoint,
 * `while (v > 0) v = max(v - decel*dt, 0)`.  Bounce.cpp:74-103: vz *= 0.8 until |vz| <= 0.01. */
static int64_t poly_stop(const tgx_params* p, const double* from, sink_t* sk, uint32_t* status, int64_t max_samples) {
    const tgx_polyline_params* q = &p->u.poly;
    double g[TGX_NCHAN];
    if (p->type == TGX_BOUNCE) {
        double vz = from[TGX_VZ];                             /* Bounce.cpp:81-83 */
        double z = from[TGX_PZ];
        double heading = from[TGX_PSI];
        while (fabs(vz) > 0.01) {                             /* :91 */
            if (sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
            vz *= 0.8;
            if (fabs(vz) < 0.01) vz = 0.0;
            bounce_goal(q->g[0], q->g[1], z, vz, heading, g);
            sink_push(sk, g);
        }
        return sk->n;
    }
    double v = sqrt(pow(from[TGX_VX], 2) + pow(from[TGX_VY], 2));   /* Square.cpp:118 */
    double heading = atan2(from[TGX_VY], from[TGX_VX]);             /* :119 */
    double decel = (p->type == TGX_M || p->type == TGX_I || p->type == TGX_T) ? 1.0 : q->decel;   /* M.cpp:101 */
    while (v > 0) {                                                 /* :125 */
        double v_new = std_max(v - decel * p->dt, 0.0);
        if (v_new == v || sk->n >= max_samples) { *status |= TGX_ST_TOO_LONG; return -1; }
        v = v_new;
        planar_goal(p, from[TGX_PX], from[TGX_PY], v, -decel, heading, g);
        sink_push(sk, g);
    }
    return sk->n;
}

/* trajectoryInsideBounds of the family: Square.cpp:139-157, Rectangle.cpp:139-159, Reciprocating.cpp:109-122,
 * Bounce.cpp:105-120, M.cpp:115-145, I.cpp:123-153, T.cpp:121-149. */
static int point_inside(const double box[6], double x, double y, double z);

static int poly_inside_bounds(const tgx_params* p, const double box[6]) {
    const tgx_polyline_params* q = &p->u.poly;
    if (p->type == TGX_RECIPROCATING)
        return point_inside(box, q->g[0], q->g[1], q->g[2]) && point_inside(box, q->g[3], q->g[4], q->g[5]);
    if (p->type == TGX_BOUNCE)
        return point_inside(box, q->g[0], q->g[1], q->g[2]) && point_inside(box, q->g[0], q->g[1], q->g[3]);
    double c = cos(q->orientation);
    double s = sin(q->orientation);
    if (p->type == TGX_SQUARE || p->type == TGX_RECTANGLE) {
        const int rect = p->type == TGX_RECTANGLE;
        double ha = q->g[0] / 2.0, hb = rect ? q->g[1] / 2.0 : ha;
        double cx = rect ? q->g[2] : q->g[1], cy = rect ? q->g[3] : q->g[2];
        double xs[4] = {cx + c * -ha - s * hb, cx + c * ha - s * hb, cx + c * ha - s * -hb, cx + c * -ha - s * -hb};
        double ys[4] = {cy + s * -ha + c * hb, cy + s * ha + c * hb, cy + s * ha + c * -hb, cy + s * -ha + c * -hb};
        for (int i = 0; i < 4; ++i)
            if (!point_inside(box, xs[i], ys[i], p->alt)) return 0;
        return 1;
    }
    double cx = q->g[0], cy = q->g[1], length = q->g[2], width = q->g[3];
    double xs[6], ys[6];
    int np;
    if (p->type == TGX_M) {                                   /* M.cpp:120-126 */
        np = 5;
        xs[0] = cx - width / 2; ys[0] = cy - length / 2;
        xs[1] = cx - width / 2; ys[1] = cy + length / 2;
        xs[2] = cx;             ys[2] = cy - length / 2;
        xs[3] = cx + width / 2; ys[3] = cy + length / 2;
        xs[4] = cx + width / 2; ys[4] = cy - length / 2;
    } else if (p->type == TGX_I) {                            /* I.cpp:128-135 */
        np = 6;
        xs[0] = cx - width / 2; ys[0] = cy + length / 2;
        xs[1] = cx + width / 2; ys[1] = cy + length / 2;
        xs[2] = cx;             ys[2] = cy + length / 2;
        xs[3] = cx;             ys[3] = cy - length / 2;
        xs[4] = cx - width / 2; ys[4] = cy - length / 2;
        xs[5] = cx + width / 2; ys[5] = cy - length / 2;
    } else {                                                  /* T.cpp:126-131 */
        np = 4;
        xs[0] = cx - width / 2; ys[0] = cy + length / 2;
        xs[1] = cx + width / 2; ys[1] = cy + length / 2;
        xs[2] = cx;             ys[2] = cy + length / 2;
        xs[3] = cx;             ys[3] = cy - length / 2;
    }
    for (int i = 0; i < np; ++i) {                            /* M.cpp:131-143 */
        double x_shift = xs[i] - cx;
        double y_shift = ys[i] - cy;
        double x_rot = c * x_shift - s * y_shift + cx;
        double y_rot = s * x_shift + c * y_shift + cy;
        if (!point_inside(box, x_rot, y_rot, p->alt)) return 0;
    }
    return 1;
}

int64_t orc_generate(const tgx_params* p, double* out, int64_t chan_stride, int64_t cap,
                     uint32_t* status, tgx_phases* ph, int64_t max_samples) {
    uint32_t st = 0;
    int64_t n;
    sink_t sk = {out, chan_stride, cap, 0, {0}};
    g_ph_rows = 1;
    if ((p->type == TGX_CIRCLE || p->type == TGX_FIGURE8) && p->n_vgoals > TGX_MAX_VGOALS &&
        p->n_vgoals <= TGX_MAX_VGOALS_TOTAL && TGX_ORBIT_RECORDS(p->n_vgoals) <= g_navail)
        g_ph_rows = TGX_ORBIT_RECORDS(p->n_vgoals);
    phases_clear(ph);
    if (!params_ok(p)) {
        st |= TGX_ST_BAD_PARAM;
        n = -1;
    } else if (TGX_IS_POLYLINE(p->type)) {
        n = poly_generate(p, &sk, NULL, &st, max_samples);   /* index_msgs: orc_polyline_generate's leg_of */
    } else if (p->type == TGX_LINE) {
        n = line_generate(p, &sk, &st, ph, max_samples);
    } else if (p->type == TGX_BOOMERANG) {
        n = boomerang_generate(p, &sk, &st, ph, max_samples);
    } else {
        n = orbit_generate(p, &sk, &st, ph, max_samples);
    }
    if (n < 0) phases_clear(ph);
    g_ph_rows = 1;
    if (n > cap && out) st |= TGX_ST_TRUNCATED;
    if (status) *status = st;
    return n;
}

/* ---- generateStopTraj ------------------------------------------------------------------------------- */

int64_t orc_stop(const tgx_params* p, const double* from, double* out, int64_t chan_stride, int64_t cap,
                 uint32_t* status, tgx_phases* ph, int64_t max_samples) {
    uint32_t st = 0;
    sink_t sk = {out, chan_stride, cap, 0, {0}};
    double g[TGX_NCHAN];
    g_ph_rows = 1;
    if (ph) ph->n = 0;
    if (!params_ok(p)) {
        if (status) *status = TGX_ST_BAD_PARAM;
        return -1;
    }
    if (TGX_IS_POLYLINE(p->type)) {
        phase_add(ph, 0, TGX_PH_PRESSED_END, 0.0, 0.0);           /* Square.cpp:123 */
        if (poly_stop(p, from, &sk, &st, max_samples) < 0) {
            if (ph) ph->n = 0;
            if (status) *status = st;
            return -1;
        }
        phase_add(ph, sk.n - 1, TGX_PH_STOPPED, 0.0, 0.0);        /* :129 */
        if (sk.n > cap && out) st |= TGX_ST_TRUNCATED;
        if (status) *status = st;
        return sk.n;
    }
    /* 2D current (goal) vel: Circle.cpp:140-141, Line.cpp:124-125, Figure8.cpp:138-139 */
    double v = sqrt(pow(from[TGX_VX], 2) + pow(from[TGX_VY], 2));
    if (p->type == TGX_LINE || p->type == TGX_BOOMERANG) {
        /* Line::generateStopTraj, Line.cpp:117-152 (Boomerang.cpp:169-203 is a verbatim copy) */
        const tgx_line_params* l = &p->u.line;
        double theta = atan2(from[TGX_VY], from[TGX_VX]);         /* :126-127 */
        phase_add(ph, 0, TGX_PH_PRESSED_END, 0.0, 0.0);           /* :132 */
        v = std_max(v - l->a3 * p->dt, 0.0);                      /* :133 */
        line_goal(p, from[TGX_PX], from[TGX_PY], v, -l->a3, theta, g);
        sink_push(&sk, g);
        while (v > 0) {                                           /* :136 */
            double v_new = std_max(v - l->a3 * p->dt, 0.0);
            if (v_new == v || sk.n >= max_samples) { st |= TGX_ST_TOO_LONG; break; }
            v = v_new;
            line_goal(p, sk.last[TGX_PX], sk.last[TGX_PY], v, -l->a3, theta, g);
            sink_push(&sk, g);
        }
    } else {
        /* Circle::generateStopTraj, Circle.cpp:132-169; Figure8::generateStopTraj, Figure8.cpp:130-167
         * (Figure8 recovers theta with the circle's atan2(p - c), :140-141). */
        const tgx_orbit_params* o = &p->u.orbit;
        double theta = atan2(from[TGX_PY] - o->cy, from[TGX_PX] - o->cx);
        phase_add(ph, 0, TGX_PH_PRESSED_END, 0.0, 0.0);           /* :148 */
        while (v > 0) {                                           /* :150 */
            double v_new = std_max(v - o->accel * p->dt, 0.0);
            if (v_new == v || sk.n >= max_samples) { st |= TGX_ST_TOO_LONG; break; }
            v = v_new;
            double omega = v / o->r;
            theta += omega * p->dt;
            orbit_goal(p, v, theta, g);
            sink_push(&sk, g);
        }
    }
    if (st & TGX_ST_TOO_LONG) {
        if (ph) ph->n = 0;
        if (status) *status = st;
        return -1;
    }
    /* index_msgs_tmp[goals_tmp.size() - 1]: size_t(0) - 1 converts to int key -1 when nothing was pushed. */
    phase_add(ph, sk.n - 1, TGX_PH_STOPPED, 0.0, 0.0);
    if (sk.n > cap && out) st |= TGX_ST_TRUNCATED;
    if (status) *status = st;
    return sk.n;
}

/* ---- trajectoryInsideBounds ------------------------------------------------------------------------- */

/* Trajectory::isPointInsideBounds, Trajectory.hpp:50-57 (closed intervals). */
static int point_inside(const double box[6], double x, double y, double z) {
    if (x < box[0] || x > box[1]) return 0;
    if (y < box[2] || y > box[3]) return 0;
    if (z < box[4] || z > box[5]) return 0;
    return 1;
}

int orc_inside_bounds(const tgx_params* p, const double box[6]) {
    if (TGX_IS_POLYLINE(p->type)) return poly_inside_bounds(p, box);
    if (p->type == TGX_LINE || p->type == TGX_BOOMERANG) {
        /* Line::trajectoryInsideBounds, Line.cpp:154-173 (Boomerang.cpp:205-224 is a verbatim copy) */
        const tgx_line_params* l = &p->u.line;
        if (orc_line_d2(p) < 0) return 0;
        return point_inside(box, l->A[0], l->A[1], l->A[2]) && point_inside(box, l->B[0], l->B[1], l->B[2]);
    }
    /* Circle.cpp:171-179, Figure8.cpp:169-177 */
    const tgx_orbit_params* o = &p->u.orbit;
    return point_inside(box, o->cx - o->r, o->cy - o->r, p->alt) &&
           point_inside(box, o->cx + o->r, o->cy + o->r, p->alt);
}

/* ---- batch drivers ---------------------------------------------------------------------------------- */

typedef struct {
    int mode;                 /* 0 generate, 1 feasibility, 2 timing */
    const tgx_params* p;
    int64_t n_total;
    int64_t lo, hi;
    double* out;
    int64_t traj_stride, chan_stride, cap;
    int32_t* counts;
    uint32_t* status;
    const tgx_limits* limits;
    uint8_t* flags;
    double* max_v;
    double* max_a;
    int64_t max_samples;
    int64_t total;
    double checksum;
} job_t;

static void reduce_norms(const double* row, int64_t stride, int64_t n, double* mv, double* ma) {
    double bv = 0.0, ba = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        double vx = row[TGX_VX * stride + k], vy = row[TGX_VY * stride + k], vz = row[TGX_VZ * stride + k];
        double ax = row[TGX_AX * stride + k], ay = row[TGX_AY * stride + k], az = row[TGX_AZ * stride + k];
        double nv = sqrt(vx * vx + vy * vy + vz * vz);
        double na = sqrt(ax * ax + ay * ay + az * az);
        if (nv > bv) bv = nv;
        if (na > ba) ba = na;
    }
    *mv = bv;
    *ma = ba;
}

static void* job_run(void* arg) {
    job_t* j = (job_t*)arg;
    double* scratch = NULL;
    int64_t scratch_cap = 0;
    j->total = 0;
    j->checksum = 0.0;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        uint32_t st = 0;
        int64_t n;
        g_navail = j->n_total - i;
        if (j->p[i].type == TGX_VGOALS_MORE) {
            /* a continuation record (tgx.h): an entry without a trajectory; an orphan is a bad record */
            int owned = 0;
            for (int q = 1; q < TGX_MAX_VGOALS_TOTAL / TGX_MAX_VGOALS && i - q >= 0; ++q) {
                const tgx_params* h = &j->p[i - q];
                if (h->type == TGX_VGOALS_MORE) continue;
                owned = (h->type == TGX_CIRCLE || h->type == TGX_FIGURE8) && TGX_ORBIT_RECORDS(h->n_vgoals) > q;
                break;
            }
            if (j->counts) j->counts[i] = 0;
            if (j->status) j->status[i] = owned ? 0u : (uint32_t)TGX_ST_BAD_PARAM;
            if (j->max_v) j->max_v[i] = 0.0;
            if (j->max_a) j->max_a[i] = 0.0;
            if (j->flags) j->flags[i] = owned ? 1 : 0;
            continue;
        }
        if (j->mode == 0) {
            double* dst = j->out ? j->out + i * j->traj_stride : NULL;
            n = orc_generate(&j->p[i], dst, j->chan_stride, j->cap, &st, NULL, j->max_samples);
        } else {
            /* count first so the scratch row is large enough, then generate into it */
            n = orc_generate(&j->p[i], NULL, 0, 0, &st, NULL, j->max_samples);
            if (n > 0) {
                if (n > scratch_cap) {
                    free(scratch);
                    scratch_cap = n + 1024;
                    scratch = (double*)malloc((size_t)scratch_cap * TGX_NCHAN * sizeof(double));
                    if (!scratch) { scratch_cap = 0; n = -1; }
                }
                if (n > 0) {
                    st = 0;
                    n = orc_generate(&j->p[i], scratch, scratch_cap, scratch_cap, &st, NULL, j->max_samples);
                }
            }
            if (j->mode == 1) {
                double mv = 0.0, ma = 0.0;
                if (n > 0) reduce_norms(scratch, scratch_cap, n, &mv, &ma);
                /* trajectoryInsideBounds tests the geometry alone (Circle.cpp:171-179, Line.cpp:154-173): it is
                 * reported for records the samplers reject too (polyline records: only accepted ones) */
                if (j->limits && j->limits->check_box &&
                    (!(st & TGX_ST_BAD_PARAM) || !TGX_IS_POLYLINE(j->p[i].type)) && j->p[i].type >= TGX_CIRCLE &&
                    j->p[i].type <= TGX_T && !orc_inside_bounds(&j->p[i], j->limits->box)) {
                    st |= TGX_ST_OUTSIDE_BOUNDS;
                    /* Line::trajectoryInsideBounds reports "not feasible" when d2 < 0 (Line.cpp:165-168) */
                    if ((j->p[i].type == TGX_LINE || j->p[i].type == TGX_BOOMERANG) && orc_line_d2(&j->p[i]) < 0)
                        st |= TGX_ST_LINE_D2_NEGATIVE;
                }
                if (j->limits && mv > j->limits->v_max) st |= TGX_ST_VMAX_EXCEEDED;
                if (j->limits && ma > j->limits->a_max) st |= TGX_ST_AMAX_EXCEEDED;
                if (j->max_v) j->max_v[i] = mv;
                if (j->max_a) j->max_a[i] = ma;
                if (j->flags) j->flags[i] = (st == 0) ? 1 : 0;
            } else if (n > 0) {
                j->checksum += scratch[TGX_PX * scratch_cap + (n - 1)] + scratch[TGX_PSI * scratch_cap + n / 2];
            }
        }
        if (j->counts) j->counts[i] = (int32_t)(n < 0 ? 0 : n);
        if (j->status) j->status[i] = st;
        if (n > 0) j->total += n;
    }
    free(scratch);
    g_navail = INT64_MAX;
    return NULL;
}

static int64_t run_jobs(job_t* proto, int64_t n, int nthreads, double* checksum) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    if ((int64_t)nthreads > n) nthreads = (int)(n > 0 ? n : 1);
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -1; }
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].n_total = n;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    job_run(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    int64_t total = 0;
    double cs = 0.0;
    for (int t = 0; t < nthreads; ++t) { total += jobs[t].total; cs += jobs[t].checksum; }
    if (checksum) *checksum = cs;
    free(jobs);
    free(th);
    return total;
}

int orc_generate_batch(const tgx_params* p, int64_t n, double* out, int64_t traj_stride,
                       int64_t chan_stride, int64_t cap, int32_t* counts, uint32_t* status,
                       int64_t max_samples, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.mode = 0; j.p = p; j.out = out; j.traj_stride = traj_stride; j.chan_stride = chan_stride;
    j.cap = cap; j.counts = counts; j.status = status; j.max_samples = max_samples;
    return run_jobs(&j, n, nthreads, NULL) < 0 ? -1 : 0;
}

int orc_feasibility_batch(const tgx_params* p, int64_t n, const tgx_limits* limits, uint8_t* flags,
                          double* max_v, double* max_a, int32_t* counts, uint32_t* status,
                          int64_t max_samples, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.mode = 1; j.p = p; j.limits = limits; j.flags = flags; j.max_v = max_v; j.max_a = max_a;
    j.counts = counts; j.status = status; j.max_samples = max_samples;
    return run_jobs(&j, n, nthreads, NULL) < 0 ? -1 : 0;
}

int64_t orc_time_batch(const tgx_params* p, int64_t n, int64_t max_samples, int nthreads, double* checksum) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.mode = 2; j.p = p; j.max_samples = max_samples;
    return run_jobs(&j, n, nthreads, checksum);
}

/* ---- consumer side: what the node publishes while following ----------------------- */

/* TrajectoryGenerator::saturate, TrajectoryGenerator.cpp:773-780. */
static double saturate(double val, double low, double high) {
    if (val > high)
        val = high;
    else if (val < low)
        val = low;
    return val;
}

/* goal_ = traj_goals_[pub_index_] (TrajectoryGenerator.cpp:557) then the safety bounds (:602-604), for every sample of
 * one trajectory given as SoA planes samples[c * chan_stride + k].  box may be NULL (no saturation). */
void orc_pack_goals(const double* samples, int64_t chan_stride, int64_t n, int32_t traj, const double* box,
                    tgx_goal_record* out) {
    for (int64_t k = 0; k < n; ++k) {
        tgx_goal_record r;
        memset(&r, 0, sizeof(r));
        double px = samples[TGX_PX * chan_stride + k], py = samples[TGX_PY * chan_stride + k];
        double pz = samples[TGX_PZ * chan_stride + k];
        r.p[0] = box ? saturate(px, box[0], box[1]) : px;
        r.p[1] = box ? saturate(py, box[2], box[3]) : py;
        r.p[2] = box ? saturate(pz, box[4], box[5]) : pz;
        for (int c = 0; c < 3; ++c) {
            r.v[c] = samples[(TGX_VX + c) * chan_stride + k];
            r.a[c] = samples[(TGX_AX + c) * chan_stride + k];
            r.j[c] = samples[(TGX_JX + c) * chan_stride + k];
        }
        r.psi = samples[TGX_PSI * chan_stride + k];
        r.dpsi = samples[TGX_DPSI * chan_stride + k];
        r.traj = traj;
        r.k = (int32_t)k;
        r.power = 1;                                   /* goal.power = true, Circle.cpp:127 */
        r.mode_xy = 0;                                 /* MODE_POSITION_CONTROL, never touched by the samplers */
        r.mode_z = 0;
        /* which components the saturation moved (a NaN passes through saturate() and is not "clamped") */
        r.clamped = box ? (uint8_t)(((px > box[1] || px < box[0]) ? 1 : 0) | ((py > box[3] || py < box[2]) ? 2 : 0) |
                                    ((pz > box[5] || pz < box[4]) ? 4 : 0))
                        : 0;
        r.last = (uint8_t)(k == n - 1);
        out[k] = r;
    }
}

/* ---- node-side transitions ------------------------------------------------------------ */

/* TrajectoryGenerator::wrap, TrajectoryGenerator.cpp:782-788. */
static double wrap_pi(double val) {
    if (val > M_PI) val -= 2.0 * M_PI;
    if (val < -M_PI) val += 2.0 * M_PI;
    return val;
}

static int transition_ok(const tgx_transition_params* t) {
    if (!(isfinite(t->dt) && t->dt > 0.0) || t->ticks < 0) return 0;
    for (int i = 0; i < 3; ++i)
        if (!isfinite(t->start[i]) || !isfinite(t->dest[i])) return 0;
    if (!isfinite(t->start_v[0]) || !isfinite(t->start_v[1]) || !isfinite(t->start_psi)) return 0;
    if (!(isfinite(t->vel) && t->vel > 0.0)) return 0;
    if (t->kind == TGX_TR_TAKEOFF) return 1;
    if (t->kind == TGX_TR_LANDING) return isfinite(t->vel_yaw) && t->vel_yaw > 0.0;
    if (t->kind == TGX_TR_GOTO)
        return isfinite(t->dest_yaw) && isfinite(t->vel_yaw) && t->vel_yaw > 0.0 && isfinite(t->dist_thresh) &&
               t->dist_thresh >= 0.0 && isfinite(t->yaw_thresh) && t->yaw_thresh >= 0.0;
    return 0;
}

/* The goals pubCB publishes tick by tick in TAKING_OFF (TrajectoryGenerator.cpp:531-548), INIT_POS_TRAJ / INIT_POS
 * (:549-554, :574-586 through simpleInterpolation :637-764) and LANDING (:588-599), each followed by the saturation of
 * goal_.p to the room box (:602-604), under perfect tracking (pose on tick k = goal published on tick k-1).
 * Returns the number of ticks; the first `cap` are written to out (may be NULL). */
int64_t orc_transition(const tgx_transition_params* t, int32_t traj, const double* box, tgx_goal_record* out,
                       int64_t cap, uint32_t* status, int64_t max_samples) {
    uint32_t st = 0;
    int64_t k = 0;
    if (!transition_ok(t)) {
        if (status) *status = TGX_ST_BAD_PARAM;
        return 0;
    }
    double px = t->start[0], py = t->start[1], pz = t->start[2];
    double vx = t->start_v[0], vy = t->start_v[1], psi = t->start_psi, dpsi = 0.0;
    int power = 1;
    double pose_z = pz;
    int64_t limit = t->ticks > 0 ? t->ticks : max_samples;
    int done = 0;
    while (!done) {
        if (k >= limit) {
            if (t->ticks == 0) st |= TGX_ST_TOO_LONG;
            break;
        }
        int ends = 0;
        if (t->kind == TGX_TR_TAKEOFF) {
            double takeoff_alt = t->dest[2];          /* :535 */
            double eps = 0.10;
            if (fabs(takeoff_alt - pose_z) < eps && pz >= takeoff_alt) ends = 1;   /* :540-543 */
            else pz = saturate(pz + t->vel * t->dt, 0.0, takeoff_alt);             /* :547 */
        } else if (t->kind == TGX_TR_GOTO) {
            double Dx = t->dest[0] - px;              /* :642-644 */
            double Dy = t->dest[1] - py;
            double dist = sqrt(Dx * Dx + Dy * Dy);
            double delta_yaw = t->dest_yaw - psi;     /* :646-647 */
            delta_yaw = wrap_pi(delta_yaw);
            int dist_far = dist > t->dist_thresh;     /* :649-651 */
            int yaw_far = fabs(delta_yaw) > t->yaw_thresh;
            ends = !dist_far && !yaw_far;
            int accel_for_vel = 1;                    /* `bool accel_for_vel = 0.1;` (:653) */
            double npx, npy, nvx, nvy, npsi, ndpsi;
            if (dist_far) {                           /* :657-670 */
                double c = Dx / dist;
                double s = Dy / dist;
                npx = px + c * t->vel * t->dt;
                npy = py + s * t->vel * t->dt;
                nvx = std_min(vx + accel_for_vel * t->dt, c * t->vel);
                nvy = std_min(vy + accel_for_vel * t->dt, s * t->vel);
            } else {                                  /* :671-683 */
                npx = t->dest[0];
                npy = t->dest[1];
                nvx = std_max(0.0, vx - accel_for_vel * t->dt);
                nvy = std_max(0.0, vy - accel_for_vel * t->dt);
            }
            if (yaw_far) {                            /* :685-693 */
                int sgn = delta_yaw >= 0 ? 1 : -1;
                double vel_yaw = sgn * t->vel_yaw;
                npsi = psi + vel_yaw * t->dt;
                ndpsi = vel_yaw;
            } else {
                npsi = t->dest_yaw;
                ndpsi = 0;
            }
            px = npx; py = npy; pz = t->dest[2]; vx = nvx; vy = nvy; psi = npsi; dpsi = ndpsi;
        } else {
            double vel_land = pose_z > (t->dest[2] + 0.4) ? t->vel : t->vel_yaw;   /* :590 */
            pz = pz - vel_land * t->dt;               /* :591 */
            if (pz < 0) {                             /* :593-598 */
                power = 0;
                ends = 1;
            }
        }
        int clamped = 0;
        if (box) {                                    /* :602-604 */
            clamped = ((px > box[1] || px < box[0]) ? 1 : 0) | ((py > box[3] || py < box[2]) ? 2 : 0) |
                      ((pz > box[5] || pz < box[4]) ? 4 : 0);
            px = saturate(px, box[0], box[1]);
            py = saturate(py, box[2], box[3]);
            pz = saturate(pz, box[4], box[5]);
        }
        done = ends && t->ticks == 0;
        if (out && k < cap) {
            tgx_goal_record r;
            memset(&r, 0, sizeof(r));
            r.p[0] = px; r.p[1] = py; r.p[2] = pz;
            r.v[0] = vx; r.v[1] = vy;
            r.psi = psi; r.dpsi = dpsi;
            r.traj = traj; r.k = (int32_t)k;
            r.power = (uint8_t)power;
            r.clamped = (uint8_t)clamped;
            r.last = (uint8_t)(done || (t->ticks > 0 && k + 1 == limit));
            out[k] = r;
        }
        pose_z = pz;
        ++k;
    }
    if (out && k > cap) st |= TGX_ST_TRUNCATED;
    if (status) *status = st;
    return k;
}

/* ---- checksum ---------------------------------------------------------------------------------------- */

uint64_t orc_fnv1a64(const double* x, int64_t n, uint64_t seed) {
    uint64_t h = seed;
    for (int64_t i = 0; i < n; ++i) {
        double d = x[i];
        if (d == 0.0) d = 0.0;   /* canonicalise -0.0 */
        uint64_t bits;
        memcpy(&bits, &d, sizeof(bits));
        for (int b = 0; b < 8; ++b) {
            h ^= (bits >> (8 * b)) & 0xffu;
            h *= 0x100000001b3ULL;
        }
    }
    return h;
}

/*
 * ttfleet_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) of the fleet step:
 * what the reference does around the solver call for one robot, restated in plain C.
 *
 *   InterfaceMpc.get_local_ref_traj / get_action      /root/reference/src/interface_mpc.py:72-92
 *   TrajectoryGenerator.get_local_ref_traj             src/mpc_traj_tracker/trajectory_generator.py:203-230
 *   TrajectoryGenerator.check_termination_condition    :158-164
 *   TrajectoryGenerator.run_step (speed reference, parameter assembly)   :233-260
 *   TrajectoryGenerator.run_solver (state update with the first control) :291-294
 *   unicycle_model, RK4 branch                         src/pkg_motion_model/motion_model.py:153-176
 *   est_dyn_obs_positions                              src/main.py:80-89
 *
 * PINNED: tests/golden/fleet_step.npz is produced by running the reference's own Python for
 * these functions (tools/gen_golden_fleet.py, solver replaced by a recorder); the oracle
 * reproduces the packed vectors exactly and the states to 1 ulp (libm mode).
 *
 * use_libm = 1: hypot / sin / cos from libm, as numpy / math do in the reference.
 * use_libm = 0: the device's arithmetic -- sqrt(fma(dx,dx,dy*dy)) and the kernels' tt_sincos
 *               -- which the CUDA path reproduces bit for bit.
 */
#include <math.h>
#include <string.h>

#include "ttmpc_oracle.h"

static double hyp(double dx, double dy, int use_libm) {
  return use_libm ? hypot(dx, dy) : sqrt(fma(dx, dx, dy * dy));
}
static void sc(double x, double *s, double *c, int use_libm) {
  if (use_libm) { *s = sin(x); *c = cos(x); }
  else ttmpc_oracle_sincos(x, s, c);
}

static int np_of(const ttmpc_config *g) {
  const int N = g->N_hor;
  return 2 * g->ns + g->nu + g->nq + g->ns * N + N + g->ns * N * g->Nother +
         g->Nstcobs * g->nstcobs + g->Ndynobs * g->ndynobs * N + 2 * N;
}

/* shapely Polygon.contains(Point): strictly inside (even-odd rule); Polygon.distance(Point):
 * 0 inside, else the distance to the closest edge. */
static int poly_contains(const double *xy, int nv, double px, double py) {
  int in = 0;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    const double xi = xy[2 * i], yi = xy[2 * i + 1], xj = xy[2 * j], yj = xy[2 * j + 1];
    if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) in = !in;
  }
  return in;
}
static double poly_distance(const double *xy, int nv, double px, double py) {
  if (poly_contains(xy, nv, px, py)) return 0.0;
  double best = INFINITY;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    const double ax = xy[2 * j], ay = xy[2 * j + 1], dx = xy[2 * i] - ax, dy = xy[2 * i + 1] - ay;
    const double len2 = dx * dx + dy * dy;
    double t = 0.0;
    if (len2 > 0.0) { t = ((px - ax) * dx + (py - ay) * dy) / len2; t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }
    const double qx = ax + t * dx - px, qy = ay + t * dy - py;
    const double d = sqrt(qx * qx + qy * qy);
    if (d < best) best = d;
  }
  return best;
}
/* exported for tests/test_geometry_independent.py (checked against dense sampling + winding number) */
int ttfleet_oracle_poly_contains(const double *xy, int nv, double px, double py) { return poly_contains(xy, nv, px, py); }
double ttfleet_oracle_poly_distance(const double *xy, int nv, double px, double py) { return poly_distance(xy, nv, px, py); }
/* HintSwitcher.switch (main_pre.py:35-52) for robot e; ref rows are the ORIGINAL local reference */
static int hint_switch(const ttmpc_fleet *f, int e, int N, const double *ref, int L, int idx) {
  int *st = f->sw_state + 2 * e;   /* switch_on, detach_cnt */
  const double px = f->state[3 * e], py = f->state[3 * e + 1];
  const int n_stc = f->sw_poly_xy ? f->sw_max_poly : 0;
  const int n_dyn = f->dyn_cur ? f->n_dyn_live : 0;
  const double *pxy = f->sw_poly_xy ? f->sw_poly_xy + (f->sw_poly_shared ? 0 : (size_t)e * f->sw_max_poly * f->sw_max_pv * 2) : 0;
  const int *pnv = f->sw_poly_nv ? f->sw_poly_nv + (f->sw_poly_shared ? 0 : (size_t)e * f->sw_max_poly) : 0;
  int cnt_flag = 0;
  for (int k = 0; k < N; k++) {
    int r = idx + k; if (r > L - 1) r = L - 1;
    const double ox = ref[3 * r], oy = ref[3 * r + 1];
    for (int o = 0; o < n_stc + n_dyn; o++) {
      double rect[8];
      const double *xy; int nv;
      if (o < n_stc) { nv = pnv[o]; xy = pxy + (size_t)o * f->sw_max_pv * 2; if (nv < 3) continue; }
      else {  /* circle_to_rect (main.py:91-95) */
        const double *c = f->dyn_cur + ((size_t)e * f->n_dyn_live + (o - n_stc)) * 2, rr = f->sw_dyn_radius;
        rect[0] = c[0] - rr; rect[1] = c[1] - rr; rect[2] = c[0] + rr; rect[3] = c[1] - rr;
        rect[4] = c[0] + rr; rect[5] = c[1] + rr; rect[6] = c[0] - rr; rect[7] = c[1] + rr;
        xy = rect; nv = 4;
      }
      const double dist = poly_distance(xy, nv, px, py);
      if (poly_contains(xy, nv, ox, oy)) {
        if (dist < f->sw_switch_distance && !st[0]) { st[0] = 1; return 1; }
      } else if (dist > f->sw_detach_distance && st[0]) {
        if (st[1] > f->sw_detach_steps) { st[0] = 0; st[1] = 0; }
        else if (!cnt_flag) { st[1] += 1; cnt_flag = 1; }
      }
    }
  }
  return st[0];
}

void ttfleet_oracle_pack(const ttmpc_config *g, const ttmpc_fleet *f, double *p_all, int use_libm) {
  const int N = g->N_hor, np = np_of(g);
  for (int e = 0; e < f->n; e++) {
    double *p = p_all + (size_t)e * np;
    const double *st = f->state + 3 * e, *goal = f->goal + 3 * e, *lu = f->last_u + 2 * e;
    const double *ref = f->ref_traj + (size_t)e * f->ref_stride * 3;
    const int L = f->ref_len[e];
    int idx = f->idx_ref[e];
    const int was_running = f->status[e] == TTMPC_FLEET_RUNNING;
    if (f->status[e] == TTMPC_FLEET_RUNNING) {
      /* get_local_ref_traj: closest point in [idx - action_steps, idx + 5 action_steps) */
      int lo = idx - 1 * f->action_steps; if (lo < 0) lo = 0;
      int hi = idx + 5 * f->action_steps; if (hi > L) hi = L;
      double best = INFINITY; int arg = lo;
      for (int i = lo; i < hi; i++) {
        const double d = hyp(st[0] - ref[3 * i], st[1] - ref[3 * i + 1], use_libm);
        if (d < best) { best = d; arg = i; }
      }
      idx = arg;
      f->idx_ref[e] = idx;
      /* get_action: check_termination_condition(state, last_action, goal) */
      const int close = fabs(st[0] - goal[0]) <= 0.05 && fabs(st[1] - goal[1]) <= 0.05;
      if (close && fabs(lu[0]) < 0.05) f->status[e] = TTMPC_FLEET_REACHED;
    }
    int o = 0;
    for (int i = 0; i < 3; i++) p[o++] = st[i];
    /* main.py:200 evaluates the switch before get_action's goal test */
    if (f->sw_state && f->hint && f->use_hint && was_running)
      f->use_hint[e] = hint_switch(f, e, N, ref, L, idx);
    const int hinted = f->hint && f->use_hint && f->use_hint[e];
    const double *hint = hinted ? f->hint + (size_t)e * N * 2 : 0;
    {
      int r = idx + N - 1; if (r > L - 1) r = L - 1;
      for (int i = 0; i < 3; i++) p[o++] = (hinted && i < 2) ? hint[2 * (N - 1) + i] : ref[3 * r + i];
    }
    p[o++] = lu[0]; p[o++] = lu[1];
    for (int i = 0; i < 10; i++) p[o++] = f->tuning[i];
    for (int k = 0; k < N; k++) {
      int r = idx + k; if (r > L - 1) r = L - 1;
      for (int i = 0; i < 3; i++) p[o++] = (hinted && i < 2) ? hint[2 * k + i] : ref[3 * r + i];
    }
    {
      const double dist = hyp(st[0] - goal[0], st[1] - goal[1], use_libm);
      double v = f->base_speed;
      if (!(dist >= f->base_speed * N * g->ts)) {
        v = dist / N / g->ts;
        if (!(v > f->low_speed)) v = f->low_speed;  /* max(speed_ref, low_speed) */
      }
      for (int k = 0; k < N; k++) p[o++] = v;
    }
    {
      const int no = g->ns * N * g->Nother;
      if (f->other) memcpy(p + o, f->other + (size_t)e * no, sizeof(double) * no);
      else memset(p + o, 0, sizeof(double) * no);
      o += no;
    }
    {
      const int ns = g->Nstcobs * g->nstcobs;
      memcpy(p + o, f->stc + (f->stc_shared ? 0 : (size_t)e * ns), sizeof(double) * ns);
      o += ns;
    }
    {
      const int nd = g->Ndynobs * g->ndynobs * N;
      if (f->dyn) memcpy(p + o, f->dyn + (size_t)e * nd, sizeof(double) * nd);
      else {
        memset(p + o, 0, sizeof(double) * nd);
        if (f->dyn_cur) {
          for (int j = 0; j < f->n_dyn_live; j++) {
            const double *c = f->dyn_cur + ((size_t)e * f->n_dyn_live + j) * 2;
            const double *l = f->dyn_last + ((size_t)e * f->n_dyn_live + j) * 2;
            const double dx = c[0] - l[0], dy = c[1] - l[1];
            for (int i = 0; i < N; i++) {
              double *row = p + o + ((size_t)j * N + i) * 6;
              row[0] = c[0] + dx * (double)(i + 1); row[1] = c[1] + dy * (double)(i + 1);
              row[2] = f->dyn_size; row[3] = f->dyn_size; row[4] = 0.0; row[5] = 1.0;
            }
          }
        }
      }
      o += nd;
    }
    for (int k = 0; k < N; k++) p[o++] = f->stc_weight;
    for (int k = 0; k < N; k++) p[o++] = f->dyn_weight;
  }
}

void ttfleet_oracle_advance(const ttmpc_config *g, const ttmpc_fleet *f, const double *u_all,
                            const int *exit_status, int use_libm) {
  const int N = g->N_hor;
  const double ts = g->ts;
  for (int e = 0; e < f->n; e++) {
    if (f->dyn_cur) {
      for (int j = 0; j < f->n_dyn_live; j++) {
        double *c = f->dyn_cur + ((size_t)e * f->n_dyn_live + j) * 2;
        double *l = f->dyn_last + ((size_t)e * f->n_dyn_live + j) * 2;
        const double *d = f->dyn_disp + ((size_t)e * f->n_dyn_live + j) * 2;
        l[0] = c[0]; l[1] = c[1];
        c[0] = c[0] + d[0]; c[1] = c[1] + d[1];
      }
    }
    if (f->status[e] != TTMPC_FLEET_RUNNING) continue;
    if (exit_status && exit_status[e] == TTMPC_NOT_FINITE) { f->status[e] = TTMPC_FLEET_FAILED; continue; }
    const double *u = u_all + (size_t)e * 2 * N;
    double *st = f->state + 3 * e;
    const double v = u[0], w = u[1];
    /* numpy RK4, operation for operation: d_state_f(st) = ts * [v cos th, v sin th, w] */
    double s, c;
    sc(st[2], &s, &c, use_libm);
    const double k1x = ts * (v * c), k1y = ts * (v * s), k1t = ts * w;
    sc(st[2] + 0.5 * k1t, &s, &c, use_libm);
    const double k2x = ts * (v * c), k2y = ts * (v * s), k2t = ts * w;
    sc(st[2] + 0.5 * k2t, &s, &c, use_libm);
    const double k3x = ts * (v * c), k3y = ts * (v * s), k3t = ts * w;
    sc(st[2] + k3t, &s, &c, use_libm);
    const double k4x = ts * (v * c), k4y = ts * (v * s), k4t = ts * w;
    const double sixth = 1.0 / 6.0;
    st[0] = st[0] + sixth * (((k1x + 2.0 * k2x) + 2.0 * k3x) + k4x);
    st[1] = st[1] + sixth * (((k1y + 2.0 * k2y) + 2.0 * k3y) + k4y);
    st[2] = st[2] + sixth * (((k1t + 2.0 * k2t) + 2.0 * k3t) + k4t);
    f->last_u[2 * e] = v; f->last_u[2 * e + 1] = w;
  }
}

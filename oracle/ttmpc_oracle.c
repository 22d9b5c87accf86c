/*
 * ttmpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See ttmpc_oracle.h for the parity status of each part.
 *
 * Part 1: problem functions -- restates MpcModule.build
 *         (/root/reference/src/mpc_traj_tracker/mpc/mpc_generator.py:155-283)
 *         and unicycle_model (src/pkg_motion_model/motion_model.py:153-176).
 * Part 2: the OpEn solver (third-party, optimization_engine 0.7.x + lbfgs 0.2.x):
 *         PANOC (panoc_engine.rs / panoc_cache.rs / panoc_optimizer.rs),
 *         L-BFGS with C-BFGS safeguard (lbfgs crate), ALM/PM outer loop
 *         (alm_optimizer.rs), local Lipschitz estimate (lipschitz_estimator.rs).
 *         PARITY UNPINNED: restated from the published algorithm; no output of the real solver on
 *         this problem exists here.  Anchored (round 2) on the known answers the two crates' OWN unit
 *         tests hold, restated with their test problems: the lbfgs crate's correctneess_buff_1
 *         direction (to 1e-14), the fixed points mocks::SOLUTION_A / SOLUTION_HARD that PANOC must
 *         reach on mocks::my_cost / hard_quadratic_cost over a Euclidean ball, the local Lipschitz
 *         estimate of mocks::lipschitz_mock (tests/test_oracle_solver.py; every constant is also
 *         checked against what it must satisfy -- hand arithmetic, KKT conditions).  These pin the
 *         building blocks and the fixed points, not the iterate path on the NMPC problem.
 */
#include "ttmpc_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 32          /* horizon */
#define MAXNU (2 * MAXN) /* decision variables */
#define MAXDYN 64
#define MAXEDGE 8
#define MAXMEM 16

/* ------------------------------------------------------------------ */
/* Part 1: problem functions                                           */
/* ------------------------------------------------------------------ */

typedef struct {
  int s, q, r, vref, c, os, od, qstc, qdyn, np;
} offs_t;

static offs_t offsets(const ttmpc_config *g) {
  offs_t o;
  int N = g->N_hor;
  o.s = 0;
  o.q = 2 * g->ns + g->nu;
  o.r = o.q + g->nq;
  o.vref = o.r + g->ns * N;
  o.c = o.vref + N;
  o.os = o.c + g->ns * N * g->Nother;
  o.od = o.os + g->Nstcobs * g->nstcobs;
  o.qstc = o.od + g->Ndynobs * g->ndynobs * N;
  o.qdyn = o.qstc + N;
  o.np = o.qdyn + N;
  return o;
}

/* unicycle_model with rk4=True (motion_model.py:153-176), literal op order */
static void unicycle_rk4(double ts, const double s[3], double v, double w, double out[3]) {
  double k1[3], k2[3], k3[3], k4[3], th;
  th = s[2];
  k1[0] = ts * (v * cos(th)); k1[1] = ts * (v * sin(th)); k1[2] = ts * w;
  th = s[2] + 0.5 * k1[2];
  k2[0] = ts * (v * cos(th)); k2[1] = ts * (v * sin(th)); k2[2] = ts * w;
  th = s[2] + 0.5 * k2[2];
  k3[0] = ts * (v * cos(th)); k3[1] = ts * (v * sin(th)); k3[2] = ts * w;
  th = s[2] + k3[2];
  k4[0] = ts * (v * cos(th)); k4[1] = ts * (v * sin(th)); k4[2] = ts * w;
  for (int i = 0; i < 3; i++)
    out[i] = s[i] + (1.0 / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* dist_to_lineseg squared (mpc_generator.py:27-35); returns d2, fills the
 * closest-point vector (tx,ty), t_hat and the segment for the gradient.   */
static double seg_dist2(double X, double Y, double s1x, double s1y, double s2x,
                        double s2y, double *tx, double *ty, double *that,
                        double *sx_, double *sy_, double *den_) {
  double sx = s2x - s1x, sy = s2y - s1y;
  double den = sx * sx + sy * sy + 1e-16;
  double t_hat = ((X - s1x) * sx + (Y - s1y) * sy) / den;
  double t = fmin(fmax(t_hat, 0.0), 1.0);
  double cx = s1x + t * sx - X, cy = s1y + t * sy - Y;
  *tx = cx; *ty = cy; *that = t_hat; *sx_ = sx; *sy_ = sy; *den_ = den;
  return cx * cx + cy * cy; /* sq(sqrt(.)) is simplified away by casadi SX */
}

typedef struct {
  /* forward tape of one evaluation */
  double st[MAXN + 1][3];           /* states, st[0] = start */
  double dxdv[MAXN], dydv[MAXN];    /* d state_{k+1} / d v_k  */
  double dxdw[MAXN], dydw[MAXN];
  double dxdth[MAXN], dydth[MAXN];  /* d pos_{k+1} / d theta_k */
  double S;                         /* static hard sum */
  double D[MAXDYN];                 /* dynamic hard sums */
  double F2[MAXDYN];
  double f;
} tape_t;

/* forward evaluation: f, F1, F2 and (optionally) the tape for the adjoint */
static void forward(const ttmpc_config *g, const double *u, const double *p,
                    double *f_out, double *F1, double *F2, tape_t *tp) {
  const offs_t o = offsets(g);
  const int N = g->N_hor, ns = g->ns, ne = g->nstcobs / 3;
  const double ts = g->ts;
  const double *s = p + o.s, *q = p + o.q, *r = p + o.r, *vref = p + o.vref;
  const double *c = p + o.c, *os = p + o.os, *od = p + o.od, *qdyn = p + o.qdyn;
  const double x_goal = s[3], y_goal = s[4], theta_goal = s[5];
  const double v_init = s[6], w_init = s[7];
  const double qvel = q[1], rv = q[3], rw = q[4], qN = q[5], qthetaN = q[6],
               qrpd = q[7], acc_pen = q[8], wacc_pen = q[9];
  double cost = 0.0, S = 0.0, D[MAXDYN];
  double st[3] = {s[0], s[1], s[2]};
  for (int j = 0; j < g->Ndynobs; j++) D[j] = 0.0;
  if (tp) { tp->st[0][0] = st[0]; tp->st[0][1] = st[1]; tp->st[0][2] = st[2]; }

  for (int kt = 0; kt < N; kt++) {
    const double v = u[2 * kt], w = u[2 * kt + 1];
    double nx[3];
    unicycle_rk4(ts, st, v, w, nx);
    if (tp) {
      /* analytic Jacobian pieces of the RK4 (Simpson) step */
      double tha = st[2], thb = st[2] + 0.5 * (ts * w), thc = st[2] + ts * w;
      double ca = cos(tha), sa = sin(tha), cb = cos(thb), sb = sin(thb),
             cc = cos(thc), sc = sin(thc);
      double h6 = ts / 6.0;
      tp->dxdv[kt] = h6 * (ca + 4 * cb + cc);
      tp->dydv[kt] = h6 * (sa + 4 * sb + sc);
      tp->dxdth[kt] = -h6 * v * (sa + 4 * sb + sc);
      tp->dydth[kt] = h6 * v * (ca + 4 * cb + cc);
      tp->dxdw[kt] = -h6 * v * ts * (2 * sb + sc);
      tp->dydw[kt] = h6 * v * ts * (2 * cb + cc);
      tp->st[kt + 1][0] = nx[0]; tp->st[kt + 1][1] = nx[1]; tp->st[kt + 1][2] = nx[2];
    }
    st[0] = nx[0]; st[1] = nx[1]; st[2] = nx[2];
    const double X = st[0], Y = st[1];

    /* cost_refpath_deviation(state_next, path_ref[kt:], qrpd) (l.124-139):
       path_ref has N+1 points, the last one duplicated (l.190-191)          */
    {
      double dmin = 0.0;
      for (int j = kt; j < N; j++) {
        int j2 = (j + 1 < N) ? j + 1 : N - 1;
        double tx, ty, th, sx, sy, den;
        double d2 = seg_dist2(X, Y, r[j * ns], r[j * ns + 1], r[j2 * ns],
                              r[j2 * ns + 1], &tx, &ty, &th, &sx, &sy, &den);
        if (j == kt) dmin = d2; else dmin = fmin(dmin, d2);
      }
      cost += dmin * qrpd;
    }
    /* cost_refvalue_deviation(u_t[0], speed ref, qvel) (l.203) */
    cost += qvel * ((v - vref[kt]) * (v - vref[kt]));
    /* cost_control_action(u_t, [rv, rw]) (l.204) */
    cost += rv * (v * v) + rw * (w * w);
    /* cost_fleet_collision (l.207-211): safe distance = vehicle_width, weight 1000 */
    {
      double acc = 0.0, d = g->vehicle_width;
      for (int j = 0; j < g->Nother; j++) {
        double cx = c[j * ns * N + kt * ns], cy = c[j * ns * N + kt * ns + 1];
        double dd = (X - cx) * (X - cx) + (Y - cy) * (Y - cy);
        acc += fmax(0.0, d * d - dd);
      }
      cost += 1000.0 * acc;
    }
    /* static obstacles (l.214-220): penalty += max(0, prod_e max(0, b-a0 x-a1 y)^2) */
    for (int i = 0; i < g->Nstcobs; i++) {
      const double *b = os + i * g->nstcobs, *a0 = b + ne, *a1 = b + 2 * ne;
      double inside = 1.0;
      for (int e = 0; e < ne; e++) {
        double res = a0[e] * (-X) + a1[e] * (-Y) + b[e] * 1.0;
        double m = fmax(0.0, res);
        inside *= m * m;
      }
      S += fmax(0.0, inside);
    }
    /* dynamic obstacles (l.225-237) */
    {
      double soft = 0.0;
      for (int j = 0; j < g->Ndynobs; j++) {
        const double *e = od + j * g->ndynobs * N + kt * g->ndynobs;
        double cx = e[0], cy = e[1], rx = e[2], ry = e[3], ang = e[4], alpha = e[5];
        double A = (X - cx) * cos(ang) + (Y - cy) * sin(ang);
        double B = (X - cx) * sin(ang) - (Y - cy) * cos(ang);
        double in1 = 1 - (A * A) / ((rx + 1e-6) * (rx + 1e-6)) -
                     (B * B) / ((ry + 1e-6) * (ry + 1e-6));
        D[j] += fmax(0.0, in1);
        double rxm = rx + g->social_margin, rym = ry + g->social_margin;
        double in2 = 1 - (A * A) / ((rxm + 1e-6) * (rxm + 1e-6)) -
                     (B * B) / ((rym + 1e-6) * (rym + 1e-6));
        double m = fmax(0.0, in2);
        soft += ((m * m) * alpha) * qdyn[kt];
      }
      cost += soft;
    }
  }
  /* terminal cost (l.242) */
  cost += qN * ((st[0] - x_goal) * (st[0] - x_goal) + (st[1] - y_goal) * (st[1] - y_goal)) +
          qthetaN * ((st[2] - theta_goal) * (st[2] - theta_goal));
  /* accelerations (l.250-264) */
  {
    double sa = 0.0, sw = 0.0;
    for (int k = 0; k < N; k++) {
      double vp = k ? u[2 * (k - 1)] : v_init, wp = k ? u[2 * (k - 1) + 1] : w_init;
      double a = (u[2 * k] - vp) / ts, aw = (u[2 * k + 1] - wp) / ts;
      if (F1) { F1[k] = a; F1[N + k] = aw; }
      sa += a * a; sw += aw * aw;
    }
    cost += sa * acc_pen;
    cost += sw * wacc_pen;
  }
  /* penalty_constraints is scalar + vector broadcast (l.218,234): F2_j = S + D_j */
  for (int j = 0; j < g->Ndynobs; j++) {
    if (F2) F2[j] = S + D[j];
    if (tp) { tp->D[j] = D[j]; tp->F2[j] = S + D[j]; }
  }
  if (tp) { tp->S = S; tp->f = cost; }
  if (f_out) *f_out = cost;
}

void ttmpc_oracle_eval(const ttmpc_config *g, const double *u, const double *p,
                       double *f, double *F1, double *F2) {
  forward(g, u, p, f, F1, F2, NULL);
}

static double clip(double z, double lo, double hi) { return fmin(fmax(z, lo), hi); }

/* psi = f + c/2 dist^2_C(F1 + y/max(c,1)) + c/2 |F2|^2
 * (opengen builder __construct_function_psi)                               */
static double psi_from(const ttmpc_config *g, double f, const double *F1,
                       const double *F2, double c, const double *y) {
  const int N = g->N_hor;
  double psi = f, d2 = 0.0, n2 = 0.0;
  for (int i = 0; i < 2 * N; i++) {
    double lo = i < N ? g->lin_acc_min : -g->ang_acc_max;
    double hi = i < N ? g->lin_acc_max : g->ang_acc_max;
    double z = F1[i] + (y ? y[i] : 0.0) / fmax(c, 1.0);
    double e = z - clip(z, lo, hi);
    d2 += e * e;
  }
  psi += c * d2 / 2;
  for (int j = 0; j < g->Ndynobs; j++) n2 += F2[j] * F2[j];
  psi += c * n2 / 2;
  return psi;
}

double ttmpc_oracle_psi(const ttmpc_config *g, const double *u, const double *p,
                        double c, const double *y) {
  double f, F1[MAXNU], F2[MAXDYN];
  forward(g, u, p, &f, F1, F2, NULL);
  return psi_from(g, f, F1, F2, c, y);
}

/* hand-written reverse mode of psi */
void ttmpc_oracle_psi_grad(const ttmpc_config *g, const double *u, const double *p,
                           double c, const double *y, double *grad) {
  const offs_t o = offsets(g);
  const int N = g->N_hor, ns = g->ns, ne = g->nstcobs / 3;
  const double ts = g->ts;
  const double *s = p + o.s, *q = p + o.q, *r = p + o.r, *vref = p + o.vref;
  const double *cc_ = p + o.c, *os = p + o.os, *od = p + o.od, *qdyn = p + o.qdyn;
  const double qvel = q[1], rv = q[3], rw = q[4], qN = q[5], qthetaN = q[6],
               qrpd = q[7], acc_pen = q[8], wacc_pen = q[9];
  tape_t tp;
  double F1[MAXNU];
  forward(g, u, p, NULL, F1, NULL, &tp);

  double sumF2 = 0.0;
  for (int j = 0; j < g->Ndynobs; j++) sumF2 += tp.F2[j];

  for (int i = 0; i < 2 * N; i++) grad[i] = 0.0;

  /* direct control terms */
  double ea[MAXN + 1], ew[MAXN + 1], aa[MAXN + 1], aw[MAXN + 1];
  for (int k = 0; k < N; k++) {
    aa[k] = F1[k]; aw[k] = F1[N + k];
    double z = aa[k] + (y ? y[k] : 0.0) / fmax(c, 1.0);
    ea[k] = z - clip(z, g->lin_acc_min, g->lin_acc_max);
    z = aw[k] + (y ? y[N + k] : 0.0) / fmax(c, 1.0);
    ew[k] = z - clip(z, -g->ang_acc_max, g->ang_acc_max);
  }
  aa[N] = aw[N] = ea[N] = ew[N] = 0.0;
  for (int k = 0; k < N; k++) {
    double v = u[2 * k], w = u[2 * k + 1];
    grad[2 * k] += 2 * qvel * (v - vref[k]) + 2 * rv * v;
    grad[2 * k + 1] += 2 * rw * w;
    grad[2 * k] += 2 * acc_pen * (aa[k] - aa[k + 1]) / ts + c * (ea[k] - ea[k + 1]) / ts;
    grad[2 * k + 1] += 2 * wacc_pen * (aw[k] - aw[k + 1]) / ts + c * (ew[k] - ew[k + 1]) / ts;
  }

  /* adjoint sweep */
  double lx = 0.0, ly = 0.0, lt = 0.0;
  for (int kt = N - 1; kt >= 0; kt--) {
    const double X = tp.st[kt + 1][0], Y = tp.st[kt + 1][1], TH = tp.st[kt + 1][2];
    double gx = 0.0, gy = 0.0, gt = 0.0;
    /* reference path: gradient of the selected (first-min) segment */
    {
      double dmin = 0.0, bx = 0.0, by = 0.0;
      for (int j = kt; j < N; j++) {
        int j2 = (j + 1 < N) ? j + 1 : N - 1;
        double tx, ty, th, sx, sy, den;
        double d2 = seg_dist2(X, Y, r[j * ns], r[j * ns + 1], r[j2 * ns],
                              r[j2 * ns + 1], &tx, &ty, &th, &sx, &sy, &den);
        if (j == kt || !(dmin <= d2)) { /* fmin(x,y): x kept when x<=y */
          double pass = (th >= 0.0 && th <= 1.0) ? 1.0 : 0.0;
          double cs_ = (tx * sx + ty * sy) * pass / den;
          dmin = d2;
          bx = 2 * (cs_ * sx - tx);
          by = 2 * (cs_ * sy - ty);
        }
      }
      gx += qrpd * bx; gy += qrpd * by;
    }
    /* fleet */
    {
      double d = g->vehicle_width;
      for (int j = 0; j < g->Nother; j++) {
        double cx = cc_[j * ns * N + kt * ns], cy = cc_[j * ns * N + kt * ns + 1];
        double dd = (X - cx) * (X - cx) + (Y - cy) * (Y - cy);
        if (d * d - dd > 0.0) {
          gx += 1000.0 * (-2 * (X - cx));
          gy += 1000.0 * (-2 * (Y - cy));
        }
      }
    }
    /* static obstacles: weight c * sum_j F2_j */
    if (c != 0.0) {
      for (int i = 0; i < g->Nstcobs; i++) {
        const double *b = os + i * g->nstcobs, *a0 = b + ne, *a1 = b + 2 * ne;
        double m[MAXEDGE], inside = 1.0;
        for (int e = 0; e < ne; e++) {
          double res = a0[e] * (-X) + a1[e] * (-Y) + b[e];
          m[e] = fmax(0.0, res);
          inside *= m[e] * m[e];
        }
        if (inside > 0.0) {
          double dX = 0.0, dY = 0.0;
          for (int e = 0; e < ne; e++) {
            double rest = 1.0; /* leave-one-out product */
            for (int e2 = 0; e2 < ne; e2++) if (e2 != e) rest *= m[e2] * m[e2];
            dX += rest * 2 * m[e] * (-a0[e]);
            dY += rest * 2 * m[e] * (-a1[e]);
          }
          gx += c * sumF2 * dX; gy += c * sumF2 * dY;
        }
      }
    }
    /* dynamic obstacles */
    for (int j = 0; j < g->Ndynobs; j++) {
      const double *e = od + j * g->ndynobs * N + kt * g->ndynobs;
      double cx = e[0], cy = e[1], rx = e[2], ry = e[3], ang = e[4], alpha = e[5];
      double ca = cos(ang), sa = sin(ang);
      double A = (X - cx) * ca + (Y - cy) * sa;
      double B = (X - cx) * sa - (Y - cy) * ca;
      double Rx2 = (rx + 1e-6) * (rx + 1e-6), Ry2 = (ry + 1e-6) * (ry + 1e-6);
      double in1 = 1 - (A * A) / Rx2 - (B * B) / Ry2;
      if (in1 > 0.0 && c != 0.0) {
        double wgt = c * tp.F2[j];
        gx += wgt * (-2 * A * ca / Rx2 - 2 * B * sa / Ry2);
        gy += wgt * (-2 * A * sa / Rx2 + 2 * B * ca / Ry2);
      }
      double rxm = rx + g->social_margin, rym = ry + g->social_margin;
      double Rxm2 = (rxm + 1e-6) * (rxm + 1e-6), Rym2 = (rym + 1e-6) * (rym + 1e-6);
      double in2 = 1 - (A * A) / Rxm2 - (B * B) / Rym2;
      if (in2 > 0.0) {
        double wgt = qdyn[kt] * alpha * 2 * in2;
        gx += wgt * (-2 * A * ca / Rxm2 - 2 * B * sa / Rym2);
        gy += wgt * (-2 * A * sa / Rxm2 + 2 * B * ca / Rym2);
      }
    }
    /* terminal */
    if (kt == N - 1) {
      gx += 2 * qN * (X - s[3]);
      gy += 2 * qN * (Y - s[4]);
      gt += 2 * qthetaN * (TH - s[5]);
    }
    lx += gx; ly += gy; lt += gt;
    grad[2 * kt] += lx * tp.dxdv[kt] + ly * tp.dydv[kt];
    grad[2 * kt + 1] += lx * tp.dxdw[kt] + ly * tp.dydw[kt] + lt * ts;
    lt += lx * tp.dxdth[kt] + ly * tp.dydth[kt];
  }
}

void ttmpc_oracle_rollout(const ttmpc_config *g, const double *u, const double *p,
                          double *states) {
  double st[3] = {p[0], p[1], p[2]};
  for (int k = 0; k < g->N_hor; k++) {
    double nx[3];
    unicycle_rk4(g->ts, st, u[2 * k], u[2 * k + 1], nx);
    for (int i = 0; i < 3; i++) states[3 * k + i] = st[i] = nx[i];
  }
}

/* ------------------------------------------------------------------ */
/* Part 1b: WARP-ordered evaluation                                    */
/*                                                                     */
/* Same functions as Part 1, evaluated with the operation order of the */
/* CUDA kernel (csrc/ttmpc_device.cuh): 32 lanes, lane k owns step k,  */
/* Kogge-Stone prefix/suffix scans for the rollout and its adjoint,    */
/* butterfly all-reduces, explicit fma(), and tt_sincos instead of     */
/* libm.  Compiled with -ffp-contract=off this reproduces the GPU      */
/* arithmetic bit for bit; tests check it against Part 1 (1e-12) and   */
/* against the GPU (exact).                                            */
/* ------------------------------------------------------------------ */

static void tt_sincos(double x, double *s, double *c) {
  if (!(fabs(x) < 1.0e9)) { *s = x * 0.0 + NAN; *c = *s; return; }
  const double kd = rint(x * 6.36619772367581382433e-01);
  double r = fma(-kd, 1.5707963267948966e+00, x);
  r = fma(-kd, 6.123233995736766e-17, r);
  r = fma(-kd, -1.4973849048591698e-33, r);
  const int q = (int)((long long)kd & 3);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double sr = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  switch (q) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}
void ttmpc_oracle_sincos(double x, double *s, double *c) { tt_sincos(x, s, c); }

static double relu_(double x) { return x > 0.0 ? x : 0.0; }
static double clamp01_(double x) { const double t = x > 0.0 ? x : 0.0; return t < 1.0 ? t : 1.0; }
#define WL 32
static double w_sum(const double *v) { /* butterfly all-reduce, offsets 16..1 */
  double a[WL], b[WL];
  memcpy(a, v, sizeof(a));
  for (int o = 16; o > 0; o >>= 1) {
    for (int i = 0; i < WL; i++) b[i] = a[i] + a[i ^ o];
    memcpy(a, b, sizeof(a));
  }
  return a[0];
}
static void w_scan(double *v) { /* inclusive Kogge-Stone prefix sum */
  double b[WL];
  for (int o = 1; o < WL; o <<= 1) {
    for (int i = 0; i < WL; i++) b[i] = (i >= o) ? v[i] + v[i - o] : v[i];
    memcpy(v, b, sizeof(b));
  }
}
static void w_suffix(double *v) { /* inclusive suffix sum */
  double b[WL];
  for (int o = 1; o < WL; o <<= 1) {
    for (int i = 0; i < WL; i++) b[i] = (i + o < WL) ? v[i] + v[i + o] : v[i];
    memcpy(v, b, sizeof(b));
  }
}
static double pdot(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

typedef struct { /* what stage_scene builds on the device */
  double seg[5][MAXN];
  double os[64 * 3 * MAXEDGE];
  double dyn[10][MAXDYN][MAXN];
  double D[MAXDYN];
  long long n_body;
} wstage_t;

static void w_stage(const ttmpc_config *g, const double *p, wstage_t *W) {
  const offs_t o = offsets(g);
  const int N = g->N_hor, ne = g->nstcobs / 3;
  const double *r = p + o.r, *os = p + o.os, *od = p + o.od, *qdyn = p + o.qdyn;
  for (int j = 0; j < N; j++) {
    int j2 = (j + 1 < N) ? j + 1 : N - 1;
    double s1x = r[3 * j], s1y = r[3 * j + 1];
    double sx = r[3 * j2] - s1x, sy = r[3 * j2 + 1] - s1y;
    double den = fma(sy, sy, sx * sx) + 1e-16;
    W->seg[0][j] = s1x; W->seg[1][j] = s1y; W->seg[2][j] = sx; W->seg[3][j] = sy;
    W->seg[4][j] = 1.0 / den;
  }
  for (int i = 0; i < g->Nstcobs * g->nstcobs; i++) {
    int e = i % g->nstcobs;
    W->os[i] = (e < ne) ? os[i] : -os[i];
  }
  for (int j = 0; j < g->Ndynobs; j++)
    for (int k = 0; k < N; k++) {
      const double *e = od + ((size_t)j * N + k) * 6;
      double cx = e[0], cy = e[1], rx = e[2], ry = e[3], ang = e[4], alpha = e[5];
      double sa, ca;
      tt_sincos(ang, &sa, &ca);
      double Rx = rx + 1e-6, Ry = ry + 1e-6;
      double Rxm = rx + g->social_margin + 1e-6, Rym = ry + g->social_margin + 1e-6;
      double rmax = fmax(fmax(fabs(Rx), fabs(Ry)), fmax(fabs(Rxm), fabs(Rym)));
      W->dyn[0][j][k] = cx; W->dyn[1][j][k] = cy;
      W->dyn[2][j][k] = rmax * rmax * (1.0 + 1e-9);
      W->dyn[3][j][k] = ca; W->dyn[4][j][k] = sa;
      W->dyn[5][j][k] = 1.0 / (Rx * Rx); W->dyn[6][j][k] = 1.0 / (Ry * Ry);
      W->dyn[7][j][k] = 1.0 / (Rxm * Rxm); W->dyn[8][j][k] = 1.0 / (Rym * Rym);
      W->dyn[9][j][k] = alpha * qdyn[k];
    }
  W->n_body = 0;
}

typedef struct { double psi, f, f2sq, S; } wout_t;

/* mirror of eval_psi<GRAD> in csrc/ttmpc_device.cuh; u interleaved (v0 w0 v1 w1 ...),
 * y = [ya_0..ya_{N-1}, yw_0..yw_{N-1}] (may be NULL), grad interleaved or NULL */
static wout_t w_eval(const ttmpc_config *g, const double *p, wstage_t *W, const double *u,
                     double c, const double *y, double *grad, double *st_out) {
  const offs_t o = offsets(g);
  const int N = g->N_hor, ne = g->nstcobs / 3, GRAD = grad != NULL;
  const double ts = g->ts, inv_ts = 1.0 / g->ts, h6 = g->ts / 6.0,
               veh_d2 = g->vehicle_width * g->vehicle_width;
  const double *s = p + o.s, *q = p + o.q;
  const double x0 = s[0], y0 = s[1], th0 = s[2], xg = s[3], yg = s[4], thg = s[5];
  const double v_init = s[6], w_init = s[7];
  const double qvel = q[1], rv = q[3], rw = q[4], qN = q[5], qthetaN = q[6], qrpd = q[7],
               acc_pen = q[8], wacc_pen = q[9];
  double v[WL], w[WL], tw[WL], th_in[WL], sa[WL], ca[WL], sb[WL], cb[WL], sc[WL], cc[WL];
  double Cs[WL], Ss[WL], hv[WL], dx[WL], dy[WL], X[WL], Y[WL], TH[WL];
  double cost[WL], gx[WL], gy[WL], gt[WL], S_loc[WL], gSx[WL], gSy[WL];
  double aa[WL], aw[WL], ea[WL], ew[WL], alm[WL], vr[WL];
  for (int k = 0; k < WL; k++) {
    v[k] = k < N ? u[2 * k] : 0.0;
    w[k] = k < N ? u[2 * k + 1] : 0.0;
    tw[k] = ts * w[k];
    th_in[k] = tw[k];
    cost[k] = gx[k] = gy[k] = gt[k] = S_loc[k] = gSx[k] = gSy[k] = 0.0;
    aa[k] = aw[k] = ea[k] = ew[k] = alm[k] = vr[k] = 0.0;
  }
  w_scan(th_in);
  for (int k = 0; k < WL; k++) {
    double th_ex = k ? th_in[k - 1] : 0.0;
    double tha = th0 + th_ex;
    double thb = fma(0.5, tw[k], tha), thc = tha + tw[k];
    tt_sincos(tha, &sa[k], &ca[k]);
    tt_sincos(thb, &sb[k], &cb[k]);
    tt_sincos(thc, &sc[k], &cc[k]); /* N == 32 only, see below */
  }
  /* end-of-step heading = start heading of the next step: for N < 32 the kernel takes sin / cos
     of theta_k + ts w_k from lane k + 1 (whose start heading is th0 + scan_k) instead of a third
     sincos (lane 31 reads itself; it is beyond the horizon and multiplied by v = 0) */
  if (N < WL)
    for (int k = 0; k < WL; k++) { sc[k] = sa[k + 1 < WL ? k + 1 : k]; cc[k] = ca[k + 1 < WL ? k + 1 : k]; }
  for (int k = 0; k < WL; k++) {
    Cs[k] = fma(4.0, cb[k], ca[k]) + cc[k];
    Ss[k] = fma(4.0, sb[k], sa[k]) + sc[k];
    hv[k] = h6 * v[k];
    dx[k] = hv[k] * Cs[k]; dy[k] = hv[k] * Ss[k];
    X[k] = dx[k]; Y[k] = dy[k];
  }
  w_scan(X); w_scan(Y);
  for (int k = 0; k < WL; k++) { X[k] = x0 + X[k]; Y[k] = y0 + Y[k]; TH[k] = th0 + th_in[k]; }
  if (st_out)
    for (int k = 0; k < N; k++) { st_out[3 * k] = X[k]; st_out[3 * k + 1] = Y[k]; st_out[3 * k + 2] = TH[k]; }

  unsigned long long hard_mask = 0;
  for (int j = 0; j < g->Ndynobs; j++) W->D[j] = 0.0;
  double in1_all[MAXDYN][WL];
  for (int k = 0; k < N; k++) { /* per-lane work, lanes are independent here */
    /* reference path */
    double dmin = INFINITY; int jmin = k;
    for (int j = k; j < N; j++) {
      const double s1x = W->seg[0][j], s1y = W->seg[1][j], sx = W->seg[2][j], sy = W->seg[3][j],
                   inv = W->seg[4][j];
      const double px = X[k] - s1x, py = Y[k] - s1y;
      const double t_hat = fma(py, sy, px * sx) * inv;
      const double t = clamp01_(t_hat);
      const double qx = fma(t, sx, -px), qy = fma(t, sy, -py);
      const double d2 = fma(qy, qy, qx * qx);
      if (!(dmin <= d2)) { dmin = d2; jmin = j; }
    }
    cost[k] = dmin * qrpd;
    if (GRAD) {
      const int j = jmin;
      const double s1x = W->seg[0][j], s1y = W->seg[1][j], sx = W->seg[2][j], sy = W->seg[3][j],
                   inv = W->seg[4][j];
      const double px = X[k] - s1x, py = Y[k] - s1y;
      const double t_hat = fma(py, sy, px * sx) * inv;
      const double t = clamp01_(t_hat);
      const double qx = fma(t, sx, -px), qy = fma(t, sy, -py);
      const double pass = (t_hat >= 0.0 && t_hat <= 1.0) ? 1.0 : 0.0;
      const double cs = fma(qy, sy, qx * sx) * pass * inv;
      gx[k] = qrpd * (2.0 * fma(cs, sx, -qx));
      gy[k] = qrpd * (2.0 * fma(cs, sy, -qy));
    }
    /* speed reference + control action */
    vr[k] = p[o.vref + k];
    {
      const double dv_ = v[k] - vr[k];
      cost[k] += qvel * (dv_ * dv_);
      cost[k] += fma(rw, w[k] * w[k], rv * (v[k] * v[k]));
    }
    /* fleet */
    {
      const double *cp = p + o.c + 3 * k;
      double acc = 0.0, fx = 0.0, fy = 0.0;
      for (int j = 0; j < g->Nother; j++) {
        const double ox = cp[(size_t)j * 3 * N], oy = cp[(size_t)j * 3 * N + 1];
        const double ex = X[k] - ox, ey = Y[k] - oy;
        const double e = veh_d2 - fma(ey, ey, ex * ex);
        if (e > 0.0) {
          acc += e;
          if (GRAD) { fx = fma(-2.0, ex, fx); fy = fma(-2.0, ey, fy); }
        }
      }
      cost[k] += 1000.0 * acc;
      if (GRAD) { gx[k] = fma(1000.0, fx, gx[k]); gy[k] = fma(1000.0, fy, gy[k]); }
    }
    /* dynamic obstacles */
    {
      double soft = 0.0;
      for (int j = 0; j < g->Ndynobs; j++) {
        const double ex = X[k] - W->dyn[0][j][k], ey = Y[k] - W->dyn[1][j][k];
        const int pass = fma(ey, ey, ex * ex) < W->dyn[2][j][k];
        double in1 = 0.0;
        if (pass) {
          W->n_body++;
          const double ca_ = W->dyn[3][j][k], sa_ = W->dyn[4][j][k];
          const double A = fma(ey, sa_, ex * ca_), B = fma(-ey, ca_, ex * sa_);
          const double A2 = A * A, B2 = B * B;
          in1 = fma(-B2, W->dyn[6][j][k], fma(-A2, W->dyn[5][j][k], 1.0));
          const double iRxm = W->dyn[7][j][k], iRym = W->dyn[8][j][k];
          const double in2 = fma(-B2, iRym, fma(-A2, iRxm, 1.0));
          if (in2 > 0.0) {
            const double ws = W->dyn[9][j][k];
            soft = fma(in2 * in2, ws, soft);
            if (GRAD) {
              const double wg = ws * (2.0 * in2);
              const double tA = A * iRxm, tB = B * iRym;
              gx[k] = fma(wg, -2.0 * fma(tB, sa_, tA * ca_), gx[k]);
              gy[k] = fma(wg, -2.0 * fma(-tB, ca_, tA * sa_), gy[k]);
            }
          }
        }
        in1_all[j][k] = in1 > 0.0 ? in1 : 0.0;
        if (in1 > 0.0) hard_mask |= 1ull << j;
      }
      cost[k] += soft;
    }
    /* terminal; the kernel skips the block when both weights are zero (every term is a signed zero then) */
    if (k == N - 1 && (qN != 0.0 || qthetaN != 0.0)) {
      const double dxg = X[k] - xg, dyg = Y[k] - yg, dtg = TH[k] - thg;
      cost[k] += fma(qthetaN, dtg * dtg, qN * fma(dyg, dyg, dxg * dxg));
      if (GRAD) {
        gx[k] = fma(2.0 * qN, dxg, gx[k]);
        gy[k] = fma(2.0 * qN, dyg, gy[k]);
        gt[k] = 2.0 * qthetaN * dtg;
      }
    }
    /* static obstacles */
    for (int i = 0; i < g->Nstcobs; i++) {
      const double *b = W->os + i * g->nstcobs, *na0 = b + ne, *na1 = b + 2 * ne;
      double m[MAXEDGE], sq[MAXEDGE], inside = 1.0;
      for (int e = 0; e < ne; e++) {
        const double res = fma(na1[e], Y[k], fma(na0[e], X[k], b[e]));
        m[e] = relu_(res);
        sq[e] = m[e] * m[e];
        inside *= sq[e];
      }
      if (inside > 0.0) {
        S_loc[k] += inside;
        if (GRAD)
          for (int e = 0; e < ne; e++) {
            double rest = 1.0;
            for (int e2 = 0; e2 < ne; e2++) if (e2 != e) rest *= sq[e2];
            const double coef = rest * (2.0 * m[e]);
            gSx[k] = fma(coef, na0[e], gSx[k]);
            gSy[k] = fma(coef, na1[e], gSy[k]);
          }
      }
    }
    /* accelerations + ALM */
    {
      const double vp = k ? v[k - 1] : v_init, wp = k ? w[k - 1] : w_init;
      aa[k] = (v[k] - vp) * inv_ts; aw[k] = (w[k] - wp) * inv_ts;
      cost[k] += fma(aw[k] * aw[k], wacc_pen, (aa[k] * aa[k]) * acc_pen);
      const double icm = 1.0 / fmax(c, 1.0);
      double z = fma(y ? y[k] : 0.0, icm, aa[k]);
      ea[k] = z - clip(z, g->lin_acc_min, g->lin_acc_max);
      z = fma(y ? y[N + k] : 0.0, icm, aw[k]);
      ew[k] = z - clip(z, -g->ang_acc_max, g->ang_acc_max);
      alm[k] = fma(ew[k], ew[k], ea[k] * ea[k]);
    }
  }
  /* D_j: butterfly sum over lanes, only for obstacles with a positive term */
  for (int j = 0; j < g->Ndynobs; j++)
    if (hard_mask >> j & 1ull) {
      double t[WL];
      for (int k = 0; k < WL; k++) t[k] = k < N ? in1_all[j][k] : 0.0;
      W->D[j] = w_sum(t);
    }
  const double f = w_sum(cost), d2 = w_sum(alm), S = w_sum(S_loc);
  double f2sq = 0.0, sumF2 = 0.0;
  for (int j = 0; j < g->Ndynobs; j++) {
    const double F2j = S + W->D[j];
    f2sq = fma(F2j, F2j, f2sq);
    sumF2 += F2j;
  }
  wout_t out;
  out.f = f; out.f2sq = f2sq; out.S = S;
  out.psi = f + c * d2 / 2 + c * f2sq / 2;
  if (GRAD) {
    if (c != 0.0) {
      const double cs_ = c * sumF2;
      for (int k = 0; k < N; k++) {
        gx[k] = fma(cs_, gSx[k], gx[k]);
        gy[k] = fma(cs_, gSy[k], gy[k]);
        for (int j = 0; j < g->Ndynobs; j++) {
          if (!(hard_mask >> j & 1ull)) continue;
          const double ex = X[k] - W->dyn[0][j][k], ey = Y[k] - W->dyn[1][j][k];
          if (fma(ey, ey, ex * ex) < W->dyn[2][j][k]) {
            const double ca_ = W->dyn[3][j][k], sa_ = W->dyn[4][j][k];
            const double iRx = W->dyn[5][j][k], iRy = W->dyn[6][j][k];
            const double A = fma(ey, sa_, ex * ca_), B = fma(-ey, ca_, ex * sa_);
            const double in1 = fma(-(B * B), iRy, fma(-(A * A), iRx, 1.0));
            if (in1 > 0.0) {
              const double wg = c * (S + W->D[j]);
              const double tA = A * iRx, tB = B * iRy;
              gx[k] = fma(wg, -2.0 * fma(tB, sa_, tA * ca_), gx[k]);
              gy[k] = fma(wg, -2.0 * fma(-tB, ca_, tA * sa_), gy[k]);
            }
          }
        }
      }
    }
    double lx[WL], ly[WL], m[WL], lt[WL];
    memcpy(lx, gx, sizeof(lx)); memcpy(ly, gy, sizeof(ly));
    w_suffix(lx); w_suffix(ly);
    for (int k = 0; k < WL; k++) {
      m[k] = fma(ly[k], dx[k], -(lx[k] * dy[k]));
      lt[k] = gt[k] + m[k];
    }
    w_suffix(lt);
    for (int k = 0; k < WL; k++) lt[k] = lt[k] - m[k];
    for (int k = 0; k < N; k++) {
      const double dxdv = h6 * Cs[k], dydv = h6 * Ss[k];
      const double hvt = hv[k] * ts;
      const double dxdw = -(hvt * fma(2.0, sb[k], sc[k])), dydw = hvt * fma(2.0, cb[k], cc[k]);
      const int last = k == N - 1;
      const double aa_n = last ? 0.0 : aa[k + 1], aw_n = last ? 0.0 : aw[k + 1];
      const double ea_n = last ? 0.0 : ea[k + 1], ew_n = last ? 0.0 : ew[k + 1];
      double dv = 2 * qvel * (v[k] - vr[k]) + 2 * rv * v[k];
      double dw = 2 * rw * w[k];
      dv += (2 * acc_pen * (aa[k] - aa_n) + c * (ea[k] - ea_n)) * inv_ts;
      dw += (2 * wacc_pen * (aw[k] - aw_n) + c * (ew[k] - ew_n)) * inv_ts;
      grad[2 * k] = dv + lx[k] * dxdv + ly[k] * dydv;
      grad[2 * k + 1] = dw + lx[k] * dxdw + ly[k] * dydw + lt[k] * ts;
    }
  }
  return out;
}

/* public: warp-ordered evaluation (testing) */
void ttmpc_oracle_eval_warp(const ttmpc_config *g, const double *u, const double *p, double c,
                            const double *y, double *f, double *F2, double *psi, double *grad) {
  wstage_t *W = (wstage_t *)malloc(sizeof(wstage_t));
  w_stage(g, p, W);
  wout_t o = w_eval(g, p, W, u, c, y, grad, NULL);
  if (f) *f = o.f;
  if (psi) *psi = o.psi;
  if (F2) for (int j = 0; j < g->Ndynobs; j++) F2[j] = o.S + W->D[j];
  free(W);
}

/* ------------------------------------------------------------------ */
/* Part 2: OpEn solver restatement (PARITY UNPINNED)                   */
/* ------------------------------------------------------------------ */

typedef struct {
  const ttmpc_config *g;
  const double *p;
  double c;        /* xi[0] */
  double *y;       /* xi[1..] */
  long long n_cost, n_grad;
  int warp;        /* 0: reference order (Part 1, libm)  1: GPU order (Part 1b) */
  wstage_t *W;
  int mock;        /* 0: the NMPC problem; 1, 2: the crate's own unit-test problems (see ttmpc_oracle_panoc_mock) */
} prob_t;

/* The two test problems of optimization_engine's src/mocks.rs (my_cost / my_gradient over a Euclidean ball of
 * radius 0.2, hard_quadratic_cost / hard_quadratic_gradient over a ball of radius 0.05), restated from the crate;
 * the crate's unit tests (t_panoc_basic and friends) assert that PANOC reaches mocks::SOLUTION_A / SOLUTION_HARD
 * on them.  tests/test_oracle_solver.py runs the PANOC engine below on both and checks the constants (and,
 * independently of anybody's memory of them, the KKT conditions of the two problems at the point reached). */
static double mock_cost(int which, const double *u) {
  if (which == 3) return 0.0;
  if (which == 1)
    return 0.5 * (u[0] * u[0] + 2. * u[1] * u[1] + 2.0 * u[0] * u[1]) + u[0] - u[1] + 3.0;
  return (4. * u[0] * u[0]) / 2. + 5.5 * u[1] * u[1] + 500.5 * u[2] * u[2] + 5. * u[0] * u[1] +
         25. * u[0] * u[2] + 5. * u[1] * u[2] + u[0] + u[1] + u[2];
}
static void mock_grad(int which, const double *u, double *g) {
  if (which == 3) { /* mocks::lipschitz_mock */
    g[0] = 3.0 * u[0]; g[1] = 2.0 * u[1]; g[2] = 4.5;
  } else if (which == 1) {
    g[0] = u[0] + u[1] + 1.0;
    g[1] = u[0] + 2. * u[1] - 1.0;
  } else {
    g[0] = 4. * u[0] + 5. * u[1] + 25. * u[2] + 1.;
    g[1] = 5. * u[0] + 11. * u[1] + 5. * u[2] + 1.;
    g[2] = 25. * u[0] + 5. * u[1] + 1001. * u[2] + 1.;
  }
}
/* constraints::Ball2::new(None, r).project */
static void mock_project(int which, double *u) {
  const int n = which == 1 ? 2 : 3;
  const double r = which == 1 ? 0.2 : 0.05;
  double nrm = 0.0;
  for (int i = 0; i < n; i++) nrm += u[i] * u[i];
  nrm = sqrt(nrm);
  if (nrm > r) for (int i = 0; i < n; i++) u[i] *= r / nrm;
}

static void f_cost(prob_t *pb, const double *u, double *out) {
  if (pb->mock) *out = mock_cost(pb->mock, u);
  else if (pb->warp) *out = w_eval(pb->g, pb->p, pb->W, u, pb->c, pb->y, NULL, NULL).psi;
  else *out = ttmpc_oracle_psi(pb->g, u, pb->p, pb->c, pb->y);
  pb->n_cost++;
}
static void f_grad(prob_t *pb, const double *u, double *out) {
  if (pb->mock) mock_grad(pb->mock, u, out);
  else if (pb->warp) w_eval(pb->g, pb->p, pb->W, u, pb->c, pb->y, out, NULL);
  else ttmpc_oracle_psi_grad(pb->g, u, pb->p, pb->c, pb->y, out);
  pb->n_grad++;
}
static int g_warp_mode = 0; /* set per solve; the vector helpers and L-BFGS read it */
/* a*b + c.  The CUDA kernel fuses these products (explicit fma(), mirrored in WARP order); Rust
 * never contracts a multiply-add, so the REFERENCE order rounds the product first, exactly like
 * OpEn's `*u - gamma * *grad`, `u - temp_ * fpr - tau * dir`, `out[i] += s * a[i]`. */
static double mad(double a, double b, double c) { return g_warp_mode ? fma(a, b, c) : a * b + c; }
/* og.constraints.Rectangle(umin, umax) (mpc_generator.py:245-247) */
static void project_u(const ttmpc_config *g, double *u) {
  for (int k = 0; k < g->N_hor; k++) {
    u[2 * k] = clip(u[2 * k], g->lin_vel_min, g->lin_vel_max);
    u[2 * k + 1] = clip(u[2 * k + 1], -g->ang_vel_max, g->ang_vel_max);
  }
}

static double dot(int n, const double *a, const double *b) {
  if (g_warp_mode) {
    double t[WL];
    for (int k = 0; k < WL; k++)
      t[k] = (2 * k + 1 < n) ? pdot(a[2 * k], a[2 * k + 1], b[2 * k], b[2 * k + 1]) : 0.0;
    return w_sum(t);
  }
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static double norm2(int n, const double *a) { return sqrt(dot(n, a, a)); }
static double norm2sq_diff(int n, const double *a, const double *b) {
  double d[MAXNU];
  for (int i = 0; i < n; i++) d[i] = a[i] - b[i];
  return dot(n, d, d);
}

/* ---- lbfgs crate: Lbfgs with C-BFGS (Li & Fukushima) safeguard ---- */
typedef struct {
  int n, mem, active, head, first_old;
  double gamma, cbfgs_alpha, cbfgs_eps, sy_eps;
  double s[MAXMEM + 1][MAXNU], y[MAXMEM + 1][MAXNU];
  double rho[MAXMEM + 1], alpha[MAXMEM];
  double old_state[MAXNU], old_g[MAXNU];
} lbfgs_t;

static int lb_idx(const lbfgs_t *l, int i) { return (l->head + i) % (l->mem + 1); }
static void lb_init(lbfgs_t *l, int n, int mem) {
  memset(l, 0, sizeof(*l));
  l->n = n; l->mem = mem; l->gamma = 1.0; l->first_old = 1;
  /* PANOCCache::new: with_cbfgs_alpha(1.0).with_cbfgs_epsilon(1e-8).with_sy_epsilon(1e-10) */
  l->cbfgs_alpha = 1.0; l->cbfgs_eps = 1e-8; l->sy_eps = 1e-10;
}
static void lb_reset(lbfgs_t *l) { l->active = 0; l->first_old = 1; }
/* lbfgs crate apply_hessian: the literal two-loop recursion.  The CUDA kernel runs exactly this
 * (csrc/ttmpc_solve.cu lbfgs_apply); in WARP order dot() is the butterfly all-reduce and mad() fuses. */
static void lb_apply(lbfgs_t *l, double *q) {
  if (l->active == 0) return;
  const int n = l->n;
  for (int i = 0; i < l->active; i++) {
    int k = lb_idx(l, i);
    double a = l->rho[k] * dot(n, l->s[k], q);
    l->alpha[i] = a;
    for (int t = 0; t < n; t++) q[t] = mad(-a, l->y[k][t], q[t]);
  }
  for (int t = 0; t < n; t++) q[t] *= l->gamma;
  for (int i = l->active - 1; i >= 0; i--) {
    int k = lb_idx(l, i);
    double beta = l->rho[k] * dot(n, l->y[k], q);
    double cf = l->alpha[i] - beta;
    for (int t = 0; t < n; t++) q[t] = mad(cf, l->s[k][t], q[t]);
  }
}
/* returns 1 if accepted; norm_g = |g| (the caller already has it) */
static int lb_update(lbfgs_t *l, const double *g, const double *state, double norm_g) {
  const int n = l->n;
  if (l->first_old) {
    l->first_old = 0;
    memcpy(l->old_state, state, n * sizeof(double));
    memcpy(l->old_g, g, n * sizeof(double));
    return 1;
  }
  int last = lb_idx(l, l->mem);
  for (int t = 0; t < n; t++) {
    l->s[last][t] = state[t] - l->old_state[t];
    l->y[last][t] = g[t] - l->old_g[t];
  }
  double ys = dot(n, l->s[last], l->y[last]);
  double ss = dot(n, l->s[last], l->s[last]);
  double yy = dot(n, l->y[last], l->y[last]);
  l->rho[last] = 1.0 / ys;
  if (ss <= DBL_MIN || (l->sy_eps > 0.0 && ys <= l->sy_eps)) return 0;
  if (l->cbfgs_eps > 0.0 && l->cbfgs_alpha > 0.0) {
    double lhs = ys / ss;
    double rhs = l->cbfgs_eps * norm_g; /* pow(|g|, cbfgs_alpha) with alpha = 1 */
    if (!(lhs > rhs && isfinite(lhs) && isfinite(rhs))) return 0;
  }
  memcpy(l->old_state, state, n * sizeof(double));
  memcpy(l->old_g, g, n * sizeof(double));
  /* rotate_right(1): the scratch slot becomes slot 0 */
  l->head = (l->head + l->mem) % (l->mem + 1);
  int k0 = lb_idx(l, 0);
  l->gamma = (1.0 / l->rho[k0]) / yy;
  l->active = (l->mem < l->active + 1) ? l->mem : l->active + 1;
  return 1;
}

/* ---- PANOC (panoc_engine.rs) ---- */
#define GAMMA_L_COEFF 0.95
#define DELTA_LIPSCHITZ 1e-12
#define EPSILON_LIPSCHITZ 1e-6
#define LIPSCHITZ_UPDATE_EPSILON 1e-6
#define MAX_LIPSCHITZ_UPDATE_ITERATIONS 10
#define MAX_LIPSCHITZ_CONSTANT 1e9
#define MAX_LINESEARCH_ITERATIONS 10
#define MIN_L_ESTIMATE 1e-10

typedef struct {
  int n;
  lbfgs_t lb;
  double grad[MAXNU], grad_prev[MAXNU], u_half[MAXNU], gstep[MAXNU],
      dir[MAXNU], u_plus[MAXNU], fpr[MAXNU];
  double rhs_ls, lhs_ls, gamma, tolerance, norm_fpr, tau, L, sigma, cost;
  double akkt_tol;
  int iteration;
} panoc_t;

static void pc_reset(panoc_t *c) {
  lb_reset(&c->lb);
  c->lhs_ls = c->rhs_ls = 0.0; c->tau = 1.0; c->L = 0.0; c->sigma = 0.0;
  c->cost = 0.0; c->iteration = 0; c->gamma = 0.0;
}
static void pc_set_akkt(panoc_t *c, double tol) {
  c->akkt_tol = tol;
  for (int i = 0; i < c->n; i++) c->grad_prev[i] = 0.0; /* fresh zero vector */
}
static int pc_exit(const panoc_t *c) {
  if (!(c->norm_fpr < c->tolerance)) return 0;
  double r[MAXNU];
  for (int i = 0; i < c->n; i++) r[i] = mad(c->gamma, c->grad[i] - c->grad_prev[i], c->fpr[i]);
  return sqrt(dot(c->n, r, r)) < c->akkt_tol;
}
static void pe_gradient_step(panoc_t *c, const double *u) {
  for (int i = 0; i < c->n; i++) c->gstep[i] = mad(-c->gamma, c->grad[i], u[i]);
}
static void pe_half_step(panoc_t *c, const prob_t *pb) {
  memcpy(c->u_half, c->gstep, c->n * sizeof(double));
  if (pb->mock) mock_project(pb->mock, c->u_half);
  else project_u(pb->g, c->u_half);
}
static void pe_fpr(panoc_t *c, const double *u) {
  for (int i = 0; i < c->n; i++) c->fpr[i] = u[i] - c->u_half[i];
  c->norm_fpr = norm2(c->n, c->fpr);
}
static void pe_init(panoc_t *c, prob_t *pb, double *u) {
  const int n = c->n;
  pc_reset(c);
  f_cost(pb, u, &c->cost);
  /* LipschitzEstimator: h_i = max(delta, eps*u_i); L = |grad(u+h)-grad(u)|/|h| */
  {
    double h[MAXNU] = {0.0}, up[MAXNU], g2[MAXNU];
    f_grad(pb, u, c->grad);
    for (int i = 0; i < n; i++) {
      h[i] = (EPSILON_LIPSCHITZ * u[i] > DELTA_LIPSCHITZ) ? EPSILON_LIPSCHITZ * u[i]
                                                         : DELTA_LIPSCHITZ;
      up[i] = u[i] + h[i];
    }
    double nh = dot(n, h, h);
    f_grad(pb, up, g2);
    c->L = sqrt(norm2sq_diff(n, g2, c->grad)) / sqrt(nh);
  }
  c->gamma = GAMMA_L_COEFF / fmax(c->L, MIN_L_ESTIMATE);
  c->sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * c->gamma);
  pe_gradient_step(c, u);
  pe_half_step(c, pb);
}
static double pe_lip_rhs(panoc_t *c) {
  double ip = dot(c->n, c->grad, c->fpr);
  return c->cost + LIPSCHITZ_UPDATE_EPSILON * fabs(c->cost) - ip +
         (GAMMA_L_COEFF / (2.0 * c->gamma)) * (c->norm_fpr * c->norm_fpr);
}
static void pe_update_lipschitz(panoc_t *c, prob_t *pb, const double *u) {
  double cost_half = 0.0;
  f_cost(pb, c->u_half, &cost_half);
  f_cost(pb, u, &c->cost);
  int it = 0;
  while (cost_half > pe_lip_rhs(c) && it < MAX_LIPSCHITZ_UPDATE_ITERATIONS &&
         c->L < MAX_LIPSCHITZ_CONSTANT) {
    lb_reset(&c->lb);
    c->L *= 2.0;
    c->gamma /= 2.0;
    pe_gradient_step(c, u);
    pe_half_step(c, pb);
    f_cost(pb, c->u_half, &cost_half);
    pe_fpr(c, u);
    it++;
  }
  c->sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * c->gamma);
}
static int pe_ls_condition(panoc_t *c, prob_t *pb, const double *u) {
  const int n = c->n;
  const double tau = c->tau, one_m = 1.0 - tau;
  for (int i = 0; i < n; i++) c->u_plus[i] = mad(-tau, c->dir[i], mad(-one_m, c->fpr[i], u[i]));
  f_cost(pb, c->u_plus, &c->cost);
  f_grad(pb, c->u_plus, c->grad);
  for (int i = 0; i < n; i++) c->gstep[i] = mad(-c->gamma, c->grad[i], c->u_plus[i]);
  pe_half_step(c, pb);
  const double dd = norm2sq_diff(n, c->u_half, c->gstep), g2 = dot(n, c->grad, c->grad);
  c->lhs_ls = c->cost - 0.5 * c->gamma * g2 + 0.5 * dd / c->gamma;
  return c->lhs_ls > c->rhs_ls;
}
/* returns 1 to continue */
static int pe_step(panoc_t *c, prob_t *pb, double *u) {
  const int n = c->n;
  if (c->iteration >= 1) memcpy(c->grad_prev, c->grad, n * sizeof(double));
  pe_fpr(c, u);
  if (pc_exit(c)) return 0;
  pe_update_lipschitz(c, pb, u);
  /* lbfgs_direction */
  lb_update(&c->lb, c->fpr, u, c->norm_fpr);
  if (c->iteration > 0) {
    memcpy(c->dir, c->fpr, n * sizeof(double));
    lb_apply(&c->lb, c->dir);
  }
  if (c->iteration == 0) {
    /* update_no_linesearch */
    memcpy(u, c->u_half, n * sizeof(double));
    f_cost(pb, u, &c->cost);
    f_grad(pb, u, c->grad);
    pe_gradient_step(c, u);
    pe_half_step(c, pb);
  } else {
    /* linesearch */
    const double dist2 = norm2sq_diff(n, c->gstep, c->u_half), gg = dot(n, c->grad, c->grad);
    const double fbe = c->cost - 0.5 * c->gamma * gg + 0.5 * dist2 / c->gamma;
    c->rhs_ls = fbe - c->sigma * (c->norm_fpr * c->norm_fpr);
    c->tau = 1.0;
    int nls = 0;
    while (pe_ls_condition(c, pb, u) && nls < MAX_LINESEARCH_ITERATIONS) {
      c->tau /= 2.0;
      nls++;
    }
    if (nls == MAX_LINESEARCH_ITERATIONS) c->tau = 0.0;
    memcpy(u, c->u_plus, n * sizeof(double));
  }
  c->iteration++;
  return 1;
}
/* PANOCOptimizer::solve; returns exit status, *iters = num_iter */
static int panoc_solve(panoc_t *c, prob_t *pb, double *u, int max_iter, int *iters) {
  pe_init(c, pb, u);
  int num_iter = 0, cont_iters = 1;
  int flag = pe_step(c, pb, u);
  while (flag && cont_iters) {
    num_iter++;
    cont_iters = num_iter < max_iter;
    flag = pe_step(c, pb, u);
  }
  *iters = num_iter;
  for (int i = 0; i < c->n; i++)
    if (!isfinite(u[i])) return TTMPC_NOT_FINITE;
  memcpy(u, c->u_half, c->n * sizeof(double));
  return cont_iters ? TTMPC_CONVERGED : TTMPC_NOT_CONVERGED_ITERATIONS;
}

/* The L-BFGS restatement on a caller-supplied history: n_updates calls of update_hessian(g_i, x_i) (the first
 * one only stores the point, like in the crate), then apply_hessian(q).  cbfgs = 0: Lbfgs::new defaults (no
 * C-BFGS test), 1: the settings PANOCCache::new chooses.  tests/test_oracle_solver.py checks the known answer of
 * the lbfgs crate's own unit test (correctneess_buff_1) with it.  alpha0 / rho0 = alpha and rho of the newest pair. */
int ttmpc_oracle_lbfgs_kat(int n, int mem, int n_updates, const double *g, const double *x, double *q,
                           double *alpha0, double *rho0, int cbfgs) {
  if (n < 1 || n > MAXNU || mem < 1 || mem > MAXMEM) return -1;
  lbfgs_t *l = (lbfgs_t *)calloc(1, sizeof(lbfgs_t));
  g_warp_mode = 0;
  lb_init(l, n, mem);
  if (!cbfgs) { l->cbfgs_alpha = 0.0; l->cbfgs_eps = 0.0; }
  int accepted = 0;
  for (int i = 0; i < n_updates; i++)
    accepted += lb_update(l, g + (size_t)i * n, x + (size_t)i * n, norm2(n, g + (size_t)i * n));
  lb_apply(l, q);
  if (alpha0) *alpha0 = l->alpha[0];
  if (rho0) *rho0 = l->rho[lb_idx(l, 0)];
  const int active = l->active;
  free(l);
  return 100 * accepted + active;
}

/* PANOCEngine::init's local Lipschitz estimate (LipschitzEstimator with PANOC's delta = 1e-12, epsilon = 1e-6) of
 * mocks::lipschitz_mock at u[3]; the crate's t_test_lip_estimator_mock expects 1.336306209562 at (1, 2, 3). */
double ttmpc_oracle_lipschitz_mock(const double *u3) {
  panoc_t *pc = (panoc_t *)calloc(1, sizeof(panoc_t));
  prob_t pb = {NULL, NULL, 0.0, NULL, 0, 0, 0, NULL, 3};
  double u[3] = {u3[0], u3[1], u3[2]};
  g_warp_mode = 0;
  pc->n = 3;
  lb_init(&pc->lb, 3, 3);
  pc_reset(pc);
  pc->akkt_tol = INFINITY;
  pe_init(pc, &pb, u);
  const double L = pc->L;
  free(pc);
  return L;
}

/* PANOCOptimizer::solve on one of the crate's unit-test problems (which = 1: my_cost, 2 variables, ball 0.2;
 * which = 2: hard_quadratic_cost, 3 variables, ball 0.05), reference operation order, no AKKT test (a plain
 * PANOCCache has none).  Returns the exit status; u in/out. */
int ttmpc_oracle_panoc_mock(int which, double *u, double tolerance, int lbfgs_memory, int max_iter,
                            int *iters, double *norm_fpr, long long *n_cost, long long *n_grad) {
  if (which != 1 && which != 2) return -1;
  if (lbfgs_memory < 1 || lbfgs_memory > MAXMEM) return -1;
  panoc_t *pc = (panoc_t *)calloc(1, sizeof(panoc_t));
  prob_t pb = {NULL, NULL, 0.0, NULL, 0, 0, 0, NULL, which};
  g_warp_mode = 0;
  pc->n = which == 1 ? 2 : 3;
  lb_init(&pc->lb, pc->n, lbfgs_memory);
  pc->tolerance = tolerance;
  pc_reset(pc);
  pc->akkt_tol = INFINITY; /* akkt_tolerance: None */
  int it = 0;
  const int st = panoc_solve(pc, &pb, u, max_iter, &it);
  if (iters) *iters = it;
  if (norm_fpr) *norm_fpr = pc->norm_fpr;
  if (n_cost) *n_cost = pb.n_cost;
  if (n_grad) *n_grad = pb.n_grad;
  free(pc);
  return st;
}

/* ---- ALM / PM outer loop (alm_optimizer.rs) ---- */
static int solve_mode(const ttmpc_config *g, const double *p, double *u, double *y, double c0,
                      ttmpc_oracle_status *st, int warp) {
  const int N = g->N_hor, n = 2 * N, n1 = 2 * N, n2 = g->Ndynobs;
  panoc_t *pc = (panoc_t *)calloc(1, sizeof(panoc_t));
  prob_t pb = {g, p, c0, y, 0, 0, warp, NULL, 0};
  double y_plus[MAXNU], w1[MAXNU], w2[MAXDYN];
  double delta_y_norm = 0.0, delta_y_norm_plus = 0.0, f2_norm = 0.0, f2_norm_plus = 0.0;
  double last_fpr = 0.0, f_final = 0.0;
  int alm_iter = 0, inner_count = 0, num_outer = 0, exit_status = TTMPC_CONVERGED;
  const double SMALL_EPSILON = DBL_EPSILON;
  g_warp_mode = warp;
  if (warp) { pb.W = (wstage_t *)malloc(sizeof(wstage_t)); w_stage(g, p, pb.W); }

  pc->n = n;
  lb_init(&pc->lb, n, g->lbfgs_memory);
  pc->tolerance = g->tolerance;
  pc_reset(pc);
  pc_set_akkt(pc, g->initial_tolerance);

  for (int outer = 1; outer <= g->max_outer_iterations; outer++) {
    num_outer++;
    /* step(): project y on Y = [-1e12, 1e12]^n1 */
    for (int i = 0; i < n1; i++) y[i] = clip(y[i], -1e12, 1e12);
    int it = 0;
    int inner_status = panoc_solve(pc, &pb, u, g->max_inner_iterations, &it);
    inner_count += it;
    if (inner_status == TTMPC_NOT_FINITE) { exit_status = TTMPC_NOT_FINITE; break; }
    last_fpr = pc->norm_fpr;
    /* update Lagrange multipliers: y+ = y + c (F1(u) - Proj_C(F1(u) + y/max(c,1))) */
    if (warp) {
      wout_t e = w_eval(g, p, pb.W, u, 0.0, NULL, NULL, NULL);
      pb.n_cost++;
      f2_norm_plus = sqrt(e.f2sq);
      f_final = e.f;
      const double inv_ts = 1.0 / g->ts;
      for (int k = 0; k < N; k++) {
        w1[k] = (u[2 * k] - (k ? u[2 * (k - 1)] : p[6])) * inv_ts;
        w1[N + k] = (u[2 * k + 1] - (k ? u[2 * (k - 1) + 1] : p[7])) * inv_ts;
      }
    } else {
      ttmpc_oracle_eval(g, u, p, &f_final, w1, w2);
      pb.n_cost++;
      f2_norm_plus = norm2(n2, w2);
    }
    for (int i = 0; i < n1; i++) {
      double lo = i < N ? g->lin_acc_min : -g->ang_acc_max;
      double hi = i < N ? g->lin_acc_max : g->ang_acc_max;
      double t = warp ? clip(fma(y[i], 1.0 / fmax(pb.c, 1.0), w1[i]), lo, hi)
                      : clip(w1[i] + y[i] / fmax(pb.c, 1.0), lo, hi);
      y_plus[i] = y[i] + pb.c * (w1[i] - t);
    }
    if (warp) { /* lane k pairs the k-th linear and angular rows */
      double t[WL];
      for (int k = 0; k < WL; k++) {
        double da = k < N ? y_plus[k] - y[k] : 0.0, dw = k < N ? y_plus[N + k] - y[N + k] : 0.0;
        t[k] = pdot(da, dw, da, dw);
      }
      delta_y_norm_plus = sqrt(w_sum(t));
    } else {
      delta_y_norm_plus = sqrt(norm2sq_diff(n1, y_plus, y));
    }
    /* exit criterion */
    int crit1 = alm_iter > 0 && delta_y_norm_plus <= pb.c * g->delta_tolerance + SMALL_EPSILON;
    int crit2 = f2_norm_plus <= g->delta_tolerance + SMALL_EPSILON;
    int crit3 = pc->akkt_tol <= g->tolerance + SMALL_EPSILON;
    if (crit1 && crit2 && crit3) { exit_status = inner_status; break; }
    /* penalty stall criterion */
    /* is_penalty_stall_criterion: iteration 0, or BOTH infeasibilities (ALM rows and
       penalty rows; n1 > 0 and n2 > 0 here) decreased sufficiently */
    int crit_alm = delta_y_norm_plus <= g->sufficient_decrease_coeff * delta_y_norm + SMALL_EPSILON;
    int crit_pm = n2 == 0 || f2_norm_plus <= g->sufficient_decrease_coeff * f2_norm + SMALL_EPSILON;
    int stall = alm_iter == 0 || (crit_alm && crit_pm);
    if (!stall) pb.c *= g->penalty_update_factor;
    /* inner tolerance update */
    pc_set_akkt(pc, fmax(pc->akkt_tol * g->inner_tolerance_update_factor, g->tolerance));
    /* final cache update */
    alm_iter++;
    delta_y_norm = delta_y_norm_plus;
    f2_norm = f2_norm_plus;
    memcpy(y, y_plus, n1 * sizeof(double));
    pc_reset(pc);
  }
  if (exit_status != TTMPC_NOT_FINITE && num_outer == g->max_outer_iterations)
    exit_status = TTMPC_NOT_CONVERGED_ITERATIONS;
  /* y is left as the solver cache holds it (xi[1..]): on a converged exit the
     step returns before final_cache_update, so it is the PREVIOUS y; the python
     Solver object keeps exactly this vector between run() calls.             */

  if (st) {
    st->exit_status = exit_status;
    st->outer_iters = num_outer;
    st->inner_iters = inner_count;
    st->last_fpr = last_fpr;
    st->delta_y_norm = delta_y_norm_plus;
    st->f2_norm = f2_norm_plus;
    st->penalty = pb.c;
    st->cost = f_final;
    st->n_cost_evals = pb.n_cost;
    st->n_grad_evals = pb.n_grad;
  }
  if (pb.W) free(pb.W);
  free(pc);
  return 0;
}

int ttmpc_oracle_solve(const ttmpc_config *g, const double *p, double *u, double *y, double c0,
                       ttmpc_oracle_status *st) {
  return solve_mode(g, p, u, y, c0, st, 0);
}
int ttmpc_oracle_solve_warp(const ttmpc_config *g, const double *p, double *u, double *y,
                            double c0, ttmpc_oracle_status *st) {
  return solve_mode(g, p, u, y, c0, st, 1);
}

typedef struct {
  const ttmpc_config *g;
  int n, lo, hi, use_u0, use_y0, warp;
  const double *p, *c0;
  const ttmpc_result *res;
} job_t;

static void batch_range(job_t *jb) {
  const ttmpc_config *g = jb->g;
  const int N = g->N_hor, nu = 2 * N, np = offsets(g).np;
  for (int i = jb->lo; i < jb->hi; i++) {
    double u[MAXNU], y[MAXNU];
    ttmpc_oracle_status st;
    for (int t = 0; t < nu; t++) {
      u[t] = jb->use_u0 ? jb->res->u[(size_t)i * nu + t] : 0.0;
      y[t] = (jb->use_y0 && jb->res->y) ? jb->res->y[(size_t)i * nu + t] : 0.0;
    }
    double c0 = jb->c0 ? jb->c0[i] : g->initial_penalty;
    solve_mode(g, jb->p + (size_t)i * np, u, y, c0, &st, jb->warp);
    const ttmpc_result *r = jb->res;
    memcpy(r->u + (size_t)i * nu, u, nu * sizeof(double));
    if (r->y) memcpy(r->y + (size_t)i * nu, y, nu * sizeof(double));
    if (r->cost) r->cost[i] = st.cost;
    if (r->exit_status) r->exit_status[i] = st.exit_status;
    if (r->outer_iters) r->outer_iters[i] = st.outer_iters;
    if (r->inner_iters) r->inner_iters[i] = st.inner_iters;
    if (r->last_fpr) r->last_fpr[i] = st.last_fpr;
    if (r->f1_infeas) r->f1_infeas[i] = st.delta_y_norm / st.penalty;
    if (r->f2_norm) r->f2_norm[i] = st.f2_norm;
    if (r->penalty) r->penalty[i] = st.penalty;
    if (r->pred_states) {
      if (jb->warp) {
        wstage_t *W = (wstage_t *)malloc(sizeof(wstage_t));
        w_stage(g, jb->p + (size_t)i * np, W);
        w_eval(g, jb->p + (size_t)i * np, W, u, 0.0, NULL, NULL, r->pred_states + (size_t)i * N * 3);
        free(W);
      } else {
        ttmpc_oracle_rollout(g, u, jb->p + (size_t)i * np, r->pred_states + (size_t)i * N * 3);
      }
    }
    if (r->evals) { r->evals[4 * i] = st.n_cost_evals; r->evals[4 * i + 1] = st.n_grad_evals; r->evals[4 * i + 2] = 0; r->evals[4 * i + 3] = 0; }
  }
}

/* The vector helpers read the process-wide g_warp_mode, so a batch runs in ONE
 * ordering; worker PROCESSES (fork) give the parallelism, threads would share
 * the flag safely too because every job of a batch uses the same mode.        */
static void *batch_worker(void *arg) { batch_range((job_t *)arg); return NULL; }

int ttmpc_oracle_solve_batch_mode(const ttmpc_config *g, int n, const double *p, int use_u0,
                                  int use_y0, const double *c0, const ttmpc_result *res,
                                  int threads, int warp) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads > n) threads = n > 0 ? n : 1;
  pthread_t th[256];
  job_t jobs[256];
  int per = (n + threads - 1) / threads;
  g_warp_mode = warp;
  for (int t = 0; t < threads; t++) {
    int lo = t * per, hi = (t + 1) * per < n ? (t + 1) * per : n;
    job_t jb = {g, n, lo, hi > lo ? hi : lo, use_u0, use_y0, warp, p, c0, res};
    jobs[t] = jb;
    if (threads == 1) batch_range(&jobs[t]);
    else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  if (threads > 1)
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  return 0;
}
int ttmpc_oracle_solve_batch(const ttmpc_config *g, int n, const double *p, int use_u0,
                             int use_y0, const double *c0, const ttmpc_result *res, int threads) {
  return ttmpc_oracle_solve_batch_mode(g, n, p, use_u0, use_y0, c0, res, threads, 0);
}

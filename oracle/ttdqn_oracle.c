/*
 * ttdqn_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restates, for one environment:
 *   SectorAndRayObservation.external_obs
 *     (/root/reference/src/pkg_dqn/environment/components/ext_obsv_sector_and_ray.py:33-83)
 *   normalize_distance (src/pkg_dqn/environment/components/utils.py:10-15)
 *   SB3 DQN MultiInputPolicy forward: CombinedExtractor (flatten + concat of the
 *   Dict keys in sorted order: external, internal) -> q_net MLP 46-16-16-9 with
 *   ReLU -> argmax (main.py:148 model.predict(obsv, deterministic=True)).
 *
 * Geometry semantics (shapely is not available here, PARITY UNPINNED for the
 * geometric part): for sector triangle T (apex = agent A) and ray R,
 *   filled polygon G : dist(A, T^G) = 0 if A in G else min over edges e of G of
 *                      dist(A, e^T);   dist(A, G^R) = 0 if A in G else first hit.
 *   ring (LineString): same without the "inside" rule.
 * A closest point of the closed set T^G to A always lies on the boundary of G
 * when A is outside G (T is convex and contains A), so clipping G's edges to T
 * is exact.
 */
#include <math.h>
#include <string.h>

#include "ttmpc_oracle.h"

static double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

/* clip segment P0->P1 against CCW triangle V[3]; returns 0 if empty */
static int clip_to_triangle(const double V[3][2], double p0x, double p0y, double p1x,
                            double p1y, double *t0, double *t1) {
  double lo = 0.0, hi = 1.0;
  double dx = p1x - p0x, dy = p1y - p0y;
  for (int i = 0; i < 3; i++) {
    int j = (i + 1) % 3;
    double ex = V[j][0] - V[i][0], ey = V[j][1] - V[i][1];
    /* inside: cross(e, P - Vi) >= 0 ;  f(t) = f0 + t*f1 */
    double f0 = cross2(ex, ey, p0x - V[i][0], p0y - V[i][1]);
    double f1 = cross2(ex, ey, dx, dy);
    if (f1 == 0.0) {
      if (f0 < 0.0) return 0;
    } else {
      double t = -f0 / f1;
      if (f1 > 0.0) { if (t > lo) lo = t; }
      else          { if (t < hi) hi = t; }
    }
    if (lo > hi) return 0;
  }
  *t0 = lo; *t1 = hi;
  return 1;
}

static double point_seg_dist_clipped(double ax, double ay, double p0x, double p0y,
                                     double p1x, double p1y, double t0, double t1) {
  double dx = p1x - p0x, dy = p1y - p0y;
  double dd = dx * dx + dy * dy;
  double t = t0;
  if (dd > 0.0) {
    t = ((ax - p0x) * dx + (ay - p0y) * dy) / dd;
    t = fmin(fmax(t, t0), t1);
  }
  double qx = p0x + t * dx - ax, qy = p0y + t * dy - ay;
  return sqrt(qx * qx + qy * qy);
}

/* first hit of ray A + s*(rx,ry), s in [0,L], with segment P0->P1; INFINITY if none */
static double ray_seg_hit(double ax, double ay, double rx, double ry, double L,
                          double p0x, double p0y, double p1x, double p1y) {
  double dx = p1x - p0x, dy = p1y - p0y;
  double wx = p0x - ax, wy = p0y - ay;
  double den = cross2(rx, ry, dx, dy);
  if (den != 0.0) {
    double s = cross2(wx, wy, dx, dy) / den;
    double t = cross2(wx, wy, rx, ry) / den;
    if (t >= 0.0 && t <= 1.0 && s >= 0.0 && s <= L) return s;
    return INFINITY;
  }
  if (cross2(wx, wy, rx, ry) != 0.0) return INFINITY; /* parallel, not collinear */
  double s0 = wx * rx + wy * ry, s1 = (p1x - ax) * rx + (p1y - ay) * ry;
  double lo = fmin(s0, s1), hi = fmax(s0, s1);
  if (hi < 0.0 || lo > L) return INFINITY;
  return fmax(lo, 0.0);
}

static int point_in_ring(double ax, double ay, const double *xy, int nv) {
  int in = 0;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    double xi = xy[2 * i], yi = xy[2 * i + 1], xj = xy[2 * j], yj = xy[2 * j + 1];
    if (((yi > ay) != (yj > ay)) && (ax < (xj - xi) * (ay - yi) / (yj - yi) + xi)) in = !in;
  }
  return in;
}

void ttdqn_oracle_observe(const ttdqn_scene_layout *lay, const double *agent,
                          const double *poly_xy, const int *poly_off, const int *is_solid,
                          int n_poly, double *seg_dist, double *ray_dist) {
  const int ns = lay->num_segments;
  const double L = lay->ray_length;
  const double ax = agent[0], ay = agent[1], th = agent[2];
  const double width = 2 * M_PI / ns;
  for (int i = 0; i < ns; i++) {
    double angle = th + i * width;
    double a1 = angle - width / 2, a2 = angle + width / 2;
    double V[3][2] = {{ax, ay},
                      {ax + L * cos(a1), ay + L * sin(a1)},
                      {ax + L * cos(a2), ay + L * sin(a2)}};
    double rx = cos(angle), ry = sin(angle);
    double cseg = INFINITY, cray = INFINITY;
    for (int g = 0; g < n_poly; g++) {
      const double *xy = poly_xy + 2 * poly_off[g];
      int nv = poly_off[g + 1] - poly_off[g];
      if (nv < 2) continue;
      double dseg = INFINITY, dray = INFINITY;
      if (is_solid[g] && point_in_ring(ax, ay, xy, nv)) {
        dseg = 0.0; dray = 0.0;
      } else {
        for (int e = 0; e < nv; e++) {
          int e2 = (e + 1) % nv;
          double p0x = xy[2 * e], p0y = xy[2 * e + 1], p1x = xy[2 * e2], p1y = xy[2 * e2 + 1];
          double t0, t1;
          if (clip_to_triangle(V, p0x, p0y, p1x, p1y, &t0, &t1)) {
            double d = point_seg_dist_clipped(ax, ay, p0x, p0y, p1x, p1y, t0, t1);
            if (d < dseg) dseg = d;
          }
          double s = ray_seg_hit(ax, ay, rx, ry, L, p0x, p0y, p1x, p1y);
          if (s < dray) dray = s;
        }
      }
      if (dseg < cseg) cseg = dseg;
      /* the reference only looks for a ray hit when the sector intersection is non-empty */
      if (dseg < INFINITY && dray < cray) cray = dray;
    }
    seg_dist[i] = cseg;
    ray_dist[i] = cray;
  }
}

/* normalize_distance on float32, same op order as numpy on a float32 array */
static float normalize_distance_f32(float d, float max_distance) {
  float t = -2.0f * d;
  t = t / max_distance;
  t = (float)exp((double)t); /* correctly rounded fp32 exp on both sides */
  t = 1.0f + t;
  t = 2.0f / t;
  return t - 1.0f;
}

static void linear_f32(const float *w, const float *b, const float *x, int n_in, int n_out,
                       float *y, int relu) {
  for (int o = 0; o < n_out; o++) {
    double a = (double)b[o]; /* fp64 accumulation, one rounding to fp32 per neuron */
    for (int i = 0; i < n_in; i++) a += (double)w[o * n_in + i] * (double)x[i];
    float acc = (float)a;
    y[o] = (relu && acc < 0.0f) ? 0.0f : acc;
  }
}

void ttdqn_oracle_observe_act(const ttdqn_scene_layout *lay, const ttdqn_qnet *qn, int n_envs,
                              const double *agent, const double *poly_xy, const int *poly_off,
                              const int *is_solid, const int *n_poly, const float *internal,
                              float *old_ext, float *ext_out, float *q_out, int *action,
                              double *seg_out, double *ray_out) {
  const int ns = lay->num_segments;
  const int n_ext = lay->use_memory ? 4 * ns : 2 * ns;
  for (int e = 0; e < n_envs; e++) {
    double seg[64], ray[64];
    float ext[256], x[512], h1[256], h2[256], q[64];
    ttdqn_oracle_observe(lay, agent + 3 * e, poly_xy + (size_t)e * lay->max_vert * 2,
                         poly_off + (size_t)e * (lay->max_poly + 1),
                         is_solid + (size_t)e * lay->max_poly, n_poly[e], seg, ray);
    for (int i = 0; i < ns; i++) {
      ext[i] = normalize_distance_f32((float)seg[i], (float)lay->max_distance);
      ext[ns + i] = normalize_distance_f32((float)ray[i], (float)lay->max_distance);
    }
    if (lay->use_memory) {
      float *old = old_ext + (size_t)e * 2 * ns;
      for (int i = 0; i < 2 * ns; i++) ext[2 * ns + i] = old[i];
      for (int i = 0; i < 2 * ns; i++) old[i] = ext[i];
    }
    for (int i = 0; i < n_ext; i++) x[i] = ext[i];
    for (int i = 0; i < lay->n_internal; i++) x[n_ext + i] = internal[(size_t)e * lay->n_internal + i];
    linear_f32(qn->w0, qn->b0, x, qn->n_in, qn->n_h1, h1, 1);
    linear_f32(qn->w1, qn->b1, h1, qn->n_h1, qn->n_h2, h2, 1);
    linear_f32(qn->w2, qn->b2, h2, qn->n_h2, qn->n_out, q, 0);
    int best = 0;
    for (int o = 1; o < qn->n_out; o++) if (q[o] > q[best]) best = o;
    if (ext_out) memcpy(ext_out + (size_t)e * n_ext, ext, n_ext * sizeof(float));
    if (q_out) memcpy(q_out + (size_t)e * qn->n_out, q, qn->n_out * sizeof(float));
    if (action) action[e] = best;
    if (seg_out) memcpy(seg_out + (size_t)e * ns, seg, ns * sizeof(double));
    if (ray_out) memcpy(ray_out + (size_t)e * ns, ray, ns * sizeof(double));
  }
}

/* ------------------------------------------------------------------------------------
 * Internal observation of the ray model and the DQN hint trajectory (rl_ref).
 *
 *   SpeedObservation / AngularVelocityObservation        components/int_obsv_speed.py, int_obsv_angular_velocity.py
 *   ReferencePathSampleObservation(1, 0, offset)         components/int_obsv_reference_path_sample.py:27-39
 *   ReferencePathCornerObservation(corner_samples)       components/int_obsv_reference_path_corner.py:21-45
 *   env.path_progress = path.project(agent.point)        environment.py:115
 *   MobileRobot.step / step_with_ref_speed               environment/agent.py:86-145
 *   rl_ref loop                                           src/main.py:184-193
 *
 * PINNED (tests/golden/dqn_loop.npz, tools/gen_golden_dqn_loop.py): the reference's own
 * MobileRobot and observation components run on a stand-in `shapely` module, so everything
 * except LineString.project / interpolate is the reference's code.  project / interpolate are
 * GEOS' length-indexed-line operations restated here (parity unpinned for those two: first
 * closest segment wins, clamped segment fraction, linear interpolation inside a segment).
 * ------------------------------------------------------------------------------------ */
double ttdqn_oracle_project(const double *xy, int n, double px, double py) {
  double best = INFINITY, best_s = 0.0, cum = 0.0;
  for (int i = 0; i + 1 < n; i++) {
    const double ax = xy[2 * i], ay = xy[2 * i + 1], dx = xy[2 * i + 2] - ax, dy = xy[2 * i + 3] - ay;
    const double len2 = dx * dx + dy * dy, len = sqrt(len2);
    double t = 0.0;
    if (len2 > 0.0) {
      t = ((px - ax) * dx + (py - ay) * dy) / len2;
      t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    }
    const double qx = ax + t * dx - px, qy = ay + t * dy - py;
    const double d = sqrt(qx * qx + qy * qy);
    if (d < best) { best = d; best_s = cum + t * len; }
    cum += len;
  }
  return best_s;
}
void ttdqn_oracle_interpolate(const double *xy, int n, double s, double *x, double *y) {
  if (s <= 0.0 || n < 2) { *x = xy[0]; *y = xy[1]; return; }
  double cum = 0.0;
  for (int i = 0; i + 1 < n; i++) {
    const double ax = xy[2 * i], ay = xy[2 * i + 1], dx = xy[2 * i + 2] - ax, dy = xy[2 * i + 3] - ay;
    const double len = sqrt(dx * dx + dy * dy);
    if (s <= cum + len && len > 0.0) {
      const double t = (s - cum) / len;
      *x = ax + t * dx; *y = ay + t * dy;
      return;
    }
    cum += len;
  }
  *x = xy[2 * (n - 1)]; *y = xy[2 * (n - 1) + 1];
}
static void rel_obs(double px, double py, const double *agent, double max_distance, float *o) {
  const double dx = px - agent[0], dy = py - agent[1];
  const double rel = atan2(dy, dx) - agent[2];
  o[0] = (float)cos(rel);
  o[1] = (float)sin(rel);
  o[2] = (float)(2.0 / (1.0 + exp(-2.0 * sqrt(dx * dx + dy * dy) / max_distance)) - 1.0);
}
/* agent5 = x y theta v w; obs [2 + 3 + 3*corner_samples] fp32 */
void ttdqn_oracle_internal_obs(int corner_samples, double offset, double max_distance, const double *agent5,
                               const double *path_xy, int n_nodes, float *obs, double *progress) {
  obs[0] = (float)(2.0 * (agent5[3] - (-0.5)) / (1.5 - (-0.5)) - 1.0);
  /* the reference normalises the angular VELOCITY with the angular ACCELERATION bounds */
  obs[1] = (float)(2.0 * (agent5[4] - (-3.0)) / (3.0 - (-3.0)) - 1.0);
  const double s = ttdqn_oracle_project(path_xy, n_nodes, agent5[0], agent5[1]);
  if (progress) *progress = s;
  double px, py;
  ttdqn_oracle_interpolate(path_xy, n_nodes, s + 0 * 0.0 + offset, &px, &py);
  rel_obs(px, py, agent5, max_distance, obs + 2);
  double length = 0.0;
  int i = 0;
  while (length < s) {
    const double dx = path_xy[2 * (i + 1)] - path_xy[2 * i], dy = path_xy[2 * (i + 1) + 1] - path_xy[2 * i + 1];
    length += sqrt(dx * dx + dy * dy);
    i++;
  }
  for (int j = 0; j < corner_samples; j++) {
    if (i > n_nodes - 1) i = n_nodes - 1;
    rel_obs(path_xy[2 * i], path_xy[2 * i + 1], agent5, max_distance, obs + 5 + 3 * j);
    i++;
  }
}
/* rl_ref [steps][2]; use_libm = 0: the kernels' tt_sincos (bit-exact with the GPU) */
void ttdqn_oracle_rl_ref(int steps, double ts, double ref_speed, const double *agent5, int action,
                         double *rl_ref, int use_libm) {
  double x = agent5[0], y = agent5[1], th = agent5[2], v = agent5[3], w = agent5[4], s, c;
  for (int j = 0; j < steps; j++) {
    double sp;
    if (j == 0) {  /* MobileRobot.step(action_index, ts) */
      if (action / 3 == 0) v += ts * 1.0;
      if (action / 3 == 2) v += ts * -1.0;
      if (action % 3 == 0) w += ts * 3.0;
      if (action % 3 == 2) w += ts * -3.0;
      if (v > 1.5) v = 1.5;
      if (v < -0.5) v = -0.5;
      if (w > 0.5) w = 0.5;
      if (w < -0.5) w = -0.5;
      th += ts * w;
      sp = v;
    } else {       /* step_with_ref_speed(ts, ref_speed) */
      w *= 0.95;
      th += ts * w;
      sp = ref_speed <= 0.0 ? 1.5 : ref_speed;
    }
    if (use_libm) { s = sin(th); c = cos(th); } else ttmpc_oracle_sincos(th, &s, &c);
    x += (ts * sp) * c; y += (ts * sp) * s;
    rl_ref[2 * j] = x; rl_ref[2 * j + 1] = y;
  }
}

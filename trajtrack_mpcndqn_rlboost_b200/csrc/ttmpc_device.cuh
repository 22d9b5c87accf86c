// ttmpc_device.cuh -- device side of the batched NMPC planner (sm_100a).
//
// One WARP solves one scene; lane k owns horizon step k: its controls
// (v_k, w_k), every PANOC vector's two entries, the state s_{k+1}, and the
// k-th column of the obstacle tables.  The sequential unicycle rollout and its
// adjoint become warp prefix / suffix scans; every inner product is a butterfly
// all-reduce, so all solver control flow is warp-uniform.
//
// Reference for the maths:
//   cost / constraints : /root/reference/src/mpc_traj_tracker/mpc/mpc_generator.py:155-283
//   dynamics           : /root/reference/src/pkg_motion_model/motion_model.py:153-176
//   solver             : OpEn (optimization_engine 0.7.x) PANOC + ALM/PM, as
//                        called at src/mpc_traj_tracker/trajectory_generator.py:284
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ttmpc.h"

namespace ttmpc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_MEM = 16;
constexpr int MAX_EDGE = 8;
constexpr int DYN_FIELDS = 10;  // per (obstacle, step): see build_dyn_table

struct DevCfg {
  int N, Nother, Nstc, ne, nstcobs, Ndyn, mem, max_inner, max_outer;
  int off_s, off_q, off_r, off_vref, off_c, off_os, off_od, off_qdyn, np;
  int smem_per_warp;  // bytes
  int warps_per_block;
  double ts, veh_d2, margin;
  double vmin, vmax, wmax, amin, amax, awmax;
  double tol, init_tol, delta_tol, c0, pen_factor, tol_factor, suff_dec;
};

// ---------------------------------------------------------------- warp utils
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ void wsum3(double &a, double &b, double &c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ta = __shfl_xor_sync(FULL, a, o);
    double tb = __shfl_xor_sync(FULL, b, o);
    double tc = __shfl_xor_sync(FULL, c, o);
    a += ta; b += tb; c += tc;
  }
}
// inclusive prefix sum over lanes (Kogge-Stone)
__device__ __forceinline__ double wscan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// inclusive suffix sum over lanes
__device__ __forceinline__ double wsuffix(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}
__device__ __forceinline__ double clipd(double z, double lo, double hi) {
  return fmin(fmax(z, lo), hi);
}

// ---------------------------------------------------------------- per-warp state
// Everything a warp needs that is uniform across lanes lives in shared memory
// (broadcast reads), so the eval routine can be a real (non-inlined) function.
struct WarpCtx {
  const double *p;   // this scene's packed parameter vector (global)
  double *dyn;       // this warp's dynamic-obstacle table (global scratch, L2 resident)
  double x0, y0, th0, xg, yg, thg, v_init, w_init;
  double qvel, rv, rw, qN, qthetaN, qrpd, acc_pen, wacc_pen;
  long long n_cost, n_grad, n_body;  // evaluation counters (lane 0 view)
};

struct WarpSmem {
  WarpCtx *ctx;
  double *seg;   // [5][N]: s1x s1y sx sy inv_den
  double *os;    // [Nstc*nstcobs] raw static half-spaces
  double *D;     // [Ndyn] per-obstacle hard sums of the last evaluation
  double2 *lbs;  // [(mem+1)][N]
  double2 *lby;  // [(mem+1)][N]
  double *rho;   // [mem+1]
  double *alpha; // [mem]
};

__host__ __device__ inline int smem_bytes_per_warp(int N, int Nstc, int nstcobs, int Ndyn, int mem) {
  size_t b = 0;
  b += (sizeof(WarpCtx) + 15) / 16 * 16;
  b += sizeof(double) * 5 * N;
  b += sizeof(double) * Nstc * nstcobs;
  b += sizeof(double) * ((Ndyn + 1) / 2 * 2);
  b += sizeof(double2) * (size_t)(mem + 1) * N * 2;
  b += sizeof(double) * (mem + 2) / 2 * 2;
  b += sizeof(double) * (mem + 1) / 2 * 2;
  return (int)((b + 15) / 16 * 16);
}

__device__ __forceinline__ WarpSmem carve(unsigned char *base, const DevCfg &g) {
  WarpSmem w;
  unsigned char *q = base;
  w.ctx = reinterpret_cast<WarpCtx *>(q); q += (sizeof(WarpCtx) + 15) / 16 * 16;
  w.lbs = reinterpret_cast<double2 *>(q); q += sizeof(double2) * (size_t)(g.mem + 1) * g.N;
  w.lby = reinterpret_cast<double2 *>(q); q += sizeof(double2) * (size_t)(g.mem + 1) * g.N;
  w.seg = reinterpret_cast<double *>(q); q += sizeof(double) * 5 * g.N;
  w.os = reinterpret_cast<double *>(q); q += sizeof(double) * g.Nstc * g.nstcobs;
  w.D = reinterpret_cast<double *>(q); q += sizeof(double) * ((g.Ndyn + 1) / 2 * 2);
  w.rho = reinterpret_cast<double *>(q); q += sizeof(double) * ((g.mem + 2) / 2 * 2);
  w.alpha = reinterpret_cast<double *>(q);
  return w;
}

// Dynamic-obstacle table, field-major so that lane k reads element [f][j][k]
// coalesced.  Built once per solve from the raw (cx cy rx ry angle alpha) records
// (mpc_generator.py:225-237): the sincos and the four divisions leave the hot loop.
//   0 cx  1 cy  2 R2 (rejection radius^2)  3 cos  4 sin
//   5 1/(rx+1e-6)^2  6 1/(ry+1e-6)^2  7 1/(rx+m+1e-6)^2  8 1/(ry+m+1e-6)^2  9 alpha*qdyn[k]
__device__ __forceinline__ size_t dyn_table_doubles(const DevCfg &g) {
  return (size_t)DYN_FIELDS * g.Ndyn * g.N;
}
__device__ __forceinline__ double &dynf(double *t, const DevCfg &g, int f, int j, int k) {
  return t[((size_t)f * g.Ndyn + j) * g.N + k];
}

__device__ inline void stage_scene(const DevCfg &g, const WarpSmem &sm, const double *p,
                                   double *dyn_scratch, int lane) {
  WarpCtx *c = sm.ctx;
  if (lane == 0) {
    const double *s = p + g.off_s, *q = p + g.off_q;
    c->p = p; c->dyn = dyn_scratch;
    c->x0 = s[0]; c->y0 = s[1]; c->th0 = s[2];
    c->xg = s[3]; c->yg = s[4]; c->thg = s[5];
    c->v_init = s[6]; c->w_init = s[7];
    c->qvel = q[1]; c->rv = q[3]; c->rw = q[4]; c->qN = q[5]; c->qthetaN = q[6];
    c->qrpd = q[7]; c->acc_pen = q[8]; c->wacc_pen = q[9];
    c->n_cost = 0; c->n_grad = 0; c->n_body = 0;
  }
  // reference-path segments: path_ref has N+1 points, last duplicated (l.190-191)
  const double *r = p + g.off_r;
  for (int j = lane; j < g.N; j += 32) {
    int j2 = (j + 1 < g.N) ? j + 1 : g.N - 1;
    double s1x = r[3 * j], s1y = r[3 * j + 1];
    double sx = r[3 * j2] - s1x, sy = r[3 * j2 + 1] - s1y;
    double den = sx * sx + sy * sy + 1e-16;
    sm.seg[0 * g.N + j] = s1x; sm.seg[1 * g.N + j] = s1y;
    sm.seg[2 * g.N + j] = sx;  sm.seg[3 * g.N + j] = sy;
    sm.seg[4 * g.N + j] = 1.0 / den;
  }
  const double *os = p + g.off_os;
  for (int i = lane; i < g.Nstc * g.nstcobs; i += 32) sm.os[i] = os[i];
  // dynamic obstacle table
  const double *od = p + g.off_od, *qdyn = p + g.off_qdyn;
  const int npair = g.Ndyn * g.N;
  for (int t = lane; t < npair; t += 32) {
    int j = t / g.N, k = t - j * g.N;
    const double *e = od + (size_t)t * 6;  // obstacle-major, then step: contiguous records
    double cx = e[0], cy = e[1], rx = e[2], ry = e[3], ang = e[4], alpha = e[5];
    double sa, ca;
    sincos(ang, &sa, &ca);
    double Rx = rx + 1e-6, Ry = ry + 1e-6;
    double Rxm = rx + g.margin + 1e-6, Rym = ry + g.margin + 1e-6;
    double rmax = fmax(fmax(fabs(Rx), fabs(Ry)), fmax(fabs(Rxm), fabs(Rym)));
    dynf(dyn_scratch, g, 0, j, k) = cx;
    dynf(dyn_scratch, g, 1, j, k) = cy;
    dynf(dyn_scratch, g, 2, j, k) = rmax * rmax * (1.0 + 1e-9);
    dynf(dyn_scratch, g, 3, j, k) = ca;
    dynf(dyn_scratch, g, 4, j, k) = sa;
    dynf(dyn_scratch, g, 5, j, k) = 1.0 / (Rx * Rx);
    dynf(dyn_scratch, g, 6, j, k) = 1.0 / (Ry * Ry);
    dynf(dyn_scratch, g, 7, j, k) = 1.0 / (Rxm * Rxm);
    dynf(dyn_scratch, g, 8, j, k) = 1.0 / (Rym * Rym);
    dynf(dyn_scratch, g, 9, j, k) = alpha * qdyn[k];
  }
  __syncwarp();
}

// ---------------------------------------------------------------- evaluation
struct EvalOut {
  double psi;   // f + c/2 dist^2_C(F1 + y/max(c,1)) + c/2 |F2|^2
  double f;     // original cost (psi at c = 0)
  double f2sq;  // |F2|^2
  double S;     // static hard sum (F2_j = S + D_j, D in smem)
  double gv, gw;  // this lane's gradient entries (GRAD only)
};

// Evaluate psi (and its gradient when GRAD) at this lane's (v, w).
// ya / yw are this lane's multipliers for the linear / angular acceleration rows.
template <bool GRAD>
__device__ __noinline__ EvalOut eval_psi(const DevCfg *gp, unsigned char *smem_base, double v,
                                         double w, double c, double ya, double yw,
                                         double *st_out) {
  const DevCfg &g = *gp;
  const WarpSmem sm = carve(smem_base, g);
  const int lane = threadIdx.x & 31;
  double gv = 0.0, gw = 0.0;
  const WarpCtx *cx = sm.ctx;
  const int N = g.N;
  const bool act = lane < N;
  const double ts = g.ts;

  // ---- rollout (motion_model.py:153-176): theta and position as prefix sums
  const double tw = ts * w;
  const double dth = (1.0 / 6.0) * (tw + 2 * tw + 2 * tw + tw);
  const double th_in = wscan(dth, lane);
  double th_ex = __shfl_up_sync(FULL, th_in, 1);
  if (lane == 0) th_ex = 0.0;
  const double tha = cx->th0 + th_ex;
  const double thb = tha + 0.5 * tw, thc = tha + tw;
  double sa, ca, sb, cb, sc, cc;
  sincos(tha, &sa, &ca);
  sincos(thb, &sb, &cb);
  sincos(thc, &sc, &cc);
  const double k1x = ts * (v * ca), k2x = ts * (v * cb), k4x = ts * (v * cc);
  const double k1y = ts * (v * sa), k2y = ts * (v * sb), k4y = ts * (v * sc);
  const double dx = (1.0 / 6.0) * (k1x + 2 * k2x + 2 * k2x + k4x);
  const double dy = (1.0 / 6.0) * (k1y + 2 * k2y + 2 * k2y + k4y);
  const double X = cx->x0 + wscan(dx, lane);
  const double Y = cx->y0 + wscan(dy, lane);
  const double TH = cx->th0 + th_in;
  if (st_out && act) { st_out[3 * lane] = X; st_out[3 * lane + 1] = Y; st_out[3 * lane + 2] = TH; }

  double cost = 0.0;         // this lane's share of f
  double gx = 0.0, gy = 0.0; // d psi / d position_{k+1}
  double S_loc = 0.0, gSx = 0.0, gSy = 0.0;

  // ---- reference-path deviation (l.124-139, 202)
  {
    double dmin = 0.0; int jmin = lane;
    for (int j = 0; j < N; j++) {
      const double s1x = sm.seg[j], s1y = sm.seg[N + j], sx = sm.seg[2 * N + j],
                   sy = sm.seg[3 * N + j], inv = sm.seg[4 * N + j];
      if (j >= lane) {
        double t_hat = ((X - s1x) * sx + (Y - s1y) * sy) * inv;
        double t = fmin(fmax(t_hat, 0.0), 1.0);
        double qx = s1x + t * sx - X, qy = s1y + t * sy - Y;
        double d2 = qx * qx + qy * qy;
        if (j == lane || !(dmin <= d2)) { dmin = d2; jmin = j; }
      }
    }
    if (act) {
      cost += dmin * cx->qrpd;
      if (GRAD) {
        const int j = jmin;
        const double s1x = sm.seg[j], s1y = sm.seg[N + j], sx = sm.seg[2 * N + j],
                     sy = sm.seg[3 * N + j], inv = sm.seg[4 * N + j];
        double t_hat = ((X - s1x) * sx + (Y - s1y) * sy) * inv;
        double t = fmin(fmax(t_hat, 0.0), 1.0);
        double qx = s1x + t * sx - X, qy = s1y + t * sy - Y;
        double pass = (t_hat >= 0.0 && t_hat <= 1.0) ? 1.0 : 0.0;
        double cs = (qx * sx + qy * sy) * pass * inv;
        gx += cx->qrpd * (2 * (cs * sx - qx));
        gy += cx->qrpd * (2 * (cs * sy - qy));
      }
    }
  }
  // ---- speed reference + control action (l.203-204)
  if (act) {
    const double vr = cx->p[g.off_vref + lane];
    cost += cx->qvel * ((v - vr) * (v - vr));
    cost += cx->rv * (v * v) + cx->rw * (w * w);
  }
  // ---- fleet collision (l.207-211)
  if (act) {
    const double *cp = cx->p + g.off_c + 3 * lane;
    double acc = 0.0, fx = 0.0, fy = 0.0;
    for (int j = 0; j < g.Nother; j++) {
      const double ox = cp[(size_t)j * 3 * N], oy = cp[(size_t)j * 3 * N + 1];
      const double ex = X - ox, ey = Y - oy;
      const double e = g.veh_d2 - (ex * ex + ey * ey);
      if (e > 0.0) {
        acc += e;
        if (GRAD) { fx += -2 * ex; fy += -2 * ey; }
      }
    }
    cost += 1000.0 * acc;
    if (GRAD) { gx += 1000.0 * fx; gy += 1000.0 * fy; }
  }
  // ---- static obstacles (l.214-220): hard penalty only
  if (act) {
    const int ne = g.ne;
    for (int i = 0; i < g.Nstc; i++) {
      const double *b = sm.os + i * g.nstcobs, *a0 = b + ne, *a1 = b + 2 * ne;
      double m[MAX_EDGE];
      double inside = 1.0;
#pragma unroll
      for (int e = 0; e < MAX_EDGE; e++) {
        if (e < ne) {
          double res = a0[e] * (-X) + a1[e] * (-Y) + b[e];
          m[e] = fmax(0.0, res);
          inside *= m[e] * m[e];
        }
      }
      if (inside > 0.0) {
        S_loc += inside;
        if (GRAD) {
#pragma unroll
          for (int e = 0; e < MAX_EDGE; e++) {
            if (e < ne) {
              double rest = 1.0;
#pragma unroll
              for (int e2 = 0; e2 < MAX_EDGE; e2++)
                if (e2 < ne && e2 != e) rest *= m[e2] * m[e2];
              gSx += rest * 2 * m[e] * (-a0[e]);
              gSy += rest * 2 * m[e] * (-a1[e]);
            }
          }
        }
      }
    }
  }
  // ---- dynamic obstacles (l.225-237): hard penalty D_j and soft cost
  unsigned hard_mask_lo = 0, hard_mask_hi = 0;  // obstacles with a positive hard term
  {
    const double *T = cx->dyn;
    double soft = 0.0;
    int bodies = 0;
    for (int j = 0; j < g.Ndyn; j++) {
      bool pass = false;
      double ex = 0.0, ey = 0.0;
      if (act) {
        ex = X - T[((size_t)0 * g.Ndyn + j) * N + lane];
        ey = Y - T[((size_t)1 * g.Ndyn + j) * N + lane];
        pass = (ex * ex + ey * ey) < T[((size_t)2 * g.Ndyn + j) * N + lane];
      }
      double in1 = 0.0;
      if (pass) {
        bodies++;
        const double ca_ = T[((size_t)3 * g.Ndyn + j) * N + lane];
        const double sa_ = T[((size_t)4 * g.Ndyn + j) * N + lane];
        const double A = ex * ca_ + ey * sa_, B = ex * sa_ - ey * ca_;
        const double A2 = A * A, B2 = B * B;
        in1 = 1 - A2 * T[((size_t)5 * g.Ndyn + j) * N + lane] -
              B2 * T[((size_t)6 * g.Ndyn + j) * N + lane];
        const double iRxm = T[((size_t)7 * g.Ndyn + j) * N + lane];
        const double iRym = T[((size_t)8 * g.Ndyn + j) * N + lane];
        const double in2 = 1 - A2 * iRxm - B2 * iRym;
        if (in2 > 0.0) {
          const double ws = T[((size_t)9 * g.Ndyn + j) * N + lane];
          soft += (in2 * in2) * ws;
          if (GRAD) {
            const double wg = ws * 2 * in2;
            gx += wg * (-2 * A * ca_ * iRxm - 2 * B * sa_ * iRym);
            gy += wg * (-2 * A * sa_ * iRxm + 2 * B * ca_ * iRym);
          }
        }
      }
      const bool hard = in1 > 0.0;
      if (__any_sync(FULL, hard)) {
        const double Dj = wsum(hard ? in1 : 0.0);
        if (lane == 0) sm.D[j] = Dj;
        if (j < 32) hard_mask_lo |= 1u << j; else hard_mask_hi |= 1u << (j - 32);
      } else if (lane == 0) {
        sm.D[j] = 0.0;
      }
    }
    cost += soft;
    bodies = __reduce_add_sync(FULL, bodies);
    if (lane == 0) sm.ctx->n_body += bodies;
  }
  // ---- terminal cost (l.242)
  double gt = 0.0;
  if (lane == N - 1) {
    cost += cx->qN * ((X - cx->xg) * (X - cx->xg) + (Y - cx->yg) * (Y - cx->yg)) +
            cx->qthetaN * ((TH - cx->thg) * (TH - cx->thg));
    if (GRAD) {
      gx += 2 * cx->qN * (X - cx->xg);
      gy += 2 * cx->qN * (Y - cx->yg);
      gt = 2 * cx->qthetaN * (TH - cx->thg);
    }
  }
  // ---- accelerations: cost (l.250-264) and the ALM set C = acc bounds
  double vp = __shfl_up_sync(FULL, v, 1), wp = __shfl_up_sync(FULL, w, 1);
  if (lane == 0) { vp = cx->v_init; wp = cx->w_init; }
  double aa = 0.0, aw = 0.0, ea = 0.0, ew = 0.0, alm = 0.0;
  if (act) {
    aa = (v - vp) / ts; aw = (w - wp) / ts;
    cost += (aa * aa) * cx->acc_pen + (aw * aw) * cx->wacc_pen;
    const double cm = fmax(c, 1.0);
    double z = aa + ya / cm;
    ea = z - clipd(z, g.amin, g.amax);
    z = aw + yw / cm;
    ew = z - clipd(z, -g.awmax, g.awmax);
    alm = ea * ea + ew * ew;
  }
  // ---- reductions: f, ALM distance, static sum
  double f = cost, d2 = alm, S = S_loc;
  wsum3(f, d2, S);
  __syncwarp();
  double f2sq = 0.0, sumF2 = 0.0;
  for (int j = 0; j < g.Ndyn; j++) {
    const double F2j = S + sm.D[j];
    f2sq += F2j * F2j;
    sumF2 += F2j;
  }
  EvalOut out;
  out.f = f; out.f2sq = f2sq; out.S = S;
  out.psi = f + c * d2 / 2 + c * f2sq / 2;

  if (GRAD) {
    // hard-penalty gradient: c * sum_j F2_j * (grad S + grad D_j)
    if (c != 0.0) {
      gx += c * sumF2 * gSx;
      gy += c * sumF2 * gSy;
      unsigned mlo = hard_mask_lo, mhi = hard_mask_hi;
      const double *T = cx->dyn;
      while (mlo | mhi) {
        int j;
        if (mlo) { j = __ffs(mlo) - 1; mlo &= mlo - 1; }
        else     { j = 32 + __ffs(mhi) - 1; mhi &= mhi - 1; }
        if (act) {
          const double ex = X - T[((size_t)0 * g.Ndyn + j) * N + lane];
          const double ey = Y - T[((size_t)1 * g.Ndyn + j) * N + lane];
          const double ca_ = T[((size_t)3 * g.Ndyn + j) * N + lane];
          const double sa_ = T[((size_t)4 * g.Ndyn + j) * N + lane];
          const double iRx = T[((size_t)5 * g.Ndyn + j) * N + lane];
          const double iRy = T[((size_t)6 * g.Ndyn + j) * N + lane];
          const double A = ex * ca_ + ey * sa_, B = ex * sa_ - ey * ca_;
          const double in1 = 1 - (A * A) * iRx - (B * B) * iRy;
          if (in1 > 0.0 && (ex * ex + ey * ey) < T[((size_t)2 * g.Ndyn + j) * N + lane]) {
            const double wg = c * (S + sm.D[j]);
            gx += wg * (-2 * A * ca_ * iRx - 2 * B * sa_ * iRy);
            gy += wg * (-2 * A * sa_ * iRx + 2 * B * ca_ * iRy);
          }
        }
      }
    }
    // adjoint of the rollout: suffix scans
    const double h6 = ts / 6.0;
    const double Cs = ca + 4 * cb + cc, Ss = sa + 4 * sb + sc;
    const double dxdv = h6 * Cs, dydv = h6 * Ss;
    const double dxdth = -h6 * v * Ss, dydth = h6 * v * Cs;
    const double dxdw = -h6 * v * ts * (2 * sb + sc), dydw = h6 * v * ts * (2 * cb + cc);
    const double lx = wsuffix(gx, lane), ly = wsuffix(gy, lane);
    const double m = lx * dxdth + ly * dydth;
    const double lt = wsuffix(gt + m, lane) - m;
    // direct control terms
    const double aa_n = __shfl_down_sync(FULL, aa, 1), aw_n = __shfl_down_sync(FULL, aw, 1);
    const double ea_n = __shfl_down_sync(FULL, ea, 1), ew_n = __shfl_down_sync(FULL, ew, 1);
    if (act) {
      const bool last = lane == N - 1;
      const double vr = cx->p[g.off_vref + lane];
      double dv = 2 * cx->qvel * (v - vr) + 2 * cx->rv * v;
      double dw = 2 * cx->rw * w;
      dv += 2 * cx->acc_pen * (aa - (last ? 0.0 : aa_n)) / ts + c * (ea - (last ? 0.0 : ea_n)) / ts;
      dw += 2 * cx->wacc_pen * (aw - (last ? 0.0 : aw_n)) / ts + c * (ew - (last ? 0.0 : ew_n)) / ts;
      gv = dv + lx * dxdv + ly * dydv;
      gw = dw + lx * dxdw + ly * dydw + lt * ts;
    }
  }
  out.gv = gv; out.gw = gw;
  return out;
}

}  // namespace ttmpc

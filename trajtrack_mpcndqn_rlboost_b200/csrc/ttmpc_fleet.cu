// ttmpc_fleet.cu -- the caller side of the solve, batched on the device: one control step of
// n independent robots as the reference's InterfaceMpc.get_action / TrajectoryGenerator.run_step
// perform it (src/interface_mpc.py:72-92, src/mpc_traj_tracker/trajectory_generator.py:203-294).
// See include/ttmpc.h ("Fleet step") for the contract; oracle/ttfleet_oracle.c is the CPU
// restatement these kernels match bit for bit (use_libm = 0).
//
// Both kernels are copy / scalar work: pack writes the n x np parameter block once, coalesced
// (HBM-bound, 21 KB per robot), advance touches 5 doubles per robot.
#include <cuda_runtime.h>

#include "ttmpc_device.cuh"
#include "ttmpc_launch.cuh"

namespace ttmpc {

__device__ __forceinline__ double hyp2(double dx, double dy) { return sqrt(fma(dx, dx, dy * dy)); }

// shapely Polygon.contains(Point) (strictly inside, even-odd rule) / Polygon.distance(Point)
__device__ __forceinline__ bool poly_contains(const double *xy, int nv, double px, double py) {
  bool in = false;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    const double xi = xy[2 * i], yi = xy[2 * i + 1], xj = xy[2 * j], yj = xy[2 * j + 1];
    if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) in = !in;
  }
  return in;
}
__device__ __forceinline__ double poly_distance(const double *xy, int nv, double px, double py) {
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
constexpr int SW_MAX_OBS = 64;  // obstacles per robot the parallel switch path handles (more: serial path)
// obstacle o of robot e in the list "static polygons + circle_to_rect(moving obstacle)"; rect: 8 doubles scratch
__device__ __forceinline__ const double *sw_obstacle(const ttmpc_fleet &f, int e, int o, int n_stc, double *rect, int *nv) {
  if (o < n_stc) {
    const double *pxy = f.sw_poly_xy + (f.sw_poly_shared ? 0 : (size_t)e * f.sw_max_poly * f.sw_max_pv * 2);
    const int *pnv = f.sw_poly_nv + (f.sw_poly_shared ? 0 : (size_t)e * f.sw_max_poly);
    *nv = pnv[o];
    return pxy + (size_t)o * f.sw_max_pv * 2;
  }
  const double *c = f.dyn_cur + ((size_t)e * f.n_dyn_live + (o - n_stc)) * 2, rr = f.sw_dyn_radius;  // main.py:91-95
  rect[0] = c[0] - rr; rect[1] = c[1] - rr; rect[2] = c[0] + rr; rect[3] = c[1] - rr;
  rect[4] = c[0] + rr; rect[5] = c[1] + rr; rect[6] = c[0] - rr; rect[7] = c[1] + rr;
  *nv = 4;
  return rect;
}
// the state machine of HintSwitcher.switch on precomputed predicates: dist[o] (distance of the
// robot to obstacle o, < 0: unused slot) and in[k * n_obs + o] (reference point k inside obstacle o)
__device__ int hint_switch_sm(const ttmpc_fleet &f, int e, int N, int n_obs, const double *dist,
                              const unsigned char *in) {
  int *st = f.sw_state + 2 * e;
  bool cnt_flag = false;
  for (int k = 0; k < N; k++)
    for (int o = 0; o < n_obs; o++) {
      if (dist[o] < 0.0) continue;
      if (in[k * n_obs + o]) {
        if (dist[o] < f.sw_switch_distance && !st[0]) { st[0] = 1; return 1; }
      } else if (dist[o] > f.sw_detach_distance && st[0]) {
        if (st[1] > f.sw_detach_steps) { st[0] = 0; st[1] = 0; }
        else if (!cnt_flag) { st[1] += 1; cnt_flag = true; }
      }
    }
  return st[0];
}
// HintSwitcher.switch (main_pre.py:35-52) of robot e, on the ORIGINAL local reference (serial form)
__device__ int hint_switch(const ttmpc_fleet &f, int e, int N, const double *ref, int L, int idx) {
  int *st = f.sw_state + 2 * e;  // switch_on, detach_cnt
  const double px = f.state[3 * e], py = f.state[3 * e + 1];
  const int n_stc = f.sw_poly_xy ? f.sw_max_poly : 0, n_dyn = f.dyn_cur ? f.n_dyn_live : 0;
  const double *pxy = f.sw_poly_xy ? f.sw_poly_xy + (f.sw_poly_shared ? 0 : (size_t)e * f.sw_max_poly * f.sw_max_pv * 2) : nullptr;
  const int *pnv = f.sw_poly_nv ? f.sw_poly_nv + (f.sw_poly_shared ? 0 : (size_t)e * f.sw_max_poly) : nullptr;
  bool cnt_flag = false;
  for (int k = 0; k < N; k++) {
    int r = idx + k; if (r > L - 1) r = L - 1;
    const double ox = ref[3 * r], oy = ref[3 * r + 1];
    for (int o = 0; o < n_stc + n_dyn; o++) {
      double rect[8];
      const double *xy; int nv;
      if (o < n_stc) { nv = pnv[o]; xy = pxy + (size_t)o * f.sw_max_pv * 2; if (nv < 3) continue; }
      else {  // circle_to_rect (main.py:91-95)
        const double *c = f.dyn_cur + ((size_t)e * f.n_dyn_live + (o - n_stc)) * 2, rr = f.sw_dyn_radius;
        rect[0] = c[0] - rr; rect[1] = c[1] - rr; rect[2] = c[0] + rr; rect[3] = c[1] - rr;
        rect[4] = c[0] + rr; rect[5] = c[1] + rr; rect[6] = c[0] - rr; rect[7] = c[1] + rr;
        xy = rect; nv = 4;
      }
      const double dist = poly_distance(xy, nv, px, py);
      if (poly_contains(xy, nv, ox, oy)) {
        if (dist < f.sw_switch_distance && !st[0]) { st[0] = 1; return 1; }
      } else if (dist > f.sw_detach_distance && st[0]) {
        if (st[1] > f.sw_detach_steps) { st[0] = 0; st[1] = 0; }
        else if (!cnt_flag) { st[1] += 1; cnt_flag = true; }
      }
    }
  }
  return st[0];
}

// One CTA per robot.  Thread 0 does the scalar decisions (closest reference point, goal test,
// speed reference); all threads then write the packed vector, element o of the row by thread
// o mod blockDim (coalesced 8-byte stores).
__global__ void __launch_bounds__(128) fleet_pack_kernel(const ttmpc_fleet f, const FleetDims d,
                                                         double *__restrict__ p_all) {
  const int e = blockIdx.x;
  if (e >= f.n) return;
  __shared__ int s_idx, s_hint, s_sw;
  __shared__ double s_speed;
  __shared__ double s_dist[SW_MAX_OBS];
  __shared__ unsigned char s_in[32 * SW_MAX_OBS];
  const double *st = f.state + 3 * e, *goal = f.goal + 3 * e, *lu = f.last_u + 2 * e;
  const double *ref = f.ref_traj + (size_t)e * f.ref_stride * 3;
  const int L = f.ref_len[e], N = d.N;
  if (threadIdx.x == 0) {
    int idx = f.idx_ref[e];
    const double x = st[0], y = st[1];
    const bool was_running = f.status[e] == TTMPC_FLEET_RUNNING;
    if (f.status[e] == TTMPC_FLEET_RUNNING) {
      // get_local_ref_traj (trajectory_generator.py:214-219): first minimum in the window
      int lo = idx - 1 * f.action_steps; if (lo < 0) lo = 0;
      int hi = idx + 5 * f.action_steps; if (hi > L) hi = L;
      double best = INFINITY; int arg = lo;
      for (int i = lo; i < hi; i++) {
        const double dist = hyp2(x - ref[3 * i], y - ref[3 * i + 1]);
        if (dist < best) { best = dist; arg = i; }
      }
      idx = arg;
      f.idx_ref[e] = idx;
      // check_termination_condition (:158-164): np.allclose(atol=0.05, rtol=0) and |v| < 0.05
      const bool close = fabs(x - goal[0]) <= 0.05 && fabs(y - goal[1]) <= 0.05;
      if (close && fabs(lu[0]) < 0.05) f.status[e] = TTMPC_FLEET_REACHED;
    }
    s_idx = idx;
    s_sw = (f.sw_state && f.hint && f.use_hint && was_running) ? 1 : 0;
    // speed reference (:248-255)
    const double dist = hyp2(x - goal[0], y - goal[1]);
    double v = f.base_speed;
    if (!(dist >= f.base_speed * N * d.ts)) {
      v = dist / N / d.ts;
      if (!(v > f.low_speed)) v = f.low_speed;
    }
    s_speed = v;
  }
  __syncthreads();
  const int idx = s_idx;
  // hybrid mode: main.py:200 evaluates HintSwitcher.switch before get_action's goal test.  The
  // geometric predicates are independent: one obstacle distance / one (point, obstacle)
  // containment per thread, then thread 0 walks the state machine over them.
  if (s_sw) {
    const int n_stc = f.sw_poly_xy ? f.sw_max_poly : 0, n_obs = n_stc + (f.dyn_cur ? f.n_dyn_live : 0);
    if (n_obs <= SW_MAX_OBS) {
      double rect[8]; int nv;
      for (int o = threadIdx.x; o < n_obs; o += blockDim.x) {
        const double *xy = sw_obstacle(f, e, o, n_stc, rect, &nv);
        s_dist[o] = nv < 3 ? -1.0 : poly_distance(xy, nv, st[0], st[1]);
      }
      for (int q = threadIdx.x; q < N * n_obs; q += blockDim.x) {
        const int k = q / n_obs, o = q - k * n_obs;
        const double *xy = sw_obstacle(f, e, o, n_stc, rect, &nv);
        int r = idx + k; if (r > L - 1) r = L - 1;
        s_in[q] = (nv >= 3 && poly_contains(xy, nv, ref[3 * r], ref[3 * r + 1])) ? 1 : 0;
      }
      __syncthreads();
      if (threadIdx.x == 0) f.use_hint[e] = hint_switch_sm(f, e, N, n_obs, s_dist, s_in);
    } else if (threadIdx.x == 0) {
      f.use_hint[e] = hint_switch(f, e, N, ref, L, idx);
    }
  }
  if (threadIdx.x == 0) s_hint = (f.hint && f.use_hint && f.use_hint[e]) ? 1 : 0;
  __syncthreads();
  double *p = p_all + (size_t)e * d.np;
  const int o_refs = 18, o_speed = o_refs + 3 * N, o_other = o_speed + N, o_stc = o_other + d.n_other,
            o_dyn = o_stc + d.n_stc, o_ws = o_dyn + d.n_dyn, o_wd = o_ws + N;
  const double *stc = f.stc + (f.stc_shared ? 0 : (size_t)e * d.n_stc);
  const bool hinted = s_hint != 0;  // hybrid mode: DQN hint positions
  const double *hint = f.hint + (size_t)e * N * 2;
  for (int o = threadIdx.x; o < d.np; o += blockDim.x) {
    double v;
    if (o < 3) v = st[o];
    else if (o < 6) {
      int r = idx + N - 1; if (r > L - 1) r = L - 1;
      v = (hinted && o < 5) ? hint[2 * (N - 1) + (o - 3)] : ref[3 * r + (o - 3)];
    }
    else if (o < 8) v = lu[o - 6];
    else if (o < o_refs) v = f.tuning[o - 8];
    else if (o < o_speed) {
      const int k = (o - o_refs) / 3, c = (o - o_refs) - 3 * k;
      int r = idx + k; if (r > L - 1) r = L - 1;
      v = (hinted && c < 2) ? hint[2 * k + c] : ref[3 * r + c];
    }
    else if (o < o_other) v = s_speed;
    else if (o < o_stc) v = f.other ? f.other[(size_t)e * d.n_other + (o - o_other)] : 0.0;
    else if (o < o_dyn) v = stc[o - o_stc];
    else if (o < o_ws) {
      const int q = o - o_dyn;
      if (f.dyn) v = f.dyn[(size_t)e * d.n_dyn + q];
      else {
        v = 0.0;
        const int j = q / (6 * N);
        if (f.dyn_cur && j < f.n_dyn_live) {  // est_dyn_obs_positions (main.py:80-89)
          const int i = (q - j * 6 * N) / 6, c = q - j * 6 * N - 6 * i;
          const double *cur = f.dyn_cur + ((size_t)e * f.n_dyn_live + j) * 2;
          const double *last = f.dyn_last + ((size_t)e * f.n_dyn_live + j) * 2;
          if (c < 2) v = cur[c] + (cur[c] - last[c]) * (double)(i + 1);
          else if (c < 4) v = f.dyn_size;
          else v = c == 4 ? 0.0 : 1.0;
        }
      }
    }
    else if (o < o_wd) v = f.stc_weight;
    else v = f.dyn_weight;
    p[o] = v;
  }
}

// One thread per robot: obstacles move, RUNNING robots take the first control.
__global__ void __launch_bounds__(128) fleet_advance_kernel(const ttmpc_fleet f, const FleetDims d,
                                                            const double *__restrict__ u_all,
                                                            const int *__restrict__ exit_status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= f.n) return;
  if (f.dyn_cur) {
    for (int j = 0; j < f.n_dyn_live; j++) {
      double *c = f.dyn_cur + ((size_t)e * f.n_dyn_live + j) * 2;
      double *l = f.dyn_last + ((size_t)e * f.n_dyn_live + j) * 2;
      const double *dd = f.dyn_disp + ((size_t)e * f.n_dyn_live + j) * 2;
      l[0] = c[0]; l[1] = c[1];
      c[0] = c[0] + dd[0]; c[1] = c[1] + dd[1];
    }
  }
  if (f.status[e] != TTMPC_FLEET_RUNNING) return;
  if (exit_status && exit_status[e] == TTMPC_NOT_FINITE) { f.status[e] = TTMPC_FLEET_FAILED; return; }
  const double *u = u_all + (size_t)e * 2 * d.N;
  double *st = f.state + 3 * e;
  const double v = u[0], w = u[1], ts = d.ts;
  // unicycle_model, RK4 branch, in numpy's operation order (motion_model.py:166-174)
  double s, c;
  tt_sincos(st[2], &s, &c);
  const double k1x = ts * (v * c), k1y = ts * (v * s), k1t = ts * w;
  tt_sincos(st[2] + 0.5 * k1t, &s, &c);
  const double k2x = ts * (v * c), k2y = ts * (v * s), k2t = ts * w;
  tt_sincos(st[2] + 0.5 * k2t, &s, &c);
  const double k3x = ts * (v * c), k3y = ts * (v * s), k3t = ts * w;
  tt_sincos(st[2] + k3t, &s, &c);
  const double k4x = ts * (v * c), k4y = ts * (v * s), k4t = ts * w;
  const double sixth = 1.0 / 6.0;
  st[0] = st[0] + sixth * (((k1x + 2.0 * k2x) + 2.0 * k3x) + k4x);
  st[1] = st[1] + sixth * (((k1y + 2.0 * k2y) + 2.0 * k3y) + k4y);
  st[2] = st[2] + sixth * (((k1t + 2.0 * k2t) + 2.0 * k3t) + k4t);
  f.last_u[2 * e] = v; f.last_u[2 * e + 1] = w;
}

cudaError_t launch_fleet_pack(const ttmpc_fleet &f, const FleetDims &d, double *p, cudaStream_t st) {
  fleet_pack_kernel<<<f.n, 128, 0, st>>>(f, d, p);
  return cudaGetLastError();
}
cudaError_t launch_fleet_advance(const ttmpc_fleet &f, const FleetDims &d, const double *u,
                                 const int *exit_status, cudaStream_t st) {
  fleet_advance_kernel<<<(f.n + 127) / 128, 128, 0, st>>>(f, d, u, exit_status);
  return cudaGetLastError();
}

}  // namespace ttmpc

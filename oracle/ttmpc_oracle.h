/*
 * ttmpc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C fp64 restatement of the reference's NMPC path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product (trajtrack_mpcndqn_rlboost_b200) never does.
 *
 * PARITY STATUS
 *   problem functions (f, grad f, F1, F2): PINNED -- checked against the
 *     reference's own MpcModule.build (src/mpc_traj_tracker/mpc/mpc_generator.py)
 *     executed in-container with a torch-backed casadi shim
 *     (tools/gen_golden_problem.py -> tests/golden/problem_*.npz).
 *   PANOC / ALM / L-BFGS: PARITY UNPINNED -- the algorithm lives in the
 *     third-party Rust crate `optimization_engine` (pulled by opengen==0.7.1,
 *     requirements.txt:25; crate 0.7.x, lbfgs 0.2.x), absent from
 *     /root/reference and not buildable here (no Rust).  Restated from the
 *     published algorithm (Stella et al. 2017; Sopasakis et al. 2020) and the
 *     crate's documented constants.  Anchored on the known answers of the two
 *     crates' own unit tests (lbfgs correctneess_buff_1; PANOC's fixed points
 *     mocks::SOLUTION_A / SOLUTION_HARD; the Lipschitz estimate of
 *     mocks::lipschitz_mock): ttmpc_oracle_lbfgs_kat / _panoc_mock /
 *     _lipschitz_mock below, tests/test_oracle_solver.py.  Building blocks and
 *     fixed points only -- the iterate path on the NMPC problem is unpinned.
 *   sector/ray observation: PARITY UNPINNED (shapely absent; checked by an
 *     independent method, tests/test_geometry_independent.py); Q-network: PINNED
 *     against torch with the reference's Model/ray/best_model.zip weights.
 */
#ifndef TTMPC_ORACLE_H
#define TTMPC_ORACLE_H
#include "../include/ttmpc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttmpc_oracle_status {
  int exit_status;
  int outer_iters;
  int inner_iters;
  double last_fpr;
  double delta_y_norm;
  double f2_norm;
  double penalty;
  double cost;
  long long n_cost_evals;
  long long n_grad_evals;
} ttmpc_oracle_status;

/* f(u;p), F1 [2N], F2 [Ndynobs].  Any output may be NULL. */
void ttmpc_oracle_eval(const ttmpc_config *cfg, const double *u, const double *p,
                       double *f, double *F1, double *F2);
/* psi(u; xi=(c,y), p) and its gradient (hand-written adjoint). */
double ttmpc_oracle_psi(const ttmpc_config *cfg, const double *u, const double *p,
                        double c, const double *y);
void ttmpc_oracle_psi_grad(const ttmpc_config *cfg, const double *u, const double *p,
                           double c, const double *y, double *grad);
/* One full OpEn-style solve.  u: in = initial guess, out = solution.
 * y: in = initial Lagrange multipliers, out = final (len 2N).          */
int ttmpc_oracle_solve(const ttmpc_config *cfg, const double *p, double *u,
                       double *y, double c0, ttmpc_oracle_status *st);
/* Same solve with the arithmetic ORDER of the CUDA kernel (32-lane scans and
 * butterfly sums, explicit fma, tt_sincos): reproduces the GPU bit for bit.   */
int ttmpc_oracle_solve_warp(const ttmpc_config *cfg, const double *p, double *u,
                            double *y, double c0, ttmpc_oracle_status *st);
void ttmpc_oracle_eval_warp(const ttmpc_config *cfg, const double *u, const double *p,
                            double c, const double *y, double *f, double *F2,
                            double *psi, double *grad);
void ttmpc_oracle_sincos(double x, double *s, double *c);
/* the PANOC engine of the restatement on the two unit-test problems of optimization_engine's mocks.rs
 * (known answers mocks::SOLUTION_A / SOLUTION_HARD); which = 1 or 2, u has 2 or 3 entries */
/* the L-BFGS restatement on a supplied history (known answer of the lbfgs crate's unit test); returns
 * 100 * accepted updates + active pairs */
int ttmpc_oracle_lbfgs_kat(int n, int mem, int n_updates, const double *g, const double *x, double *q,
                           double *alpha0, double *rho0, int cbfgs);
double ttmpc_oracle_lipschitz_mock(const double *u3);
int ttmpc_oracle_panoc_mock(int which, double *u, double tolerance, int lbfgs_memory, int max_iter,
                            int *iters, double *norm_fpr, long long *n_cost, long long *n_grad);
int ttmpc_oracle_solve_batch_mode(const ttmpc_config *cfg, int n, const double *p,
                                  int use_u0, int use_y0, const double *c0,
                                  const ttmpc_result *res, int threads, int warp);
/* Batched convenience (sequential loop, or `threads` pthreads).
 * Same per-scene arrays as ttmpc_result (host pointers).               */
int ttmpc_oracle_solve_batch(const ttmpc_config *cfg, int n, const double *p,
                             int use_u0, int use_y0, const double *c0,
                             const ttmpc_result *res, int threads);
/* Rollout of u from p.s: states [N][3] (trajectory_generator.py:296-301). */
void ttmpc_oracle_rollout(const ttmpc_config *cfg, const double *u, const double *p,
                          double *states);

/* Fleet step (caller side of the solve), host pointers in ttmpc_fleet; see ttfleet_oracle.c.
 * use_libm = 1: libm hypot/sin/cos like the reference's Python; 0: the device's arithmetic. */
void ttfleet_oracle_pack(const ttmpc_config *cfg, const ttmpc_fleet *fleet, double *p, int use_libm);
void ttfleet_oracle_advance(const ttmpc_config *cfg, const ttmpc_fleet *fleet, const double *u,
                            const int *exit_status, int use_libm);

/* shapely Polygon.contains(Point) / Polygon.distance(Point) as HintSwitcher uses them (main_pre.py:35-52) */
int ttfleet_oracle_poly_contains(const double *xy, int nv, double px, double py);
double ttfleet_oracle_poly_distance(const double *xy, int nv, double px, double py);

/* DQN companion oracle: one env. */
void ttdqn_oracle_observe(const ttdqn_scene_layout *lay, const double *agent,
                          const double *poly_xy, const int *poly_off,
                          const int *is_solid, int n_poly, double *seg_dist,
                          double *ray_dist);
void ttdqn_oracle_observe_act(const ttdqn_scene_layout *lay, const ttdqn_qnet *qnet,
                              int n_envs, const double *agent, const double *poly_xy,
                              const int *poly_off, const int *is_solid,
                              const int *n_poly, const float *internal,
                              float *old_ext, float *ext, float *q, int *action,
                              double *seg_dist, double *ray_dist);

/* Internal observation of the ray model and the DQN hint trajectory (see ttdqn_oracle.c). */
double ttdqn_oracle_project(const double *xy, int n, double px, double py);
void ttdqn_oracle_interpolate(const double *xy, int n, double s, double *x, double *y);
void ttdqn_oracle_internal_obs(int corner_samples, double offset, double max_distance,
                               const double *agent5, const double *path_xy, int n_nodes,
                               float *obs, double *progress);
void ttdqn_oracle_rl_ref(int steps, double ts, double ref_speed, const double *agent5, int action,
                         double *rl_ref, int use_libm);

#ifdef __cplusplus
}
#endif
#endif

/*
 * ttmpc.h -- C-ABI of the B200-native batched NMPC planner + DQN companion.
 *
 * This is the drop-in boundary for ONE path of Woodenonez/TrajTrack-MPCnDQN-RLBoost:
 * the call that the reference makes into its OpEn-generated (Rust, PyO3) solver
 *
 *     solution = self.solver.run(parameters, initial_guess)
 *         -- src/mpc_traj_tracker/trajectory_generator.py:284
 *     class Solver: def run(self, p, initial_guess, initial_lagrange_multipliers,
 *                           initial_penalty) -> SolverStatus
 *         -- src/mpc_traj_tracker/trajectory_generator.py:27-29
 *
 * and, for the DQN-boosted loop, the observation + policy call
 *
 *     obsv['external'] = SectorAndRayObservation.external_obs()
 *         -- src/pkg_dqn/environment/components/ext_obsv_sector_and_ray.py:33-83
 *     action_index, _ = model.predict(obsv, deterministic=True)
 *         -- src/main.py:148 (SB3 DQN MultiInputPolicy, net_arch [16,16])
 *
 * Everything is plain pointers and sizes.  No torch types, no C++ types.
 * Pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 * All entry points return 0 on success, a negative ttmpc error code otherwise;
 * ttmpc_last_error() gives a human-readable string for the calling thread.
 *
 * The packed parameter vector p of one scene is laid out exactly as the
 * reference assembles it (trajectory_generator.py:251-254, mpc_generator.py:175-184):
 *
 *   s     [2*ns+nu]              x y theta | x_goal y_goal theta_goal | v_init w_init
 *   q     [nq = 10]              qpos qvel qtheta rv rw qN qthetaN qrpd acc_pen wacc_pen
 *   r     [ns*N + N]             (x y theta) * N reference states, then N speed refs
 *   c     [ns*N*Nother]          predicted states of other robots, robot-major
 *   o_s   [Nstcobs*nstcobs]      per obstacle: b[ne] a0[ne] a1[ne]   (ne = nstcobs/3)
 *   o_d   [Ndynobs*ndynobs*N]    per obstacle, per step: cx cy rx ry angle alpha
 *   q_stc [N]                    static-obstacle weights (unused by the cost, kept)
 *   q_dyn [N]                    dynamic-obstacle soft weights
 *
 * Batches are scene-major: p[scene][np], u[scene][nu*N], y[scene][2*N].
 */
#ifndef TTMPC_H
#define TTMPC_H

#ifdef __cplusplus
extern "C" {
#endif

#define TTMPC_VERSION 100

/* exit_status codes; names are the strings the OpEn python binding returns
 * (format!("{:?}", ExitStatus)), which the reference matches against
 * config.bad_exit_codes (config/mpc_default.yaml:58).  */
enum {
  TTMPC_CONVERGED = 0,                 /* "Converged"              */
  TTMPC_NOT_CONVERGED_ITERATIONS = 1,  /* "NotConvergedIterations" */
  TTMPC_NOT_CONVERGED_OUT_OF_TIME = 2, /* "NotConvergedOutOfTime"  */
  TTMPC_NOT_FINITE = 3                 /* SolverError::NotFiniteComputation: the
                                          binding returns None -> reference raises */
};

enum {
  TTMPC_OK = 0,
  TTMPC_ERR_BAD_CONFIG = -1,
  TTMPC_ERR_BAD_ARG = -2,
  TTMPC_ERR_CUDA = -3,
  TTMPC_ERR_UNSUPPORTED = -4
};

/* Problem + solver configuration.
 * Problem fields mirror config/mpc_default.yaml; solver fields mirror
 * opengen.config.SolverConfiguration as used at mpc_generator.py:268-276
 * (initial_penalty 10, everything else opengen 0.7.1 defaults).          */
typedef struct ttmpc_config {
  /* dimensions */
  int N_hor;    /* horizon, 1..32 (one lane per step)                  */
  int nu;       /* must be 2 (v, omega)                                */
  int ns;       /* must be 3 (x, y, theta)                             */
  int nq;       /* must be 10                                          */
  int Nother;   /* other-robot slots                                   */
  int Nstcobs;  /* static obstacle slots                               */
  int nstcobs;  /* doubles per static obstacle = 3*edges, edges<=8     */
  int Ndynobs;  /* dynamic obstacle slots                              */
  int ndynobs;  /* must be 6                                           */
  int _pad0;
  /* model */
  double ts;
  double vehicle_width;  /* fleet safe distance                         */
  double social_margin;  /* soft-ellipse inflation                      */
  double lin_vel_min, lin_vel_max, ang_vel_max; /* box on u              */
  double lin_acc_min, lin_acc_max, ang_acc_max; /* set C of the ALM map  */
  /* solver (OpEn) */
  double tolerance;              /* epsilon, 1e-4                        */
  double initial_tolerance;      /* initial inner AKKT tolerance, 1e-4   */
  double delta_tolerance;        /* constraints tolerance, 1e-4          */
  double initial_penalty;        /* 10 (mpc_generator.py:269)            */
  double penalty_update_factor;  /* 5                                    */
  double inner_tolerance_update_factor; /* 0.1                           */
  double sufficient_decrease_coeff;     /* 0.1                           */
  int lbfgs_memory;              /* 10, max 16                           */
  int max_inner_iterations;      /* 500                                  */
  int max_outer_iterations;      /* 10                                   */
  int max_duration_ms;           /* wall-clock budget of ONE scene's solve, 5000 =
                                    MAX_SOVLER_TIME (mpc_generator.py:22,270: with_max_duration_micros);
                                    0 = no limit.  Measured on the device from the moment the scene's
                                    solve starts; checked where OpEn checks it (after every PANOC step,
                                    before every outer iteration) -> TTMPC_NOT_CONVERGED_OUT_OF_TIME.
                                    Like in the reference, a solve that hits the budget is not
                                    reproducible; the CPU oracle ignores the field.               */
} ttmpc_config;

/* Fill cfg with config/mpc_default.yaml + opengen defaults. */
void ttmpc_default_config(ttmpc_config *cfg);
/* Number of doubles in one packed parameter vector / decision vector / F1 / F2. */
int ttmpc_num_params(const ttmpc_config *cfg);
int ttmpc_num_decision(const ttmpc_config *cfg);
int ttmpc_num_alm(const ttmpc_config *cfg);     /* n1 = 2*N_hor */
int ttmpc_num_penalty(const ttmpc_config *cfg); /* n2 = Ndynobs */
const char *ttmpc_last_error(void);
const char *ttmpc_exit_status_name(int code);
int ttmpc_version(void);
/* The CUDA device of the CALLING THREAD (cudaSetDevice / cudaGetDevice): every entry point works on
 * the calling thread's current device, so a host thread that serves GPU k calls ttmpc_set_device(k)
 * once before its first solve (new threads start on device 0).                                  */
int ttmpc_set_device(int device);
int ttmpc_get_device(int *device);

/* Per-scene outputs of a batched solve.  Any pointer may be NULL (skipped)
 * except u.  Mirrors the fields of OpEn's python OptimizerSolution that the
 * reference reads (trajectory_generator.py:286-289) plus the diagnostics.  */
typedef struct ttmpc_result {
  double *u;            /* [n][nu*N]  in: initial guess (see use_u0), out: solution */
  double *cost;         /* [n]  f(u*) (psi with c = 0)                              */
  int *exit_status;     /* [n]                                                      */
  int *outer_iters;     /* [n]                                                      */
  int *inner_iters;     /* [n]                                                      */
  double *last_fpr;     /* [n]  last_problem_norm_fpr                               */
  double *f1_infeas;    /* [n]  delta_y_norm / c                                    */
  double *f2_norm;      /* [n]                                                      */
  double *penalty;      /* [n]  final c                                             */
  double *y;            /* [n][2*N] in: initial multipliers (see use_y0), out: final */
  double *pred_states;  /* [n][N][ns] rollout of u* from p.s, i.e. from the state the solve STARTED in.
                           NOTE: the reference's list (trajectory_generator.py:296-301) starts from the
                           state AFTER u[0] was applied (taken_states[-1]) and re-applies u from there, so
                           it is this rollout shifted by one step; the Python TrajectoryGenerator mirror
                           rebuilds the reference's list on the host from u.                          */
  long long *evals;     /* [n][4] cost-only evals, cost+gradient evals, solve time [ns],
                           solve start [ns, %globaltimer] (diagnostics)              */
} ttmpc_result;

/* ------------------------------------------------------------------------
 * Batched NMPC solve, device-resident.  All pointers in res and d_p are
 * device pointers.  Asynchronous on `stream` (a cudaStream_t passed as void*).
 *   use_u0 = 0: initial guess is all-zero, like solver.run(p) with
 *               initial_guess=None; res->u is output only.
 *   use_y0 = 0: multipliers start at zero.  (The reference's Solver object
 *               keeps them between run() calls; the host mirror does that.)
 *   d_c0      : optional per-scene initial penalty (NULL -> cfg value).
 * Batches in flight: solves issued on the SAME stream run one after the other;
 * solves issued on DIFFERENT streams (each with its own result buffers; 16 stream
 * bindings per device, a 17th stream drains the device once and rebinds) may overlap -- every stream has its own scene queue and scratch
 * tables inside the library, and the CTAs of the next batch become resident as
 * the CTAs of the running one drain.  Results do not depend on what else runs.
 * ------------------------------------------------------------------------ */
int ttmpc_solve_batch_device(const ttmpc_config *cfg, int n_scenes,
                             const double *d_p, int use_u0, int use_y0,
                             const double *d_c0, const ttmpc_result *res,
                             void *stream);

/* Same call with HOST buffers: stages p/u/y through pinned memory, launches,
 * copies results back, synchronises.  This is what the python Solver.run
 * mirror and the `e2e` bench leg use.  Thread-safe: calls from several host
 * threads each take their own staging buffers / streams (8 sets per device, a
 * ninth caller waits) and overlap on the GPU.                              */
int ttmpc_solve_batch_host(const ttmpc_config *cfg, int n_scenes,
                           const double *h_p, int use_u0, int use_y0,
                           const double *h_c0, const ttmpc_result *res);

/* Evaluate the problem functions on device for a batch (testing / parity):
 * f, F1 [n][2N], F2 [n][Ndynobs], and psi/grad psi at xi = (c, y).
 * Any output may be NULL.  d_c [n], d_y [n][2N] may be NULL (c=0, y=0).   */
int ttmpc_eval_batch_device(const ttmpc_config *cfg, int n_scenes,
                            const double *d_p, const double *d_u,
                            const double *d_c, const double *d_y, double *d_f,
                            double *d_F1, double *d_F2, double *d_psi,
                            double *d_grad, void *stream);
int ttmpc_eval_batch_host(const ttmpc_config *cfg, int n_scenes,
                          const double *h_p, const double *h_u,
                          const double *h_c, const double *h_y, double *h_f,
                          double *h_F1, double *h_F2, double *h_psi,
                          double *h_grad);

/* ------------------------------------------------------------------------
 * DQN companion: sector+ray ("lidar") observation against the obstacle set,
 * then Q-network inference and argmax, one environment per warp.
 *
 * Geometry of one env (all fp64, world frame, already padded as the reference
 * pads them: obstacle.py:158-163,248-254):
 *   polygons: n_poly closed rings, ring i has poly_nv[i] vertices stored at
 *             poly_xy[poly_off[i] .. ], (x,y) pairs.  is_solid[i] != 0 for an
 *             obstacle (filled Polygon), 0 for the boundary ring (LineString).
 * Batched layout: fixed capacity per env: max_poly rings, max_vert vertices.
 * ------------------------------------------------------------------------ */
typedef struct ttdqn_scene_layout {
  int num_segments; /* 8 (rays_reward1.py:20)                              */
  int max_poly;     /* ring slots per env                                  */
  int max_vert;     /* vertex slots per env (sum over rings)               */
  int n_internal;   /* 14 internal observation entries                     */
  int use_memory;   /* 1: external obs = [seg, ray, old_seg, old_ray]      */
  int _pad;
  double ray_length;   /* L = 1000 (ext_obsv_sector_and_ray.py:34)         */
  double max_distance; /* normalize_distance max_distance = 10 (utils.py:10) */
} ttdqn_scene_layout;

/* Q-network weights, fp32 row-major [out][in] like torch.nn.Linear.
 * Input is concat(external[4*num_segments or 2*num_segments], internal[n_internal])
 * (SB3 CombinedExtractor concatenates Dict keys in sorted order:
 *  'external' then 'internal').                                           */
typedef struct ttdqn_qnet {
  int n_in, n_h1, n_h2, n_out; /* 46, 16, 16, 9 */
  const float *w0, *b0, *w1, *b1, *w2, *b2;
} ttdqn_qnet;

void ttdqn_default_layout(ttdqn_scene_layout *lay);
/* Last failure of a ttdqn_* entry point on the calling thread (the same text is also what
 * ttmpc_last_error() returns, so one query serves the whole library).                     */
const char *ttdqn_last_error(void);

/* Device-resident batched observe + act.
 *   d_agent    [n][3]   x y theta (fp64)
 *   d_poly_xy  [n][max_vert][2], d_poly_off [n][max_poly+1] (prefix offsets),
 *   d_is_solid [n][max_poly], d_n_poly [n]
 *   d_internal [n][n_internal] fp32 internal observation
 *   d_old_ext  [n][2*num_segments] fp32 in/out memory of previous (seg, ray) obs
 *              (NULL when use_memory == 0)
 * outputs (any may be NULL): d_ext [n][n_ext] fp32, d_q [n][n_out] fp32,
 *   d_action [n] int32, d_seg_dist / d_ray_dist [n][num_segments] fp64 raw distances.
 * qnet pointers are device pointers.                                      */
int ttdqn_observe_act_device(const ttdqn_scene_layout *lay, const ttdqn_qnet *qnet,
                             int n_envs, const double *d_agent,
                             const double *d_poly_xy, const int *d_poly_off,
                             const int *d_is_solid, const int *d_n_poly,
                             const float *d_internal, float *d_old_ext,
                             float *d_ext, float *d_q, int *d_action,
                             double *d_seg_dist, double *d_ray_dist, void *stream);
int ttdqn_observe_act_host(const ttdqn_scene_layout *lay, const ttdqn_qnet *qnet,
                           int n_envs, const double *h_agent,
                           const double *h_poly_xy, const int *h_poly_off,
                           const int *h_is_solid, const int *h_n_poly,
                           const float *h_internal, float *h_old_ext,
                           float *h_ext, float *h_q, int *h_action,
                           double *h_seg_dist, double *h_ray_dist);

/* ------------------------------------------------------------------------
 * Fleet step: the caller side of the solve, batched on the device.
 *
 * One control step of n independent robots, each following the reference's
 * InterfaceMpc.get_action -> TrajectoryGenerator.run_step sequence
 * (src/interface_mpc.py:80-92, src/mpc_traj_tracker/trajectory_generator.py:203-274):
 *
 *   pack    : termination test (check_termination_condition, :158-164), closest point of
 *             the global reference trajectory in the window [idx-1, idx+5) and the N
 *             reference states from there (get_local_ref_traj, :203-230), speed
 *             reference from the distance to the goal (:248-255), dynamic-obstacle rows
 *             extrapolated from two consecutive positions (main.py:80-89,
 *             est_dyn_obs_positions), and the packed parameter vector in the block
 *             order of :257-260.
 *   solve   : ttmpc_solve_batch_device.
 *   advance : state <- unicycle RK4 step with the first control (run_solver, :291-294;
 *             motion_model.py:153-176), last action, solver failure flag
 *             (NotFiniteComputation -> the reference raises), obstacle positions moved by
 *             their per-step displacement.
 *
 * All pointers are device pointers unless the struct member says otherwise; in/out
 * members persist between steps, so K steps run without touching the host.
 * ------------------------------------------------------------------------ */
enum {
  TTMPC_FLEET_RUNNING = 0,  /* get_action returned an action                      */
  TTMPC_FLEET_REACHED = 1,  /* check_termination_condition: get_action -> None    */
  TTMPC_FLEET_FAILED = 2    /* solver error (the reference raises RuntimeError)   */
};
typedef struct ttmpc_fleet {
  int n;            /* robots                                                         */
  int ref_stride;   /* rows per robot in ref_traj                                     */
  double *state;    /* [n][3] x y theta, in/out                                       */
  const double *goal;     /* [n][3]                                                   */
  double *last_u;   /* [n][2] last applied action, in/out (zeros before the first)    */
  int *idx_ref;     /* [n] index into the global reference trajectory, in/out         */
  int *status;      /* [n] TTMPC_FLEET_*, in/out (robots not RUNNING are left alone)  */
  const double *ref_traj; /* [n][ref_stride][3] get_global_ref_traj output            */
  const int *ref_len;     /* [n] valid rows                                           */
  const double *stc;      /* static half-space rows: [n][Nstcobs*nstcobs], or one
                             shared row block when stc_shared != 0                   */
  int stc_shared;
  int n_dyn_live;   /* moving obstacles described by dyn_cur / dyn_last (<= Ndynobs)  */
  int action_steps; /* config.action_steps; only 1 is supported                       */
  int _pad;
  const double *other;    /* [n][ns*N*Nother] other-robot predictions or NULL (zeros) */
  const double *dyn;      /* [n][Ndynobs*ndynobs*N] ready-made rows or NULL           */
  double *dyn_cur;  /* [n][n_dyn_live][2] current obstacle positions (in/out) or NULL */
  double *dyn_last; /* [n][n_dyn_live][2] positions one step earlier (in/out)         */
  const double *dyn_disp; /* [n][n_dyn_live][2] displacement per step (advance)       */
  double dyn_size;  /* rx = ry of the extrapolated rows (main.py:32 DYN_OBS_SIZE)     */
  double tuning[10];      /* q block (set_work_mode, trajectory_generator.py:117-131) */
  double base_speed, low_speed;
  double stc_weight, dyn_weight;
  /* hybrid mode (main.py:194-201): where use_hint[e] != 0 the N reference positions are the
   * DQN hint hint[e] (rl_ref) and the headings stay those of the original local reference
   * (InterfaceMpc.get_local_ref_traj(rl_ref) + ref_traj_filter(decay = 1)).  NULL = never. */
  const double *hint;     /* [n][N][2] */
  int *use_hint;          /* [n] input -- or OUTPUT when sw_state != NULL             */
  /* HintSwitcher.switch (main_pre.py:27-52), evaluated by the pack kernel when sw_state != NULL
   * with the robot position, the ORIGINAL local reference and the obstacle list
   * "processed static polygons + circle_to_rect(moving obstacle)" (main.py:91-95,200):
   * shapely Polygon.contains / distance restated (interior strictly, 0 inside).      */
  int *sw_state;            /* [n][2] switch_on, detach_cnt (in/out) or NULL           */
  const double *sw_poly_xy; /* [n or 1][sw_max_poly][sw_max_pv][2] static polygons      */
  const int *sw_poly_nv;    /* [n or 1][sw_max_poly] vertex counts (0 = unused slot)    */
  int sw_max_poly, sw_max_pv, sw_poly_shared, sw_detach_steps;
  double sw_switch_distance, sw_detach_distance; /* HintSwitcher(10, 2, 10) in main.py:129 */
  double sw_dyn_radius;     /* circle_to_rect radius (main.py:91: DYN_OBS_SIZE)         */
} ttmpc_fleet;

/* d_p [n][np] is written for every robot (robots that are not RUNNING keep packing from
 * their frozen state, so the batch stays dense). */
int ttmpc_fleet_pack_device(const ttmpc_config *cfg, const ttmpc_fleet *fleet, double *d_p,
                            void *stream);
/* d_u [n][nu*N] solutions, d_exit_status [n] of the solve that used d_p. */
int ttmpc_fleet_advance_device(const ttmpc_config *cfg, const ttmpc_fleet *fleet,
                               const double *d_u, const int *d_exit_status, void *stream);
/* pack + solve (cold start, like run_step with initial_guess=None; multipliers carried in
 * res->y when use_y0 != 0, like the reference's Solver object) + advance, on `stream`. */
int ttmpc_fleet_step_device(const ttmpc_config *cfg, const ttmpc_fleet *fleet, double *d_p,
                            int use_y0, const ttmpc_result *res, void *stream);

/* ------------------------------------------------------------------------
 * DQN side of the hybrid loop (src/main.py:174-193), device-resident.
 *
 * ttdqn_internal_obs_device: the `internal` observation of the ray model
 *   (variants/rays_reward1.py:27-31): SpeedObservation, AngularVelocityObservation,
 *   ReferencePathSampleObservation(1, 0, sample_offset), ReferencePathCornerObservation(
 *   corner_samples) -- components/int_obsv_*.py -- with path_progress =
 *   path.project(agent) (environment.py:115).
 *     d_agent5   [n][5]  x y theta v w
 *     d_path_xy  [n][max_nodes][2] reference-path polyline, d_path_n [n] node counts (>= 2)
 *     d_internal [n][5 + 3*corner_samples] fp32 out, d_progress [n] out (may be NULL)
 * ttdqn_rl_ref_device: the DQN "hint" trajectory: MobileRobot.step(action, ts) followed by
 *   steps-1 x step_with_ref_speed(ts, ref_speed) on a copy of the agent
 *   (environment/agent.py:86-145); d_rl_ref [n][steps][2] positions.
 * ------------------------------------------------------------------------ */
int ttdqn_internal_obs_device(int n_envs, int max_nodes, int corner_samples, double sample_offset,
                              double max_distance, const double *d_agent5, const double *d_path_xy,
                              const int *d_path_n, float *d_internal, double *d_progress, void *stream);
int ttdqn_rl_ref_device(int n_envs, int steps, double ts, double ref_speed, const double *d_agent5,
                        const int *d_action, double *d_rl_ref, void *stream);

/* Cumulative device-side counters since the last reset: [0] cost-only
 * evaluations, [1] cost+gradient evaluations, [2] dynamic-obstacle bodies that
 * passed the bounding test, [3] PANOC iterations.  Synchronises the device.   */
int ttmpc_read_stats(unsigned long long out[4], int reset);
/* Launch geometry the solve kernel uses for n_scenes (reporting only). */
int ttmpc_launch_info(const ttmpc_config *cfg, int n_scenes, int *grid, int *block,
                      int *smem_bytes, int *blocks_per_sm, int *sm_count);

/* Single-warp latency probe (diagnostics): cycles per call of {cost eval, gradient
 * eval, butterfly sum, L-BFGS apply, fp64 divide, sqrt, dependent DFMA, checksum}. */
int ttmpc_probe_latency(const ttmpc_config *cfg, const double *h_p, long long out[8], int reps);

/* FP64 FMA-pipe peak probe: runs a dependent-chain-free DFMA kernel on every
 * SM and returns achieved TFLOP/s (used by bench.py as the roofline
 * denominator for the FP64-bound solve kernel).                            */
int ttmpc_measure_fp64_peak(double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TTMPC_H */

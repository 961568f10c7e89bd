/*
 * icem_b200 -- C ABI of the B200-native iCEM planner (libicem_b200.so).
 *
 * The reference (martius-lab/iCEM) has no FFI: its drop-in boundary is the Python `Controller`
 * plugin API (`beginning_of_rollout` / `get_action`, icem/misc/base_types.py:42-59,
 * icem/controllers/abstract_controller.py:43-58).  `icem_b200/controller.py::MpcICemB200` is the
 * thin Python class that mirrors that API and forwards to the entry points below through ctypes;
 * INTEGRATION.md shows the binding.  Every entry point cites the reference interface it replaces
 * (paths relative to /root/reference/icem/).
 *
 * Conventions: plain pointers and sizes, no torch types; all functions return 0 on success and a
 * non-zero code on failure, with a thread-local message available from icem_last_error();
 * all `float*`/`double*`/`int32_t*` arguments are HOST pointers unless the name ends in `_dev`.
 */
#ifndef ICEM_B200_H
#define ICEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICEM_ABI_VERSION 9

/* status codes */
enum { ICEM_OK = 0, ICEM_ERR_INVALID = 1, ICEM_ERR_CUDA = 2, ICEM_ERR_STATE = 3, ICEM_ERR_UNSUPPORTED = 4,
       ICEM_ERR_COMM = 5 };

/* forward models (what `forward_model.predict_n_steps` runs, controllers/mpc.py:56-67) */
enum {
  ICEM_DYN_DENSE_TANH = 0,       /* obs' = tanh(W_o obs + W_a act + b): single dense layer batched model
                                    (models/abstract_models.py:17-53 with a dense `predict`) */
  ICEM_DYN_HALFCHEETAH = 1,      /* ground-truth planar articulated body (stands in for gym HalfCheetah-v3) */
  ICEM_DYN_HUMANOID_STANDUP = 2, /* ground-truth 3-D articulated body (stands in for gym HumanoidStandup-v2) */
  ICEM_DYN_MLP = 3,              /* dense MLP forward model (tensor-core rollout) */
  ICEM_DYN_ARTICULATED = 4       /* any articulated body given through icem_set_articulated_model (Hopper, Ant, ...) */
};

/* per-step cost functions (env.cost_fn, controllers/abstract_controller.py:70,78-80) */
enum {
  ICEM_COST_HALFCHEETAH = 0,     /* environments/mujoco.py:67-99  */
  ICEM_COST_HUMANOID_STANDUP = 1,/* environments/mujoco.py:259-277 */
  ICEM_COST_LOCOMOTION = 2,      /* Ant (environments/mujoco.py:153-176), Hopper (:196-231), Humanoid (:314-343):
                                    -w_fwd * x_velocity + w_unhealthy * unhealthy(obs) + w_ctrl * |a|^2 with
                                    x_velocity = (next_obs[0] - obs[0]) / dt (Ant, Hopper; reads next_obs, so all h
                                    steps are simulated) or obs[cost_velocity_index] (Humanoid); parameters in
                                    icem_config_t.cost_* */
  ICEM_COST_REACHER = 3,         /* Reacher (environments/mujoco.py:346-368): |fingertip - target|, the last three
                                    entries of gym's observation.  On the device the observation is the state
                                    (q0, q1 = arm hinges, q2, q3 = target slides), so the distance is formed from the
                                    planar forward kinematics with the constants in icem_config_t.cost_reach */
  ICEM_COST_GOAL_DISTANCE = 4    /* goal-space envs (environments/abstract_environments.py:115-123
                                    MaskedGoalSpaceEnvironmentInterface.cost_fn; environments/robotics.py:150-164
                                    FetchPickAndPlace.cost_fn): d = |obs[goal..goal+3) - obs[achieved..achieved+3)|,
                                    shaped: e = |obs[0..3) - obs[3..6)| (end effector to box);
                                    dense: d + 0.1 e, sparse: [d > threshold] + 0.1 [e > threshold].  Reads the
                                    observation only: for the batched models (ICEM_DYN_MLP, ICEM_DYN_DENSE_TANH) */
};

/* action sampler / planner family */
enum {
  ICEM_PLANNER_ICEM = 0,     /* MpcICem  (controllers/icem.py): colored / white Gaussian noise + clip            */
  ICEM_PLANNER_CEM_STD = 1,  /* MpcCemStd (controllers/mpc.py:142-327): truncated normal sampling, no decay,
                                 no elite reuse; bounds update, execute_best_elite / shift_means switches        */
  ICEM_PLANNER_RANDOM = 2    /* MpcRandom (controllers/mpc.py:86-138): random shooting, one population of
                                 piecewise-constant uniform actions, execute the best trajectory's first action  */
};

/* cost_along_trajectory (controllers/abstract_controller.py:82-91) */
enum { ICEM_REDUCE_SUM = 0, ICEM_REDUCE_BEST = 1, ICEM_REDUCE_FINAL = 2 };

/* Keyword surface of MpcICem (controllers/icem.py:22,213-233; controllers/mpc.py:22;
 * controllers/abstract_controller.py:64-65) plus what the env contributes. */
typedef struct icem_config {
  int32_t abi_version;            /* = ICEM_ABI_VERSION */
  int32_t device;                 /* CUDA device ordinal */
  int32_t horizon;                /* h */
  int32_t act_dim;                /* d = env.action_space.shape[0] */
  int32_t num_simulated_trajectories; /* N (GLOBAL population of iteration 0) */
  int32_t opt_iterations;
  int32_t elites_size;
  int32_t use_mean_actions;
  int32_t keep_previous_elites;
  int32_t shift_elites_over_time;
  int32_t cost_along_trajectory;  /* ICEM_REDUCE_* */
  int32_t dynamics;               /* ICEM_DYN_* */
  int32_t cost;                   /* ICEM_COST_* */
  int32_t cost_penalise_flipping; /* HalfCheetah env_params.penalise_flipping */
  int32_t obs_dim;                /* observation width the cost reads (17/18 cheetah; >=3 humanoid) */
  int32_t colorednoise_v2;        /* 0: colorednoise 1.x spectrum (default); 1: 2.x sqrt(2) DC/Nyquist scaling */
  int32_t keep_iteration_actions; /* debug/parity: keep every iteration's population for icem_get_actions */
  int32_t world_size;             /* number of ranks sharing one plan (1 = single GPU) */
  int32_t rank;
  int32_t planner;                /* ICEM_PLANNER_* */
  int32_t execute_best_elite;     /* CEM_STD (mpc.py:237-240): 1 = first action of the best trajectory, 0 = mean[0] */
  int32_t shift_means;            /* CEM_STD (mpc.py:243-248): 1 = time-shift the mean, 0 = reset it to zeros        */
  int32_t bounds_like_levine;     /* CEM_STD (mpc.py:290-301): clamp std to half the distance to the bounds, +-2 sigma */
  int32_t action_change_frequency;/* RANDOM (mpc.py:91,95-101): a drawn action is held for this many further sample()
                                     calls; calls run over the whole population row by row, and on across plan steps */
  int32_t num_problems;           /* independent MPC problems batched in ONE handle (0 or 1 = a single problem): same
                                     settings and model, own start state / mean / std / elites each, Philox seed + i;
                                     every kernel of a plan step covers all problems (grid.y = problem).  See
                                     icem_plan_batch. */
  int32_t cost_z_index;           /* LOCOMOTION: observation index of the height (Hopper 1, Ant 2) */
  int32_t cost_z_strict;          /* LOCOMOTION: 1 = z_lo < z < z_hi (Hopper, mujoco.py:205), 0 = z_lo <= z <= z_hi (Ant, :149) */
  int32_t cost_velocity_index1;   /* LOCOMOTION: 0 = x velocity by finite difference of obs[0]; k > 0 = read obs[k - 1]
                                     (Humanoid: observation[nq], mujoco.py:333) */
  int32_t cost_reserved;          /* keeps the doubles 8-byte aligned; must be 0 */
  int32_t cost_goal_index;        /* GOAL_DISTANCE: first observation index of the desired goal (goal_idx[0]) */
  int32_t cost_achieved_index;    /* GOAL_DISTANCE: first observation index of the achieved goal (Fetch: 3 = box, FetchReach: 0) */
  int32_t cost_goal_sparse;       /* GOAL_DISTANCE: 1 = thresholded 0/1 cost (env_params.sparse) */
  int32_t cost_goal_shaped;       /* GOAL_DISTANCE: 1 = add 0.1 * end-effector-to-box term (FetchPickAndPlace shaped_reward) */
  double factor_decrease_num;     /* gamma */
  double alpha;
  double init_std;
  double fraction_elites_reused;
  double noise_beta;
  double cost_dt;                 /* LOCOMOTION: env.dt = timestep * frame_skip */
  double cost_ctrl_weight;        /* LOCOMOTION: _ctrl_cost_weight (Hopper 1e-3, Ant 0.5) */
  double cost_unhealthy_weight;   /* LOCOMOTION: 200 (Hopper, mujoco.py:227) / 100 (Ant, :171) */
  double cost_z_lo, cost_z_hi;    /* LOCOMOTION: _healthy_z_range */
  double cost_state_bound;        /* LOCOMOTION: Hopper |states[..., 2:]| < bound (_healthy_state_range); <= 0: none */
  double cost_forward_weight;     /* LOCOMOTION: weight of the velocity term (Humanoid _forward_reward_weight 1.25); 0 = 1 */
  double cost_goal_threshold;     /* GOAL_DISTANCE: env_params.threshold */
  double cost_reach[4];           /* REACHER: link 1 length, link 2 length to the fingertip, world x, y of the target at
                                     q2 = q3 = 0 (gym reacher.xml: 0.1, 0.11, 0, 0 -- the slides' `ref` cancels the body offset) */
  uint64_t seed;                  /* Philox key (production noise) */
  const float* action_low;        /* [d] env.action_space.low  (float32 like gym.spaces.Box) */
  const float* action_high;       /* [d] */
} icem_config_t;

typedef struct icem_planner icem_planner_t;

const char* icem_last_error(void);
int icem_abi_version(void);
/* sizeof(icem_config_t) / sizeof(icem_articulated_model_t) as the library was compiled: a binding checks its own
 * mirror of the structs against these before the first call (a stale library would read shifted fields) */
int icem_config_sizeof(void);
int icem_articulated_model_sizeof(void);
/* number of kernels of this library launched by the calling process so far (bench `gpu_launches`) */
uint64_t icem_kernel_launch_count(void);

/* ---- planner lifetime: MpcICem.__init__ (controllers/icem.py:22-29, :235-247) ------------------------- */
int icem_create(const icem_config_t* cfg, icem_planner_t** out);
int icem_destroy(icem_planner_t* p);

/* ---- forward-model parameters (the object `main.py:105-109` builds and hands to the controller) -------- */
/* ICEM_DYN_DENSE_TANH: w_obs[obs_dim*obs_dim] row-major, w_act[obs_dim*act_dim], bias[obs_dim] (may be NULL) */
int icem_set_dense_model(icem_planner_t* p, int32_t obs_dim, const float* w_obs, const float* w_act,
                         const float* bias);
/* ICEM_DYN_MLP: n_layers dense layers, layer l: weight[out_l*in_l] row-major + bias[out_l]; tanh between
 * layers, last layer linear, output = next observation delta (obs' = obs + mlp([obs, act])) */
int icem_set_mlp_model(icem_planner_t* p, int32_t n_layers, const int32_t* layer_dims /* n_layers+1 */,
                       const float* const* weights, const float* const* biases);

/* Integrator of the articulated ground-truth models (MuJoCo's <option integrator=...>, reached by the reference through
 * environments/mujoco.py:101-131 `do_simulation`): semi-implicit Euler with joint springs / dampers in the system
 * matrix, or the four-stage Runge-Kutta of mj_RungeKutta that gym's half_cheetah.xml / humanoidstandup.xml select
 * (4 dynamics evaluations per substep, all forces explicit). */
enum { ICEM_INTEGRATOR_EULER = 0, ICEM_INTEGRATOR_RK4 = 1 };

/* ICEM_DYN_HALFCHEETAH / ICEM_DYN_HUMANOID_STANDUP: tables of the articulated-body model the kernels simulate
 * (the role of the MuJoCo model the reference's GroundTruthModel owns, models/gt_model.py:24-57).  All arrays are
 * HOST pointers, copied by the call.  Bodies are listed parent-before-child, dofs in body order, contact spheres
 * sorted by body.  dof_type: 0 slide, 1 hinge, 2 free-joint translation, 3 free-joint rotation (body-frame angular
 * velocity; the three share dof_qadr = start of the unit quaternion w,x,y,z).  State layout = [qpos(nq), qvel(nv)];
 * the observation the cost reads is state[obs_offset:]. */
typedef struct icem_articulated_model {
  int32_t nb, nq, nv, nu, nc;
  int32_t nsub;                   /* substeps per control step (env frame_skip) */
  int32_t obs_offset;
  float dt;                       /* substep length */
  float gravity, ctrl_limit;
  float contact_stiffness, contact_damping, contact_damping_max, friction_viscous, friction;
  const int32_t* body_parent;     /* [nb], -1 = world */
  const int32_t* body_dof_start;  /* [nb] */
  const int32_t* body_dof_count;  /* [nb] */
  const float* body_pos;          /* [nb*3] frame offset in the parent frame */
  const float* body_mass;         /* [nb] */
  const float* body_com;          /* [nb*3] */
  const float* body_inertia;      /* [nb*6] xx yy zz xy xz yz about the com, body axes */
  const int32_t* dof_body;        /* [nv] */
  const int32_t* dof_type;        /* [nv] */
  const int32_t* dof_qadr;        /* [nv] */
  const int32_t* dof_parent;      /* [nv] previous dof on the path to the root, -1 = none */
  const int32_t* dof_limited;     /* [nv] */
  const int32_t* dof_act;         /* [nv] control index driving this dof, -1 = none */
  const float* dof_axis;          /* [nv*3] unit, body frame */
  const float* dof_anchor;        /* [nv*3] body frame */
  const float* dof_stiffness; const float* dof_damping; const float* dof_armature;   /* [nv] each */
  const float* dof_lo; const float* dof_hi; const float* dof_klim; const float* dof_blim; const float* dof_gear;
  const int32_t* con_body;        /* [nc] */
  const float* con_pos;           /* [nc*3] body frame */
  const float* con_radius;        /* [nc] */
  int32_t integrator;             /* ICEM_INTEGRATOR_*: how a substep advances (q, qvel) */
} icem_articulated_model_t;
int icem_set_articulated_model(icem_planner_t* p, const icem_articulated_model_t* model);

/* ---- MpcICem.beginning_of_rollout (controllers/icem.py:31-43; controllers/mpc.py:69-73) ---------------- */
int icem_begin_rollout(icem_planner_t* p);

/* ---- MpcICem.get_action (controllers/icem.py:106-189): one plan step ------------------------------------
 * state: forward-model start state (GT state without the leading time entry for the articulated models,
 * the observation for ICEM_DYN_DENSE_TANH / ICEM_DYN_MLP), float64 like the reference's arrays.
 * action_out[d]: the executed action (first action of the best trajectory of the last CEM iteration).
 * Returns ICEM_ERR_STATE ("beginning_of_rollout() needs to be called before") if not reset. */
int icem_plan(icem_planner_t* p, const double* state, int32_t state_dim, double* action_out);

/* The two halves of icem_plan, for MANY planner handles driven by one host thread (one independent MPC problem
 * per handle, each on its own CUDA stream): launch every handle's plan step, then collect them -- the device runs
 * the problems side by side.  This is the GPU form of the reference's episode-level parallelism
 * (misc/rollout_utils.py:129-152 spawns one process per evaluation rollout); at the reference's default budget of
 * 97 trajectories per step one problem occupies a few percent of a B200.  Exactly one plan step may be in flight
 * per handle (ICEM_ERR_STATE otherwise). */
int icem_plan_async(icem_planner_t* p, const double* state, int32_t state_dim);
int icem_plan_finish(icem_planner_t* p, double* action_out);

/* One plan step of EVERY problem of a handle created with num_problems = B: states[B][state_dim] in,
 * actions_out[B][act_dim] out, one CUDA graph launch for all of them.  Problem i plans exactly like a single-problem
 * handle created with seed + i.  The getters (icem_get_mean / _std / _elites / _iteration / _costs / _actions) read
 * the problem selected with icem_set_active_problem (default 0). */
int icem_plan_batch(icem_planner_t* p, const double* states, int32_t state_dim, int32_t num_states,
                    double* actions_out);
int icem_num_problems(icem_planner_t* p);
int icem_set_active_problem(icem_planner_t* p, int32_t problem);

/* Same plan step with the start state already resident on the device (no host<->device copy): the state is
 * the one left by the previous icem_plan / icem_advance_state_device.  Asynchronous; pair with icem_sync. */
int icem_plan_device(icem_planner_t* p);
/* state_dev <- f(state_dev, last executed action): closed loop entirely on the device (bench `value`). */
int icem_advance_state_device(icem_planner_t* p);
int icem_sync(icem_planner_t* p);
/* CUDA-event time of the last completed plan step on the planner's stream (ms) and of its rollout kernels */
int icem_last_plan_ms(icem_planner_t* p, float* total_ms, float* rollout_kernels_ms);

/* ---- parity mode: identical Gaussian draws --------------------------------------------------
 * Provide the UNIT normal draws the next icem_plan consumes for CEM iteration `iteration`, in the reference's
 * order (oracle/shims/colorednoise.py): zr[rows][d][K] then zi[rows][d][K] (K = h/2+1), rows = this rank's
 * fresh trajectories followed by its shifted-elite rows; for noise_beta == 0: z[rows][h][d] in zr, zi NULL;
 * for ICEM_PLANNER_CEM_STD: the uniform draws u[rows][h][d] of truncnorm.rvs as SIGNED TAIL PROBABILITIES in zr
 * (w = u for u < 1/2, w = -(1 - u) otherwise: keeps the precision of the small tail in float32), zi NULL;
 * for ICEM_PLANNER_RANDOM: the uniform draw behind every action entry, u[rows][h][d] in zr (entries of one held
 * action repeat), zi NULL: action = low + (high - low) * u like gym's Box.sample.
 * Cleared after the plan step that consumed them. */
int icem_inject_noise(icem_planner_t* p, int32_t iteration, int32_t rows, const float* zr, const float* zi);

/* ---- observable planner state (attributes MpcICem exposes: mean, std, elite_samples) ------------------- */
int icem_get_mean(icem_planner_t* p, float* mean_out /* [h*d] */);
int icem_get_std(icem_planner_t* p, float* std_out /* [h*d] */);
int icem_num_elites(icem_planner_t* p);
int icem_get_elites(icem_planner_t* p, float* actions_out /* [k*h*d] */, float* costs_out /* [k] */,
                    int32_t* idx_out /* [k] */);
/* per-iteration record of the last plan step */
int icem_population_size(icem_planner_t* p, int32_t iteration, int32_t first_step, int32_t* global_out,
                         int32_t* local_out);
int icem_get_iteration(icem_planner_t* p, int32_t iteration, float* mean_out, float* std_out,
                       float* elite_costs_out, int32_t* elite_idx_out);
int icem_get_costs(icem_planner_t* p, int32_t iteration, float* costs_out, int32_t n /* local rows */);
int icem_get_actions(icem_planner_t* p, int32_t iteration, float* actions_out, int32_t n /* local rows */);

/* ---- forward_model.predict / env.step on the device model (controllers/icem.py:186-188) ---------------- */
int icem_sim_step(icem_planner_t* p, const double* state, int32_t state_dim, const double* action,
                  double* next_state, double* obs_out, int32_t obs_dim, double* reward_out);
/* n independent transitions in ONE launch (one CTA each): states[n][state_dim], actions[n][act_dim] ->
 * next_states[n][state_dim].  The env.step of many episodes at once (icem_b200/batched.py). */
int icem_sim_step_batch(icem_planner_t* p, int32_t n, const double* states, int32_t state_dim, const double* actions,
                        double* next_states);
int icem_state_dim(icem_planner_t* p);
int icem_observe(icem_planner_t* p, const double* state, int32_t state_dim, double* obs_out, int32_t obs_dim);

/* ---- single operators with HOST buffers (kernel-level parity tests; each is one kernel of the plan step) - */
/* controllers/icem.py:61-82: actions[n][h][d] = clip(colored(z) * std + mean, low, high) */
int icem_op_sample(icem_planner_t* p, int32_t n, const float* zr, const float* zi, const float* mean,
                   const float* std, float* actions_out);
/* controllers/mpc.py:56-67 + controllers/abstract_controller.py:74-91: costs[n] for given action sequences */
int icem_op_rollout_cost(icem_planner_t* p, int32_t n, const double* state, int32_t state_dim,
                         const float* actions, float* costs_out);
/* The observations a FEW given action sequences visit (what the reference keeps for every rollout in its
 * RolloutBuffer, misc/rolloutbuffer.py:10-54, models/abstract_models.py:28-53): obs_out[n][h+1][obs_dim], entry t =
 * observation before action t, entry h = the final predicted observation.  Used lazily for `elite_samples`
 * (controllers/icem.py:201), `visualize_plan` (controllers/abstract_controller.py:93-128); the planner itself never
 * materialises observations.  One small launch per step: meant for k elites, not for populations. */
int icem_op_rollout_observations(icem_planner_t* p, int32_t n, const double* state, int32_t state_dim,
                                 const float* actions, int32_t obs_dim, double* obs_out);
/* controllers/icem.py:199: k smallest by ascending (cost, index); NaN sorts last */
int icem_op_topk(icem_planner_t* p, int32_t n, const float* costs, int32_t k, int32_t* idx_out, float* costs_out);

/* ---- multi-GPU: one NCCL all-gather of per-shard elites per CEM iteration ------------------------------ */
#define ICEM_UNIQUE_ID_BYTES 128
int icem_comm_get_unique_id(char id_out[ICEM_UNIQUE_ID_BYTES]);
int icem_comm_init(icem_planner_t* p, const char id[ICEM_UNIQUE_ID_BYTES]);

/* ---- trainer of the dense MLP forward model ------------------------------------------------------------------ */
/* Replaces the hook `forward_model.train(rollout_buffer)` (icem/main.py:209-210) for the MLP model the tensor-core
 * rollout consumes: next_obs = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3, fitted to transitions
 * (inputs[n][in] = [obs, act], targets[n][out] = next_obs - obs) by minibatch Adam on the mean squared error, fp32,
 * on the device.  The reference ships no trainable model (icem/models/__init__.py:5-8), so the semantics are those of
 * torch.nn.MSELoss (mean) + torch.optim.Adam, which is also the oracle (oracle/mlp_train_torch.py).
 * weights[l] is row-major [out_l][in_l] (torch.nn.Linear layout), l = 0..2; the handle owns device copies. */
typedef struct icem_mlp_trainer icem_mlp_trainer_t;
int icem_mlp_trainer_create(int32_t device, int32_t in_dim, int32_t hidden, int32_t out_dim, icem_mlp_trainer_t** out);
int icem_mlp_trainer_destroy(icem_mlp_trainer_t* t);
/* reset_optimizer != 0 also zeroes the Adam moments and the step count */
int icem_mlp_trainer_set_weights(icem_mlp_trainer_t* t, const float* const* weights, const float* const* biases,
                                 int32_t reset_optimizer);
int icem_mlp_trainer_get_weights(icem_mlp_trainer_t* t, float* const* weights, float* const* biases);
/* the whole data set goes to the device once (the reference keeps every rollout of a run in its RolloutBuffer) */
int icem_mlp_trainer_set_data(icem_mlp_trainer_t* t, int64_t n, const float* inputs, const float* targets);
/* n_steps Adam steps; step s trains on rows indices[s * batch .. (s + 1) * batch) of the data set;
 * losses_out[n_steps] (may be NULL) = each step's minibatch loss BEFORE its update */
int icem_mlp_trainer_fit(icem_mlp_trainer_t* t, int32_t n_steps, int32_t batch, const int32_t* indices, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float* losses_out);
/* outputs[n][out] = the network's delta prediction for inputs[n][in] (fp32 forward pass of the trainer) */
int icem_mlp_trainer_predict(icem_mlp_trainer_t* t, int32_t n, const float* inputs, float* outputs);

/* ---- bench helpers ------------------------------------------------------------------------------------- */
/* Run `steps` device-resident closed-loop plan steps (plan_device + advance_state_device) after `warmup`,
 * flushing L2 between steps when flush_l2 != 0; CUDA events on the planner's stream.
 * total_ms: timed region; rollout_ms: summed duration of the fused sample->rollout->cost kernels in it;
 * rollout_launches: how many of them. */
int icem_bench_device(icem_planner_t* p, int32_t steps, int32_t warmup, int32_t flush_l2, float* total_ms,
                      float* rollout_ms, int32_t* rollout_launches);

/* Time ONE kernel of the plan step in isolation on `n` trajectories (production Philox noise), `reps` launches after
 * 3 warm-ups, CUDA events on the planner's stream; *ms_avg = average launch duration.
 *   op 0: sampler only (colored-noise synthesis + clip -> actions[n][h*d] in HBM; controllers/icem.py:61-82)
 *   op 1: fused sample -> rollout -> cost (one CEM iteration's dominant kernel)
 *   op 2: elite selection + refit on n costs (controllers/icem.py:194-211)
 *   op 3: rollout + cost of GIVEN action tiles (the tensor-core kernel for ICEM_DYN_MLP)
 *   op 4: sharded planners only (every rank must call it): one NCCL all-gather of the per-rank elite records + the
 *         merge / refit kernel, i.e. the exchange step of a CEM iteration; n is ignored
 * flush_l2 != 0 writes a 256 MiB buffer between launches. */
int icem_bench_op(icem_planner_t* p, int32_t op, int32_t n, int32_t reps, int32_t flush_l2, float* ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* ICEM_B200_H */

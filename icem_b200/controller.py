"""MpcICemB200 -- drop-in for the reference's `MpcICem` (icem/controllers/icem.py:16-247).

Same constructor keywords (settings JSON `controller_params`), same plugin methods
(`beginning_of_rollout` / `get_action` / `end_of_rollout`), same observable attributes (`mean`, `std`,
`elite_samples`, `model_evals_per_timestep`) and the same error behaviour; the CEM loop itself runs as CUDA kernels
behind the C ABI (include/icem_b200.h).  There is NO CPU path: a forward model without `cuda_spec()` or a missing
library / device raises.
"""
from warnings import warn

import numpy as np

from . import api
from .planner import IcemError, Planner, PlannerSettings

_ref = api.reference_bases()
_Bases = (_ref["mbc"], _ref["sc"]) if _ref else (api.ModelBasedController, api.StatefulController)

_ENV_COSTS = {"HalfCheetahMaybeWithPosition": "halfcheetah", "HumanoidStandup": "humanoid_standup"}


def _get_logger(name):
    try:
        import allogger
        return allogger.get_logger(scope=name, default_outputs=["tensorboard"])
    except ImportError:
        return None


class MpcICemB200(*_Bases):
    def __init__(self, *, action_sampler_params, horizon, num_simulated_trajectories, factor_decrease_num=1,
                 verbose=False, seed=None, device=None, world_size=None, rank=None, num_problems=1, **kwargs):
        super().__init__(**kwargs)     # ModelBasedController: forward_model, env, cost_along_trajectory, ...
        # controllers/mpc.py:22-36
        self.horizon = horizon
        self.num_sim_traj = num_simulated_trajectories
        self.factor_decrease_num = factor_decrease_num
        if num_simulated_trajectories < 2:
            raise ValueError("At least two trajectories needed!")
        self.verbose = verbose
        self.model_dir = None                 # controllers/mpc.py:34
        self.forward_model_state = None
        self._parse_action_sampler_params(**action_sampler_params)
        self._check_validity_parameters()
        self.logger = _get_logger(self.__class__.__name__)
        self.was_reset = False
        if getattr(self, "use_env_reward_as_cost", False) and not getattr(self.env, "reward_is_negative_cost", False):
            # abstract_controller.py:75-76 scores -rewards of the simulated steps.  The device kernels evaluate
            # env.cost_fn; that is the same number exactly when the env's reward is defined as -cost_fn (true for
            # the stand-in envs of icem_b200.envs, which declare it), not for gym's own reward terms.
            raise NotImplementedError("use_env_reward_as_cost needs an env whose reward is -cost_fn "
                                      "(env.reward_is_negative_cost); gym's reward terms have no device kernel")

        fm = self.forward_model
        if not getattr(fm, "is_cuda_model", False) or not hasattr(fm, "cuda_spec"):
            raise TypeError(f"MpcICemB200 needs a CUDA-capable forward model (got {type(fm).__name__}); "
                            "use forward_model 'CudaGroundTruthModel' / 'CudaDenseTanhModel'. There is no CPU fallback.")
        spec = fm.cuda_spec()
        cost, penalise, *cost_extra = self._cost_spec()
        if device is None or world_size is None or rank is None:
            from .distributed import default_placement
            dev_, ws_, rk_ = default_placement()
            device = dev_ if device is None else device
            world_size = ws_ if world_size is None else world_size
            rank = rk_ if rank is None else rank
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))   # follows np.random.seed(Seeding.SEED), misc/seeding.py:18
        if world_size > 1:
            from .distributed import agree_on_seed
            seed = agree_on_seed(seed)                       # one Philox stream for all shards: rank 0's
        self._world_size = world_size
        self._planner = Planner(PlannerSettings(
            horizon=horizon, num_simulated_trajectories=num_simulated_trajectories,
            action_low=self.env.action_space.low, action_high=self.env.action_space.high,
            dynamics=spec["dynamics"], cost=cost, obs_dim=spec["obs_dim"], penalise_flipping=penalise,
            cost_along_trajectory=self.cost_along_trajectory, alpha=self.alpha, elites_size=self.elites_size,
            opt_iterations=self.opt_iter, init_std=self.init_std, seed=seed, device=device, world_size=world_size,
            rank=rank, num_problems=num_problems, cost_params=cost_extra[0] if cost_extra else None,
            articulated_model=spec.get("articulated_model"),
            obs_offset=0 if spec["dynamics"] == "articulated" else None, **self._sampler_settings()))
        if spec.get("dense") is not None:
            self._planner.set_dense_model(*spec["dense"])
        if spec.get("mlp") is not None:
            self._planner.set_mlp_model(*spec["mlp"])
        self._model_version = getattr(fm, "version", 0)
        if world_size > 1:
            from .distributed import init_planner_comm
            init_planner_comm(self._planner)
        self.expected_cost = None

    def _sampler_settings(self):
        return dict(planner="icem", factor_decrease_num=self.factor_decrease_num,
                    use_mean_actions=self.use_mean_actions, keep_previous_elites=self.keep_previous_elites,
                    shift_elites_over_time=self.shift_elites_over_time,
                    fraction_elites_reused=self.fraction_elites_reused, noise_beta=self.noise_beta)

    def _evals_per_timestep(self):      # controllers/icem.py:38-41
        return sum([max(self.elites_size * 2, int(self.num_sim_traj / (self.factor_decrease_num ** i)))
                    for i in range(0, self.opt_iter)]) * self.horizon

    _banner = "iCEM"

    # ---- settings (controllers/icem.py:213-247) ------------------------------------------------
    def _parse_action_sampler_params(self, *, alpha, elites_size, opt_iterations, init_std, use_mean_actions,
                                     keep_previous_elites, shift_elites_over_time, fraction_elites_reused,
                                     noise_beta=1):
        self.alpha = alpha
        self.elites_size = elites_size
        self.opt_iter = opt_iterations
        self.init_std = init_std
        self.use_mean_actions = use_mean_actions
        self.keep_previous_elites = keep_previous_elites
        self.shift_elites_over_time = shift_elites_over_time
        self.fraction_elites_reused = fraction_elites_reused
        self.noise_beta = noise_beta

    def _check_validity_parameters(self):
        self.num_elites = min(self.elites_size, self.num_sim_traj // 2)
        if self.num_elites < 2:
            warn('Number of trajectories is too low for given elites_frac. Setting num_elites to 2.')
            self.num_elites = 2
        space = self.env.action_space
        if type(space).__name__ == "Discrete":
            raise NotImplementedError("CEM ERROR: Implement categorical distribution for discrete envs.")
        elif type(space).__name__ == "Box":
            self.dim_samples = (self.horizon, space.shape[0])
        else:
            raise NotImplementedError

    def _cost_spec(self):
        env = self.env
        if hasattr(env, "cuda_cost_spec"):
            return env.cuda_cost_spec()
        for cls in type(env).__mro__:
            if cls.__name__ in _ENV_COSTS:
                return _ENV_COSTS[cls.__name__], bool(getattr(env, "penalise_flipping", False))
        raise NotImplementedError(f"no CUDA cost function for env {type(env).__name__}")

    # ---- plugin API ----------------------------------------------------------------------------
    def beginning_of_rollout(self, *, observation, state=None, mode):
        # controllers/mpc.py:69-73
        if state is not None and isinstance(self.forward_model, (api.AbstractGroundTruthModel,) + (
                (_ref["gt"],) if _ref else ())):
            self.forward_model_state = state
        else:
            self.forward_model_state = self.forward_model.reset(observation)
        if getattr(self.forward_model, "version", 0) != self._model_version:
            # the model was re-trained since the last rollout (main.py:209-210): the planner takes the new weights
            spec = self.forward_model.cuda_spec()
            if spec.get("mlp") is not None:
                self._planner.set_mlp_model(*spec["mlp"])
            self._model_version = self.forward_model.version
        self._planner.begin_rollout()
        self.was_reset = True
        self._elite_cache = None
        self._steps_since_reset = 0
        # controllers/icem.py:38-43 / controllers/mpc.py:172-175
        self.model_evals_per_timestep = self._evals_per_timestep()
        print(f"{self._banner} using {self.model_evals_per_timestep} evaluations per step "
              f"and {self.model_evals_per_timestep / self.horizon} trajectories per step")

    def end_of_rollout(self, total_time, total_return, mode):
        pass

    def get_action(self, obs, state, mode="train"):
        self.begin_get_action(obs, state, mode)
        return self.finish_get_action()

    # get_action in two halves: many controllers (independent episodes) launch their plan steps first and collect
    # them afterwards, so the device works on all of them at once (icem_b200/batched.py)
    def begin_get_action(self, obs, state, mode="train"):
        if not self.was_reset:
            raise AttributeError("beginning_of_rollout() needs to be called before")
        if self.verbose:
            print(f"-------------------- {self.mean[0][0:6]}")
            if mode != "expert":       # icem.py:114-115
                self.check_model_consistency()
        self.forward_model_state = self.forward_model.got_actual_observation_and_env_state(
            observation=obs, env_state=state, model_state=self.forward_model_state)
        start = self.forward_model.start_state(obs, self.forward_model_state)
        self._last_start = np.array(start, dtype=np.float64)
        self._obs_dim = int(np.asarray(obs).shape[-1])
        self._pending = (obs, state, self._steps_since_reset == 0)
        if self._world_size > 1 and self._steps_since_reset == 0:
            from .distributed import assert_same_on_all_ranks
            assert_same_on_all_ranks(self._last_start)       # every shard must roll out from the same state
        try:
            self._planner.plan_async(start)
        except IcemError as e:
            if "beginning_of_rollout" in str(e):
                raise AttributeError(str(e))
            raise

    def plan_batch(self, start_states, obs_dim=None):
        """One plan step of EVERY problem of a controller built with num_problems = B (icem_plan_batch): the same
        bookkeeping as get_action -- step counter, elite cache, expected cost of the active problem."""
        if not self.was_reset:
            raise AttributeError("beginning_of_rollout() needs to be called before")
        start_states = np.asarray(start_states, np.float64)
        actions = self._planner.plan_batch(start_states)
        self._last_start = np.array(start_states[self._planner.active_problem], dtype=np.float64)
        if obs_dim is not None:
            self._obs_dim = int(obs_dim)
        self._steps_since_reset += 1
        self._elite_cache = None
        if self.logger is not None:
            rec = self._planner.iteration_record(self.opt_iter - 1)
            self.expected_cost = float(rec["elite_costs"][0])
            self.logger.log(self._logged_cost(self.expected_cost), key="Expected_trajectory_cost")
        return actions

    def finish_get_action(self):
        obs, state, first = self._pending
        executed_action = self._planner.plan_finish()
        self._steps_since_reset += 1
        self._elite_cache = None
        if self.verbose:
            self._print_iterations(first)
        rec = self._planner.iteration_record(self.opt_iter - 1) if (self.logger is not None or self.verbose) else None
        if rec is not None:
            self.expected_cost = float(rec["elite_costs"][0])
            if self.logger is not None:
                self.logger.log(self._logged_cost(self.expected_cost), key="Expected_trajectory_cost")   # icem.py:177
        if self.do_visualize_plan:      # icem.py:179-183: the best trajectory of the last iteration = elite 0
            best = self.elite_samples[0]
            self.visualize_plan(obs=best["observations"], state=self.forward_model_state, acts=best["actions"])
        # for stateful models, actually simulate step (icem.py:185-188); the result is overwritten by the next
        # call whenever the env state is supplied, so it is only evaluated when it will be used
        if self.forward_model_state is not None and state is None:
            _, self.forward_model_state, _ = self.forward_model.predict(
                observations=obs, states=self.forward_model_state, actions=executed_action)
        return executed_action

    def _logged_cost(self, cost):
        return cost          # MpcICem logs min(costs) as it is (icem.py:177)

    def _print_iterations(self, first):
        p = self._planner
        for i in range(self.opt_iter):
            _, n_local = p.population_size(i, first)
            costs = p.costs(i, n_local).astype(np.float64)
            rec = p.iteration_record(i)

            def display_cost(cost):
                return cost / self.horizon if self.cost_along_trajectory == "sum" else cost
            print(self._iter_format
                  .format(i, n_local, display_cost(min(np.amin(costs), rec["elite_costs"][0])),
                          display_cost(np.mean(costs)), display_cost(np.amax(costs)), rec["elite_idx"][0:6]))

    _iter_format = 'iter {}:{} --- best cost: {:.2f} --- mean: {:.2f} --- worst: {:.2f}  elites: {}...'

    # ---- observable attributes -----------------------------------------------------------------
    @property
    def mean(self):
        return self._planner.mean().astype(np.float64)

    @property
    def std(self):
        return self._planner.std().astype(np.float64)

    @property
    def elite_samples(self):
        """RolloutBuffer of the k elites of the last CEM iteration, best first (icem.py:201).  The planner keeps
        only the elites' action sequences on the device; the remaining fields the reference's rollouts carry
        (models/abstract_models.py:28-29: observations, next_observations, rewards) are materialised lazily, on
        first access after a plan step, by re-running just these k sequences on the device model
        (icem_op_rollout_observations) -- SURVEY 8f-2."""
        if getattr(self, "_steps_since_reset", 0) == 0:
            return (_ref["buffer"]() if _ref else api.EliteBuffer())
        if self._elite_cache is None:
            acts, costs, _ = self._planner.elites()
            pad = int(getattr(self.env, "obs_pad", 0))     # observation entries the device model does not carry
            if hasattr(self.env, "observation_from_state"):
                # the env's observation is a function of the device state, not a slice of it (Reacher: cos / sin /
                # fingertip - target, environments/mujoco.py:346-368)
                states = self._planner.rollout_observations(self._last_start, acts, self._last_start.shape[-1])
                obs = self.env.observation_from_state(states)
            else:
                obs = self._planner.rollout_observations(self._last_start, acts, self._obs_dim - pad)
            if pad:
                obs = np.concatenate([obs, np.zeros(obs.shape[:-1] + (pad,))], axis=-1)
            rollouts = []
            for a, o in zip(acts.astype(np.float64), obs):
                step_costs = np.asarray(self.cost_fn(o[:-1], a, o[1:]), dtype=np.float64)
                fields = dict(observations=o[:-1], next_observations=o[1:], actions=a, rewards=-step_costs)
                rollouts.append(_ref["rollout"].from_dict(**fields) if _ref else api.EliteRollout(**fields))
            self._elite_cache = _ref["buffer"](rollouts=rollouts) if _ref else api.EliteBuffer(rollouts)
            self._elite_costs = costs.astype(np.float64)
        return self._elite_cache

    def check_model_consistency(self):
        """controllers/mpc.py:39-47: warn when the model state and the real env state have drifted apart."""
        env = self.env
        if self.forward_model_state is None or not hasattr(env, "compute_state_difference"):
            return
        diff = env.compute_state_difference(self.forward_model_state, env.get_GT_state())
        if diff > 1e-5:
            print(f"Warning: internal GT model and actual env are not in sync: Difference: {diff}")
            print("env state:", env.get_GT_state())
            print("model_state:", self.forward_model_state)

    def close(self):
        self._planner.close()



class MpcCemStdB200(MpcICemB200):
    """Drop-in for the reference's vanilla CEM `MpcCemStd` (icem/controllers/mpc.py:142-327): truncated-normal
    sampling inside the action bounds, constant population, no elite reuse / mean injection, `_update_bounds`
    (optionally "like Levine"), `execute_best_elite` and `shift_means` switches -- on the same kernels
    (SURVEY 8f-1).  Same constructor keywords as the reference class."""
    _banner = "CEM-Standard"
    _iter_format = 'iter {}:{} --- best cost: {:.2f} --- mean: {:.2f} --- worst: {:.2f}  elites: {}...'

    def _parse_action_sampler_params(self, *, alpha, elites_size, opt_iterations, init_std, shift_means,
                                     execute_best_elite, bounds_like_levine):
        self.alpha = alpha
        self.elites_size = elites_size
        self.opt_iter = opt_iterations
        self.init_std = init_std
        self.execute_best_elite = execute_best_elite
        self.like_levine = bounds_like_levine
        self.shift_means = shift_means

    def _sampler_settings(self):
        return dict(planner="cem_std", factor_decrease_num=1.0, use_mean_actions=False, keep_previous_elites=False,
                    shift_elites_over_time=False, fraction_elites_reused=0.0, noise_beta=0.0,
                    execute_best_elite=self.execute_best_elite, shift_means=self.shift_means,
                    bounds_like_levine=self.like_levine)

    def _evals_per_timestep(self):      # controllers/mpc.py:172
        return self.num_sim_traj * self.opt_iter * self.horizon

    def _logged_cost(self, cost):       # MpcCemStd logs display_cost(min(costs)): per step for "sum" (mpc.py:217-218,251)
        return cost / self.horizon if self.cost_along_trajectory == "sum" else cost


class MpcRandomB200(MpcICemB200):
    """Drop-in for the reference's random-shooting `MpcRandom` (icem/controllers/mpc.py:86-138): one population of
    uniformly drawn, piecewise-constant action sequences per step (a drawn action is held for
    `action_change_frequency` further `sample()` calls, and the calls run on across rows and across plan steps),
    execute the first action of the cheapest trajectory.  Same constructor keywords.  As shipped the
    reference class cannot be instantiated (it inherits the abstract `StatefulController.end_of_rollout` and never
    defines it); this class provides the no-op the other controllers have."""
    _banner = "MPC-Random"

    def _parse_action_sampler_params(self, *, action_change_frequency):
        self.action_change_frequency = action_change_frequency
        assert self.action_change_frequency < self.horizon      # mpc.py:92
        self.alpha, self.elites_size, self.opt_iter, self.init_std = 0.0, 2, 1, 0.5

    def _sampler_settings(self):
        return dict(planner="random", factor_decrease_num=1.0, use_mean_actions=False, keep_previous_elites=False,
                    shift_elites_over_time=False, fraction_elites_reused=0.0, noise_beta=0.0,
                    action_change_frequency=self.action_change_frequency)

    def _evals_per_timestep(self):
        return self.num_sim_traj * self.horizon

"""Drive the UNMODIFIED reference controller (`/root/reference/icem/controllers/icem.py`) with a
batched NumPy forward model and record its intermediates.

TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (this container).  Used by
`oracle/make_golden.py` (fixtures), `tests/test_oracle.py` (live pin of oracle/icem_np.py) and
`bench.py --impl reference` / `cpu_baseline` when the reference is present.
"""
import numpy as np

from oracle import ref_loader


def build_reference_controller(model, cost_name, cfg_kwargs, action_low, action_high,
                               penalise_flipping=True, controller="MpcICem"):
    """model: an oracle/dynamics_np.py model.  cost_name: "halfcheetah" | "humanoid_standup".
    cfg_kwargs: the `controller_params` dict of the settings JSON.  Returns (controller, recorder)."""
    ref = ref_loader.load_reference()
    import environments.mujoco as ref_mujoco   # the reference's own cost functions
    from gym import spaces

    class _Env(ref.abstract_environments.GroundTruthSupportEnv):
        """Stand-in env: only the attributes the controller reads."""

        def __init__(self, *, name):
            super().__init__(name=name)
            self.action_space = spaces.Box(np.asarray(action_low, np.float32),
                                           np.asarray(action_high, np.float32))
            self.penalise_flipping = penalise_flipping

        def cost_fn(self, observation, action, next_obs):
            if cost_name == "halfcheetah":
                return ref_mujoco.HalfCheetahMaybeWithPosition.cost_fn(self, observation, action, next_obs)
            if cost_name == "humanoid_standup":
                return ref_mujoco.HumanoidStandup.cost_fn(self, observation, action, next_obs)
            raise KeyError(cost_name)

        def set_GT_state(self, state):
            raise NotImplementedError

        def get_GT_state(self):
            raise NotImplementedError

        def set_state_from_observation(self, observation):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def reset(self):
            raise NotImplementedError

    class _Model(ref.abstract_models.ForwardModelWithDefaults):
        """Batched model through the reference's default time-major rollout
        (models/abstract_models.py:17-53); rewards must be [p,1]."""

        def predict(self, *, observations, states, actions):
            nxt = model.step(observations, actions)
            return nxt, states, np.zeros((len(observations), 1))

        def train(self, buffer):
            pass

        def save(self, path):
            pass

        def load(self, path):
            pass

    env = _Env(name="standin")
    cls = ref.icem.MpcICem if controller == "MpcICem" else getattr(ref.mpc, controller)
    if controller == "MpcRandom":
        # As shipped, MpcRandom inherits the abstract StatefulController.end_of_rollout and cannot be instantiated;
        # add that one no-op (what MpcICem / MpcCemStd define) and nothing else.  Its sampler parameters are read by
        # attribute (mpc.py:91), as smart_settings objects allow.
        import types

        class _Runnable(cls):
            def end_of_rollout(self, total_time, total_return, mode):
                pass
        cfg_kwargs = dict(cfg_kwargs)
        cfg_kwargs["action_sampler_params"] = types.SimpleNamespace(**cfg_kwargs["action_sampler_params"])
        ctrl = _Runnable(env=env, forward_model=_Model(env=env), **cfg_kwargs)
        rec = {"iterations": []}
        orig_cost = ctrl.trajectory_cost_fn

        def recording_cost(cost_fn, rollout_buffer):
            costs = np.array(orig_cost(cost_fn, rollout_buffer))
            rec["iterations"].append(dict(population=len(costs), costs=costs.copy(),
                                          elite_idx=np.array([int(np.argmin(costs))]),
                                          actions=rollout_buffer.as_array("actions").copy()))
            return costs
        ctrl.trajectory_cost_fn = recording_cost
        return ctrl, rec
    ctrl = cls(env=env, forward_model=_Model(env=env), **cfg_kwargs)

    rec = {"iterations": []}
    orig_update = ctrl.update_distributions

    def recording_update(sampled_trajectories, costs):
        costs = np.array(costs)
        orig_update(sampled_trajectories, costs)
        # elite order as the reference's own (unstable) argsort produced it
        rec["iterations"].append(dict(
            population=len(costs), costs=costs.copy(),
            elite_idx=np.array(costs).argsort()[: ctrl.num_elites],
            mean=ctrl.mean.copy(), std=ctrl.std.copy(),
            elite_actions=ctrl.elite_samples.as_array("actions").copy()))

    ctrl.update_distributions = recording_update
    return ctrl, rec


def run_reference_episode(model, cost_name, cfg_kwargs, action_low, action_high, start_obs,
                          seed, num_steps, penalise_flipping=True, controller="MpcICem"):
    """np.random.seed(seed); beginning_of_rollout; `num_steps` x (get_action; obs <- model.step)."""
    if controller == "MpcRandom":
        np.random.seed(seed)          # the constructor already draws (random.py:8, mpc.py:90)
    ctrl, rec = build_reference_controller(model, cost_name, cfg_kwargs, action_low, action_high,
                                           penalise_flipping, controller)
    if controller != "MpcRandom":
        np.random.seed(seed)
    obs = np.asarray(start_obs, dtype=np.float64).copy()
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    steps = []
    for _ in range(num_steps):
        rec["iterations"] = []
        action = np.array(ctrl.get_action(obs, None), dtype=np.float64)
        steps.append(dict(action=action, iterations=rec["iterations"],
                          mean_after_shift=ctrl.mean.copy() if hasattr(ctrl, "mean") else None,
                          std_after_reset=ctrl.std.copy() if hasattr(ctrl, "std") else None,
                          start_obs=obs.copy()))
        obs = model.step(obs[None], action[None])[0]
    return steps, float(np.random.randn())

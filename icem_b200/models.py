"""CUDA-capable forward models for the B200 controller.

The reference hands the controller a `ForwardModel` object built from a registry string
(icem/main.py:105-109, icem/models/__init__.py:5-26).  These classes are what a settings file names in
`"forward_model"` for the B200 path (`launch.py` registers them in `models_dict`); they do not simulate on the
host -- they describe the device model (`cuda_spec`) and the controller keeps every rollout on the GPU.
"""
import numpy as np

from . import api

_ref = api.reference_bases()
_FMBase = _ref["fm"] if _ref else api.ForwardModel
_GTBase = _ref["gt"] if _ref else api.AbstractGroundTruthModel


class CudaDenseTanhModel(_FMBase):
    """obs' = tanh(W_o obs + W_a act + b): the dense single-layer batched model (the path
    `ForwardModelWithDefaults.predict_n_steps`, icem/models/abstract_models.py:31-53, takes)."""
    is_cuda_model = True

    def __init__(self, *, env, w_obs, w_act, bias=None, **kwargs):
        super().__init__(env=env, **kwargs)
        self.w_obs = np.asarray(w_obs, np.float32)
        self.w_act = np.asarray(w_act, np.float32)
        self.bias = None if bias is None else np.asarray(bias, np.float32)
        self.is_trained = True

    def cuda_spec(self):
        return dict(dynamics="dense_tanh", dense=(self.w_obs, self.w_act, self.bias), obs_dim=self.w_obs.shape[0])

    def start_state(self, observation, model_state):
        return np.asarray(observation, np.float64)

    # ForwardModel API -------------------------------------------------------------------------
    def train(self, buffer):
        pass

    def reset(self, observation):
        return None

    def got_actual_observation_and_env_state(self, *, observation, env_state=None, model_state=None):
        return None

    def predict(self, *, observations, states, actions):
        # float64 host evaluation of ONE transition for API completeness (post-plan bookkeeping,
        # icem/controllers/icem.py:186-188, is skipped for state-less models because reset() returns None)
        o = np.asarray(observations, np.float64)
        nxt = np.tanh(o @ self.w_obs.T.astype(np.float64) + np.asarray(actions, np.float64) @ self.w_act.T.astype(
            np.float64) + (0 if self.bias is None else self.bias.astype(np.float64)))
        return nxt, states, np.zeros(o.shape[:-1] + (1,))

    def predict_n_steps(self, *, start_observations, start_states, policy, horizon):
        raise NotImplementedError("CudaDenseTanhModel rollouts run inside MpcICemB200 on the device")

    def rollout_generator(self, *a, **k):
        raise NotImplementedError

    def rollout_field_names(self):
        return "observations", "next_observations", "actions", "rewards"

    def save(self, path):
        pass

    def load(self, path):
        pass


class CudaMlpModel(CudaDenseTanhModel):
    """Dense MLP forward model rolled out on the tensor cores (csrc/mlp_rollout.cuh):
    obs' = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3.  The reference ships no learned model
    (icem/models/__init__.py:5-8); this is the hook `forward_model_from_string` would resolve for one."""

    def __init__(self, *, env, weights=None, biases=None, hidden=256, train_params=None, init_seed=0, device=0,
                 **kwargs):
        _FMBase.__init__(self, env=env, **kwargs)
        if weights is None:       # untrained model: random initialisation, to be fitted by train()
            from .workloads import mlp_model_weights
            obs_dim, act_dim = env.observation_space.shape[0], env.action_space.shape[0]
            weights, biases = mlp_model_weights(obs_dim, act_dim, hidden, init_seed)
        self.weights = [np.asarray(w, np.float32) for w in weights]
        self.biases = [np.asarray(b, np.float32) for b in biases]
        self.is_trained = True
        # forward_model.train(rollout_buffer) (icem/main.py:209-210): Adam on the device, see trainer.py
        self.train_params = dict(epochs=20, batch_size=256, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                                 seed=0)
        self.train_params.update(train_params or {})
        self.device = device
        self.version = 0            # bumped by train(); the controller re-uploads the weights when it changed
        self.train_losses = None
        self._trainer = None

    def train(self, buffer):
        """Fit the model to every transition of `buffer` (the reference passes its whole RolloutBuffer each training
        iteration, main.py:202-210) and keep the Adam moments across calls."""
        from .trainer import epoch_indices, transitions_from_buffer
        x, t = transitions_from_buffer(buffer)
        tp = self.train_params
        self._device_net().set_data(x, t)
        idx = epoch_indices(x.shape[0], int(tp["batch_size"]), int(tp["epochs"]), int(tp["seed"]) + self.version)
        self.train_losses = self._trainer.fit(idx, lr=tp["lr"], betas=tp["betas"], eps=tp["eps"],
                                              weight_decay=tp["weight_decay"])
        self.weights, self.biases = self._trainer.get_weights()
        self.version += 1
        k = max(1, len(self.train_losses) // 10)
        print(f"CudaMlpModel.train: {x.shape[0]} transitions, {len(self.train_losses)} Adam steps on the device, "
              f"loss {float(np.mean(self.train_losses[:k])):.3e} -> {float(np.mean(self.train_losses[-k:])):.3e}")
        return self.train_losses

    # CheckpointManager.store_forward_model / load_forward_model (icem/misc/initialization.py:146-162)
    def save(self, path):
        np.savez(path if str(path).endswith(".npz") else str(path) + ".npz", version=self.version,
                 **{f"w{l}": w for l, w in enumerate(self.weights)}, **{f"b{l}": b for l, b in enumerate(self.biases)})

    def load(self, path):
        import os
        f = path if str(path).endswith(".npz") else str(path) + ".npz"
        if not os.path.exists(f):
            raise FileNotFoundError(f)
        d = np.load(f)
        self.weights = [np.asarray(d[f"w{l}"], np.float32) for l in range(3)]
        self.biases = [np.asarray(d[f"b{l}"], np.float32) for l in range(3)]
        self.version = int(d["version"]) + 1          # the controller re-uploads
        if self._trainer is not None:
            self._trainer.set_weights(self.weights, self.biases)

    def cuda_spec(self):
        return dict(dynamics="mlp", dense=None, mlp=(self.weights, self.biases), obs_dim=self.weights[-1].shape[0])

    def __getstate__(self):          # the device handle does not travel (copy / pickle): it is re-created on demand
        d = dict(self.__dict__)
        d["_trainer"] = None
        return d

    def _device_net(self):
        if self._trainer is None:
            from .trainer import MlpTrainer
            self._trainer = MlpTrainer(self.weights[0].shape[1], self.weights[0].shape[0], self.weights[2].shape[0],
                                       device=self.device)
            self._trainer.set_weights(self.weights, self.biases)
        return self._trainer

    def predict(self, *, observations, states, actions):
        """One transition (or a batch) through the fp32 network ON THE DEVICE (icem_mlp_trainer_predict); the planner
        itself rolls the model out on the tensor cores and never calls this."""
        o = np.asarray(observations, np.float64)
        x = np.concatenate([o, np.asarray(actions, np.float64)], axis=-1)
        delta = self._device_net().predict(x.reshape(-1, x.shape[-1])).astype(np.float64).reshape(o.shape)
        return o + delta, states, np.zeros(o.shape[:-1] + (1,))


class CudaGroundTruthModel(_GTBase):
    """Ground-truth model of a device-simulated stand-in env (envs.py): the role of
    `ParallelGroundTruthModel` (icem/models/gt_par_model.py:17-100) with the worker pool replaced by the GPU."""
    is_cuda_model = True

    def __init__(self, *, env, num_parallel=None, **kwargs):
        super().__init__(env=env, **kwargs)
        if not hasattr(env, "cuda_dynamics"):
            raise NotImplementedError("Environment does not support the CUDA ground truth forward model")
        self.is_trained = True

    def cuda_spec(self):
        model = self.env.cuda_articulated_model() if hasattr(self.env, "cuda_articulated_model") else None
        return dict(dynamics=self.env.cuda_dynamics, dense=None, obs_dim=self.env.observation_space.shape[0],
                    articulated_model=model)

    def start_state(self, observation, model_state):
        if model_state is None:
            raise ValueError("CudaGroundTruthModel needs the env state (rollout_params.use_env_states: true)")
        return np.asarray(model_state, np.float64)[1:]     # [time, qpos, qvel] -> (qpos, qvel)

    def close(self):
        pass

    def train(self, buffer):
        pass

    def set_state(self, state):
        raise NotImplementedError

    def get_state(self, observation):
        # The reference's GroundTruthModel (models/gt_model.py:45-55) rebuilds the simulator state from an observation
        # through env.set_state_from_observation; the device-simulated stand-ins cannot (HalfCheetah's observation
        # drops the x position, HumanoidStandup's is a function of the state, not the state).
        if hasattr(self.env, "state_from_observation"):
            return self.env.state_from_observation(observation)
        raise ValueError("CudaGroundTruthModel needs the env state (rollout_params.use_env_states: true): this env "
                         "cannot rebuild its state from an observation")

    def reset(self, observation):
        return self.get_state(observation)

    def got_actual_observation_and_env_state(self, *, observation, env_state=None, model_state=None):
        # icem/models/gt_par_model.py:60-64
        return self.reset(observation) if env_state is None else env_state

    def predict(self, *, observations, states, actions):
        # one transition on the device model WITHOUT touching the live env (the reference keeps a separate
        # simulated_env for this, models/gt_model.py:28-33): the stand-in's stateless `simulate_state`
        if hasattr(self.env, "simulate_state"):
            return self.env.simulate_state(states, actions)
        return self.env.simulate(states, actions)

    def predict_n_steps(self, *, start_observations, start_states, policy, horizon):
        raise NotImplementedError("CudaGroundTruthModel rollouts run inside MpcICemB200 on the device")

    def rollout_generator(self, *a, **k):
        raise NotImplementedError

    def rollout_field_names(self):
        return "observations", "next_observations", "actions", "rewards"

    def save(self, path):
        pass

    def load(self, path):
        pass

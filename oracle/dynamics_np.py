"""NumPy float64 forward models used by the oracle.  TEST INFRASTRUCTURE ONLY.

Interface: `rollout(start_state, actions[p,h,d]) -> observations[p,h,obs_dim]` returning the
PRE-action observation of every step (the only thing the reference's cost functions read,
SURVEY F9), as `ForwardModelWithDefaults.predict_n_steps` /
`GroundTruthModel.predict_n_steps` record it (models/abstract_models.py:17-53,
models/gt_model.py:84-102).
"""
import numpy as np


class DenseTanhModel:
    """obs' = tanh(W_o obs + W_a act + b): the toy of SURVEY Appendix C and the single-layer
    case of a dense learned model driven through `ForwardModelWithDefaults.predict`."""

    def __init__(self, w_obs, w_act, bias=None):
        self.w_obs = np.asarray(w_obs, dtype=np.float64)
        self.w_act = np.asarray(w_act, dtype=np.float64)
        self.bias = np.zeros(self.w_obs.shape[0]) if bias is None else np.asarray(bias, np.float64)
        self.obs_dim = self.w_obs.shape[0]
        self.act_dim = self.w_act.shape[1]
        self.state_dim = self.obs_dim

    @classmethod
    def appendix_c(cls, obs_dim=17, act_dim=6, seed=7):
        rs = np.random.RandomState(seed)
        a = 0.95 * np.eye(obs_dim) + 0.02 * rs.randn(obs_dim, obs_dim)
        b = 0.1 * rs.randn(obs_dim, act_dim)
        return cls(a, b)

    def step(self, obs, act):
        return np.tanh(obs @ self.w_obs.T + act @ self.w_act.T + self.bias)

    def observe(self, state):
        return np.asarray(state, dtype=np.float64)

    def rollout(self, start_state, actions):
        p, h, _ = actions.shape
        obs = np.broadcast_to(np.asarray(start_state, np.float64), (p, self.obs_dim)).copy()
        out = np.empty((p, h, self.obs_dim))
        for t in range(h):
            out[:, t] = obs
            obs = self.step(obs, actions[:, t])
        return out


def bf16_round(x):
    """Round float64/float32 values to the nearest bfloat16 (ties to even), returned as float64."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64).reshape(a.shape)


def f16_round(x):
    """Round to the nearest IEEE half (ties to even), returned as float64."""
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float64)


class MlpModel:
    """obs' = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3 with the tensor-core path's operand roundings
    restated: weights, the layer-1 input and both hidden activations are rounded to the operand type of csrc/
    mlp_rollout.cuh (`operand="f16"`, IEEE half, the shipped kernel; "bf16" = the round-1 kernel), products accumulate
    in >= fp32, biases and the state stay fp32/f64.  Not restated: the accumulation order inside the tensor core and
    tanh.approx (rel. error ~5e-4), so agreement is at operand resolution, not bit level (tests/test_gpu_mlp.py states
    the tolerance).  The PRECISION reference of the path is MlpModelF32 below, not this class."""

    def __init__(self, weights, biases, operand="f16"):
        self._round = f16_round if operand == "f16" else bf16_round
        bf16_round_ = self._round
        self.w = [bf16_round_(w) for w in weights]
        self.b = [np.asarray(b, np.float32).astype(np.float64) for b in biases]
        self.obs_dim = self.w[2].shape[0]
        self.act_dim = self.w[0].shape[1] - self.obs_dim
        self.state_dim = self.obs_dim

    def step(self, obs, act):
        rnd = self._round
        x = rnd(np.concatenate([obs, act], axis=-1))
        h1 = rnd(np.tanh(x @ self.w[0].T + self.b[0]))
        h2 = rnd(np.tanh(h1 @ self.w[1].T + self.b[1]))
        return obs + h2 @ self.w[2].T + self.b[2]

    def observe(self, state):
        return np.asarray(state, dtype=np.float64)

    def rollout(self, start_state, actions):
        p, h, _ = actions.shape
        obs = np.broadcast_to(np.asarray(start_state, np.float32).astype(np.float64), (p, self.obs_dim)).copy()
        out = np.empty((p, h, self.obs_dim))
        for t in range(h):
            out[:, t] = obs
            if t + 1 < h:
                obs = self.step(obs, np.asarray(actions[:, t], np.float32).astype(np.float64))
                obs = obs.astype(np.float32).astype(np.float64)      # the device keeps the state in fp32
        return out


class MlpModelF32:
    """The same MLP WITHOUT any reduced-precision restatement: float32 weights, activations and accumulation, exact
    tanh -- what a PyTorch fp32 `nn.Sequential(Linear, Tanh, Linear, Tanh, Linear)` computes (checked against torch
    in tests/test_oracle_mlp.py) and what a learned model driven through the reference's
    `ForwardModelWithDefaults.predict_n_steps` (models/abstract_models.py:31-53) would feed the controller.  This is
    the precision reference for the tensor-core path: the device kernel rounds operands to bfloat16, this model does
    not, and scripts/mlp_precision_report.py measures what that costs the planner (elite overlap, executed action)."""

    def __init__(self, weights, biases):
        self.w = [np.asarray(w, np.float32) for w in weights]
        self.b = [np.asarray(b, np.float32) for b in biases]
        self.obs_dim = self.w[2].shape[0]
        self.act_dim = self.w[0].shape[1] - self.obs_dim
        self.state_dim = self.obs_dim

    def step(self, obs, act):
        x = np.concatenate([np.asarray(obs, np.float32), np.asarray(act, np.float32)], axis=-1)
        h1 = np.tanh(x @ self.w[0].T + self.b[0])
        h2 = np.tanh(h1 @ self.w[1].T + self.b[1])
        return (np.asarray(obs, np.float32) + h2 @ self.w[2].T + self.b[2]).astype(np.float64)

    def observe(self, state):
        return np.asarray(state, dtype=np.float64)

    def rollout(self, start_state, actions, chunk=16384):
        p, h, _ = actions.shape
        out = np.empty((p, h, self.obs_dim))
        for lo in range(0, p, chunk):
            hi = min(p, lo + chunk)
            obs = np.broadcast_to(np.asarray(start_state, np.float32), (hi - lo, self.obs_dim)).copy()
            for t in range(h):
                out[lo:hi, t] = obs
                if t + 1 < h:
                    obs = self.step(obs, actions[lo:hi, t]).astype(np.float32)
        return out

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

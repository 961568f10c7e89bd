"""Device trainer of the MLP forward model (csrc/mlp_train.cuh) behind `forward_model.train(rollout_buffer)`
(icem/main.py:209-210).  ctypes -> C ABI (`icem_mlp_trainer_*`, include/icem_b200.h); no CPU path."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f32, fptr


def epoch_indices(n, batch, epochs, seed, drop_last=True):
    """Minibatch rows of `epochs` passes over n transitions: one permutation per epoch from RandomState(seed), the tail
    that does not fill a batch dropped (or, when n < batch, one batch drawn with replacement per epoch).
    Returns int32 [n_steps, batch]."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(epochs):
        if n < batch:
            out.append(rs.randint(0, n, size=(1, batch)))
            continue
        perm = rs.permutation(n)
        steps = n // batch if drop_last else -(-n // batch)
        for s in range(steps):
            rows = perm[s * batch:(s + 1) * batch]
            if rows.shape[0] < batch:
                rows = np.concatenate([rows, perm[:batch - rows.shape[0]]])
            out.append(rows[None])
    return np.ascontiguousarray(np.concatenate(out, axis=0), dtype=np.int32)


def transitions_from_buffer(buffer):
    """([obs, act] inputs, next_obs - obs targets) of every transition of a RolloutBuffer (misc/rolloutbuffer.py) or
    of any iterable of rollouts with `observations` / `actions` / `next_observations` fields."""
    obs, act, nxt = [], [], []
    for r in buffer:
        obs.append(np.asarray(r["observations"], np.float64))
        act.append(np.asarray(r["actions"], np.float64))
        nxt.append(np.asarray(r["next_observations"], np.float64))
    if not obs:
        raise ValueError("empty rollout buffer: nothing to train on")
    obs, act, nxt = np.concatenate(obs), np.concatenate(act), np.concatenate(nxt)
    return np.concatenate([obs, act], axis=-1), nxt - obs


class MlpTrainer:
    """fp32 minibatch Adam on the MSE of the delta prediction, on the device."""

    def __init__(self, in_dim, hidden, out_dim, device=0):
        self._lib = _lib.load()
        self.dims = (int(in_dim), int(hidden), int(hidden), int(out_dim))
        self._h = C.c_void_p()
        check(self._lib.icem_mlp_trainer_create(int(device), self.dims[0], int(hidden), self.dims[3], C.byref(self._h)))
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.icem_mlp_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shapes(self):
        i, h, _, o = self.dims
        return [(h, i), (h, h), (o, h)], [(h,), (h,), (o,)]

    def set_weights(self, weights, biases, reset_optimizer=True):
        ws, bs = self._shapes()
        w = [f32(x) for x in weights]
        b = [f32(x) for x in biases]
        for a, s in zip(w + b, ws + bs):
            if a.shape != s:
                raise ValueError(f"layer parameter of shape {a.shape}, expected {s}")
        wp = (C.POINTER(C.c_float) * 3)(*[fptr(a) for a in w])
        bp = (C.POINTER(C.c_float) * 3)(*[fptr(a) for a in b])
        check(self._lib.icem_mlp_trainer_set_weights(self._h, wp, bp, int(bool(reset_optimizer))))

    def get_weights(self):
        ws, bs = self._shapes()
        w = [np.empty(s, np.float32) for s in ws]
        b = [np.empty(s, np.float32) for s in bs]
        wp = (C.POINTER(C.c_float) * 3)(*[fptr(a) for a in w])
        bp = (C.POINTER(C.c_float) * 3)(*[fptr(a) for a in b])
        check(self._lib.icem_mlp_trainer_get_weights(self._h, wp, bp))
        return w, b

    def set_data(self, inputs, targets):
        x, t = f32(inputs), f32(targets)
        if x.ndim != 2 or t.ndim != 2 or x.shape[0] != t.shape[0] or x.shape[1] != self.dims[0] or t.shape[1] != self.dims[3]:
            raise ValueError(f"inputs {x.shape} / targets {t.shape} do not fit a {self.dims[0]} -> {self.dims[3]} model")
        check(self._lib.icem_mlp_trainer_set_data(self._h, x.shape[0], fptr(x), fptr(t)))
        self.n = x.shape[0]

    def fit(self, indices, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """One Adam step per row of `indices` [n_steps, batch]; returns the per-step minibatch losses."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        if idx.ndim != 2:
            raise ValueError("indices must be [n_steps, batch]")
        losses = np.empty(idx.shape[0], np.float32)
        check(self._lib.icem_mlp_trainer_fit(self._h, idx.shape[0], idx.shape[1], _lib.iptr(idx), float(lr),
                                             float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                             fptr(losses)))
        return losses

    def predict(self, inputs):
        x = f32(inputs)
        out = np.empty((x.shape[0], self.dims[3]), np.float32)
        check(self._lib.icem_mlp_trainer_predict(self._h, x.shape[0], fptr(x), fptr(out)))
        return out

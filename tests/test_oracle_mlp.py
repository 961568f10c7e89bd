"""The fp32 MLP oracle (oracle/dynamics_np.py::MlpModelF32) is a plain PyTorch fp32 nn.Sequential (the
precision reference of the tensor-core rollout), and the bf16-restating oracle stays within bf16 resolution of it."""
import numpy as np
import torch

from icem_b200 import workloads
from oracle.dynamics_np import MlpModel, MlpModelF32


def _torch_model(ws, bs):
    layers = []
    for i, (w, b) in enumerate(zip(ws, bs)):
        lin = torch.nn.Linear(w.shape[1], w.shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.tensor(w))
            lin.bias.copy_(torch.tensor(b))
        layers.append(lin)
        if i + 1 < len(ws):
            layers.append(torch.nn.Tanh())
    return torch.nn.Sequential(*layers)


def test_fp32_oracle_is_a_torch_fp32_sequential():
    ws, bs = workloads.mlp_model_weights(18, 6, 256, 21)
    mod = MlpModelF32(ws, bs)
    net = _torch_model(ws, bs)
    rs = np.random.RandomState(0)
    obs = (0.5 * rs.randn(512, 18)).astype(np.float32)
    act = rs.uniform(-1, 1, (512, 6)).astype(np.float32)
    with torch.no_grad():
        ref = torch.tensor(obs) + net(torch.tensor(np.concatenate([obs, act], axis=-1)))
    got = mod.step(obs, act)
    assert np.abs(got - ref.numpy()).max() <= 2e-6
    # h = 12 rollout through the oracle's rollout() == stepping torch 11 times
    acts = rs.uniform(-1, 1, (64, 12, 6)).astype(np.float32)
    start = (0.3 * rs.randn(18)).astype(np.float32)
    o = torch.tensor(np.broadcast_to(start, (64, 18)).copy())
    with torch.no_grad():
        for t in range(11):
            o = o + net(torch.cat([o, torch.tensor(acts[:, t])], dim=-1))
    assert np.abs(mod.rollout(start, acts)[:, 11] - o.numpy()).max() <= 2e-5


def test_bf16_restatement_stays_within_bf16_resolution_of_fp32():
    ws, bs = workloads.mlp_model_weights(18, 6, 256, 21)
    f32, b16 = MlpModelF32(ws, bs), MlpModel(ws, bs)
    rs = np.random.RandomState(1)
    acts = rs.uniform(-1, 1, (256, 12, 6)).astype(np.float32)
    start = (0.3 * rs.randn(18)).astype(np.float32)
    d = np.abs(f32.rollout(start, acts) - b16.rollout(start, acts.astype(np.float64)))
    assert np.median(d[:, 11]) <= 2e-2 and d.max() <= 0.25, (np.median(d[:, 11]), d.max())

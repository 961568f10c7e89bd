"""GPU tests of the device MLP trainer (csrc/mlp_train.cuh) behind `forward_model.train(rollout_buffer)`
(icem/main.py:209-210), called through the C ABI.

Oracle: oracle/mlp_train_torch.py -- PyTorch fp32 nn.Sequential + MSELoss + Adam on the CPU, same minibatches (the
reference has no trainable model: parity unpinned against it by construction).  Tolerances: both sides are fp32 with
different summation orders; Adam divides by sqrt(v), so a gradient entry near zero moves its weight by up to lr per
step either way.  Stated below: per-step losses relative 2e-4, weights |d| <= 0.05 * lr * steps at the worst entry
and 1e-3 * lr * steps in the median."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(n, in_dim, hidden, out_dim, seed):
    from icem_b200 import workloads
    rs = np.random.RandomState(seed)
    ws, bs = workloads.mlp_model_weights(out_dim, in_dim - out_dim, hidden, seed)
    x = rs.uniform(-1, 1, (n, in_dim)).astype(np.float32)
    teacher_w, teacher_b = workloads.mlp_model_weights(out_dim, in_dim - out_dim, hidden, seed + 100)
    h = x.astype(np.float64)
    for l in range(3):
        h = h @ teacher_w[l].T.astype(np.float64) + teacher_b[l]
        if l < 2:
            h = np.tanh(h)
    return ws, bs, x, h.astype(np.float32)


@pytest.mark.parametrize("n,in_dim,hidden,out_dim,batch,steps,wd", [(3000, 24, 256, 18, 256, 40, 0.0),
                                                                  (5000, 24, 64, 18, 1000, 25, 0.0),
                                                                  (700, 11, 128, 7, 100, 30, 1e-2),
                                                                  (4100, 24, 256, 18, 4096, 6, 0.0)])
def test_adam_steps_match_torch_fp32(n, in_dim, hidden, out_dim, batch, steps, wd):
    from icem_b200.trainer import MlpTrainer, epoch_indices
    from oracle import mlp_train_torch
    ws, bs, x, t = _problem(n, in_dim, hidden, out_dim, 3)
    idx = epoch_indices(n, batch, 1 + steps * batch // n, seed=5)[:steps]
    assert idx.shape == (steps, batch)
    lr = 1e-3
    tr = MlpTrainer(in_dim, hidden, out_dim)
    tr.set_weights(ws, bs)
    tr.set_data(x, t)
    losses = tr.fit(idx, lr=lr, weight_decay=wd)
    w_dev, b_dev = tr.get_weights()
    w_ref, b_ref, l_ref = mlp_train_torch.fit(ws, bs, x, t, idx, lr=lr, weight_decay=wd)
    assert np.all(np.isfinite(losses))
    assert np.abs(losses / l_ref - 1).max() <= 2e-4, np.abs(losses / l_ref - 1).max()
    assert losses[-1] < losses[0]
    for a, b in zip(w_dev + b_dev, w_ref + b_ref):
        d = np.abs(a - b)
        assert d.max() <= 0.05 * lr * steps, d.max()
        assert np.median(d) <= 1e-3 * lr * steps, np.median(d)
    # the trainer's forward pass == the fp32 network with the trained weights
    pred = tr.predict(x[:300])
    h = x[:300].astype(np.float64)
    for l in range(3):
        h = h @ w_dev[l].T.astype(np.float64) + b_dev[l]
        if l < 2:
            h = np.tanh(h)
    assert np.abs(pred - h).max() <= 2e-5
    # a second fit continues the optimiser state (Adam moments, step count): two halves == one run
    tr2 = MlpTrainer(in_dim, hidden, out_dim)
    tr2.set_weights(ws, bs)
    tr2.set_data(x, t)
    la = tr2.fit(idx[: steps // 2], lr=lr, weight_decay=wd)
    lb = tr2.fit(idx[steps // 2:], lr=lr, weight_decay=wd)
    np.testing.assert_array_equal(np.concatenate([la, lb]), losses)          # deterministic, bit for bit
    for a, b in zip(sum(tr2.get_weights(), []), w_dev + b_dev):
        np.testing.assert_array_equal(a, b)
    tr.close()
    tr2.close()


def test_trainer_errors():
    from icem_b200.planner import IcemError
    from icem_b200.trainer import MlpTrainer
    tr = MlpTrainer(8, 64, 4)
    with pytest.raises(IcemError, match="set_data"):
        tr.fit(np.zeros((1, 4), np.int32))
    tr.set_data(np.zeros((10, 8)), np.zeros((10, 4)))
    with pytest.raises(IcemError, match="index"):
        tr.fit(np.full((1, 4), 10, np.int32))
    with pytest.raises(ValueError):
        tr.set_data(np.zeros((10, 7)), np.zeros((10, 4)))
    with pytest.raises(ValueError):
        tr.set_weights([np.zeros((64, 8)), np.zeros((64, 64)), np.zeros((5, 64))], [np.zeros(64)] * 2 + [np.zeros(4)])
    tr.close()


def test_train_hook_fits_the_env_and_the_controller_takes_the_new_weights():
    """`forward_model.train(rollout_buffer)` as main.py:209-210 calls it: random-action rollouts of an env whose true
    dynamics is a teacher MLP, a randomly initialised CudaMlpModel fitted to them on the device; afterwards the
    controller's planner rolls out the TRAINED weights (the weights are re-uploaded at the next beginning_of_rollout)."""
    from icem_b200 import api, envs, workloads
    from icem_b200.controller import MpcICemB200
    from icem_b200.models import CudaMlpModel
    od, ad, hidden = 18, 6, 64
    teacher = workloads.mlp_model_weights(od, ad, hidden, 77)
    env = envs.MlpStandInEnv(act_dim=ad, bound=1.0, cost="halfcheetah", obs_dim=od, penalise_flipping=True, mlp=teacher)
    env.seed(0)
    rs = np.random.RandomState(1)
    rollouts = []
    for _ in range(40):
        obs = env.reset()
        o, a, nx = [], [], []
        for _ in range(50):
            act = rs.uniform(-1, 1, ad)
            nobs, _, _, _ = env.step(act)
            o.append(obs); a.append(act); nx.append(nobs)
            obs = nobs
        rollouts.append(api.EliteRollout(observations=np.array(o), next_observations=np.array(nx), actions=np.array(a),
                                         rewards=np.zeros(50)))
    buffer = api.EliteBuffer(rollouts)
    model = CudaMlpModel(env=env, hidden=hidden, init_seed=5,
                         train_params=dict(epochs=60, batch_size=250, lr=2e-3, seed=0))
    ctrl = MpcICemB200(env=env, forward_model=model, horizon=12, num_simulated_trajectories=256,
                       cost_along_trajectory="sum", seed=3,
                       action_sampler_params=dict(alpha=0.1, elites_size=10, opt_iterations=2, init_std=0.5,
                                                  use_mean_actions=True, keep_previous_elites=True,
                                                  shift_elites_over_time=True, fraction_elites_reused=0.3,
                                                  noise_beta=0.25))
    obs = env.reset()
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    act = rs.uniform(-1, 1, ad)
    before, _, _ = ctrl._planner.sim_step(obs, act)
    losses = model.train(buffer)
    assert losses.shape == (60 * (2000 // 250),)
    assert losses[-20:].mean() < 0.05 * losses[:5].mean(), (losses[:5].mean(), losses[-20:].mean())
    assert model.version == 1
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    after, _, _ = ctrl._planner.sim_step(obs, act)
    want, _, _ = model.predict(observations=obs, states=None, actions=act)
    assert np.abs(after - want).max() <= 5e-3                     # fp16 operands of the tensor-core path
    assert np.abs(before - want).max() > 10 * np.abs(after - want).max()
    env.set_GT_state(obs)
    true_next, _, _, _ = env.step(act)
    assert np.abs(after - true_next).max() < 0.5 * np.abs(before - true_next).max()      # it learned the env
    action = ctrl.get_action(obs, None)
    assert action.shape == (ad,) and np.all(np.abs(action) <= 1.0)
    ctrl.close()

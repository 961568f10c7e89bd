"""TEST INFRASTRUCTURE -- oracle of the device MLP trainer (icem_b200/csrc/mlp_train.cuh).

The reference calls `forward_model.train(rollout_buffer)` (icem/main.py:209-210) but ships no trainable model
(icem/models/__init__.py:5-8): there is nothing of the reference to restate, so this oracle is the plain
PyTorch fp32 formulation of the same job -- nn.Sequential(Linear, Tanh, Linear, Tanh, Linear) on [obs, act] -> delta,
nn.MSELoss (mean), torch.optim.Adam -- run on the CPU on the SAME minibatches.  PARITY UNPINNED against the reference.
Imported by tests/ only."""
import numpy as np


def fit(weights, biases, inputs, targets, indices, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """Returns (weights, biases, per-step losses) after one Adam step per row of `indices`."""
    import torch
    torch.set_num_threads(1)
    w = [torch.tensor(np.asarray(a, np.float32)) for a in weights]
    b = [torch.tensor(np.asarray(a, np.float32)) for a in biases]
    layers = []
    for l in range(3):
        lin = torch.nn.Linear(w[l].shape[1], w[l].shape[0])
        with torch.no_grad():
            lin.weight.copy_(w[l])
            lin.bias.copy_(b[l])
        layers.append(lin)
        if l < 2:
            layers.append(torch.nn.Tanh())
    net = torch.nn.Sequential(*layers)
    opt = torch.optim.Adam(net.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    x = torch.tensor(np.asarray(inputs, np.float32))
    t = torch.tensor(np.asarray(targets, np.float32))
    loss_fn = torch.nn.MSELoss()
    losses = []
    for rows in np.asarray(indices):
        r = torch.tensor(rows.astype(np.int64))
        opt.zero_grad()
        loss = loss_fn(net(x[r]), t[r])
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    lins = [m for m in net if isinstance(m, torch.nn.Linear)]
    return ([m.weight.detach().numpy().copy() for m in lins], [m.bias.detach().numpy().copy() for m in lins],
            np.asarray(losses, np.float32))

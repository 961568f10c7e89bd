"""Many independent MPC episodes planned side by side on one GPU.

The reference parallelises evaluation EPISODES over host processes (`RolloutManager.par_sample`,
icem/misc/rollout_utils.py:129-152: one spawned worker per rollout, each with its own env and controller copy).
On a B200 one plan step at the reference's default budget (40 -> 32 -> 25 trajectories, settings/defaults/
i-cem-blitz.json) keeps a handful of the 148 SMs busy for its whole latency, so the device form of that parallelism
is: one planner handle (own CUDA stream + graph) per episode, every handle's plan step launched before any is
collected (`icem_plan_async` / `icem_plan_finish`).  Episodes stay independent: each has its own env, controller
state (mean / std / elites) and Philox seed, exactly like the reference's per-process copies.
"""
import os
import time

import numpy as np

# Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8): streams that share a
# queue serialise.  Measured on B200 with the default: 4-way overlap no matter how many episodes.  32 is the
# maximum; it only takes effect if set before the process creates its CUDA context, so import this module first.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


class EpisodeBatch:
    """`envs[i]`, `controllers[i]` form episode i; all advance in lockstep like RolloutManager._sample does for one
    (icem/misc/rollout_utils.py:155-227)."""

    def __init__(self, envs, controllers):
        if len(envs) != len(controllers) or not envs:
            raise ValueError("need one controller per env")
        self.envs, self.controllers = list(envs), list(controllers)
        self.obs = [None] * len(envs)

    def __len__(self):
        return len(self.envs)

    def reset(self, mode="train"):
        for i, (env, ctrl) in enumerate(zip(self.envs, self.controllers)):
            self.obs[i] = env.reset_with_mode(mode) if hasattr(env, "reset_with_mode") else env.reset()
            ctrl.beginning_of_rollout(observation=self.obs[i], state=env.get_GT_state(), mode=mode)
        return list(self.obs)

    def plan(self, mode="train"):
        """One plan step of every episode: launch all, then collect all.  Returns actions [B, d]."""
        for env, ctrl, ob in zip(self.envs, self.controllers, self.obs):
            ctrl.begin_get_action(ob, env.get_GT_state(), mode)
        return np.stack([ctrl.finish_get_action() for ctrl in self.controllers])

    def step(self, mode="train"):
        actions = self.plan(mode)
        rewards = np.empty(len(self))
        dones = np.zeros(len(self), dtype=bool)
        for i, (env, a) in enumerate(zip(self.envs, actions)):
            self.obs[i], rewards[i], dones[i], _ = env.step(a)
        return actions, rewards, dones

    def run(self, task_horizon, mode="train"):
        """Full episodes; returns per-episode dicts like the transitions RolloutManager collects."""
        t0 = time.time()
        self.reset(mode)
        log = [dict(observations=[], actions=[], rewards=[]) for _ in self.envs]
        for _ in range(task_horizon):
            prev = list(self.obs)
            actions, rewards, _ = self.step(mode)
            for i in range(len(self)):
                log[i]["observations"].append(prev[i])
                log[i]["actions"].append(actions[i])
                log[i]["rewards"].append(rewards[i])
        for i, ctrl in enumerate(self.controllers):
            ctrl.end_of_rollout(time.time() - t0, float(np.sum(log[i]["rewards"])), mode)
        return [{k: np.asarray(v) for k, v in ep.items()} for ep in log]

    def close(self):
        for c in self.controllers:
            c.close()
        for e in self.envs:
            e.close()


class FusedEpisodeBatch(EpisodeBatch):
    """The same B episodes through ONE planner handle created with `num_problems = B` (icem_plan_batch): every
    kernel of the plan step covers all episodes (grid.y = episode) and the whole step of all episodes is one CUDA
    graph launch, so the number of concurrent episodes is not limited by the 32 hardware stream queues.  All
    episodes share settings and model; episode i draws with seed + i, exactly like a single-problem handle with
    that seed (tests/test_gpu_batched.py)."""

    def __init__(self, envs, controller):
        self.envs, self.controller = list(envs), controller
        self.controllers = [controller]
        self.obs = [None] * len(self.envs)
        if controller._planner.settings.num_problems != len(self.envs):
            raise ValueError("controller must be created with num_problems == len(envs)")

    def reset(self, mode="train"):
        for i, env in enumerate(self.envs):
            self.obs[i] = env.reset_with_mode(mode) if hasattr(env, "reset_with_mode") else env.reset()
        c = self.controller
        c.beginning_of_rollout(observation=self.obs[0], state=self.envs[0].get_GT_state(), mode=mode)
        return list(self.obs)

    def plan(self, mode="train"):
        fm = self.controller.forward_model
        states = np.stack([fm.start_state(ob, env.get_GT_state()) for env, ob in zip(self.envs, self.obs)])
        return self.controller.plan_batch(states, obs_dim=int(np.asarray(self.obs[0]).shape[-1]))

    def step(self, mode="train"):
        """Plan all episodes (one graph launch) and step all envs (one launch of B transitions, icem_sim_step_batch)
        instead of B env.step calls of one small launch + synchronisation each."""
        actions = self.plan(mode)
        envs = self.envs
        acts = np.stack([np.clip(a, e.action_space.low, e.action_space.high) for a, e in zip(actions, envs)])
        prev_obs = list(self.obs)
        nxt = self.controller._planner.sim_step_batch(np.stack([e._state for e in envs]), acts)
        rewards = np.empty(len(self))
        for i, e in enumerate(envs):
            e._state = nxt[i]
            e._t += e.dt
            self.obs[i] = e._obs()
            rewards[i] = -float(e.cost_fn(prev_obs[i], acts[i], self.obs[i]))
        return actions, rewards, np.zeros(len(self), dtype=bool)


def make_fused_episode_batch(env_name, num_episodes, controller_params, seed=0, device=0, controller_cls=None):
    from . import envs as envs_mod
    from .controller import MpcICemB200
    from .models import CudaGroundTruthModel
    cls = controller_cls or MpcICemB200
    es = []
    for i in range(num_episodes):
        env = envs_mod.make_env(env_name, device=device)
        env.seed(seed + i)
        es.append(env)
    ctrl = cls(env=es[0], forward_model=CudaGroundTruthModel(env=es[0]), seed=seed + 1000, device=device,
               world_size=1, rank=0, num_problems=num_episodes, **controller_params)
    return FusedEpisodeBatch(es, ctrl)


def make_episode_batch(env_name, num_episodes, controller_params, seed=0, device=0, controller_cls=None):
    """B stand-in envs + CUDA ground-truth models + controllers with the reference's `controller_params` dict."""
    from . import envs as envs_mod
    from .controller import MpcICemB200
    from .models import CudaGroundTruthModel
    cls = controller_cls or MpcICemB200
    es, cs = [], []
    for i in range(num_episodes):
        env = envs_mod.make_env(env_name, device=device)
        env.seed(seed + i)
        es.append(env)
        cs.append(cls(env=env, forward_model=CudaGroundTruthModel(env=env), seed=seed + 1000 + i, device=device,
                      world_size=1, rank=0, **controller_params))
    return EpisodeBatch(es, cs)

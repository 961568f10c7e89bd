"""Stand-alone copies of the reference's plugin ABC *signatures* (no logic), used only when the reference package
is not importable (e.g. on a box without /root/reference).  When the reference IS importable (the launcher put
`/root/reference/icem` on sys.path), `controller.py` / `models.py` subclass the reference's own ABCs instead, so
`issubclass(MpcICemB200, ModelBasedController)` in `main.get_controllers` (icem/main.py:43) holds.

Mirrors: icem/misc/base_types.py:42-59 (Controller), :62-118 (ForwardModel);
icem/controllers/abstract_controller.py:43-58 (StatefulController), :61-72 (ModelBasedController).
"""
from abc import ABC, abstractmethod


_REF_TOP_LEVEL = ("controllers", "environments", "misc", "models")


def forget_failed_reference_import(before):
    """The reference uses top-level package names (`environments`, `controllers`, ...).  Trying to import them when
    the reference is NOT on sys.path can pick up an unrelated namespace package of the same name from site-packages
    and leave it cached in sys.modules, which would shadow the reference if it is put on sys.path later (launcher,
    tests).  Drop whatever such an attempt added."""
    import sys
    for name in list(sys.modules):
        if name not in before and name.split(".")[0] in _REF_TOP_LEVEL:
            del sys.modules[name]


def reference_bases():
    """(Controller bases, ForwardModel base, RolloutBuffer, Rollout, AbstractGroundTruthModel) from the reference
    if importable, else None."""
    import sys
    before = set(sys.modules)
    try:
        from controllers.abstract_controller import ModelBasedController as RefMBC  # noqa
        from controllers.abstract_controller import StatefulController as RefSC  # noqa
        from misc.rolloutbuffer import Rollout, RolloutBuffer  # noqa
        from models.abstract_models import ForwardModelWithDefaults  # noqa
        from models.gt_model import AbstractGroundTruthModel  # noqa
        return dict(mbc=RefMBC, sc=RefSC, rollout=Rollout, buffer=RolloutBuffer, fm=ForwardModelWithDefaults,
                    gt=AbstractGroundTruthModel)
    except ImportError:
        forget_failed_reference_import(before)
        return None


class Controller(ABC):
    needs_training = False
    needs_data = False
    has_state = False
    required_settings = []

    def __init__(self, *, env):
        self.env = env

    @abstractmethod
    def get_action(self, obs, state, mode="train"):
        pass


class StatefulController(Controller, ABC):
    has_state = True

    @abstractmethod
    def beginning_of_rollout(self, *, observation, state=None, mode):
        pass

    @abstractmethod
    def end_of_rollout(self, total_time, total_return, mode):
        pass


class ModelBasedController(Controller, ABC):
    def __init__(self, *, forward_model, env, cost_along_trajectory, do_visualize_plan=None,
                 use_env_reward_as_cost=False, **kwargs):
        super().__init__(env=env, **kwargs)
        self.forward_model = forward_model
        self.do_visualize_plan = do_visualize_plan
        self.visualize_env = None
        self.cost_fn = self.env.cost_fn
        self.cost_along_trajectory = cost_along_trajectory
        self.use_env_reward_as_cost = use_env_reward_as_cost

    def visualize_plan(self, *, obs, state, acts):
        """Replays a planned trajectory in a second env instance, like the reference hook
        (controllers/abstract_controller.py:93-128): mode "last" shows where the plan ends, mode "all" re-simulates it
        from `state` at 25 fps and reports the first step at which the replay leaves the plan by more than 0.01."""
        mode = self.do_visualize_plan
        if not mode or not getattr(self.env, "supports_live_rendering", False):
            return
        if mode not in ("last", "all"):
            raise AttributeError("unknown mode for do_visualize_plan: Options: None, 'last','all'")
        import time

        import numpy as np
        viewer = self.visualize_env
        if viewer is None:
            viewer = self.visualize_env = type(self.env)(name=self.env.name, **getattr(self.env, "init_kwargs", {}))
            viewer.reset()
        if mode == "last":
            viewer.set_state_from_observation(obs[-1])
            viewer.step(acts[-1])
            viewer.render()
            return
        viewer.set_GT_state(state)
        mismatch_at = None
        for t, action in enumerate(acts):
            replayed = viewer.step(action)[0]
            if mismatch_at is None and t + 1 < len(obs) and np.linalg.norm(replayed - obs[t + 1]) > 0.01:
                mismatch_at = t
                print(f"simulation for visualization does not match mental model at {t}: ")
                print("orig: ", obs[t + 1])
                print("simu: ", replayed)
            viewer.render()
            time.sleep(1.0 / 25.0)


class ForwardModel(ABC):
    supports_stochastic = False

    def __init__(self, *, env):
        self.env = env

    def reset(self, observation):
        return None

    def got_actual_observation_and_env_state(self, *, observation, env_state=None, model_state=None):
        return None

    def train(self, buffer):
        pass

    def save(self, path):
        pass

    def load(self, path):
        pass


class AbstractGroundTruthModel(ForwardModel, ABC):
    pass


class EliteRollout:
    """Minimal stand-in for misc/rolloutbuffer.py::Rollout: field access by name."""

    def __init__(self, **fields):
        self._fields = fields

    def __getitem__(self, key):
        return self._fields[key]

    def __len__(self):
        return len(next(iter(self._fields.values())))


class EliteBuffer:
    """Minimal stand-in for misc/rolloutbuffer.py::RolloutBuffer: len / truthiness / iteration / as_array."""

    def __init__(self, rollouts=None):
        self.rollouts = list(rollouts or [])

    def __len__(self):
        return len(self.rollouts)

    def __iter__(self):
        return iter(self.rollouts)

    def __getitem__(self, i):
        return self.rollouts[i]

    def as_array(self, key):
        import numpy as np
        return np.concatenate([r[key][None, ...] for r in self.rollouts], axis=0)

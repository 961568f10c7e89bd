"""Import-only stubs so that `/root/reference/icem/environments/mujoco.py` can be imported and its
`cost_fn`s (pure NumPy) called unbound; the simulator itself (MuJoCo) is absent. TEST INFRASTRUCTURE ONLY."""


class MujocoEnv:
    def __init__(self, *a, **k):
        raise RuntimeError("gym shim: MuJoCo is not available in this image")


class ReacherEnv(MujocoEnv):
    pass

from . import MujocoEnv


class HumanoidEnv(MujocoEnv):
    pass

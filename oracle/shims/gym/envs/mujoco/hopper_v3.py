from . import MujocoEnv


class HopperEnv(MujocoEnv):
    pass

from . import MujocoEnv


class HalfCheetahEnv(MujocoEnv):
    pass

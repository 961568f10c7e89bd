from . import MujocoEnv


class AntEnv(MujocoEnv):
    pass

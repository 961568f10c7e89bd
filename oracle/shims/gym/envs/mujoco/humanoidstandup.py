from . import MujocoEnv


class HumanoidStandupEnv(MujocoEnv):
    pass

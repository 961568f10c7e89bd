from ..robot_env import RobotEnv


class FetchReachEnv(RobotEnv):
    pass

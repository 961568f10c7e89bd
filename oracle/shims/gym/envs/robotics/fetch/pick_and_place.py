from ..robot_env import RobotEnv


class FetchPickAndPlaceEnv(RobotEnv):
    pass

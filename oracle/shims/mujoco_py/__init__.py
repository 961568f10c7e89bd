"""Import-only stub (MuJoCo absent). TEST INFRASTRUCTURE ONLY."""


class MujocoException(Exception):
    pass

"""Minimal stand-in for `gym` (absent here). TEST INFRASTRUCTURE ONLY.
Only what the reference's controller path touches: Env base, spaces.{Box,Discrete,Dict}, utils.EzPickle."""
from . import spaces, utils  # noqa: F401


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def __init__(self, *args, **kwargs):
        pass

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass

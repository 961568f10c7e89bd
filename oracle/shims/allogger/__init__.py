"""Minimal stand-in for the third-party `allogger` package (absent in this image).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Covers exactly the call sites the
reference touches: main.py:63,74,85-86,96,102-103,180,243; controllers/icem.py:28,177;
controllers/mpc.py:35; misc/rollout_utils.py:38.
"""
import os
from collections import defaultdict
from . import utils  # noqa: F401

_LOGDIR = os.environ.get("ALLOGGER_SHIM_LOGDIR", "/tmp/allogger_shim")
_LOGGERS = {}


class _Manager:
    @staticmethod
    def dict(d=None):
        return dict(d or {})


class _Logger:
    def __init__(self, scope):
        self.scope = scope
        self.step_per_key = defaultdict(int)
        self.manager = _Manager()
        self.records = defaultdict(list)

    @property
    def logdir(self):
        return _LOGDIR

    def info(self, *a, **k):
        pass

    def log(self, value, key=None, **k):
        self.records[key].append(value)
        self.step_per_key[key] += 1


def basic_configure(logdir=None, default_outputs=None, **kwargs):
    global _LOGDIR
    if logdir is not None:
        _LOGDIR = logdir
        os.makedirs(logdir, exist_ok=True)


def get_logger(scope="root", default_outputs=None, **kwargs):
    if scope not in _LOGGERS:
        _LOGGERS[scope] = _Logger(scope)
    return _LOGGERS[scope]


def close():
    pass

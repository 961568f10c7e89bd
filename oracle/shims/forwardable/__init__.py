"""Minimal stand-in for `forwardable` (absent here). TEST INFRASTRUCTURE ONLY.
Call sites: misc/rolloutbuffer.py:5,9,14,58,65,124,126."""
import sys


def forwardable():
    def deco(cls):
        return cls
    return deco


def def_delegators(attr, names):
    """Inject forwarding methods into the *calling class body*."""
    frame = sys._getframe(1)
    for name in [n.strip() for n in names.split(",") if n.strip()]:
        def make(name):
            def fwd(self, *a, **k):
                return getattr(getattr(self, attr), name)(*a, **k)
            fwd.__name__ = name
            return fwd
        frame.f_locals[name] = make(name)

"""Import the UNMODIFIED reference (`/root/reference/icem`) inside this container.

TEST INFRASTRUCTURE ONLY.  The reference uses top-level imports (`from controllers import ...`,
icem/main.py:8-18), `from collections import Mapping` (icem/misc/helpers.py:5, removed in
py3.10) and third-party packages that are absent here; this module provides exactly the
sys.path / alias surgery needed, never edits the reference.
"""
import collections
import collections.abc
import os
import sys

def _find_reference():
    """ICEM_REFERENCE_ROOT, else the read-only checkout of this container, else the copy staged for the GPU box
    (scripts/stage_reference.sh -> baseline/_ref/icem: git-ignored, shipped by gpurun, never committed)."""
    env = os.environ.get("ICEM_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "icem")
    for cand in ("/root/reference/icem", staged):
        if os.path.isfile(os.path.join(cand, "controllers", "icem.py")):
            return cand
    return "/root/reference/icem"


REFERENCE_ROOT = _find_reference()
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "controllers", "icem.py"))


def install_shims():
    """Put the third-party stand-ins on sys.path (needed by both the reference and the launcher tests)."""
    if not hasattr(collections, "Mapping"):
        collections.Mapping = collections.abc.Mapping
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)


def install_reference():
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)


def load_reference():
    """Returns a namespace with the reference modules on the hot path."""
    install_reference()
    import importlib
    ns = type("RefNS", (), {})()
    ns.icem = importlib.import_module("controllers.icem")
    ns.mpc = importlib.import_module("controllers.mpc")
    ns.controllers = importlib.import_module("controllers")
    ns.abstract_controller = importlib.import_module("controllers.abstract_controller")
    ns.abstract_models = importlib.import_module("models.abstract_models")
    ns.models = importlib.import_module("models")
    ns.abstract_environments = importlib.import_module("environments.abstract_environments")
    ns.rolloutbuffer = importlib.import_module("misc.rolloutbuffer")
    ns.colorednoise = importlib.import_module("colorednoise")
    return ns

"""Stand-in for `imageio` (absent here); call site misc/rollout_utils.py:7,80 (never reached in tests)."""


def get_writer(*a, **k):
    raise RuntimeError("imageio shim: video recording unsupported")

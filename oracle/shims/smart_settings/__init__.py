"""Minimal stand-in for `smart_settings` (absent here). TEST INFRASTRUCTURE ONLY.
Call sites: misc/helpers.py:202 (load), main.py:7,40 (recursive_objectify)."""
import ast
import json
import os
from . import param_classes
from .param_classes import recursive_objectify


def load(path_or_literal, pre_unpack_hooks=None, **kwargs):
    if os.path.isfile(path_or_literal):
        with open(path_or_literal, "r") as f:
            d = json.load(f)
    else:
        d = ast.literal_eval(path_or_literal)
    for hook in pre_unpack_hooks or []:
        hook(d)
    return recursive_objectify(d, make_immutable=False)

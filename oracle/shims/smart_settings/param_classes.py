class AttributeDict(dict):
    """dict with attribute access; stays JSON-serialisable (main.py:103, helpers.py:206-209)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(e)

    def __setattr__(self, key, value):
        self[key] = value


def recursive_objectify(nested, make_immutable=True):
    if isinstance(nested, dict):
        return AttributeDict({k: recursive_objectify(v, make_immutable) for k, v in nested.items()})
    if isinstance(nested, (list, tuple)):
        return type(nested)(recursive_objectify(v, make_immutable) for v in nested)
    return nested

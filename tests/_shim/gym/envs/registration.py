"""TEST-ONLY: id -> entry_point registry with gym.make semantics."""
import importlib

_REGISTRY = {}


def register(id, entry_point=None, kwargs=None, **_ignored):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError("No registered env with id: {}".format(id))
    entry_point, default_kwargs = _REGISTRY[id]
    if callable(entry_point):
        ctor = entry_point
    else:
        mod_name, attr = entry_point.split(":")
        ctor = getattr(importlib.import_module(mod_name), attr)
    merged = dict(default_kwargs)
    merged.update(kwargs)
    return ctor(**merged)

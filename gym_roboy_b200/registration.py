"""`register` / `make` for the env ids of this package.

The reference registers only `msj-control-v0` with no default `simulation_client`
(gym_roboy/__init__.py:3-6), so neither `gym.make('msj-control-v0')` nor the documented
`gym.make('msj-control-v1')` (README.md:24) can construct an env.  Here `msj-control-v1` is
registered with a default `CudaSimulationClient`, in this package's own registry and -- when a
`gym` / `gymnasium` module is importable -- in theirs too.
"""
import importlib

_REGISTRY = {}


def register(id, entry_point, kwargs=None):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))
    for mod in ("gym", "gymnasium"):
        try:
            reg = importlib.import_module(mod + ".envs.registration")
            reg.register(id=id, entry_point=entry_point, kwargs=dict(kwargs or {}))
        except Exception:  # not installed, or id already registered there
            pass


def spec(id):
    if id not in _REGISTRY:
        raise KeyError("No registered env with id: {}".format(id))
    return _REGISTRY[id]


def make(id, **kwargs):
    entry_point, defaults = spec(id)
    if callable(entry_point):
        ctor = entry_point
    else:
        mod_name, attr = entry_point.split(":")
        ctor = getattr(importlib.import_module(mod_name), attr)
    merged = dict(defaults)
    merged.update(kwargs)
    return ctor(**merged)

"""`model_constructor` / `NetConstructor` with the contract of the reference's admin/model_constructor.py:5-45: a decorated
factory attaches a `NetConstructor` (factory name, module, arguments) to the network it returns as `.constructor`, which
`trainers/base_trainer.py:105` pickles into checkpoints and `admin/loading.py` uses to rebuild the network.  The attribute
names (`fun_name`, `fun_module`, `args`, `kwds`) are part of that pickle format."""
import functools
import importlib


class NetConstructor:
    """Recipe to rebuild a network: `get()` imports `fun_module` and calls `fun_name(*args, **kwds)`."""

    def __init__(self, fun_name, fun_module, args, kwds):
        self.fun_name, self.fun_module = fun_name, fun_module
        self.args, self.kwds = args, kwds

    def get(self):
        factory = getattr(importlib.import_module(self.fun_module), self.fun_name)
        return factory(*self.args, **self.kwds)


def model_constructor(f):
    """Decorator for network factories; a factory may return the network alone or first in a tuple / list."""

    def build(*args, **kwds):
        result = f(*args, **kwds)
        net = result[0] if isinstance(result, (tuple, list)) else result
        net.constructor = NetConstructor(f.__name__, f.__module__, args, kwds)
        return result

    return functools.update_wrapper(build, f)

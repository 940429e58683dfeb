"""Checkpoint loading with the contract of the reference's admin/loading.py:24-100: a checkpoint is the dict the trainer
writes (`trainers/base_trainer.py:95-115`) -- `'net'` (state_dict), `'constructor'` (a pickled `NetConstructor`: factory
name / module / arguments), optional `'net_info'` -- and `load_network` rebuilds the network from the constructor and loads
the weights.  Checkpoint selection follows the reference: a file path, a directory (latest `*.pth.tar`), or a directory plus
an epoch number (`*_ep%04d.pth.tar`, exactly one match).

Checkpoints written by the REFERENCE load here unchanged: their pickles name `admin.model_constructor.NetConstructor` and the
factory module `models.dbsr.dbsrnet`; both are resolved inside this package (`deep_rawburst_sr_b200.<module>`) while
unpickling / constructing, so `NetworkParam(network_path='dbsr_synthetic_default.pth', ...)` works with the published file.
Unpickling executes code from the file, like the reference's `torch.load`: only load checkpoints you trust."""
import importlib
import inspect
import os
import pickle
import types
from pathlib import Path

import torch

_PACKAGE = __name__.split('.')[0]


def package_module(name: str) -> str:
    """module path of the reference layout (`models.dbsr.dbsrnet`, `admin.model_constructor`) -> the same module of this package"""
    if name == _PACKAGE or name.startswith(_PACKAGE + '.'):
        return name
    top = name.split('.')[0]
    if top in ('admin', 'models', 'evaluation', 'data', 'dataset', 'utils'):
        return _PACKAGE + '.' + name
    return name


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        return super().find_class(package_module(module), name)


def _pickle_module():
    m = types.ModuleType('dbsr_b200_checkpoint_pickle')
    m.Unpickler = _Unpickler
    m.load = lambda f, **kw: _Unpickler(f, **kw).load()
    m.loads = pickle.loads
    m.dump, m.dumps, m.Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    m.__dict__.update({k: getattr(pickle, k) for k in ('HIGHEST_PROTOCOL', 'DEFAULT_PROTOCOL', 'PickleError', 'UnpicklingError')})
    return m


def read_checkpoint(path) -> dict:
    return torch.load(str(path), map_location='cpu', pickle_module=_pickle_module(), weights_only=False)


def resolve_checkpoint(network_dir=None, checkpoint=None) -> str:
    """admin/loading.py:36-66: the file itself, the latest `*.pth.tar` of a directory, or the one `*_ep%04d.pth.tar` of an epoch"""
    net_path = Path(network_dir) if network_dir is not None else None
    if net_path is not None and net_path.is_file():
        checkpoint = str(net_path)
    if checkpoint is None:
        found = sorted(net_path.glob('*.pth.tar')) if net_path is not None else []
        if not found:
            raise Exception('No matching checkpoint file found')
        return str(found[-1])
    if isinstance(checkpoint, int):
        found = sorted(net_path.glob('*_ep{:04d}.pth.tar'.format(checkpoint))) if net_path is not None else []
        if not found:
            raise Exception('No matching checkpoint file found')
        if len(found) > 1:
            raise Exception('Multiple matching checkpoint files found')
        return str(found[0])
    if isinstance(checkpoint, str):
        return os.path.expanduser(checkpoint)
    raise TypeError


def load_network(network_dir=None, checkpoint=None, constructor_fun_name=None, constructor_module=None, **kwargs):
    """-> (net, checkpoint_dict).  Extra keyword arguments replace saved constructor arguments of the same name; unknown
    ones are reported and ignored (admin/loading.py:78-84)."""
    checkpoint_dict = read_checkpoint(resolve_checkpoint(network_dir, checkpoint))
    net_constr = checkpoint_dict.get('constructor')
    if net_constr is None:
        raise RuntimeError('No constructor for the given network.')
    if constructor_fun_name is not None:
        net_constr.fun_name = constructor_fun_name
    if constructor_module is not None:
        net_constr.fun_module = constructor_module
    net_constr.fun_module = package_module(net_constr.fun_module)
    net_fun = getattr(importlib.import_module(net_constr.fun_module), net_constr.fun_name)
    accepted = list(inspect.signature(net_fun).parameters.keys())
    for arg, val in kwargs.items():
        if arg in accepted:
            net_constr.kwds[arg] = val
        else:
            print('WARNING: Keyword argument "{}" not found when loading network. It was ignored.'.format(arg))
    net = net_constr.get()
    net.load_state_dict(checkpoint_dict['net'])
    net.constructor = checkpoint_dict['constructor']
    if checkpoint_dict.get('net_info') is not None:
        net.info = checkpoint_dict['net_info']
    return net, checkpoint_dict


def load_weights(net, path, strict=True):
    net.load_state_dict(read_checkpoint(path)['net'], strict=strict)
    return net

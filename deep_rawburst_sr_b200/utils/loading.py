"""`load_network(net_path)` of the reference's utils/loading.py:6-19: an absolute path is used as it is, anything else is
relative to `<workspace_dir>/checkpoints`."""
import os

from ..admin import loading
from ..admin.environment import env_settings


def load_network(net_path, return_dict=False, **kwargs):
    kwargs['backbone_pretrained'] = False
    path_full = net_path if os.path.isabs(net_path) else os.path.join(env_settings().workspace_dir, 'checkpoints', net_path)
    net, checkpoint_dict = loading.load_network(path_full, **kwargs)
    return (net, checkpoint_dict) if return_dict else net

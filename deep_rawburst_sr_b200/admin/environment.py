"""`env_settings()` with the fields the forward path's callers read (reference admin/environment.py:6-50): `pretrained_nets_dir`
(models/dbsr/dbsrnet.py:59-60, evaluation/common_utils/network_param.py:82), `workspace_dir` (utils/loading.py:12),
`save_data_path` (evaluation/*/compute_score.py:45, save_results.py:41), `synburstval_dir` (dataset/synthetic_burst_val_set.py:32)
and `burstsr_dir`.  Unlike the reference this never WRITES a `local.py`: the values come from `admin/local.py` when the user
provides one (class `EnvironmentSettings`, the reference's format), else from the environment variables DBSR_<FIELD>."""
import importlib
import os

_FIELDS = ('workspace_dir', 'tensorboard_dir', 'pretrained_nets_dir', 'save_data_path', 'zurichraw2rgb_dir', 'burstsr_dir',
           'synburstval_dir')


class EnvSettings:
    def __init__(self):
        for f in _FIELDS:
            setattr(self, f, os.environ.get('DBSR_' + f.upper(), ''))


def env_settings():
    try:
        local = importlib.import_module('deep_rawburst_sr_b200.admin.local')
        return local.EnvironmentSettings()
    except ImportError:
        return EnvSettings()

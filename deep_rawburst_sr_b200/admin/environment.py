"""Minimal `env_settings()` (reference admin/environment.py:42-50): only `pretrained_nets_dir` is needed on the
forward path (models/dbsr/dbsrnet.py:59-60).  Unlike the reference this never writes a `local.py`; the directory
comes from `admin/local.py` when the user provides one, else from $DBSR_PRETRAINED_NETS_DIR."""
import importlib
import os


class EnvSettings:
    def __init__(self):
        self.pretrained_nets_dir = os.environ.get('DBSR_PRETRAINED_NETS_DIR', '')


def env_settings():
    try:
        local = importlib.import_module('deep_rawburst_sr_b200.admin.local')
        return local.EnvironmentSettings()
    except ImportError:
        return EnvSettings()

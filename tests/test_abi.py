"""CPU: the C-ABI shared library loads without a GPU and exports exactly the entry points declared in
include/dbsr_b200.h; the ctypes prototype table covers each of them (no compute calls here)."""
import ctypes
import os
import re

import pytest

from deep_rawburst_sr_b200 import _lib
from deep_rawburst_sr_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'dbsr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dbsr_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/dbsr_b200.h but not exported'
    assert sorted(_lib.PROTOTYPES.keys()) == names


def test_probes_without_gpu():
    lib = _lib.load_library()
    assert lib.dbsr_version() == 100
    # struct layout must match the header (8-byte pointer + 7 int32, padded to 40 bytes)
    assert ctypes.sizeof(_lib.NhwcView) == 40
    assert ctypes.sizeof(_lib.ConvDesc) == 3 * 40 + 8 + 8 + 8 * 4


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.DbsrB200Error):
        _lib.load_library(str(tmp_path / 'nope.so'))


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """The built cubin must contain tcgen05 MMA (UTCHMMA), TMEM loads (LDTM), TMA loads (UTMALDG), the TMA tensor stores of the
    conv epilogue (UTMASTG), the asynchronous copies of the fused warp + softmax kernel (LDGSTS) and the warp-level bf16 MMA of the
    banded cost-volume product (HMMA.16816)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', build_library()], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'HMMA.16816.F32.BF16'):
        assert mnemonic in sass, mnemonic

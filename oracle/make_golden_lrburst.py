"""Golden vectors of `single2lrburst` / `rgb2rawburst` from the UNMODIFIED reference (build container only; needs cv2):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_lrburst.py

TEST INFRASTRUCTURE ONLY.  data/synthetic_burst_generation.py is loaded by file with its two imports pre-registered
(`data.camera_pipeline` and `utils.data_format_utils`, loaded by file as well): importing the `data` PACKAGE would run
data/__init__.py -> data/loader.py, which needs `torch._six` (removed from torch 2.x) and is not on this path.  The frame
transforms are sampled by the reference from Python's `random` (seeded here); `get_tmat` is wrapped only to RECORD the matrices
it returns, so that the test can feed the same transforms to the oracle and to the CUDA kernel."""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, ROOT)

# name, seed, H, W, burst, factor, border_crop, transformation params
CASES = [
    ('lrburst_default_b14_432', 0, 432, 432, 14, 4, 24, {'max_translation': 24.0, 'max_rotation': 1.0, 'max_shear': 0.0, 'max_scale': 0.0}),
    ('lrburst_shear_scale_b5_200x264', 1, 200, 264, 5, 4, 4, {'max_translation': 8.0, 'max_rotation': 5.0, 'max_shear': 0.05, 'max_scale': 0.1, 'max_ar_factor': 0.05}),
    ('lrburst_factor2_b3_96x80', 2, 96, 80, 3, 2, None, {'max_translation': 3.0, 'max_rotation': 0.5}),
]


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def make_image(seed, H, W):
    g = torch.Generator().manual_seed(15000 + seed)
    coarse = torch.rand(1, 3, H // 8 + 2, W // 8 + 2, generator=g)
    img = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False)[0]
    return (img + 0.15 * torch.rand(3, H, W, generator=g)).clamp(0.0, 1.0).contiguous()


def main():
    sys.dont_write_bytecode = True
    data_pkg, utils_pkg = types.ModuleType('data'), types.ModuleType('utils')
    data_pkg.__path__, utils_pkg.__path__ = [], []
    sys.modules['data'], sys.modules['utils'] = data_pkg, utils_pkg
    data_pkg.camera_pipeline = load('data.camera_pipeline', os.path.join(REF, 'data', 'camera_pipeline.py'))
    utils_pkg.data_format_utils = load('utils.data_format_utils', os.path.join(REF, 'utils', 'data_format_utils.py'))
    gen = load('ref_synthetic_burst_generation', os.path.join(REF, 'data', 'synthetic_burst_generation.py'))
    recorded = []
    orig = gen.get_tmat

    def recording_get_tmat(*a, **k):
        m = orig(*a, **k)
        recorded.append(np.array(m, dtype=np.float64))
        return m

    gen.get_tmat = recording_get_tmat
    for name, seed, H, W, n, f, crop, params in CASES:
        image = make_image(seed, H, W)
        params = dict(params)
        if crop is not None:
            params['border_crop'] = crop
        recorded.clear()
        random.seed(seed)
        burst, flow = gen.single2lrburst(image, n, downsample_factor=f, transformation_params=params, interpolation_type='bilinear')
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name + '.npz'), meta=np.array([seed, H, W, n, f, -1 if crop is None else crop]),
                            t_mats=np.stack(recorded), burst_u8=np.rint(burst.numpy() * 255.0).astype(np.uint8), flow=flow.numpy().astype(np.float32))
        print(name, tuple(burst.shape), 'mean', float(burst.mean()), 'max |flow|', float(flow.abs().max()))
        assert np.array_equal(np.float32(np.rint(burst.numpy() * 255.0).astype(np.uint8)) / np.float32(255.0), burst.numpy())


if __name__ == '__main__':
    main()

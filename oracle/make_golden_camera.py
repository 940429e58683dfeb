"""Golden vectors of the inverse camera pipeline from the UNMODIFIED reference functions (build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_camera.py

TEST INFRASTRUCTURE ONLY.  Loads data/camera_pipeline.py from /root/reference (torch / random / math only, no shim) and
calls its functions in the order data/synthetic_burst_generation.py:59-99 does.  `add_noise` draws its standard-normal tensor
from torch's global CPU generator: it is seeded right before the call and the test re-draws the same tensor."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import camera_oracle as C  # noqa: E402

CASES = [('camera_s0_64x96', 0, 64, 96, 3), ('camera_s1_48x40', 1, 48, 40, 2)]


def main():
    sys.dont_write_bytecode = True
    # the file itself is loaded, not the `data` package: data/__init__.py imports data/loader.py, which needs `torch._six`
    # (removed from torch 2.x) and is not on this path
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_camera_pipeline', '/root/reference/data/camera_pipeline.py')
    rgb2raw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rgb2raw)
    for name, seed, h, w, n in CASES:
        image, rgb2cam, (rgb_gain, red_gain, blue_gain), burst_rgb, (shot, read) = C.make_inputs(seed, h, w, n)
        x = rgb2raw.invert_smoothstep(image)
        x = rgb2raw.gamma_expansion(x)
        x = rgb2raw.apply_ccm(x, rgb2cam)
        x = rgb2raw.safe_invert_gains(x, rgb_gain, red_gain, blue_gain)
        lin = x.clamp(0.0, 1.0)
        lin_nogamma = rgb2raw.safe_invert_gains(rgb2raw.apply_ccm(rgb2raw.invert_smoothstep(image), rgb2cam), rgb_gain, red_gain,
                                                blue_gain).clamp(0.0, 1.0)
        raw = rgb2raw.mosaic(burst_rgb.clone())
        torch.manual_seed(500 + seed)
        noisy = rgb2raw.add_noise(raw, shot, read).clamp(0.0, 1.0)
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name + '.npz'), meta=np.array([seed, h, w, n], dtype=np.int64),
                            linear=lin.numpy(), linear_nogamma=lin_nogamma.numpy(), raw=raw.numpy(), noisy=noisy.numpy())
        print(name, 'linear mean', float(lin.mean()), 'saturated frac', float((lin >= 1).float().mean()), 'noisy mean', float(noisy.mean()))


if __name__ == '__main__':
    main()

"""Inverse camera pipeline of the synthetic burst generator on the device (the parameter-deterministic
part of reference data/synthetic_burst_generation.py:23-103 and data/camera_pipeline.py).

`unprocess` is lines :59-79 of `rgb2rawburst` (invert_smoothstep -> gamma_expansion -> apply_ccm -> safe_invert_gains -> clamp)
as ONE kernel; `mosaic_add_noise` is lines :88-99 (mosaic -> add_noise -> clamp) as one kernel over the whole burst.  The
parameter sampling (`random_ccm`, `random_gains`, `random_noise_levels`: a few scalars from Python's / torch's host RNG) is
restated unchanged; the random affine warps and the down-sampling in between (`single2lrburst`: cv2.warpAffine / cv2.resize
on uint8 images, OpenCV's fixed-point interpolation) are host code in the reference and are not provided.  CUDA fp32 tensors
only; CPU tensors raise (no fallback)."""
import math
import random

import torch

from .. import ops


def random_ccm():
    """camera_pipeline.py:28-58: random convex combination of four XYZ -> camera matrices times RGB -> XYZ, rows normalised"""
    xyz2cams = torch.tensor([[[1.0234, -0.2969, -0.2266], [-0.5625, 1.6328, -0.0469], [-0.0703, 0.2188, 0.6406]],
                             [[0.4913, -0.0541, -0.0202], [-0.613, 1.3513, 0.2906], [-0.1564, 0.2151, 0.7183]],
                             [[0.838, -0.263, -0.0639], [-0.2887, 1.0725, 0.2496], [-0.0627, 0.1427, 0.5438]],
                             [[0.6596, -0.2079, -0.0562], [-0.4782, 1.3016, 0.1933], [-0.097, 0.1581, 0.5181]]])
    weights = torch.FloatTensor(len(xyz2cams), 1, 1).uniform_(0.0, 1.0)
    xyz2cam = (xyz2cams * weights).sum(dim=0) / weights.sum()
    rgb2xyz = torch.tensor([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]])
    rgb2cam = torch.mm(xyz2cam, rgb2xyz)
    return rgb2cam / rgb2cam.sum(dim=-1, keepdims=True)


def random_gains():
    """camera_pipeline.py:61-69"""
    rgb_gain = 1.0 / random.gauss(mu=0.8, sigma=0.1)
    return rgb_gain, random.uniform(1.9, 2.4), random.uniform(1.5, 1.9)


def random_noise_levels():
    """camera_pipeline.py:165-175"""
    log_shot_noise = random.uniform(math.log(0.0001), math.log(0.012))
    log_read_noise = 2.18 * log_shot_noise + 1.20 + random.gauss(mu=0.0, sigma=0.26)
    return math.exp(log_shot_noise), math.exp(log_read_noise)


def inverse_gains(rgb_gain, red_gain, blue_gain):
    """the fp32 gain vector of safe_invert_gains (camera_pipeline.py:125)"""
    return (torch.tensor([1.0 / red_gain, 1.0, 1.0 / blue_gain]) / rgb_gain).tolist()


def unprocess(image, rgb2cam, rgb_gain, red_gain, blue_gain, smoothstep=True, gamma=True):
    """sRGB image [3, h, w] (or a batch [b, 3, h, w] sharing the parameters) -> linear sensor-space RGB in [0, 1]"""
    return ops.unprocess_rgb(image, rgb2cam, inverse_gains(rgb_gain, red_gain, blue_gain), smoothstep, gamma)


def mosaic(image):
    """camera_pipeline.py:139-150 (mode 'rggb'); values are clamped to [0, 1], which they already are inside the generator"""
    single = image.dim() == 3
    raw = ops.mosaic_noise(image.unsqueeze(0) if single else image)
    return raw[0] if single else raw


def mosaic_add_noise(image_burst_rgb, shot_noise=0.01, read_noise=0.0005, noise=None):
    """mosaic -> add_noise -> clamp over a whole burst [n, 3, h, w] -> [n, 4, h/2, w/2].  noise: standard-normal tensor of the
    output's shape; None draws it with torch on the device (the reference draws it on the host, camera_pipeline.py:181)."""
    n, _, h, w = image_burst_rgb.shape
    if noise is None:
        noise = torch.randn(n, 4, h // 2, w // 2, dtype=torch.float32, device=image_burst_rgb.device)
    return ops.mosaic_noise(image_burst_rgb, shot_noise, read_noise, noise)

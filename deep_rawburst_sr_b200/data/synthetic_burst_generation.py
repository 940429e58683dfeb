"""Synthetic RAW burst generation on the device with the reference's interface (data/synthetic_burst_generation.py:
`rgb2rawburst` :23-103, `get_tmat` :106-128, `single2lrburst` :131-246) -- SURVEY.md 8(f) rank 4.

All random draws stay on the host, in the reference's order and from the same generators (Python `random` for gains, frame
transforms and noise levels; torch's CPU generator for the colour matrix and the noise field), so seeding both as the
reference's caller would reproduces the same burst.  The pixel work runs in three fused kernels: `dbsr_unprocess_rgb`
(inverse camera pipeline), `dbsr_single2lrburst` (uint8 quantisation + cv2.warpAffine + border crop + cv2.resize + / 255,
bit-exact against OpenCV's fixed-point arithmetic, no OpenCV needed) and `dbsr_mosaic_noise`.  Bilinear interpolation only
(`interpolation_type='lanczos'` raises); CUDA fp32 tensors only, CPU tensors raise (no fallback)."""
import math
import random

import numpy as np
import torch

from .. import ops
from . import camera_pipeline as rgb2raw


def get_tmat(image_shape, translation, theta, shear_values, scale_factors):
    """ Generates a transformation matrix corresponding to the input transformation parameters (:106-128; the rotation block is
    cv2.getRotationMatrix2D((w / 2, h / 2), theta, 1.0) written out) """
    im_h, im_w = image_shape
    t_mat = np.identity(3)
    t_mat[0, 2] = translation[0]
    t_mat[1, 2] = translation[1]
    a, b = math.cos(theta * math.pi / 180.0), math.sin(theta * math.pi / 180.0)
    cx, cy = im_w * 0.5, im_h * 0.5
    t_rot = np.array([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy], [0.0, 0.0, 1.0]])
    t_shear = np.array([[1.0, shear_values[0], -shear_values[0] * 0.5 * im_w],
                        [shear_values[1], 1.0, -shear_values[1] * 0.5 * im_h],
                        [0.0, 0.0, 1.0]])
    t_scale = np.array([[scale_factors[0], 0.0, 0.0], [0.0, scale_factors[1], 0.0], [0.0, 0.0, 1.0]])
    return (t_scale @ t_rot @ t_shear @ t_mat)[:2, :]


def sample_transforms(image_shape, burst_size, downsample_factor=1, transformation_params=None):
    """the per-frame parameter sampling of single2lrburst (:171-207), same `random` call order; returns the 2x3 matrices"""
    transformation_params = transformation_params or {}
    t_mats = []
    for i in range(burst_size):
        if i == 0:
            shift = (downsample_factor / 2.0) - 0.5
            translation, theta, shear_factor, scale_factor = (shift, shift), 0.0, (0.0, 0.0), (1.0, 1.0)
        else:
            max_translation = transformation_params.get('max_translation', 0.0)
            if max_translation <= 0.01:
                shift = (downsample_factor / 2.0) - 0.5
                translation = (shift, shift)
            else:
                translation = (random.uniform(-max_translation, max_translation), random.uniform(-max_translation, max_translation))
            max_rotation = transformation_params.get('max_rotation', 0.0)
            theta = random.uniform(-max_rotation, max_rotation)
            max_shear = transformation_params.get('max_shear', 0.0)
            shear_factor = (random.uniform(-max_shear, max_shear), random.uniform(-max_shear, max_shear))
            max_ar_factor = transformation_params.get('max_ar_factor', 0.0)
            ar_factor = np.exp(random.uniform(-max_ar_factor, max_ar_factor))
            max_scale = transformation_params.get('max_scale', 0.0)
            scale_factor = np.exp(random.uniform(-max_scale, max_scale))
            scale_factor = (scale_factor, scale_factor * ar_factor)
        t_mats.append(get_tmat(image_shape, translation, theta, shear_factor, scale_factor))
    return t_mats


def _inverse_map(t_mat):
    """the double-precision inversion cv::warpAffine applies to a forward 2x3 matrix"""
    M = np.array(t_mat, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def lrburst_from_transforms(image, t_mats, downsample_factor=1, border_crop=None, normalize=None):
    """warp / crop / down-sample `image` [3, H, W] by the given 2x3 matrices -> (burst [n, 3, h, w], flow_vectors [n, 2, h, w])"""
    if normalize is None:
        normalize = bool(image.max() < 2.0)                 # the reference's data-dependent test (:154), one host read
    inv = torch.from_numpy(np.stack([_inverse_map(m) for m in t_mats]))
    # fp32 inverse of the 3x3 forward matrices (:214-215), all frames in one batched call
    t3 = torch.zeros(len(t_mats), 3, 3)
    t3[:, :2, :] = torch.from_numpy(np.stack([np.asarray(m) for m in t_mats])).float()
    t3[:, 2, 2] = 1.0
    pos = torch.linalg.inv(t3)[:, :2, :].reshape(len(t_mats), 6)
    return ops.single2lrburst(image, inv, pos, int(downsample_factor), int(border_crop or 0), normalize)


def single2lrburst(image, burst_size, downsample_factor=1, transformation_params=None, interpolation_type='bilinear'):
    """ Generates a burst of size burst_size from the input image by applying random transformations defined by
    transformation_params, and downsampling the resulting burst by downsample_factor (:131-246) """
    if interpolation_type != 'bilinear':
        raise NotImplementedError("only interpolation_type='bilinear' (cv2.INTER_LINEAR) is implemented")
    transformation_params = transformation_params or {}
    t_mats = sample_transforms(tuple(image.shape[-2:]), burst_size, downsample_factor, transformation_params)
    return lrburst_from_transforms(image, t_mats, downsample_factor, transformation_params.get('border_crop'))


def rgb2rawburst(image, burst_size, downsample_factor=1, burst_transformation_params=None, image_processing_params=None,
                 interpolation_type='bilinear'):
    """ Generates a synthetic LR RAW burst from the input image (:23-103): inverse camera pipeline -> random burst ->
    mosaic -> noise.  Returns (image_burst, image, image_burst_rgb, flow_vectors, meta_info) like the reference. """
    if image_processing_params is None:
        image_processing_params = {}
    for k, v in {'random_ccm': True, 'random_gains': True, 'smoothstep': True, 'gamma': True, 'add_noise': True}.items():
        image_processing_params.setdefault(k, v)
    rgb2cam = rgb2raw.random_ccm() if image_processing_params['random_ccm'] else torch.eye(3).float()
    cam2rgb = rgb2cam.inverse()
    rgb_gain, red_gain, blue_gain = rgb2raw.random_gains() if image_processing_params['random_gains'] else (1.0, 1.0, 1.0)
    use_smoothstep, use_gamma = image_processing_params['smoothstep'], image_processing_params['gamma']
    image = rgb2raw.unprocess(image, rgb2cam, rgb_gain, red_gain, blue_gain, use_smoothstep, use_gamma)
    if interpolation_type != 'bilinear':
        raise NotImplementedError("only interpolation_type='bilinear' (cv2.INTER_LINEAR) is implemented")
    # single2lrburst, with its `image.max() < 2` test (:154) known to hold: the image was just clamped to [0, 1] (no host read)
    tp = burst_transformation_params or {}
    t_mats = sample_transforms(tuple(image.shape[-2:]), burst_size, downsample_factor, tp)
    image_burst_rgb, flow_vectors = lrburst_from_transforms(image, t_mats, downsample_factor, tp.get('border_crop'), normalize=True)
    if image_processing_params['add_noise']:
        shot_noise_level, read_noise_level = rgb2raw.random_noise_levels()
        n, _, h, w = image_burst_rgb.shape
        noise = torch.FloatTensor(n, 4, h // 2, w // 2).normal_()       # host draw, as add_noise does (camera_pipeline.py:181)
        image_burst = rgb2raw.mosaic_add_noise(image_burst_rgb, shot_noise_level, read_noise_level, noise.to(image.device))
    else:
        shot_noise_level, read_noise_level = 0, 0
        image_burst = rgb2raw.mosaic(image_burst_rgb)
    meta_info = {'rgb2cam': rgb2cam, 'cam2rgb': cam2rgb, 'rgb_gain': rgb_gain, 'red_gain': red_gain, 'blue_gain': blue_gain,
                 'smoothstep': use_smoothstep, 'gamma': use_gamma, 'shot_noise_level': shot_noise_level,
                 'read_noise_level': read_noise_level}
    return image_burst, image, image_burst_rgb, flow_vectors, meta_info

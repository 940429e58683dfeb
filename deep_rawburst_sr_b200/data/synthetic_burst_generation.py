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


def sample_transform_params(burst_size, downsample_factor=1, transformation_params=None):
    """the random draws of `sample_transforms` (same `random` call order) without building the matrices:
    list of (translation, theta, shear_factor, scale_factor) per frame"""
    transformation_params = transformation_params or {}
    shift = (downsample_factor / 2.0) - 0.5
    out = [((shift, shift), 0.0, (0.0, 0.0), (1.0, 1.0))]
    max_translation = transformation_params.get('max_translation', 0.0)
    max_rotation = transformation_params.get('max_rotation', 0.0)
    max_shear = transformation_params.get('max_shear', 0.0)
    max_ar_factor = transformation_params.get('max_ar_factor', 0.0)
    max_scale = transformation_params.get('max_scale', 0.0)
    for _ in range(1, burst_size):
        if max_translation <= 0.01:
            translation = (shift, shift)
        else:
            translation = (random.uniform(-max_translation, max_translation), random.uniform(-max_translation, max_translation))
        theta = random.uniform(-max_rotation, max_rotation)
        shear_factor = (random.uniform(-max_shear, max_shear), random.uniform(-max_shear, max_shear))
        ar_factor = np.exp(random.uniform(-max_ar_factor, max_ar_factor))
        scale_factor = np.exp(random.uniform(-max_scale, max_scale))
        out.append((translation, theta, shear_factor, (scale_factor, scale_factor * ar_factor)))
    return out


def get_tmat_batch(image_shape, params):
    """`get_tmat` for a list of (translation, theta, shear_values, scale_factors) in one stacked numpy pass -> [m, 2, 3] float64;
    the same products in the same order as the per-frame function (bit-identical, tests/test_host_logic.py)"""
    im_h, im_w = image_shape
    m = len(params)
    tr = np.array([p[0] for p in params], dtype=np.float64).reshape(m, 2)
    sh = np.array([p[2] for p in params], dtype=np.float64).reshape(m, 2)
    sc = np.array([p[3] for p in params], dtype=np.float64).reshape(m, 2)
    a = np.array([math.cos(p[1] * math.pi / 180.0) for p in params])
    b = np.array([math.sin(p[1] * math.pi / 180.0) for p in params])
    cx, cy = im_w * 0.5, im_h * 0.5
    t_mat = np.tile(np.identity(3), (m, 1, 1))
    t_mat[:, 0, 2], t_mat[:, 1, 2] = tr[:, 0], tr[:, 1]
    t_rot = np.zeros((m, 3, 3))
    t_rot[:, 0, 0], t_rot[:, 0, 1], t_rot[:, 0, 2] = a, b, (1 - a) * cx - b * cy
    t_rot[:, 1, 0], t_rot[:, 1, 1], t_rot[:, 1, 2] = -b, a, b * cx + (1 - a) * cy
    t_rot[:, 2, 2] = 1.0
    t_shear = np.tile(np.identity(3), (m, 1, 1))
    t_shear[:, 0, 1], t_shear[:, 0, 2] = sh[:, 0], -sh[:, 0] * 0.5 * im_w
    t_shear[:, 1, 0], t_shear[:, 1, 2] = sh[:, 1], -sh[:, 1] * 0.5 * im_h
    t_scale = np.zeros((m, 3, 3))
    t_scale[:, 0, 0], t_scale[:, 1, 1], t_scale[:, 2, 2] = sc[:, 0], sc[:, 1], 1.0
    return (t_scale @ t_rot @ t_shear @ t_mat)[:, :2, :]


def _inverse_map_batch(t_mats):
    """`_inverse_map` for stacked forward matrices [m, 2, 3] -> [m, 6] float64, the same operations element by element"""
    M = np.array(t_mats, dtype=np.float64).reshape(-1, 6).copy()
    D = M[:, 0] * M[:, 4] - M[:, 1] * M[:, 3]
    with np.errstate(divide='ignore'):
        D = np.where(D != 0, 1.0 / D, 0.0)
    A11, A22 = M[:, 4] * D, M[:, 0] * D
    M[:, 0] = A11
    M[:, 1] *= -D
    M[:, 3] *= -D
    M[:, 4] = A22
    b1 = -M[:, 0] * M[:, 2] - M[:, 1] * M[:, 5]
    b2 = -M[:, 3] * M[:, 2] - M[:, 4] * M[:, 5]
    M[:, 2], M[:, 5] = b1, b2
    return M


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


def rgb2rawburst_batch(images, burst_size, downsample_factor=1, burst_transformation_params=None, image_processing_params=None,
                       interpolation_type='bilinear', noise='device'):
    """`rgb2rawburst` for a BATCH of images [b, 3, H, W] in three kernel launches (feed the forward at
    > 3 000 bursts/s from one process).  The random parameters of every burst -- colour matrix, gains, 13 frame transforms, noise
    levels -- are drawn on the host exactly as the per-burst function draws them (same generators, same call order, burst after
    burst), packed, and uploaded in ONE copy per parameter table; the kernels read them per burst from device memory.

    noise='host' : the standard-normal field is drawn with torch's CPU generator like the reference's `add_noise`
                   (camera_pipeline.py:181); under the same seeds the batch equals `rgb2rawburst` called burst after burst, bit
                   for bit (tests/test_gpu_camera.py).  The draw costs ~0.4 ms per burst on the host.
    noise='device': the field is drawn with torch's CUDA generator (same distribution, different stream); the host work left is
                   ~0.15 ms per burst.
    Returns (image_burst [b, n, 4, h/2, w/2], image [b, 3, H, W], image_burst_rgb [b, n, 3, h, w], flow_vectors [b, n, 2, h, w],
    meta_info: list of b dicts with the keys of the per-burst function)."""
    if interpolation_type != 'bilinear':
        raise NotImplementedError("only interpolation_type='bilinear' (cv2.INTER_LINEAR) is implemented")
    if noise not in ('host', 'device'):
        raise ValueError("noise must be 'host' or 'device'")
    ops.require_device(images)
    assert images.dim() == 4 and images.shape[1] == 3
    ipp = dict(image_processing_params or {})
    for k, v in {'random_ccm': True, 'random_gains': True, 'smoothstep': True, 'gamma': True, 'add_noise': True}.items():
        ipp.setdefault(k, v)
    tp = burst_transformation_params or {}
    b, _, H, W = images.shape
    crop = int(tp.get('border_crop') or 0)
    f = int(downsample_factor)
    h, w = (H - 2 * crop) // f, (W - 2 * crop) // f
    metas, tparams, host_noise = [], [], []
    params12 = torch.empty(b, 12)
    levels = torch.zeros(b, 2)
    for i in range(b):          # scalar draws only, in the order of `rgb2rawburst`
        rgb2cam = rgb2raw.random_ccm() if ipp['random_ccm'] else torch.eye(3).float()
        rgb_gain, red_gain, blue_gain = rgb2raw.random_gains() if ipp['random_gains'] else (1.0, 1.0, 1.0)
        params12[i, :9] = rgb2cam.reshape(-1)
        params12[i, 9:] = torch.tensor(rgb2raw.inverse_gains(rgb_gain, red_gain, blue_gain))
        tparams += sample_transform_params(burst_size, f, tp)
        shot, read = (rgb2raw.random_noise_levels() if ipp['add_noise'] else (0, 0))
        levels[i, 0], levels[i, 1] = shot, read
        if ipp['add_noise'] and noise == 'host':
            host_noise.append(torch.FloatTensor(burst_size, 4, h // 2, w // 2).normal_())
        metas.append({'rgb2cam': rgb2cam, 'cam2rgb': rgb2cam.inverse(), 'rgb_gain': rgb_gain, 'red_gain': red_gain, 'blue_gain': blue_gain,
                      'smoothstep': ipp['smoothstep'], 'gamma': ipp['gamma'], 'shot_noise_level': shot, 'read_noise_level': read})
    t_mats = get_tmat_batch((H, W), tparams)                                   # [b * n, 2, 3]
    inv = torch.from_numpy(_inverse_map_batch(t_mats))
    t3 = torch.zeros(len(tparams), 3, 3)
    t3[:, :2, :] = torch.from_numpy(t_mats).float()
    t3[:, 2, 2] = 1.0
    pos = torch.linalg.inv(t3)[:, :2, :].reshape(-1, 6).contiguous()
    dev = images.device
    image = ops.unprocess_rgb_batch(images.float(), params12.to(dev, non_blocking=True), ipp['smoothstep'], ipp['gamma'])
    burst_rgb, flow = ops.single2lrburst_batch(image, inv.to(dev, non_blocking=True), pos.to(dev, non_blocking=True), burst_size, f, crop, True)
    if ipp['add_noise']:
        z = torch.cat(host_noise).to(dev, non_blocking=True) if noise == 'host' else \
            torch.randn(b * burst_size, 4, h // 2, w // 2, dtype=torch.float32, device=dev)
        raw = ops.mosaic_noise_batch(burst_rgb, levels.to(dev, non_blocking=True), burst_size, z)
    else:
        raw = ops.mosaic_noise(burst_rgb)
    return (raw.view(b, burst_size, 4, h // 2, w // 2), image, burst_rgb.view(b, burst_size, 3, h, w),
            flow.view(b, burst_size, 2, h, w), metas)

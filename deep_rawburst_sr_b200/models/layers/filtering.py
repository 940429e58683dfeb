"""gauss_1d / gauss_2d / get_gaussian_kernel / apply_kernel as in the reference (models/layers/filtering.py:20-62)."""
import math

import torch


def gauss_1d(sz, sigma, center, end_pad=0, density=False):
    k = torch.arange(-(sz - 1) / 2, (sz + 1) / 2 + end_pad).reshape(1, -1)
    gauss = torch.exp(-1.0 / (2 * sigma ** 2) * (k - center.reshape(-1, 1)) ** 2)
    if density:
        gauss /= math.sqrt(2 * math.pi) * sigma
    return gauss


def gauss_2d(sz, sigma, center, end_pad=(0, 0), density=False):
    if isinstance(sigma, (float, int)):
        sigma = (sigma, sigma)
    if isinstance(sz, int):
        sz = (sz, sz)
    if isinstance(center, (list, tuple)):
        center = torch.tensor(center).view(1, 2)
    return gauss_1d(sz[0], sigma[0], center[:, 0], end_pad[0], density).reshape(center.shape[0], 1, -1) * \
        gauss_1d(sz[1], sigma[1], center[:, 1], end_pad[1], density).reshape(center.shape[0], -1, 1)


def get_gaussian_kernel(sd, ksz=None):
    """ Returns a 2D Gaussian kernel with standard deviation sd (filtering.py:43-52) """
    if ksz is None:
        ksz = int(4 * sd + 1)
    assert ksz % 2 == 1
    K = gauss_2d(ksz, sd, (0.0, 0.0), density=True)
    K = K / K.sum()
    return K.unsqueeze(0), ksz


def apply_kernel(im, ksz, kernel):
    """ apply the provided kernel on input image (filtering.py:55-62): reflect padding + per-channel correlation.
    Written as an explicit sum over the ksz^2 taps so that the result is exact fp32 on the device (a cudnn convolution
    would run in TF32 by default); the images on this path are the 80x80 low-resolution frames. """
    import torch.nn.functional as F
    shape = im.shape
    im = im.reshape(-1, 1, *im.shape[-2:])
    r = ksz // 2
    imp = F.pad(im, [r, r, r, r], mode='reflect')
    H, W = shape[-2:]
    k = kernel.reshape(ksz, ksz).to(im.device, im.dtype)
    out = torch.zeros_like(im)
    for dy in range(ksz):
        for dx in range(ksz):
            out = out + k[dy, dx] * imp[..., dy:dy + H, dx:dx + W]
    return out.view(shape)

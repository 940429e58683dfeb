"""gauss_1d / gauss_2d / get_gaussian_kernel / apply_kernel as in the reference (models/layers/filtering.py:20-62)."""
import math

import torch


def gauss_1d(sz, sigma, center, end_pad=0, density=False):
    """rows of exp(-(k - c)^2 / (2 sigma^2)) on the symmetric grid k = -(sz-1)/2 ... (sz-1)/2 (+ end_pad), one per centre c"""
    half = (sz - 1) / 2
    grid = torch.arange(-half, half + 1 + end_pad).reshape(1, -1)
    values = torch.exp((grid - center.reshape(-1, 1)) ** 2 * (-1.0 / (2 * sigma ** 2)))
    return values / (math.sqrt(2 * math.pi) * sigma) if density else values


def gauss_2d(sz, sigma, center, end_pad=(0, 0), density=False):
    """outer product of two 1-D Gaussians per centre: [n_centres, sz_y (+pad), sz_x (+pad)]"""
    sigma = (sigma, sigma) if isinstance(sigma, (float, int)) else sigma
    sz = (sz, sz) if isinstance(sz, int) else sz
    if isinstance(center, (list, tuple)):
        center = torch.tensor(center).view(1, 2)
    n = center.shape[0]
    along_x = gauss_1d(sz[0], sigma[0], center[:, 0], end_pad[0], density).reshape(n, 1, -1)
    along_y = gauss_1d(sz[1], sigma[1], center[:, 1], end_pad[1], density).reshape(n, -1, 1)
    return along_x * along_y


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

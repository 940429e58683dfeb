"""SSIM / MS-SSIM with the reference's interface (models/loss/msssim.py: `gaussian` :10-12, `create_window` :15-19, `ssim`
:22-74, `msssim` :77-104, `SSIM` :107-130, `MSSSIM` :133-142) on one fused sm_100a kernel per level (`dbsr_ssim`): both images
are read once, the 11 x 11 Gaussian window is applied separably to the five moments in shared memory, and only the
per-image means (or, with `spatial_out`, the map) leave the SM.  The data-dependent value range (:24-35) is derived on the
device, so no call synchronises with the host.  CUDA fp32 tensors only -- CPU tensors raise (no fallback)."""
from math import exp

import torch

from ... import ops

_MAX_WINDOW = 11


def gaussian(window_size, sigma):
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel=1):
    _1D_window = gaussian(window_size, 1.5).unsqueeze(1)
    _2D_window = _1D_window.mm(_1D_window.t()).float().unsqueeze(0).unsqueeze(0)
    return _2D_window.expand(channel, 1, window_size, window_size).contiguous()


_verified_windows = {}


def _window_taps(window, window_size, height, width):
    """1-D taps of the window the reference would convolve with.  A caller-supplied `window` must be the reference's own
    `create_window(k, C)` (the only window its classes build); anything else is refused rather than approximated."""
    if window is None:
        k = min(window_size, height, width)
    else:
        if window.dim() != 4 or window.shape[1] != 1 or window.shape[2] != window.shape[3]:
            raise ValueError(f'ssim: unsupported window of shape {tuple(window.shape)}')
        k = int(window.shape[2])
        key = (window.data_ptr(), tuple(window.shape), window.device, window._version)
        if key not in _verified_windows:
            ok = torch.allclose(window[0, 0].detach().float().cpu(), create_window(k)[0, 0], rtol=0, atol=1e-8)
            _verified_windows.clear()
            _verified_windows[key] = ok
        if not _verified_windows[key]:
            raise NotImplementedError('ssim: only the Gaussian window of create_window (sigma 1.5) is implemented')
    if k > _MAX_WINDOW:
        raise NotImplementedError(f'ssim: window sizes up to {_MAX_WINDOW} are implemented (reference default 11), got {k}')
    return gaussian(k, 1.5).tolist()


def _stats(img1, img2, window_size, window, val_range, want_map=False, crop=0, fixed_window=False, valid=None):
    if img1.dim() != 4:
        raise ValueError('ssim expects [n, c, h, w] tensors')
    if fixed_window:      # the SSIM class always convolves with its window_size window (:121-130), never min(11, h, w)
        if window_size > _MAX_WINDOW:
            raise NotImplementedError(f'ssim: window sizes up to {_MAX_WINDOW} are implemented, got {window_size}')
        taps = gaussian(window_size, 1.5).tolist()
    else:
        taps = _window_taps(window, window_size, img1.shape[2] - 2 * crop, img1.shape[3] - 2 * crop)
    return ops.ssim_stats(img1.contiguous(), img2.contiguous(), taps, crop=crop, val_range=val_range, want_map=want_map, valid=valid)


def ssim(img1, img2, window_size=11, window=None, size_average=True, full=False, val_range=None, spatial_out=False):
    stats, smap = _stats(img1, img2, window_size, window, val_range, want_map=spatial_out)
    cs = stats[:, 1].mean()
    if spatial_out:
        ret = smap
    elif size_average:
        ret = stats[:, 0].mean()
    else:
        ret = stats[:, 0]
    if full:
        return ret, cs
    return ret


def msssim(img1, img2, window_size=11, size_average=True, val_range=None, normalize=False):
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float32, device=img1.device)
    levels = weights.size()[0]
    mssim, mcs = [], []
    img1, img2 = img1.contiguous(), img2.contiguous()
    for level in range(levels):
        sim, cs = ssim(img1, img2, window_size=window_size, size_average=size_average, full=True, val_range=val_range)
        mssim.append(sim)
        mcs.append(cs)
        if level + 1 < levels:       # the reference also pools after the last level (:88-89); that result is never read
            img1, img2 = ops.avgpool2_pair(img1, img2)
    mssim = torch.stack(mssim)
    mcs = torch.stack(mcs)
    if normalize:
        mssim = (mssim + 1) / 2
        mcs = (mcs + 1) / 2
    pow1 = mcs ** weights
    pow2 = mssim ** weights
    return torch.prod(pow1[:-1] * pow2[-1])


class SSIM(torch.nn.Module):
    def __init__(self, window_size=11, size_average=True, val_range=None, spatial_out=False):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.val_range = val_range
        self.spatial_out = spatial_out
        self.channel = 1
        self.window = create_window(window_size)

    def forward(self, img1, img2, crop=0):
        """`crop` (extension): boundary_ignore applied inside the kernel instead of slicing copies."""
        # (the reference's forward does not hand `val_range` on to ssim(), :129-130: the range is always data-derived)
        stats, smap = _stats(img1, img2, self.window_size, None, None, want_map=self.spatial_out, crop=crop, fixed_window=True)
        if self.spatial_out:
            return smap
        return stats[:, 0].mean() if self.size_average else stats[:, 0]


class MSSSIM(torch.nn.Module):
    def __init__(self, window_size=11, size_average=True, channel=3):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = channel

    def forward(self, img1, img2):
        return msssim(img1, img2, window_size=self.window_size, size_average=self.size_average)

"""PixShuffleUpsampler parameter container (reference models/layers/upsampling.py:22-66).  On the fast path the
1x1 conv + ReLU + PixelShuffle is one tcgen05 kernel whose epilogue stores in shuffled order, followed by the 3x3
Gaussian (`dbsr_blur3x3`)."""
import torch.nn as nn

from . import blocks
from .filtering import gauss_2d
from .initializations import ICNR


class PixShuffleUpsampler(nn.Module):
    """Parameters: `conv_layer` (1x1 conv to output_dim * r^2 channels, no bias under ICNR init, + activation); attributes read
    by the engine: `gauss_kernel` ([1, k, k] normalised Gaussian or None -- a plain attribute, not in the state_dict, as in
    the reference), `gauss_ksz`, `upsample_factor`."""

    @staticmethod
    def _get_gaussian_kernel(ksz, sd):
        if ksz % 2 != 1:
            raise AssertionError('the blur kernel size must be odd')
        weights = gauss_2d(ksz, sd, (0.0, 0.0), density=True)
        return weights / weights.sum()

    def __init__(self, input_dim, output_dim, upsample_factor=2, use_bn=False, activation='relu',
                 icnrinit=False, gauss_blur_sd=None, gauss_ksz=3):
        super().__init__()
        self.upsample_factor, self.gauss_ksz = upsample_factor, gauss_ksz
        self.conv_layer = blocks.conv_block(input_dim, output_dim * upsample_factor ** 2, 1, stride=1, padding=0,
                                            batch_norm=use_bn, activation=activation, bias=not icnrinit)
        if icnrinit:            # sub-pixel groups start identical: no checkerboard at initialisation
            conv = self.conv_layer[0]
            conv.weight.data.copy_(ICNR(conv.weight, upsample_factor))
        self.gauss_kernel = None if gauss_blur_sd is None else self._get_gaussian_kernel(gauss_ksz, gauss_blur_sd).unsqueeze(0)
        self.pix_shuffle = nn.PixelShuffle(upsample_factor)

    def forward(self, x):
        raise NotImplementedError('PixShuffleUpsampler runs as part of ResPixShuffleConv on the B200 engine')

"""PixShuffleUpsampler parameter container (reference models/layers/upsampling.py:22-66).  On the fast path the
1x1 conv + ReLU + PixelShuffle is one tcgen05 kernel whose epilogue stores in shuffled order, followed by the 3x3
Gaussian (`dbsr_blur3x3`)."""
import torch.nn as nn

from . import blocks
from .filtering import gauss_2d
from .initializations import ICNR


class PixShuffleUpsampler(nn.Module):
    @staticmethod
    def _get_gaussian_kernel(ksz, sd):
        assert ksz % 2 == 1
        K = gauss_2d(ksz, sd, (0.0, 0.0), density=True)
        K = K / K.sum()
        return K

    def __init__(self, input_dim, output_dim, upsample_factor=2, use_bn=False, activation='relu',
                 icnrinit=False, gauss_blur_sd=None, gauss_ksz=3):
        super().__init__()
        pre_shuffle_dim = output_dim * upsample_factor ** 2
        self.conv_layer = blocks.conv_block(input_dim, pre_shuffle_dim, 1, stride=1, padding=0, batch_norm=use_bn,
                                            activation=activation, bias=not icnrinit)
        if icnrinit:
            kernel = ICNR(self.conv_layer[0].weight, upsample_factor)
            self.conv_layer[0].weight.data.copy_(kernel)
        if gauss_blur_sd is not None:
            self.gauss_kernel = self._get_gaussian_kernel(gauss_ksz, gauss_blur_sd).unsqueeze(0)
        else:
            self.gauss_kernel = None
        self.gauss_ksz = gauss_ksz
        self.upsample_factor = upsample_factor
        self.pix_shuffle = nn.PixelShuffle(upsample_factor)

    def forward(self, x):
        raise NotImplementedError('PixShuffleUpsampler runs as part of ResPixShuffleConv on the B200 engine')

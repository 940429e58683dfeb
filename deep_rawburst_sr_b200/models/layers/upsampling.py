"""PixShuffleUpsampler parameter container (reference models/layers/upsampling.py:22-66).  On the fast path the
1x1 conv + ReLU + PixelShuffle is one tcgen05 kernel whose epilogue stores in shuffled order, followed by the 3x3
Gaussian (`dbsr_blur3x3`)."""
import torch
import torch.nn as nn

from ... import ops
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

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, x):
        """upsampling.py:51-66 standalone (NCHW fp32 seam): 1x1 conv (+ bias) + activation with the PixelShuffle folded into the
        store addressing (`dbsr_conv2d_direct`, shuffle_r), then the per-channel 3x3 Gaussian (`dbsr_blur3x3`).  Inside
        `ResPixShuffleConv` the same two steps run on the tensor-core kernel of the engine."""
        assert x.dim() == 4
        ops.require_device(x)
        conv = self.conv_layer[0]
        rest = list(self.conv_layer)[1:]
        if len(rest) == 0:
            act = ops.ACT_NONE
        elif len(rest) == 1 and isinstance(rest[0], nn.ReLU):
            act = ops.ACT_RELU
        else:
            raise NotImplementedError('PixShuffleUpsampler on the B200 kernels covers activation "relu" / "none" without batch norm')
        n, c, h, w = x.shape
        r = self.upsample_factor
        cout = conv.out_channels // (r * r)
        xa = ops.Act.empty(n, h, w, c, torch.float32, x.device).from_nchw(x.contiguous().float())
        ya = ops.Act.empty(n, h * r, w * r, cout, torch.float32, x.device)
        wt = conv.weight.detach().float().permute(2, 3, 1, 0).reshape(1, c, conv.out_channels).contiguous()
        b = None if conv.bias is None else conv.bias.detach().float().contiguous()
        ops.conv2d(xa, wt, b, ya, 1, 1, 1, act, None, shuffle_r=r)
        gk = getattr(self, 'gauss_kernel', None)
        if gk is not None:
            if getattr(self, 'gauss_ksz', 3) != 3:
                raise NotImplementedError('the blur kernel of the B200 path is 3x3')
            yb = ops.Act.empty(n, h * r, w * r, cout, torch.float32, x.device)
            ops.blur3x3(ya, yb, [float(v) for v in gk.reshape(-1).tolist()])
            ya = yb
        return ya.to_nchw()

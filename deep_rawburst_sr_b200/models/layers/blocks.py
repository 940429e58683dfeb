"""Parameter containers for the DBSR convolution stacks, interface-compatible with the reference's
models/layers/blocks.py (`get_activation` :19-36, `conv_block` :46-60, `ResBlock` :63-96): same constructor arguments and
the same sub-module layout (`<block>.0` = convolution, then optional batch norm / activation; `conv1` / `conv2` inside a
residual block), so `state_dict()` keys match reference checkpoints.  The fast path never calls these modules' forward:
the owning network hands its state_dict to `DBSREngine`.  Called standalone, a convolution runs the CUDA-core kernel
through the C ABI (NCHW seam), never cuDNN."""
import torch
from torch import nn

from ... import ops


class Conv2dB200(nn.Conv2d):
    """nn.Conv2d parameters; forward = dbsr_conv2d_direct (fp32, exact)."""

    def forward(self, x):
        ops.require_device(x)
        assert self.groups == 1 and self.padding_mode == 'zeros'
        k, s, d = self.kernel_size[0], self.stride[0], self.dilation[0]
        assert k in (1, 3) and self.padding[0] == d * (k - 1) // 2, 'only "same"-style padding is supported'
        n, c, h, w = x.shape
        xa = ops.Act.empty(n, h, w, c, torch.float32, x.device).from_nchw(x.contiguous().float())
        ho = (h + 2 * self.padding[0] - d * (k - 1) - 1) // s + 1
        wo = (w + 2 * self.padding[0] - d * (k - 1) - 1) // s + 1
        ya = ops.Act.empty(n, ho, wo, self.out_channels, torch.float32, x.device)
        wt = self.weight.detach().float().permute(2, 3, 1, 0).reshape(k * k, c, self.out_channels).contiguous()
        b = None if self.bias is None else self.bias.detach().float().contiguous()
        ops.conv2d(xa, wt, b, ya, k, s, d, ops.ACT_NONE)
        return ya.to_nchw()


# activation name -> factory(params, channels); 'none' yields no module at all
_ACTIVATION_FACTORIES = {
    'relu': lambda params, channels: nn.ReLU(inplace=True),
    'lrelu': lambda params, channels: nn.LeakyReLU(negative_slope=params.get('negative_slope', 0.1), inplace=True),
    'prelu': lambda params, channels: nn.PReLU(num_parameters=channels),
    'sigmoid': lambda params, channels: nn.Sigmoid(),
    'tanh': lambda params, channels: nn.Tanh(),
    'none': lambda params, channels: None,
}


def get_activation(activation, activation_params=None, num_channels=None):
    try:
        factory = _ACTIVATION_FACTORIES[activation]
    except KeyError:
        raise Exception('Unknown activation {}'.format(activation)) from None
    return factory(activation_params or {}, num_channels)


def get_attention(attention_type, num_channels=None):
    if attention_type != 'none':
        raise Exception('Unknown attention {}'.format(attention_type))
    return None


def conv_block(in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1, bias=True,
               batch_norm=False, activation='relu', padding_mode='zeros', activation_params=None):
    """[convolution, batch norm?, activation?] as one nn.Sequential (index 0 is always the convolution)"""
    stages = (Conv2dB200(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation,
                         bias=bias, padding_mode=padding_mode),
              nn.BatchNorm2d(out_planes) if batch_norm else None,
              get_activation(activation, activation_params, num_channels=out_planes))
    return nn.Sequential(*(m for m in stages if m is not None))


class ResBlock(nn.Module):
    """act(x + conv2(act(conv1(x)))) with two 3x3 convolutions (`conv1` carries the activation, `conv2` none)"""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, batch_norm=False, activation='relu',
                 padding_mode='zeros', attention='none'):
        super().__init__()
        shared = dict(kernel_size=3, padding=1, dilation=dilation, batch_norm=batch_norm, padding_mode=padding_mode)
        self.conv1 = conv_block(inplanes, planes, stride=stride, activation=activation, **shared)
        self.conv2 = conv_block(planes, planes, activation='none', **shared)
        self.downsample, self.stride = downsample, stride
        self.activation = get_activation(activation, num_channels=planes)
        self.attention = get_attention(attention_type=attention, num_channels=planes)

    def forward(self, x):
        branch = self.conv2(self.conv1(x))
        if self.attention is not None:
            branch = self.attention(branch)
        skip = x if self.downsample is None else self.downsample(x)
        total = branch + skip
        return total if self.activation is None else self.activation(total)

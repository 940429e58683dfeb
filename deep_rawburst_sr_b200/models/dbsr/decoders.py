"""ResPixShuffleConv with the reference's interface (models/dbsr/decoders.py:20-62)."""
import torch
import torch.nn as nn

from ... import ops
from ...engine import DBSREngine
from ..engine_owner import EngineOwner
from ..layers import blocks
from ..layers.upsampling import PixShuffleUpsampler


class ResPixShuffleConv(EngineOwner, nn.Module):
    """Residual decoder with sub-pixel-convolution upsampling.  forward({'fused_enc' [B, C, H, W]}) ->
    {'pred' [B, 3, r*H, r*W]}."""

    def __init__(self, input_dim, init_conv_dim, num_pre_res_blocks, post_conv_dim, num_post_res_blocks, use_bn=False,
                 activation='relu', upsample_factor=2, icnrinit=False, gauss_blur_sd=None, gauss_ksz=3):
        super().__init__()
        if use_bn or activation != 'relu' or gauss_ksz != 3:
            raise NotImplementedError('B200 engine covers use_bn=False, activation="relu", gauss_ksz=3')
        self.gauss_ksz = gauss_ksz
        self.init_layer = blocks.conv_block(input_dim, init_conv_dim, 3, stride=1, padding=1, batch_norm=use_bn,
                                            activation=activation)
        d_in = init_conv_dim
        self.pre_res_layers = nn.Sequential(*[blocks.ResBlock(d_in, d_in, stride=1, batch_norm=use_bn,
                                                              activation=activation) for _ in range(num_pre_res_blocks)])
        self.upsample_layer = PixShuffleUpsampler(d_in, post_conv_dim, upsample_factor=upsample_factor, use_bn=use_bn,
                                                  activation=activation, icnrinit=icnrinit, gauss_blur_sd=gauss_blur_sd,
                                                  gauss_ksz=gauss_ksz)
        self.post_res_layers = nn.Sequential(*[blocks.ResBlock(post_conv_dim, post_conv_dim, stride=1, batch_norm=use_bn,
                                                               activation=activation)
                                               for _ in range(num_post_res_blocks)])
        self.predictor = blocks.conv_block(post_conv_dim, 3, 1, stride=1, padding=0, batch_norm=False)
        self.precision = 'bf16'
        self._engine = None

    def gauss_taps(self):
        gk = getattr(self.upsample_layer, 'gauss_kernel', None)
        return False if gk is None else gk.reshape(3, 3)

    def engine(self, device):
        if not self._engine_is_current(device, precision=self.precision):
            sd = {'decoder.' + k: v for k, v in self.state_dict().items()}
            self._set_engine(DBSREngine(sd, device, precision=self.precision, gauss_kernel=self.gauss_taps(),
                                        parts=('decoder',)))
        return self._engine

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, x):
        feat = x['fused_enc']
        assert feat.dim() == 4
        ops.require_device(feat)
        eng = self.engine(feat.device)
        B, C, H, W = feat.shape
        ws = eng.workspace((B, H, W))
        fused = eng._buf(ws, 'fused', B, H, W, C, eng.act_dtype)
        fused.from_nchw(feat.contiguous().float())
        pred = torch.empty((B, eng.pred_w.shape[0], H * eng.up_r, W * eng.up_r), dtype=torch.float32, device=feat.device)
        eng.decode(ws, fused, pred)
        return {'pred': pred}

"""ResEncoderWarpAlignnet with the reference's interface (models/dbsr/encoders.py:21-86)."""
import torch
import torch.nn as nn

from ... import ops
from ...engine import DBSREngine
from ..engine_owner import EngineOwner
from ..layers import blocks


class ResEncoderWarpAlignnet(EngineOwner, nn.Module):
    """Encodes the burst with a residual network, estimates the flow of every frame w.r.t. the first one with the
    alignment net and warps the embeddings to the reference frame.  forward(x [B, N, 4, H, W]) ->
    {'ref_feat' [B, N-1 (expanded), C, H, W], 'oth_feat' [B, N-1, C, H, W], 'offsets' [B, N-1, 2, H, W]}."""

    def __init__(self, init_dim, num_res_blocks, out_dim, alignment_net, use_bn=False, activation='relu',
                 train_alignmentnet=True, warp_type='bilinear'):
        super().__init__()
        if use_bn or activation != 'relu' or warp_type != 'bilinear':
            raise NotImplementedError('B200 engine covers use_bn=False, activation="relu", warp_type="bilinear" '
                                      '(the configuration of train_settings/dbsr/default_synthetic.py)')
        input_channels = 4
        self.warp_type = warp_type
        self.alignment_net = alignment_net
        self.train_alignmentnet = train_alignmentnet
        self.init_layer = blocks.conv_block(input_channels, init_dim, 3, stride=1, padding=1, batch_norm=use_bn,
                                            activation=activation)
        self.res_layers = nn.Sequential(*[blocks.ResBlock(init_dim, init_dim, stride=1, batch_norm=use_bn,
                                                          activation=activation) for _ in range(num_res_blocks)])
        self.out_layer = blocks.conv_block(init_dim, out_dim, 3, stride=1, padding=1, batch_norm=use_bn,
                                           activation=activation)
        self.precision = 'bf16'
        self._engine = None

    def engine(self, device):
        if not self._engine_is_current(device, precision=self.precision):
            sd = {'encoder.' + k: v for k, v in self.state_dict().items()}
            self._set_engine(DBSREngine(sd, device, precision=self.precision, parts=('pwc', 'encoder')))
        return self._engine

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, x):
        assert x.dim() == 5
        ops.require_device(x)
        eng = self.engine(x.device)
        x = x.contiguous().float()
        B, N, _, H, W = x.shape
        import math
        Hp, Wp = int(math.ceil(H / 64.0) * 64), int(math.ceil(W / 64.0) * 64)
        ws = eng.workspace((B, N, H, W))
        enc_in = eng._buf(ws, 'enc_in', B * N, H, W, 8, eng.act_dtype)
        pwc_in = eng._buf(ws, 'pwc_in', B * N, Hp, Wp, 4, torch.float32)
        ops.prep_burst(x, enc_in, pwc_in)
        offsets = torch.empty((B * (N - 1), 2, H, W), dtype=torch.float32, device=x.device)
        eng.pwc_burst(ws, pwc_in, B, N, H, W, offsets)
        feat = eng.encode(ws, enc_in)
        all_feat = eng._buf(ws, 'all_feat', B * N, H, W, eng.feat_dim, eng.act_dtype)
        ops.warp(feat, offsets, all_feat, frames=N)
        full = all_feat.to_nchw().view(B, N, eng.feat_dim, H, W)
        ref_feat = full[:, :1].expand(-1, N - 1, -1, -1, -1)
        oth_feat = full[:, 1:]
        return {'ref_feat': ref_feat, 'oth_feat': oth_feat, 'offsets': offsets.view(B, N - 1, 2, H, W)}

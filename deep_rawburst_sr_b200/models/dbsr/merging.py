"""WeightedSum with the reference's interface (models/dbsr/merging.py:21-127)."""
import torch
import torch.nn as nn

from ... import ops
from ...engine import DBSREngine
from ..engine_owner import EngineOwner
from ..layers import blocks


class WeightedSum(EngineOwner, nn.Module):
    """Adaptive weighted-sum fusion of the aligned burst embeddings.  forward({'ref_feat','oth_feat','offsets'}) ->
    {'fused_enc' [B, C, H, W], 'fusion_weights' [B, N, C, H, W]}."""

    def __init__(self, input_dim, project_dim, offset_feat_dim, num_offset_feat_extractor_res=1,
                 num_weight_predictor_res=1, use_offset=True, offset_modulo=None, ref_offset_noise=0.0, softmax=True,
                 use_base_frame=False, use_bn=False, activation='relu'):
        super().__init__()
        unsupported = []
        if not use_offset: unsupported.append('use_offset=False')
        if ref_offset_noise > 0.0: unsupported.append('ref_offset_noise>0')
        if not softmax: unsupported.append('softmax=False')
        if not use_base_frame: unsupported.append('use_base_frame=False')
        if use_bn: unsupported.append('use_bn=True')
        if activation != 'relu': unsupported.append('activation!=relu')
        if unsupported:
            raise NotImplementedError('B200 fused merging covers the dbsrnet_cvpr2021 flag set only; unsupported: '
                                      + ', '.join(unsupported))
        self.use_offset = use_offset
        self.offset_modulo = offset_modulo
        self.ref_offset_noise = ref_offset_noise
        self.softmax = softmax
        self.use_base_frame = use_base_frame
        self.feat_project_layer = blocks.conv_block(input_dim, project_dim, 1, stride=1, padding=0, batch_norm=use_bn,
                                                    activation=activation)
        ofe = [blocks.conv_block(2, offset_feat_dim, 3, stride=1, padding=1, batch_norm=use_bn, activation=activation)]
        for _ in range(num_offset_feat_extractor_res):
            ofe.append(blocks.ResBlock(offset_feat_dim, offset_feat_dim, stride=1, batch_norm=use_bn,
                                       activation=activation))
        self.offset_feat_extractor = nn.Sequential(*ofe)
        wp = [blocks.conv_block(project_dim * 2 + offset_feat_dim * use_offset, 2 * project_dim, 3, stride=1, padding=1,
                                batch_norm=use_bn, activation=activation)]
        for _ in range(num_weight_predictor_res):
            wp.append(blocks.ResBlock(2 * project_dim, 2 * project_dim, stride=1, batch_norm=use_bn,
                                      activation=activation))
        wp.append(blocks.conv_block(2 * project_dim, input_dim, 3, stride=1, padding=1, batch_norm=use_bn,
                                    activation='none'))
        self.weight_predictor = nn.Sequential(*wp)
        self.precision = 'bf16'
        self.return_fusion_weights = True
        self._engine = None

    def engine(self, device):
        if not self._engine_is_current(device, precision=self.precision):
            sd = {'merging.' + k: v for k, v in self.state_dict().items()}
            self._set_engine(DBSREngine(sd, device, precision=self.precision, offset_modulo=self.offset_modulo,
                                        parts=('merging',)))
        return self._engine

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, x):
        ref_feat, oth_feat, offsets = x['ref_feat'], x['oth_feat'], x['offsets']
        assert ref_feat.dim() == 5
        ops.require_device(oth_feat)
        if ref_feat.shape[1] != 1:
            ref_feat = ref_feat[:, :1, ...]
        all_feat = torch.cat((ref_feat, oth_feat), dim=1).contiguous().float()
        B, N, C, H, W = all_feat.shape
        eng = self.engine(all_feat.device)
        ws = eng.workspace((B, N, H, W))
        af = eng._buf(ws, 'all_feat', B * N, H, W, C, eng.act_dtype)
        af.from_nchw(all_feat.view(B * N, C, H, W))
        weights = None
        if self.return_fusion_weights:
            weights = torch.empty((B, N, C, H, W), dtype=torch.float32, device=all_feat.device)
        fused = eng.merge(ws, af, offsets.reshape(B * (N - 1), 2, H, W).contiguous().float(), B, N, weights, aligned=True)
        return {'fused_enc': fused.to_nchw(), 'fusion_weights': weights}

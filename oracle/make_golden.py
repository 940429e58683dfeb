"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference modules
(/root/reference, read-only) on CPU.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Shims (BASELINE.json config 1): a stub `cupy` module so that
external/pwcnet/correlation/correlation.py imports, and `correlation.FunctionCorrelation` replaced by a
pure-torch cost volume restating correlation.py:69-100 (the cupy kernel is CUDA-only).  The network is
assembled by hand exactly as models/dbsr/dbsrnet.py:62-81 does, minus the pretrained PWC load (weights
are not on the box); `env_settings()` / `dbsrnet_cvpr2021()` are never called (they would write
admin/local.py into the reference tree).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('DBSR_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from oracle import dbsr_oracle as O  # noqa: E402


def import_reference():
    sys.dont_write_bytecode = True
    cupy = types.ModuleType('cupy')
    cupy.util = types.SimpleNamespace(memoize=lambda **kw: (lambda f: f))
    cupy.cuda = types.SimpleNamespace()
    sys.modules['cupy'] = cupy
    sys.path.insert(0, REF)
    from external.pwcnet.correlation import correlation as ref_corr

    def torch_cost_volume(tenFirst, tenSecond):
        n, c, H, W = tenFirst.shape
        f2p = F.pad(tenSecond, (4, 4, 4, 4))
        out = []
        for dy in range(9):
            for dx in range(9):
                out.append((tenFirst * f2p[:, :, dy:dy + H, dx:dx + W]).mean(1, keepdim=True))
        return torch.cat(out, 1)

    ref_corr.FunctionCorrelation = torch_cost_volume
    import models.dbsr.dbsrnet as ref_dbsrnet
    import models.dbsr.encoders as ref_enc
    import models.dbsr.merging as ref_mer
    import models.dbsr.decoders as ref_dec
    from models.alignment.pwcnet import PWCNet
    return dict(dbsrnet=ref_dbsrnet, enc=ref_enc, mer=ref_mer, dec=ref_dec, PWCNet=PWCNet, corr=ref_corr)


def build_reference_net(ref):
    pwc = ref['PWCNet'](load_pretrained=False)
    enc = ref['enc'].ResEncoderWarpAlignnet(O.ENC_INIT_DIM, O.ENC_NUM_RES, O.ENC_OUT_DIM, pwc, activation='relu',
                                            train_alignmentnet=False)
    mer = ref['mer'].WeightedSum(O.ENC_OUT_DIM, O.PROJ_DIM, O.OFFSET_FEAT_DIM,
                                 num_offset_feat_extractor_res=O.NUM_OFFSET_RES,
                                 num_weight_predictor_res=O.NUM_WP_RES, offset_modulo=O.OFFSET_MODULO,
                                 use_offset=True, ref_offset_noise=0.0, softmax=True, use_base_frame=True)
    dec = ref['dec'].ResPixShuffleConv(O.ENC_OUT_DIM, O.DEC_INIT_DIM, O.DEC_NUM_PRE_RES, O.DEC_POST_DIM,
                                       O.DEC_NUM_POST_RES, upsample_factor=O.UPSAMPLE, activation='relu',
                                       gauss_blur_sd=O.GAUSS_SD, icnrinit=True, gauss_ksz=O.GAUSS_KSZ)
    return ref['dbsrnet'].DBSRNet(enc, mer, dec).eval()


CASES = [
    # name, weight seed, pwc_gain, burst seed, B, N, H, W
    ('tiny_b1n3_16x16', 0, 1.0, 0, 1, 3, 16, 16),
    ('rect_b2n4_24x40', 1, 1.0, 1, 2, 4, 24, 40),
    ('stress_b1n5_32x32', 2, 2.2, 2, 1, 5, 32, 32),
    ('cfg1_b1n14_48x48', 0, 1.0, 3, 1, 14, 48, 48),
]


def main():
    torch.manual_seed(0)
    ref = import_reference()
    net = build_reference_net(ref)
    # key/shape contract
    spec = O.state_dict_spec()
    ref_sd = net.state_dict()
    assert [k for k, _ in spec] == list(ref_sd.keys()), 'state_dict key order mismatch'
    for k, s in spec:
        assert tuple(ref_sd[k].shape) == tuple(s), (k, ref_sd[k].shape, s)
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, 'state_dict_keys.txt'), 'w') as f:
        for k, s in spec:
            f.write(f"{k} {'x'.join(map(str, s))}\n")
    for name, wseed, gain, bseed, B, N, H, W in CASES:
        sd = O.make_state_dict(wseed, pwc_gain=gain)
        net.load_state_dict(sd, strict=True)
        burst = O.make_burst(bseed, B, N, H, W)
        with torch.no_grad():
            pred, aux = net(burst)
        offsets = aux['offsets']
        fw = aux['fusion_weights']
        out = {
            'meta': np.array([wseed, bseed, B, N, H, W], dtype=np.int64),
            'pwc_gain': np.array([gain], dtype=np.float64),
            'offsets': offsets.numpy().astype(np.float32),
            # fusion weights are big: keep a strided sub-sample + per-frame means
            'fusion_weights_sub': fw[:, :, ::37, ::3, ::3].contiguous().numpy().astype(np.float32),
            'fusion_weights_mean': fw.mean(dim=(2, 3, 4)).numpy().astype(np.float64),
        }
        if pred.numel() <= 3 * 256 * 256 * 2:
            out['pred'] = pred.numpy().astype(np.float32)
        else:
            out['pred_sub'] = pred[:, :, ::2, ::2].contiguous().numpy().astype(np.float32)
            out['pred_sum'] = np.array([pred.double().sum().item(), (pred.double() ** 2).sum().item()])
        np.savez_compressed(os.path.join(outdir, name + '.npz'), **out)
        print(name, 'pred', tuple(pred.shape), 'range', float(pred.min()), float(pred.max()),
              'max|flow|', float(offsets.abs().max()))

    # op-level goldens from the reference functions themselves (warp, backwarp, PWCNet resize path)
    import models.layers.warp as ref_warp
    import models.alignment.pwcnet as ref_pwc
    g = torch.Generator().manual_seed(7)
    feat = torch.randn(2, 5, 9, 11, generator=g)
    flow = (torch.rand(2, 2, 9, 11, generator=g) - 0.5) * 8.0
    flow[0, :, 0, 0] = torch.tensor([1.0, -2.0])        # integer flow
    flow[0, :, 1, 1] = torch.tensor([0.5, 0.5])         # half pixel
    flow[1, :, 2, 2] = torch.tensor([-30.0, 40.0])      # far out of bounds
    ops = {
        'feat': feat.numpy(), 'flow': flow.numpy(),
        'warp': ref_warp.warp(feat, flow).numpy(),
        'backwarp': ref_pwc.backwarp(feat, flow).numpy(),
        'interp_up': F.interpolate(feat, size=(64, 64), mode='bilinear', align_corners=False).numpy(),
        'interp_down': F.interpolate(feat, size=(5, 7), mode='bilinear', align_corners=False).numpy(),
        'mod': (torch.tensor([-0.25, 1.75, -1e-9, 0.0, 3.0]) % 1.0).numpy(),
    }
    f1 = torch.randn(2, 6, 5, 7, generator=g)
    f2 = torch.randn(2, 6, 5, 7, generator=g)
    ops['corr_f1'] = f1.numpy()
    ops['corr_f2'] = f2.numpy()
    ops['corr'] = ref['corr'].FunctionCorrelation(f1, f2).numpy()
    np.savez_compressed(os.path.join(outdir, 'ops.npz'), **ops)
    print('done')


if __name__ == '__main__':
    main()
